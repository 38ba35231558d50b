"""Device-side shard construction (np_shard_load_gpu, csrc/devload.cu): BGZF inflate, record-boundary chains between
the .bai anchors, field extraction and packing on the GPU must reproduce the host packer's shard byte for byte, and
the engine must polish the adopted device shard to the same bytes."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests.conftest import GOLDEN, REF_SAMTOOLS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(E):
    e = E.Engine(0)
    yield e
    e.close()


def _same(E, fa, bam, names, with_qual):
    host = E.Shard.load(fa, bam, names=names, with_qual=with_qual)
    dev = E.DeviceShard(fa, bam, names=names, with_qual=with_qual)
    assert dev.names == host.names
    a, b = host.arrays(), dev.arrays()
    assert set(a) == set(b), (sorted(a), sorted(b))
    for k in a:
        assert a[k].shape == b[k].shape and bytes(a[k]) == bytes(b[k]), k
    return host, dev


@pytest.mark.parametrize("with_qual", [0, 1, 2])
def test_device_shard_equals_host_shard_on_golden_bam(E, with_qual):
    fa, bam = os.path.join(GOLDEN, "td30.step2.fa"), os.path.join(GOLDEN, "td30.step2.bam")
    host, dev = _same(E, fa, bam, None, with_qual)
    assert dev.n_reads > 10000
    for nm in host.names:                       # one contig at a time: byte range from the index
        _same(E, fa, bam, [nm], with_qual)


@pytest.mark.skipif(not os.path.exists(REF_SAMTOOLS), reason="needs oracle/_ref/samtools to index the synthetic BAM")
@pytest.mark.parametrize("case", ["c30", "ragged", "lower"])
def test_device_shard_equals_host_shard_on_synthetic_bam(E, synth_files, tmp_path, case):
    fa0, bam0 = synth_files(case)
    fa, bam = str(tmp_path / "x.fa"), str(tmp_path / "x.bam")
    shutil.copy(fa0, fa); shutil.copy(bam0, bam)
    subprocess.check_call([REF_SAMTOOLS, "index", bam])
    host, dev = _same(E, fa, bam, None, 2)
    names = host.names
    _same(E, fa, bam, names[1:2], 2)
    _same(E, fa, bam, [names[-1], names[0]], 0)     # a subset out of order spans the whole file range


def test_polish_from_device_shard(E, eng):
    fa, bam = os.path.join(GOLDEN, "td30.step2.fa"), os.path.join(GOLDEN, "td30.step2.bam")
    cfg = E.default_config(fa.encode(), bam.encode())
    host = E.Shard.load(fa, bam, with_qual=2)
    dev = E.DeviceShard(fa, bam, with_qual=2)
    for task in E.TASKS:
        want = eng.polish(host, task, cfg)
        eng.adopt_device(dev.view)
        eng.run(task, cfg)
        out, off = eng.download(dev.n_contigs)
        raw = out.tobytes()
        assert {nm: raw[off[i]:off[i + 1]] for i, nm in enumerate(dev.names)} == want, task


def test_missing_index_is_reported(E, tmp_path):
    bam = str(tmp_path / "noidx.bam")
    shutil.copy(os.path.join(GOLDEN, "td30.step1.bam"), bam)
    with pytest.raises(E.NativeError):
        E.DeviceShard(os.path.join(GOLDEN, "td30.step1.fa"), bam)


@pytest.mark.skipif(not os.path.exists(REF_SAMTOOLS), reason="needs oracle/_ref/samtools to index the synthetic BAM")
def test_device_shard_edge_inputs(E, tmp_path):
    """Contigs without reads, a FASTA contig the BAM header does not know, and a shard with no reads at all."""
    kw = dict(seed=61, n_contigs=5, contig_len=0, min_len=300, max_len=20000, depth=8.0, lowercase_frac=0.02)
    fa, bam = str(tmp_path / "e.fa"), str(tmp_path / "e.bam")
    assert E.lib().np_synth_write(E.synth_params(**kw), fa.encode(), bam.encode()) == 0
    subprocess.check_call([REF_SAMTOOLS, "index", bam])
    with open(fa, "a") as f:                       # a contig that is not in the BAM header: goes last, no reads
        f.write(">extra_contig\nACGTACGTTTGACCAacgtNNACGT\n")
    host, dev = _same(E, fa, bam, None, 2)
    assert host.names[-1] == "extra_contig"
    _same(E, fa, bam, ["extra_contig"], 2)         # nothing to read from the BAM
    _same(E, fa, bam, ["extra_contig", host.names[0]], 1)
    # a BAM without any record for the requested contigs (depth 0)
    fa0, bam0 = str(tmp_path / "z.fa"), str(tmp_path / "z.bam")
    assert E.lib().np_synth_write(E.synth_params(seed=62, n_contigs=2, contig_len=1000, depth=0.0), fa0.encode(), bam0.encode()) == 0
    subprocess.check_call([REF_SAMTOOLS, "index", bam0])
    _same(E, fa0, bam0, None, 0)
