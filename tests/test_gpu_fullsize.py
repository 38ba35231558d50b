"""GPU parity at the BENCHMARKED shapes: bit-exact, per contig, through the C ABI.

* BASELINE config 2 (5 x 1 Mb, 30x 150 bp, both tasks, the very seeds / generator bench.py uses): the from-files path
  (np_files_submit / np_files_wait: GPU inflate + unpack + packing + kernels) against the compiled, unmodified
  reference binary (oracle/_ref/nextpolish1) run on the same FASTA + BAM, and the resident packed-shard path against it.
* BASELINE config 3 shape (log-uniform 20 kb - 1 Mb contigs): a slice of 100 contigs (~25 Mb, 5 M reads; the oracle port polishes
  ~3 Mbp/s on one core) against the oracle port, both tasks.
"""
import hashlib
import os
import subprocess

import pytest

import bench
from tests.conftest import REF_BIN, REF_SAMTOOLS, run_checker

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(REF_SAMTOOLS)), reason="oracle/_ref not built")
def test_c2_shape_files_and_resident_vs_reference_binary(E, tmp_path):
    tasks = [1, 2]
    rr = bench.ReferenceRunner(str(tmp_path), tasks)        # writes rank 0's bench inputs with the standalone generator
    want = {t: rr.outputs_md5(t) for t in tasks}
    assert all(len(want[t]) == bench.WORKLOAD["n_contigs"] for t in tasks)
    # (a) from the files
    fp = E.FilePipeline(0, depth=2)
    cfgs = {}
    for t in tasks:
        fa, bam = rr.files[t]
        cfgs[t] = E.default_config(fa, bam)            # read_tlen estimated from the BAM head like config_init does
        fp.submit(t, fa, bam, cfgs[t])
    for t in tasks:
        r = fp.wait_oldest()
        assert r["task"] == t
        assert r["md5"] == want[t], "from-files task %d differs from the reference binary" % t
    fp.close()
    # (b) the resident packed-shard path on the generator's own packing of the same seeds
    eng = E.Engine(0)
    for t in tasks:
        p = E.synth_params(**bench.synth_kwargs(t, bench.seed_for(0, t)))
        sh = E.Shard.synthetic(p, 0, bench.WORKLOAD["n_contigs"], with_qual=(2 if t == 2 else 0))
        got = eng.polish(sh, t, cfgs[t])
        assert {n: hashlib.md5(s).hexdigest() for n, s in got.items()} == want[t], "resident task %d differs from the reference binary" % t
        sh.close()
    eng.close()


def test_c3_shape_slice_vs_oracle(E, oracle):
    """100 log-uniform contigs (20 kb - 1 Mb, ~25 Mb) of the config-3 generator, 30x, both tasks, against the oracle port."""
    n = 100
    p1 = E.synth_params(seed=20240917 + 3, n_contigs=n, contig_len=0, min_len=20000, max_len=1000000, depth=30.0)
    eng = E.Engine(0)
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750
    lo = 0
    for task, p in ((1, p1), (2, E.synth_params(seed=20240917 + 3, n_contigs=n, contig_len=0, min_len=20000, max_len=1000000, depth=30.0,
                                                 draft_snv=1e-5, draft_indel=2e-5, lowercase_frac=6.3e-4))):
        # 25 contigs at a time bounds the oracle's memory (it materialises per-column vote lists)
        for lo in range(0, n, 25):
            sh = E.Shard.synthetic(p, lo, lo + 25, with_qual=(2 if task == 2 else 0))
            want = run_checker(oracle.np_oracle_run, sh, task, cfg)
            got = eng.polish(sh, task, cfg)
            assert got == want, (task, lo)
            sh.close()
    eng.close()
