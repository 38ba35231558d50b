"""Pins the C restatement (oracle/np_oracle.c) against the reference itself:
 - committed golden fixtures minted by tests/golden/make_golden.py from oracle/_ref (test_data
   reads mapped with the vendored bwa, ~30x);
 - md5s of the reference's output on the seeded synthetic cases;
 - live runs of oracle/_ref/nextpolish1 when that binary is present (build container, GPU box)."""
import json
import os
import subprocess

import pytest

from tests.conftest import GOLDEN, REF_BIN, REF_SAMTOOLS, md5, read_fasta, run_checker
from tests.synth_cases import CASES


@pytest.mark.parametrize("step", [1, 2])
def test_oracle_vs_golden_testdata(E, oracle, step):
    fa = os.path.join(GOLDEN, "td30.step%d.fa" % step)
    bam = os.path.join(GOLDEN, "td30.step%d.bam" % step)
    exp = read_fasta(os.path.join(GOLDEN, "td30.step%d.expected.fa" % step))
    sh = E.Shard.load(fa, bam, with_qual=True)
    cfg = E.default_config(fa, bam)
    got = run_checker(oracle.np_oracle_run, sh, step, cfg)
    assert set(exp) == {"%s_%d" % (n, step) for n in got}
    for n, s in got.items():
        assert s == exp["%s_%d" % (n, step)], n


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("step", [1, 2])
def test_oracle_vs_reference_md5(E, oracle, synth_files, case, step):
    want = json.load(open(os.path.join(GOLDEN, "synth_md5.json")))[case][str(step)]
    fa, bam = synth_files(case)
    sh = E.Shard.load(fa, bam, with_qual=True)
    cfg = E.default_config(fa, bam)
    got = run_checker(oracle.np_oracle_run, sh, step, cfg)
    assert {"%s_%d" % (n, step): md5(s) for n, s in got.items()} == want


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref not built")
@pytest.mark.parametrize("step,cmd", [(1, "scorechain"), (2, "kmercount")])
def test_oracle_vs_live_reference(E, oracle, tmp_path, step, cmd):
    # a case that is NOT in the committed md5 list
    p = E.synth_params(seed=4242 + step, n_contigs=4, contig_len=30000, depth=40.0, lowercase_frac=0.03,
                       draft_indel=0.008, read_indel=0.001)
    fa, bam = str(tmp_path / "x.fa"), str(tmp_path / "x.bam")
    assert E.lib().np_synth_write(p, fa.encode(), bam.encode()) == 0
    subprocess.check_call([REF_SAMTOOLS, "index", bam])
    out = subprocess.run([REF_BIN, cmd, fa, bam], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    ref = str(tmp_path / "ref.fa")
    open(ref, "wb").write(out)
    exp = read_fasta(ref)
    sh = E.Shard.load(fa, bam, with_qual=True)
    cfg = E.default_config(fa, bam)
    got = run_checker(oracle.np_oracle_run, sh, step, cfg)
    for n, s in got.items():
        assert s == exp["%s_%d" % (n, step)], n
