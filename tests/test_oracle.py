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


def test_oracle_snp_valid_vs_golden_testdata(E, oracle):
    """snp_valid (task 4, snpvalid.c:3-35) on the step-1 fixture against `nextpolish1 snpvalid` of the reference."""
    fa, bam = os.path.join(GOLDEN, "td30.step1.fa"), os.path.join(GOLDEN, "td30.step1.bam")
    exp = read_fasta(os.path.join(GOLDEN, "td30.step1.snpvalid.expected.fa"))
    sh = E.Shard.load(fa, bam, with_qual=True)
    got = run_checker(oracle.np_oracle_run, sh, 4, E.default_config(fa, bam))
    assert {"%s_4" % n: s for n, s in got.items()} == exp


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("step,qual_mode", [(1, 1), (2, 1), (2, 2), (4, 1)])
def test_oracle_vs_reference_md5(E, oracle, synth_files, case, step, qual_mode):
    """qual_mode 2 = sparse quality stream (only reads overlapping lowercase draft bases): the reference's
    output must still be reproduced, i.e. no other read's qualities are ever consulted."""
    want = json.load(open(os.path.join(GOLDEN, "synth_md5.json")))[case][str(step)]
    fa, bam = synth_files(case)
    sh = E.Shard.load(fa, bam, with_qual=qual_mode)
    cfg = E.default_config(fa, bam)
    got = run_checker(oracle.np_oracle_run, sh, step, cfg)
    assert {"%s_%d" % (n, step): md5(s) for n, s in got.items()} == want


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref not built")
@pytest.mark.parametrize("step,cmd", [(1, "scorechain"), (2, "kmercount"), (4, "snpvalid")])
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
    if step == 4:
        # the reference's second pass is undefined on some inputs (odd cut-point lists, snpvalid.c:37-66): the oracle reports
        # those contigs instead of guessing; every contig it does polish must match
        import ctypes as C
        import numpy as np
        oracle.np_oracle_run_contig.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        checked = 0
        for c, n in enumerate(sh.names):
            cap = int(sh.view.ctg_off[c + 1] - sh.view.ctg_off[c]) * 2 + 4096
            buf, ln = np.zeros(cap, np.uint8), C.c_int64(0)
            rc = oracle.np_oracle_run_contig(C.addressof(sh.view), c, 4, C.cast(cfg, C.c_void_p), buf.ctypes.data, cap, C.byref(ln))
            assert rc in (0, -2)
            if rc == 0:
                assert buf[:ln.value].tobytes() == exp["%s_4" % n], n
                checked += 1
        assert checked > 0
        return
    got = run_checker(oracle.np_oracle_run, sh, step, cfg)
    for n, s in got.items():
        assert s == exp["%s_%d" % (n, step)], n


REF_SO = os.path.join(os.path.dirname(REF_BIN), "nextpolish1.so")


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_oracle_vs_reference_so_with_mutated_thresholds(E, oracle, tmp_path, seed):
    """Calls the reference's own score_chain/kmer_count (oracle/_ref/nextpolish1.so, through the same
    ctypes declarations nextpolish1.py uses) with non-default Configure fields, like update_cfg does."""
    import ctypes as C
    import random
    from nextpolish_b200.binding import Configure, PolishResult
    R = C.CDLL(REF_SO)
    R.config_init.argtypes = [C.c_char_p] * 3
    R.config_init.restype = C.POINTER(Configure)
    for f in ("score_chain", "kmer_count"):
        getattr(R, f).argtypes = [C.c_char_p, C.POINTER(Configure)]
        getattr(R, f).restype = C.POINTER(PolishResult)
    R.polishresult_destory.argtypes = [C.POINTER(PolishResult)]
    rng = random.Random(seed)
    kw = dict(seed=rng.randrange(1 << 30), n_contigs=rng.choice([1, 3]), contig_len=rng.choice([3000, 8000, 20000]),
              depth=rng.choice([2, 5, 12, 30, 60]), draft_snv=rng.choice([0.001, 0.01]), draft_indel=rng.choice([0.003, 0.02]),
              read_sub=rng.choice([0.002, 0.02]), read_indel=rng.choice([0.0001, 0.003]), lowercase_frac=rng.choice([0.01, 0.05, 0.15]))
    fa, bam = str(tmp_path / "x.fa"), str(tmp_path / "x.bam")
    assert E.lib().np_synth_write(E.synth_params(**kw), fa.encode(), bam.encode()) == 0
    subprocess.check_call([REF_SAMTOOLS, "index", bam])
    sh = E.Shard.load(fa, bam, with_qual=True)
    cfg, rcfg = E.default_config(fa, bam), R.config_init(fa.encode(), bam.encode(), None)
    assert (cfg.contents.read_tlen, cfg.contents.read_len) == (rcfg.contents.read_tlen, rcfg.contents.read_len)
    vals = dict(trim_len_edge=rng.choice([1, 2, 4]), ext_len_edge=rng.choice([1, 2, 3]), min_len_ldr=rng.choice([1, 3, 6]),
                min_len_inter_kmer=rng.choice([0, 2, 5, 9]), max_len_kmer=rng.choice([10, 50, 120]), max_count_kmer=rng.choice([3, 50]),
                min_map_quality=rng.choice([0, 30]), indel_balance_factor_sgs=rng.choice([0.5, 0.25, 0.75, 0.33, 0.7]),
                min_count_ratio_skip=rng.choice([0.8, 0.6, 0.95]))
    for k, v in vals.items():
        setattr(cfg.contents, k, v)
        setattr(rcfg.contents, k, v)
    for task, fn in ((1, R.score_chain), (2, R.kmer_count)):
        want = run_checker(oracle.np_oracle_run, sh, task, cfg)
        for n in sh.names:
            res = fn(n.encode(), rcfg)
            seq = C.string_at(res.contents.contig)
            R.polishresult_destory(res)
            assert seq == want[n], (task, n, vals)
