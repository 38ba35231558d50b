"""GPU parity tests: the sm_100a engine, called through the C ABI, against the oracle (and the
reference binary when oracle/_ref is present) — bit-exact (integer / index work; the score chain's
doubles are dyadic rationals, compared through the emitted sequence)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from tests.conftest import GOLDEN, REF_BIN, REF_SAMTOOLS, md5, read_fasta, run_checker
from tests.synth_cases import CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(E):
    e = E.Engine(0)
    yield e
    e.close()


def tasks(E):
    return list(E.TASKS)


@pytest.mark.parametrize("case", sorted(CASES))
def test_gpu_matches_oracle_on_synthetic(E, oracle, eng, case):
    sh = E.Shard.synthetic(E.synth_params(**CASES[case]), 0, CASES[case]["n_contigs"], with_qual=True)
    cfg = E.default_config(b"")
    for task in tasks(E):
        want = run_checker(oracle.np_oracle_run, sh, task, cfg)
        got = eng.polish(sh, task, cfg)
        for n in want:
            assert got[n] == want[n], (task, n)


@pytest.mark.parametrize("case", sorted(CASES))
def test_gpu_matches_reference_md5(E, eng, synth_files, case):
    """Against the committed md5s of the reference binary's own output (BAM path: read_tlen is
    estimated from the BAM head exactly like config_init does)."""
    want = json.load(open(os.path.join(GOLDEN, "synth_md5.json")))[case]
    fa, bam = synth_files(case)
    sh = E.Shard.load(fa, bam, with_qual=2)          # sparse quality stream
    cfg = E.default_config(fa, bam)
    for task in tasks(E):
        got = eng.polish(sh, task, cfg)
        assert {"%s_%d" % (n, task): md5(s) for n, s in got.items()} == want[str(task)], task


@pytest.mark.parametrize("seed", [21, 22, 23, 24])
def test_gpu_matches_oracle_mutated_thresholds(E, oracle, eng, seed):
    import random
    rng = random.Random(seed)
    kw = dict(seed=rng.randrange(1 << 30), n_contigs=rng.choice([2, 5]), contig_len=rng.choice([8000, 20000]),
              depth=rng.choice([5, 12, 30, 60]), draft_indel=rng.choice([0.003, 0.02]), read_indel=rng.choice([0.0001, 0.003]),
              lowercase_frac=rng.choice([0.01, 0.05, 0.15]))
    sh = E.Shard.synthetic(E.synth_params(**kw), 0, kw["n_contigs"], with_qual=True)
    cfg = E.default_config(b"")
    c = cfg.contents
    c.read_tlen = 2000
    c.trim_len_edge, c.ext_len_edge = rng.choice([0, 1, 2, 4]), rng.choice([0, 1, 2, 3])
    c.min_len_ldr, c.min_len_inter_kmer = rng.choice([1, 3, 6]), rng.choice([0, 2, 5, 9])
    c.max_len_kmer, c.max_count_kmer, c.min_map_quality = rng.choice([10, 50, 120]), rng.choice([3, 50]), rng.choice([0, 30])
    c.indel_balance_factor_sgs, c.min_count_ratio_skip = rng.choice([0.5, 0.25, 0.75]), rng.choice([0.8, 0.6, 0.95])
    for task in tasks(E):
        want = run_checker(oracle.np_oracle_run, sh, task, cfg)
        got = eng.polish(sh, task, cfg)
        assert got == want, task


def test_gpu_matches_golden_testdata(E, eng):
    for step in tasks(E):
        fa = os.path.join(GOLDEN, "td30.step%d.fa" % step)
        bam = os.path.join(GOLDEN, "td30.step%d.bam" % step)
        exp = read_fasta(os.path.join(GOLDEN, "td30.step%d.expected.fa" % step))
        sh = E.Shard.load(fa, bam, with_qual=True)
        cfg = E.default_config(fa, bam)
        got = eng.polish(sh, step, cfg)
        for n, s in got.items():
            assert s == exp["%s_%d" % (n, step)], (step, n)


@pytest.mark.parametrize("step,fn", [(1, "score_chain"), (2, "kmer_count")])
def test_reference_abi_entry_points(E, step, fn):
    """The drop-in entry points themselves: score_chain / kmer_count(tigname, cfg) -> PolishResult*."""
    fa = os.path.join(GOLDEN, "td30.step%d.fa" % step)
    bam = os.path.join(GOLDEN, "td30.step%d.bam" % step)
    exp = read_fasta(os.path.join(GOLDEN, "td30.step%d.expected.fa" % step))
    L = E.lib()
    cfg = L.config_init(fa.encode(), bam.encode(), None)
    for name in [n[:-2] for n in exp]:
        res = getattr(L, fn)(name.encode(), cfg)
        seq = C.string_at(res.contents.contig)
        assert res.contents.length == len(seq)
        assert seq == exp["%s_%d" % (name, step)]
        L.polishresult_destory(res)
    L.config_destory(cfg)


def test_native_cli_matches_reference_binary(E, tmp_path):
    """nextpolish1 scorechain <fa> <bam> (our CLI) vs the reference binary, byte for byte."""
    cli = os.path.join(os.path.dirname(E.binding.LIB_PATH), "nextpolish1")
    for step, cmd in ((1, "scorechain"), (2, "kmercount")):
        fa = os.path.join(GOLDEN, "td30.step%d.fa" % step)
        bam = os.path.join(GOLDEN, "td30.step%d.bam" % step)
        ours = subprocess.run([cli, cmd, fa, bam], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
        assert ours == open(os.path.join(GOLDEN, "td30.step%d.expected.fa" % step), "rb").read(), cmd


def test_e2e_host_call_equals_resident_path(E, eng):
    sh = E.Shard.synthetic(E.synth_params(seed=3, n_contigs=2, contig_len=50000, depth=30.0), 0, 2)
    cfg = E.default_config(b"")
    a = eng.polish(sh, 1, cfg)
    out = np.zeros(int(sh.total_bases * 2), np.uint8)
    off = np.zeros(sh.n_contigs + 1, np.int64)
    eng.polish_host(1, sh.view, cfg, out, off)
    raw = out.tobytes()
    assert {n: raw[off[i]:off[i + 1]] for i, n in enumerate(sh.names)} == a


def test_full_size_properties(E, eng):
    """BASELINE config 2 shape (5 x 1 Mb, 30x): size-independent properties of the output —
    every contig's polished sequence is nearly the truth length (indels repaired), contains only
    ACGT/acgt, idempotent on re-run, and contig order / offsets are consistent."""
    p = E.synth_params(seed=20240917 + 2, n_contigs=5, contig_len=1000000, depth=30.0)
    sh = E.Shard.synthetic(p, 0, 5)
    cfg = E.default_config(b"")
    a = eng.polish(sh, 1, cfg)
    b = eng.polish(sh, 1, cfg)
    assert a == b
    for n, s in a.items():
        assert abs(len(s) - 1000000) < 200, (n, len(s))
        assert set(s) <= set(b"ACGTacgt"), n
        assert sum(1 for c in s if c >= 97) < 2000


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref not built")
def test_gpu_vs_live_reference_binary(E, eng, tmp_path):
    p = E.synth_params(seed=99, n_contigs=6, contig_len=0, min_len=2000, max_len=120000, depth=35.0,
                       draft_indel=0.006, read_indel=0.0008)
    fa, bam = str(tmp_path / "x.fa"), str(tmp_path / "x.bam")
    assert E.lib().np_synth_write(p, fa.encode(), bam.encode()) == 0
    subprocess.check_call([REF_SAMTOOLS, "index", bam])
    sh = E.Shard.load(fa, bam, with_qual=True)
    cfg = E.default_config(fa, bam)
    for task, cmd in ((1, "scorechain"), (2, "kmercount")):
        if task not in E.TASKS:
            continue
        out = subprocess.run([REF_BIN, cmd, fa, bam], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
        ref = str(tmp_path / ("ref%d.fa" % task))
        open(ref, "wb").write(out)
        exp = read_fasta(ref)
        got = eng.polish(sh, task, cfg)
        for n, s in got.items():
            assert s == exp["%s_%d" % (n, task)], (task, n)


def test_general_kernels_equal_fused_kernel(E, eng):
    """A/B: the fused window kernel (default) and the general global-memory kernels give identical
    sequences; the fused path resolves plain 30x data without any fallback column."""
    sh = E.Shard.synthetic(E.synth_params(seed=31, n_contigs=3, contig_len=200000, depth=30.0), 0, 3)
    cfg = E.default_config(b"")
    a = eng.polish(sh, 1, cfg)
    ws = eng.window_stats()
    assert ws["n_win"] > 0 and ws["fallback_cols"] == 0, ws
    os.environ["NEXTPOLISH_B200_GENERAL_KERNELS"] = "1"
    try:
        b = eng.polish(sh, 1, cfg)
    finally:
        del os.environ["NEXTPOLISH_B200_GENERAL_KERNELS"]
    assert a == b


def test_fused_kernel_deep_and_noisy_use_fallback_correctly(E, oracle, eng):
    p = E.synth_params(seed=32, n_contigs=2, contig_len=60000, depth=200.0, draft_snv=0.01, draft_indel=0.02, read_sub=0.02, read_indel=0.004)
    sh = E.Shard.synthetic(p, 0, 2)
    cfg = E.default_config(b"")
    want = run_checker(oracle.np_oracle_run, sh, 1, cfg)
    assert eng.polish(sh, 1, cfg) == want


def test_nextpolish1_worker_mirror(E, tmp_path, capsys):
    """python -m nextpolish_b200.nextpolish1 with the reference's flags: block file, resume, header names."""
    from nextpolish_b200 import nextpolish1
    fa = os.path.join(GOLDEN, "td30.step1.fa")
    bam = os.path.join(GOLDEN, "td30.step1.bam")
    exp = read_fasta(os.path.join(GOLDEN, "td30.step1.expected.fa"))
    names = [n[:-2] for n in exp]
    blc = str(tmp_path / "g.blc")
    open(blc, "w").write("".join("%s\t%d\n" % (n, i % 2) for i, n in enumerate(names)))
    out = str(tmp_path / "part000.fasta")
    assert nextpolish1.main(["-g", fa, "-s", bam, "-t", "1", "-b", blc, "-i", "0", "-o", out]) == 0
    got = read_fasta(out)
    assert set(got) == {names[0] + "_np1"} and got[names[0] + "_np1"] == exp[names[0] + "_1"]
    # resume: a truncated last record is re-done, finished contigs are skipped
    full = open(out).read()
    open(out, "w").write(full + ">" + names[1] + "_np1 10\nACGT")
    blc_all = str(tmp_path / "all.blc")
    open(blc_all, "w").write("".join("%s\t0\n" % n for n in names))
    assert nextpolish1.main(["-g", fa, "-s", bam, "-t", "1", "-b", blc_all, "-i", "0", "-o", out]) == 0
    got = read_fasta(out)
    assert {k: v for k, v in got.items()} == {n + "_np1": exp[n + "_1"] for n in names}
    header = [l for l in open(out) if l.startswith(">")][0].split()
    assert int(header[1]) == len(got[header[0][1:]])
    # -debug: change points on stderr, "name pos index curbase base" (nextpolish1.py:230-231), sequences unchanged
    out2 = str(tmp_path / "dbg.fa")
    capsys.readouterr()
    assert nextpolish1.main(["-g", fa, "-s", bam, "-t", "1", "-b", blc_all, "-i", "0", "-o", out2, "-debug"]) == 0
    err = capsys.readouterr().err.strip().split("\n")
    assert read_fasta(out2) == got
    assert len(err) > 10 and all(len(l.split()) == 5 and l.split()[0] in names for l in err)


def test_pipelined_host_call_equals_resident_path(E, eng):
    """np_polish_host on a shard big enough to take the two-engine pipelined route (second half's H2D copy
    overlaps the first half's kernels), both tasks, pinned and pageable host buffers."""
    import torch
    p = E.synth_params(seed=41, n_contigs=5, contig_len=400000, depth=20.0, lowercase_frac=0.002)
    sh = E.Shard.synthetic(p, 0, 5, with_qual=2)
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750
    for task in tasks(E):
        want = eng.polish(sh, task, cfg)
        out = np.zeros(int(sh.total_bases * 2), np.uint8)
        off = np.zeros(sh.n_contigs + 1, np.int64)
        eng.polish_host(task, sh.view, cfg, out, off)                     # pageable arrays, single engine
        raw = out.tobytes()
        assert {n: raw[off[i]:off[i + 1]] for i, n in enumerate(sh.names)} == want, task
        os.environ["NEXTPOLISH_B200_PIPELINE"] = "1"
        try:
            eng.polish_host(task, sh.view, cfg, out, off)                 # two engines, overlapped copy
        finally:
            del os.environ["NEXTPOLISH_B200_PIPELINE"]
        raw = out.tobytes()
        assert {n: raw[off[i]:off[i + 1]] for i, n in enumerate(sh.names)} == want, task


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_streaming_front_end_equals_resident_path(E, eng, depth):
    """np_stream_submit / np_stream_wait: jobs of both tasks on different shards, uploads overlapping the
    kernels of the jobs before them; every job's bytes equal the resident-path result."""
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750
    shards = [E.Shard.synthetic(E.synth_params(seed=300 + k, n_contigs=2 + k, contig_len=150000, depth=25.0,
                                               lowercase_frac=0.002 * k), 0, 2 + k, with_qual=2) for k in range(3)]
    jobs = [(t, sh) for sh in shards for t in tasks(E)] * 2
    want = {(t, id(sh)): eng.polish(sh, t, cfg) for t, sh in set(jobs)}
    st = E.Stream(0, depth)
    bufs, tickets = [], []
    for t, sh in jobs:
        out = np.zeros(int(sh.total_bases * 2), np.uint8)
        off = np.zeros(sh.n_contigs + 1, np.int64)
        bufs.append((out, off))
        tickets.append(st.submit(t, sh.view, cfg, out, off))
    for (t, sh), (out, off), tk in reversed(list(zip(jobs, bufs, tickets))):      # waiting out of order is allowed
        st.wait(tk)
        raw = out.tobytes()
        assert {n: raw[off[i]:off[i + 1]] for i, n in enumerate(sh.names)} == want[(t, id(sh))], (t, depth)
    assert st.launch_count() > 0
    st.close()


def test_gpu_two_bit_and_four_bit_records(E, oracle, eng, monkeypatch):
    """2-bit (default) and 4-bit record streams polish identically on the device, both tasks."""
    kw = dict(seed=124, n_contigs=3, contig_len=120000, depth=30.0, draft_indel=0.01, lowercase_frac=0.01)
    two = E.Shard.synthetic(E.synth_params(**kw), 0, 3, with_qual=2)
    monkeypatch.setenv("NEXTPOLISH_B200_4BIT", "1")
    four = E.Shard.synthetic(E.synth_params(**kw), 0, 3, with_qual=2)
    monkeypatch.delenv("NEXTPOLISH_B200_4BIT")
    assert len(two.arrays()["rec"]) < 0.75 * len(four.arrays()["rec"])
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750
    for task in tasks(E):
        want = run_checker(oracle.np_oracle_run, four, task, cfg)
        assert eng.polish(two, task, cfg) == want, task
        assert eng.polish(four, task, cfg) == want, task


@pytest.mark.parametrize("rate", [0.33, 0.7])
def test_gpu_non_dyadic_rate(E, oracle, eng, rate):
    sh = E.Shard.synthetic(E.synth_params(seed=92, n_contigs=3, contig_len=60000, depth=40.0, draft_indel=0.01, read_sub=0.01,
                                          lowercase_frac=0.01), 0, 3, with_qual=2)
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750
    cfg.contents.indel_balance_factor_sgs = rate
    for task in tasks(E):
        assert eng.polish(sh, task, cfg) == run_checker(oracle.np_oracle_run, sh, task, cfg), task


@pytest.mark.parametrize("kw", [
    dict(seed=5, n_contigs=3, contig_len=2000, depth=0.0),
    dict(seed=6, n_contigs=4, contig_len=0, min_len=40, max_len=160, depth=30.0),
    dict(seed=7, n_contigs=1, contig_len=1, depth=0.0),
    dict(seed=8, n_contigs=2, contig_len=700, depth=400.0),
    dict(seed=9, n_contigs=300, contig_len=0, min_len=200, max_len=6000, depth=20.0),
], ids=["noreads", "tiny_contigs", "one_base", "deep_tiny", "many_contigs"])
def test_gpu_edge_shapes(E, oracle, eng, kw):
    sh = E.Shard.synthetic(E.synth_params(lowercase_frac=0.05, **kw), 0, kw["n_contigs"], with_qual=True)
    cfg = E.default_config(b"")
    for task in tasks(E):
        assert eng.polish(sh, task, cfg) == run_checker(oracle.np_oracle_run, sh, task, cfg), task


# ---- task 4: snp_valid (snpvalid.c:3-35) -----------------------------------------------------------------------------
def oracle_contigs(oracle, sh, task, cfg):
    """Per-contig oracle run -> {name: bytes, or None where the reference's behaviour is undefined (odd cut-point list)}."""
    import numpy as np
    oracle.np_oracle_run_contig.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    out = {}
    for c, n in enumerate(sh.names):
        cap = int(sh.view.ctg_off[c + 1] - sh.view.ctg_off[c]) * 2 + 4096
        buf, ln = np.zeros(cap, np.uint8), C.c_int64(0)
        rc = oracle.np_oracle_run_contig(C.addressof(sh.view), c, task, C.cast(cfg, C.c_void_p), buf.ctypes.data, cap, C.byref(ln))
        assert rc in (0, -2), rc
        out[n] = buf[:ln.value].tobytes() if rc == 0 else None
    return out


@pytest.mark.parametrize("case", sorted(CASES))
def test_gpu_snp_valid_matches_oracle_and_reference_md5(E, oracle, eng, synth_files, case):
    want_md5 = json.load(open(os.path.join(GOLDEN, "synth_md5.json")))[case]["4"]
    fa, bam = synth_files(case)
    sh = E.Shard.load(fa, bam, with_qual=True)
    cfg = E.default_config(fa, bam)
    got = eng.polish(sh, 4, cfg)
    assert {"%s_4" % n: md5(s) for n, s in got.items()} == want_md5
    want = oracle_contigs(oracle, sh, 4, cfg)
    assert all(v is not None for v in want.values())
    assert got == want


@pytest.mark.parametrize("seed", [41, 42, 43, 44, 45, 46])
def test_gpu_snp_valid_mutated_thresholds(E, oracle, eng, seed):
    import random
    rng = random.Random(seed)
    kw = dict(seed=rng.randrange(1 << 30), n_contigs=rng.choice([2, 5]), contig_len=rng.choice([8000, 20000]),
              depth=rng.choice([5, 12, 30, 60]), draft_indel=rng.choice([0.003, 0.02]), read_indel=rng.choice([0.0001, 0.003]),
              lowercase_frac=rng.choice([0.01, 0.05, 0.15]))
    sh = E.Shard.synthetic(E.synth_params(**kw), 0, kw["n_contigs"], with_qual=True)
    cfg = E.default_config(b"")
    c = cfg.contents
    c.read_tlen = 2000
    c.trim_len_edge, c.ext_len_edge = rng.choice([0, 1, 2, 4]), rng.choice([0, 1, 2, 3])
    c.min_len_inter_kmer = rng.choice([0, 2, 5, 9])
    c.max_len_kmer, c.max_count_kmer, c.min_map_quality = rng.choice([10, 50, 120]), rng.choice([3, 50]), rng.choice([0, 30])
    want = oracle_contigs(oracle, sh, 4, cfg)
    got = eng.polish(sh, 4, cfg)
    checked = 0
    for n, s in want.items():
        if s is not None:                      # contigs with an odd cut-point list: undefined in the reference, not compared
            assert got[n] == s, n
            checked += 1
    assert checked > 0


def test_gpu_snp_valid_golden_testdata_engine_abi_and_cli(E, eng):
    fa, bam = os.path.join(GOLDEN, "td30.step1.fa"), os.path.join(GOLDEN, "td30.step1.bam")
    exp_path = os.path.join(GOLDEN, "td30.step1.snpvalid.expected.fa")
    exp = read_fasta(exp_path)
    sh = E.Shard.load(fa, bam, with_qual=True)
    got = eng.polish(sh, 4, E.default_config(fa, bam))
    assert {"%s_4" % n: s for n, s in got.items()} == exp
    L = E.lib()                                   # the drop-in entry point: snp_valid(tigname, cfg) -> PolishResult*
    cfg = L.config_init(fa.encode(), bam.encode(), None)
    for name in [n[:-2] for n in exp]:
        res = L.snp_valid(name.encode(), cfg)
        seq = C.string_at(res.contents.contig)
        assert res.contents.length == len(seq) and seq == exp["%s_4" % name]
        L.polishresult_destory(res)
    L.config_destory(cfg)
    cli = os.path.join(os.path.dirname(E.binding.LIB_PATH), "nextpolish1")
    ours = subprocess.run([cli, "snpvalid", fa, bam], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    assert ours == open(exp_path, "rb").read()


# ---- the reference's process model: config_init in the parent, Pool forked afterwards ----------------
def _fork_worker(args):
    """Child of a forked Pool: calls the drop-in entry point through ctypes like nextpolish1.py:181-189 does."""
    fn, name = args
    from nextpolish_b200 import engine as E
    L = E.lib()
    res = getattr(L, fn)(name.encode(), _FORK_CFG)
    seq = C.string_at(res.contents.contig)
    L.polishresult_destory(res)
    return name, seq


_FORK_CFG = None


@pytest.mark.parametrize("step,fn", [(1, "score_chain"), (2, "kmer_count")])
def test_forked_pool_after_config_init(E, step, fn):
    """Exactly the reference's worker: config_init() in the parent, multiprocessing.Pool(2) forked AFTERWARDS,
    score_chain / kmer_count called in the children, imap_unordered(chunksize=1) (nextpolish1.py:219-224).
    The parent must not have touched CUDA before the fork (it has not: config_init is host code) and every child
    creates its own engine lazily."""
    import multiprocessing as mp
    global _FORK_CFG
    fa = os.path.join(GOLDEN, "td30.step%d.fa" % step)
    bam = os.path.join(GOLDEN, "td30.step%d.bam" % step)
    exp = read_fasta(os.path.join(GOLDEN, "td30.step%d.expected.fa" % step))
    # a fresh interpreter: this test process already holds a CUDA context (other tests), which must not cross fork()
    code = r"""
import ctypes as C, multiprocessing as mp, sys, json, hashlib
sys.path.insert(0, %r)
import tests.test_gpu_parity as T
from nextpolish_b200 import engine as E
L = E.lib()
T._FORK_CFG = L.config_init(%r.encode(), %r.encode(), None)
names = %r
with mp.get_context("fork").Pool(2) as pool:
    out = dict(pool.imap_unordered(T._fork_worker, [(%r, n) for n in names], chunksize=1))
print(json.dumps({k: hashlib.md5(v).hexdigest() for k, v in out.items()}))
""" % (os.path.dirname(os.path.dirname(os.path.realpath(__file__))), fa, bam, [n[:-2] for n in exp], fn)
    r = subprocess.run([os.sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    got = json.loads(r.stdout.decode().strip().splitlines()[-1])
    assert got == {n[:-2]: md5(s) for n, s in exp.items()}
