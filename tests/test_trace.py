"""PolishPoint change trace (contig_get_contig, contig.c:743-799; what `nextpolish1.py -debug` prints when
Configure.trace_polish_open is set): oracle pinned against the reference's own .so, kernel bodies (emulated) and
the GPU engine / reference-ABI entry points against the oracle."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from nextpolish_b200.binding import Configure, PolishPoint, PolishResult
from tests.conftest import REF_BIN, REF_SAMTOOLS
from tests.synth_cases import CASES

REF_SO = os.path.join(os.path.dirname(REF_BIN), "nextpolish1.so")


def oracle_points(oracle, sh, ci, task, cfg):
    oracle.np_oracle_run_contig_points.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                                   C.c_void_p, C.c_int64, C.c_void_p]
    L = int(sh.view.ctg_off[ci + 1] - sh.view.ctg_off[ci])
    cap = 2 * L + 4096
    out = np.zeros(cap, np.uint8)
    olen, npts = C.c_int64(0), C.c_int64(0)
    pts = (PolishPoint * cap)()
    assert oracle.np_oracle_run_contig_points(C.addressof(sh.view), ci, task, C.cast(cfg, C.c_void_p), out.ctypes.data, cap,
                                              C.byref(olen), pts, cap, C.byref(npts)) == 0
    return out[:olen.value].tobytes(), [(pts[i].pos, pts[i].index, pts[i].curbase, pts[i].base) for i in range(npts.value)]


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", range(4))
def test_oracle_trace_vs_reference_so(E, oracle, tmp_path, seed):
    R = C.CDLL(REF_SO)
    R.config_init.argtypes = [C.c_char_p] * 3
    R.config_init.restype = C.POINTER(Configure)
    for f in ("score_chain", "kmer_count"):
        getattr(R, f).argtypes = [C.c_char_p, C.POINTER(Configure)]
        getattr(R, f).restype = C.POINTER(PolishResult)
    R.polishresult_destory.argtypes = [C.POINTER(PolishResult)]
    rng = random.Random(seed)
    kw = dict(seed=rng.randrange(1 << 30), n_contigs=2, contig_len=rng.choice([3000, 20000]), depth=rng.choice([5, 30, 60]),
              draft_snv=0.01, draft_indel=rng.choice([0.003, 0.02]), lowercase_frac=rng.choice([0.01, 0.1]))
    fa, bam = str(tmp_path / "x.fa"), str(tmp_path / "x.bam")
    assert E.lib().np_synth_write(E.synth_params(**kw), fa.encode(), bam.encode()) == 0
    subprocess.check_call([REF_SAMTOOLS, "index", bam])
    sh = E.Shard.load(fa, bam, with_qual=True)
    cfg, rcfg = E.default_config(fa, bam), R.config_init(fa.encode(), bam.encode(), None)
    rcfg.contents.trace_polish_open = 1
    compared = 0
    for task, fn in ((1, R.score_chain), (2, R.kmer_count)):
        for ci, n in enumerate(sh.names):
            res = fn(n.encode(), rcfg)
            seq = C.string_at(res.contents.contig)
            ref = [(res.contents.data[i].pos, res.contents.data[i].index, res.contents.data[i].curbase, res.contents.data[i].base)
                   for i in range(res.contents.datalength)]
            R.polishresult_destory(res)
            got_seq, got = oracle_points(oracle, sh, ci, task, cfg)
            if got_seq != seq:
                continue     # contig-end extension reads past the reference's arrays (documented deviation)
            assert got == ref, (task, n)
            compared += 1
    assert compared >= 3


def emu_points(emu, sh, task, cfg, variant):
    emu.np_emu_run_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    emu.np_emu_run_points.restype = C.c_int
    cap = int(sh.total_bases * 2) + 4096
    out = np.zeros(cap, np.uint8)
    off = np.zeros(sh.n_contigs + 1, np.int64)
    poff = np.zeros(sh.n_contigs + 1, np.int64)
    pts = (PolishPoint * cap)()
    assert emu.np_emu_run_points(C.addressof(sh.view), task, C.cast(cfg, C.c_void_p), out.ctypes.data, cap, off.ctypes.data, variant,
                                 pts, cap, poff.ctypes.data) == 0
    return [[(pts[i].pos, pts[i].index, pts[i].curbase, pts[i].base) for i in range(poff[k], poff[k + 1])] for k in range(sh.n_contigs)]


@pytest.mark.parametrize("case", ["c30", "noisy", "shallow", "lower", "ragged"])
def test_emulated_trace_matches_oracle(E, oracle, emu, case):
    sh = E.Shard.synthetic(E.synth_params(**CASES[case]), 0, CASES[case]["n_contigs"], with_qual=True)
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750
    for task, variants in ((1, (1, 2)), (2, (1,))):
        want = [oracle_points(oracle, sh, ci, task, cfg)[1] for ci in range(sh.n_contigs)]
        assert task == 2 or sum(len(w) for w in want) > 0
        for variant in variants:
            assert emu_points(emu, sh, task, cfg, variant) == want, (task, variant)


@pytest.mark.gpu
def test_gpu_trace_matches_oracle(E, oracle):
    sh = E.Shard.synthetic(E.synth_params(seed=77, n_contigs=4, contig_len=60000, depth=30.0, draft_indel=0.01, lowercase_frac=0.02),
                           0, 4, with_qual=2)
    dense = E.Shard.synthetic(E.synth_params(seed=77, n_contigs=4, contig_len=60000, depth=30.0, draft_indel=0.01, lowercase_frac=0.02),
                              0, 4, with_qual=1)
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750
    cfg.contents.trace_polish_open = 1
    eng = E.Engine(0)
    for task in E.TASKS:
        want = [oracle_points(oracle, dense, ci, task, cfg) for ci in range(dense.n_contigs)]
        got = eng.polish(sh, task, cfg)
        pts = eng.points(sh.n_contigs)
        for ci, nm in enumerate(sh.names):
            assert got[nm] == want[ci][0] and pts[ci] == want[ci][1], (task, nm)
    cfg.contents.trace_polish_open = 0
    eng.polish(sh, 1, cfg)
    with pytest.raises(E.NativeError):
        eng.points(sh.n_contigs)
    eng.close()


@pytest.mark.gpu
def test_reference_abi_returns_trace(E, oracle, synth_files):
    """score_chain / kmer_count through the reference ABI with trace_polish_open (nextpolish1.py -debug, :133,:230-231)."""
    L = E.lib()
    fa, bam = synth_files("lower")
    sh = E.Shard.load(fa, bam, with_qual=True)
    cfg = E.default_config(fa, bam)
    cfg.contents.trace_polish_open = 1
    for task, fn in ((1, L.score_chain), (2, L.kmer_count)):
        fn.argtypes = [C.c_char_p, C.POINTER(Configure)]
        fn.restype = C.POINTER(PolishResult)
        for ci, nm in enumerate(sh.names):
            res = fn(nm.encode(), cfg)
            seq = C.string_at(res.contents.contig)
            got = [(res.contents.data[i].pos, res.contents.data[i].index, res.contents.data[i].curbase, res.contents.data[i].base)
                   for i in range(res.contents.datalength)]
            L.polishresult_destory(res)
            want_seq, want = oracle_points(oracle, sh, ci, task, cfg)
            assert seq == want_seq and got == want, (task, nm)
