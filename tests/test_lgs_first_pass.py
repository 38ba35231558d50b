"""First pass of the long-read consensus window (nextpolish2.so).
  * the C restatement oracle/np2_oracle.c against the goldens minted from the reference (tests/golden/lgs_golden.json,
    make_golden_lgs.py) and, when oracle/_ref holds the stage door into the unmodified ctg_cns.c (libnp2_refshim.so),
    against the reference live on seeded fuzz inputs;
  * the kernel bodies of the GPU path (csrc/lgs_first_pass.h) compiled for the host (tests/emu/emu_lgs.cpp, test build
    only) against that oracle: every seeded case x read type, threads run in order and shuffled, the normal build and one
    with 8-column stretches, 4-column cut blocks and a 3-entry / 1-match gather scratch (every seam, the speculative-segment
    re-runs and the literal fall-back of the warp chain are reached), both chain kernels, batches of
    several windows, rejected inputs;
  * nextpolish2.so exports what include/nextpolish2_b200.h declares and fails loudly without a GPU."""
import ctypes as C
import hashlib
import json
import os
import random
import subprocess

import numpy as np
import pytest

from tests import lgs_cases as L
from tests.conftest import GOLDEN, ROOT


@pytest.fixture(scope="module")
def O2():
    path = os.path.join(ROOT, "oracle", "libnp2_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    return C.CDLL(path)


def td_windows():
    z = np.load(os.path.join(GOLDEN, "lgs_td_windows.npz"))
    keys = sorted({k.rsplit(".", 1)[0] for k in z.files})
    out = {}
    for k in keys:
        out[k] = dict(len=int(z[k + ".len"][0]), read_type=1, min_cov=4, aln_t_s=z[k + ".aln_t_s"], aln_len=z[k + ".aln_len"],
                      str_off=z[k + ".str_off"], t_str=z[k + ".t_str"].tobytes(), q_str=z[k + ".q_str"].tobytes())
    return out


def digest(res):
    pos, base = res
    return {"n": len(base), "pos_md5": hashlib.md5(pos.astype("<u4").tobytes()).hexdigest(), "base_md5": hashlib.md5(base).hexdigest()}


def all_golden_cases():
    for name, kw in L.CASES.items():
        for rt in (1, 2, 3, 4):
            yield "%s/rt%d" % (name, rt), (lambda kw=kw, rt=rt: L.synthetic_case(**dict(kw, read_type=rt)))
    for key, case in td_windows().items():
        for rt in (1, 2, 3, 4):
            yield "%s/rt%d" % (key, rt), (lambda case=case, rt=rt: dict(case, read_type=rt))


def test_oracle_matches_reference_goldens(O2):
    gold = json.load(open(os.path.join(GOLDEN, "lgs_golden.json")))["first_pass"]
    seen = 0
    for key, make in all_golden_cases():
        got = L.first_pass_via(O2.np2_oracle_first_pass, make())
        assert not isinstance(got, int), (key, got)
        assert digest(got) == gold[key], key
        seen += 1
    assert seen == len(gold) == 68


def test_real_windows_look_like_consensus(O2):
    """Sanity of the real-data fixture itself: ~6 kb windows at test_data depth, consensus within a few percent of the
    window length, positions non-decreasing and covering the window."""
    for key, case in td_windows().items():
        pos, base = L.first_pass_via(O2.np2_oracle_first_pass, case)
        assert len(case["aln_t_s"]) > 20
        assert abs(len(base) - case["len"]) < 0.05 * case["len"]
        assert (np.diff(pos.astype(np.int64)) >= 0).all() and pos[0] == 0 and pos[-1] == case["len"] - 1
        assert set(base.upper()) <= set(b"ACGTN")


@pytest.mark.skipif(L.ref_shim() is None, reason="oracle/_ref/libnp2_refshim.so not built (needs /root/reference)")
def test_oracle_matches_live_reference_fuzz(O2):
    S = L.ref_shim()
    rng = random.Random(2024)
    for it in range(120):
        kw = dict(seed=rng.randrange(1 << 30), length=rng.choice([20, 60, 300, 1200]), depth=rng.choice([2, 5, 15, 40]),
                  read_len=rng.choice([30, 100, 400]), sub=rng.choice([0.0, 0.02, 0.08]), ins=rng.choice([0.0, 0.03, 0.1]),
                  dele=rng.choice([0.0, 0.03, 0.1]), read_type=rng.choice([1, 2, 3, 4]), min_cov=rng.choice([0, 4, 10]),
                  long_ins=rng.choice([0, 0.003]), masked=rng.choice([0, 0, 0.02]), homopolymer=rng.random() < 0.3,
                  odd_chars=rng.random() < 0.2)
        case = L.synthetic_case(**kw)
        want = L.first_pass_via(S.np2_ref_first_pass, case)
        got = L.first_pass_via(O2.np2_oracle_first_pass, case)
        assert not isinstance(want, int) and not isinstance(got, int), kw
        assert (want[0] == got[0]).all() and want[1] == got[1], kw


def test_oracle_error_codes_and_qv(O2):
    case = L.synthetic_case(**L.CASES["ont30"])
    cap = 8000
    pos, base, qv = np.zeros(cap, np.uint32), np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
    f = O2.np2_oracle_first_pass_qv
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    args = lambda c, cap_: (c["read_type"], len(c["aln_t_s"]), c["aln_t_s"].ctypes.data, c["aln_len"].ctypes.data, c["str_off"].ctypes.data,
                            c["t_str"], c["q_str"], c["len"], c["min_cov"], pos.ctypes.data, base.ctypes.data, qv.ctypes.data, cap_)
    n = f(*args(case, cap))
    assert n == 3000 and qv[:n].max() <= 100 and np.median(qv[:n]) > 60      # 100 * links / coverage of the chosen entry
    assert f(*args(case, 10)) == -1                                           # output too small
    short = dict(case, len=case["len"] + 50)                                  # nothing covers the last column
    assert f(*args(short, cap)) == -2
    cut = dict(case, len=case["len"] - 50)                                    # alignments leave the window
    assert f(*args(cut, cap)) == -3


# ---- the GPU path's kernel bodies on the host ----------------------------------------------------------------------------
def _emu2_build(name, defines=()):
    d = os.path.join(ROOT, "tests", "_emu")
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, name)
    srcs = [os.path.join(ROOT, "tests", "emu", "emu_lgs.cpp"), os.path.join(ROOT, "nextpolish_b200", "csrc", "lgs_first_pass.h"),
            os.path.join(ROOT, "include", "nextpolish2_b200.h")]
    if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall"] + list(defines) + ["-o", so, srcs[0]])
    E = C.CDLL(so)
    E.np2_emu_first_pass_batch.restype = C.c_int64
    E.np2_emu_first_pass_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_uint64, C.c_void_p]
    return E


@pytest.fixture(scope="module", params=["product_sizes", "tiny_seams"])
def emu2(request):
    if request.param == "product_sizes":
        return _emu2_build("libnp2_emu.so")
    return _emu2_build("libnp2_emu_small.so", ("-DNP2_STRETCH=8", "-DNP2_CUT_BLOCK=4", "-DNP2_GMAX=3", "-DNP2_MAXM=1"))


THREAD_CHAIN = 1 << 32          # bit 32 of the emulator's seed argument: Chain instead of ChainWarp (tests/emu/emu_lgs.cpp)


def emu_call(E, seed, stats=None):
    st = stats if stats is not None else np.zeros(4, np.int64)
    return lambda b, p, ba, q, cap, off: E.np2_emu_first_pass_batch(b, p, ba, q, cap, off, seed, st.ctypes.data)


def same(got, want):
    return len(got[1]) == len(want[1]) and (got[0] == want[0]).all() and got[1] == want[1] and (got[2] == want[2]).all()


def test_emulated_kernels_match_oracle(O2, emu2):
    reruns = 0
    for name, kw in L.CASES.items():
        for rt in (1, 2, 3, 4):
            case = L.synthetic_case(**dict(kw, read_type=rt))
            want = L.oracle_window(O2, case)
            assert not isinstance(want, int)
            for seed in (0, 11, 11 | THREAD_CHAIN):          # threads in order / shuffled / the thread-per-segment chain kernel
                st = np.zeros(4, np.int64)
                got = L.first_pass_batch(emu_call(emu2, seed, st), [case])
                assert not isinstance(got, int), (name, rt, seed, got)
                assert same(got[0], want), (name, rt, seed)
                reruns += int(st[1])
    assert reruns > 50                                       # the "zones" cases force segments to run again with their true score


def test_emulated_kernels_real_windows_and_batches(O2, emu2):
    wins = list(td_windows().values())
    for rt in (1, 3):
        cases = [dict(w, read_type=rt) for w in wins]
        cases += [L.synthetic_case(**dict(L.CASES[n], read_type=rt)) for n in ("ont_tiny", "ont_shallow", "hifi_zones", "ont_masked")]
        want = [L.oracle_window(O2, c) for c in cases]
        st = np.zeros(4, np.int64)
        got = L.first_pass_batch(emu_call(emu2, 5, st), cases)       # one batch, several windows
        assert not isinstance(got, int)
        for g, w in zip(got, want):
            assert same(g, w), rt
        assert st[0] >= len(cases)                            # at least one chain segment per window
        for c, w in zip(cases, want):                         # and each window alone
            assert same(L.first_pass_batch(emu_call(emu2, 0), [c])[0], w)


def test_emulated_kernels_fuzz(O2, emu2):
    rng = random.Random(77)
    for it in range(60):
        kw = dict(seed=rng.randrange(1 << 30), length=rng.choice([20, 60, 300, 900]), depth=rng.choice([1, 2, 5, 15, 40]),
                  read_len=rng.choice([30, 100, 400]), sub=rng.choice([0.0, 0.02, 0.08, 0.4]), ins=rng.choice([0.0, 0.03, 0.1, 0.3]),
                  dele=rng.choice([0.0, 0.03, 0.1, 0.3]), read_type=rng.choice([1, 2, 3, 4]), min_cov=rng.choice([0, 4, 10]),
                  long_ins=rng.choice([0, 0.003]), masked=rng.choice([0, 0, 0.02]), homopolymer=rng.random() < 0.3,
                  odd_chars=rng.random() < 0.2, zones=rng.choice([0, 0, (4, 20), (6, 60)]))
        case = L.synthetic_case(**kw)
        want = L.oracle_window(O2, case)
        got = L.first_pass_batch(emu_call(emu2, rng.randrange(1, 1 << 20) | (THREAD_CHAIN if it % 3 == 0 else 0)), [case])
        if isinstance(want, int):
            assert got == want, kw
        else:
            assert not isinstance(got, int) and same(got[0], want), kw


def test_emulated_kernels_reject_what_the_oracle_rejects(O2, emu2):
    case = L.synthetic_case(**L.CASES["ont30"])
    for bad in (dict(case, len=case["len"] + 50), dict(case, len=case["len"] - 50)):        # -2: bare last column; -3: out of window
        want = L.oracle_window(O2, bad)
        assert isinstance(want, int) and want in (-2, -3)
        assert L.first_pass_batch(emu_call(emu2, 0), [bad]) == want
    gap_first = L.pack_case([(0, "ACGTACGTAC", "ACGTACGTAC"), (2, "-GTACG", "TGTACG")], 10, 1)
    assert L.oracle_window(O2, gap_first) == -3 and L.first_pass_batch(emu_call(emu2, 0), [gap_first]) == -3
    assert L.first_pass_batch(emu_call(emu2, 0), []) == []                                    # empty batch


def test_nextpolish2_library_exports_and_fails_loudly_without_gpu():
    import re
    from nextpolish_b200 import nextpolish2 as NP2
    hdr = open(os.path.join(ROOT, "include", "nextpolish2_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(np2_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(NP2.EXPORTS2), declared ^ set(NP2.EXPORTS2)
    lib = C.CDLL(NP2.LIB2_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(NP2.NativeError, match="no CPU path"):
            NP2.LgsEngine(0)


def test_production_window_consensus_on_accurate_reads(O2, emu2):
    """The reference's PRODUCTION window consensus (get_cns_from_align_tags with fast = 0: first pass, low-quality regions, POA,
    second round; goldens through np2_ref_window_prod) on windows where it finds nothing to re-polish (accurate reads; noisy reads under their own read type's rules): it equals the
    first pass with the production letter-case rule applied to the link qualities (nextpolish2.production_case) — which
    pins that rule, and the qv values on the side of the threshold they fall, against the reference.  (With noisy reads the two differ: that stage is not built.)"""
    from nextpolish_b200 import nextpolish2 as NP2
    from tests.golden.make_golden_lgs import PROD_CLEAN
    gold = json.load(open(os.path.join(GOLDEN, "lgs_golden.json")))["production_clean"]
    for name, kw in PROD_CLEAN.items():
        case = L.synthetic_case(**kw)
        for res in (L.oracle_window(O2, case), L.first_pass_batch(emu_call(emu2, 4), [case])[0]):
            pos, base, qv = res
            cased = NP2.production_case(base, qv)
            got = {"n": len(cased), "base_md5": hashlib.md5(cased).hexdigest()}
            assert got == {k: gold[name][k] for k in got}, name          # (the qv array itself is rewritten by the later stages: compared through the case only)
    from tests.golden.make_golden_lgs import PROD_CHANGED
    changed = json.load(open(os.path.join(GOLDEN, "lgs_golden.json")))["production_changed"]
    for name, kw in PROD_CHANGED.items():                    # windows the stage that is not built does change
        fp = L.oracle_window(O2, L.synthetic_case(**kw))
        cased = NP2.production_case(fp[1], fp[2])
        assert {"n": len(cased), "base_md5": hashlib.md5(cased).hexdigest()} != changed[name], name
