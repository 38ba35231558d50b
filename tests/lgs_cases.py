"""Inputs of the long-read first pass (reference source/lib/ctg_cns.c): gapped alignment strings of reads
against one consensus window, seeded.  A case is a dict
    len, read_type (1 ont, 2 clr, 3 hifi, 4 rs), min_cov, aln_t_s[n], aln_len[n], str_off[n], t_str, q_str
where alignment i occupies columns [str_off[i], str_off[i] + aln_len[i]) of the two strings ('-' = gap, 'M' = masked
column) and starts at window position aln_t_s[i].  As in ctg_cns_core (ctg_cns.c:3456-3468) the first alignment is the
window against itself."""
import ctypes as C
import os
import random

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
REF_SHIM = os.path.join(ROOT, "oracle", "_ref", "libnp2_refshim.so")

CASES = {
    # name: (seed, len, depth, read_len, sub, ins, dele, read_type, extras)
    "ont30": dict(seed=1, length=3000, depth=30, read_len=900, sub=0.03, ins=0.02, dele=0.03, read_type=1),
    "ont_noisy": dict(seed=2, length=2500, depth=40, read_len=700, sub=0.05, ins=0.05, dele=0.05, read_type=1, long_ins=0.002),
    "ont_shallow": dict(seed=3, length=2000, depth=4, read_len=600, sub=0.04, ins=0.03, dele=0.03, read_type=1),
    "clr30": dict(seed=4, length=3000, depth=30, read_len=900, sub=0.02, ins=0.06, dele=0.03, read_type=2),
    "hifi20": dict(seed=5, length=4000, depth=20, read_len=1500, sub=0.002, ins=0.002, dele=0.002, read_type=3),
    "rs25": dict(seed=6, length=2500, depth=25, read_len=800, sub=0.02, ins=0.07, dele=0.04, read_type=4),
    "ont_masked": dict(seed=7, length=2000, depth=25, read_len=700, sub=0.03, ins=0.03, dele=0.03, read_type=1, masked=0.01),
    "hifi_masked": dict(seed=8, length=2000, depth=15, read_len=900, sub=0.003, ins=0.003, dele=0.003, read_type=3, masked=0.01),
    "clr_hp": dict(seed=9, length=2500, depth=35, read_len=800, sub=0.02, ins=0.05, dele=0.05, read_type=2, homopolymer=True),
    "ont_tiny": dict(seed=10, length=40, depth=6, read_len=30, sub=0.05, ins=0.05, dele=0.05, read_type=1),
    "ont_zones": dict(seed=12, length=1500, depth=12, read_len=400, sub=0.5, ins=0.3, dele=0.3, read_type=1, zones=(5, 40)),
    "clr_zones": dict(seed=13, length=1500, depth=8, read_len=300, sub=0.6, ins=0.3, dele=0.3, read_type=2, zones=(4, 30)),
    "hifi_zones": dict(seed=14, length=1200, depth=6, read_len=300, sub=0.6, ins=0.2, dele=0.4, read_type=3, zones=(6, 50)),
    "rs_zones": dict(seed=15, length=1200, depth=10, read_len=300, sub=0.5, ins=0.4, dele=0.3, read_type=4, zones=(5, 25)),
    "ont_lower_n": dict(seed=11, length=1500, depth=20, read_len=500, sub=0.03, ins=0.03, dele=0.03, read_type=1, odd_chars=True),
}


def synthetic_case(seed=1, length=3000, depth=30, read_len=900, sub=0.03, ins=0.02, dele=0.03, read_type=1, min_cov=4,
                   long_ins=0.0, masked=0.0, homopolymer=False, odd_chars=False, zones=0):
    """zones = (clean, dirty): the error rates (given very high) apply only to the last `dirty` of every clean + dirty
    window positions, the others are error free: clean zones give the score chain its cut
    columns, the dirty ones drive scores to the reference's clamp at 0 right behind a cut (the case in which a chain
    segment cannot be computed relative to its cut and has to be run again with the true score)."""
    rng = random.Random(seed)
    if homopolymer:
        draft = []
        while len(draft) < length:
            draft.extend(rng.choice("ACGT") * rng.choice([1, 1, 1, 2, 3, 5, 8]))
        draft = "".join(draft[:length])
    else:
        draft = "".join(rng.choice("ACGT") for _ in range(length))
    if odd_chars:                                  # lowercase and N in the window itself (base_to_int, ctg_cns.c:58-67)
        d = list(draft)
        for _ in range(length // 50):
            i = rng.randrange(length)
            d[i] = rng.choice([d[i].lower(), "N", "n"])
        draft = "".join(d)
    alns = [(0, draft, draft)]
    n_reads = max(1, int(depth * length / read_len))
    starts = sorted(rng.randrange(-read_len // 2, max(length - 8, -read_len // 2 + 1)) for _ in range(n_reads))      # BAM order
    for s0 in starts:
        s = max(0, s0)
        e = min(length, s0 + int(read_len * rng.uniform(0.6, 1.4)))
        if e - s < 8:
            continue
        t, q = [], []
        p = s
        while p < e:
            first_or_last = p == s or p == e - 1 or (zones and p % (zones[0] + zones[1]) < zones[0])
            r = rng.random()
            if not first_or_last and masked and r < masked:
                run = min(rng.randrange(1, 12), e - 1 - p)
                t.extend("M" * run), q.extend("M" * run)
                p += run
                continue
            r = rng.random()
            if not first_or_last and r < ins:
                run = rng.randrange(20, 200) if rng.random() < long_ins / max(ins, 1e-9) else rng.choice([1, 1, 1, 2, 3])
                for _ in range(run):
                    t.append("-"), q.append(rng.choice("ACGT"))
                # an insertion column never ends the alignment: fall through to the target column below
            r = rng.random()
            c = draft[p]
            if not first_or_last and r < dele:
                t.append(c), q.append("-")
            elif not first_or_last and r < dele + sub:
                t.append(c), q.append(rng.choice("ACGT"))
            else:
                t.append(c), q.append(c)
            p += 1
        alns.append((s, "".join(t), "".join(q)))
    return pack_case(alns, length, read_type, min_cov)


def pack_case(alns, length, read_type, min_cov=4):
    """alns: [(aln_t_s, t_aln_str, q_aln_str)] in tags_list order."""
    off, t_all, q_all = [], [], []
    pos = 0
    for _, t, q in alns:
        assert len(t) == len(q) and len(t) > 0
        off.append(pos)
        t_all.append(t), q_all.append(q)
        pos += len(t)
    return dict(len=length, read_type=read_type, min_cov=min_cov,
                aln_t_s=np.array([a[0] for a in alns], np.uint32), aln_len=np.array([len(a[1]) for a in alns], np.uint32),
                str_off=np.array(off, np.uint64), t_str="".join(t_all).encode(), q_str="".join(q_all).encode())


def first_pass_via(fn, case):
    """Calls an np2_*_first_pass-shaped C function (reference shim, oracle): -> (pos uint32[n], base bytes) or the
    negative return code."""
    cap = int(case["len"]) * 3 + int(case["aln_len"].sum()) + 16
    pos = np.zeros(cap, np.uint32)
    base = np.zeros(cap, np.uint8)
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                   C.c_void_p, C.c_void_p, C.c_int]
    n = fn(case["read_type"], len(case["aln_t_s"]), case["aln_t_s"].ctypes.data, case["aln_len"].ctypes.data,
           case["str_off"].ctypes.data, case["t_str"], case["q_str"], case["len"], case["min_cov"],
           pos.ctypes.data, base.ctypes.data, cap)
    if n < 0:
        return n
    return pos[:n].copy(), base[:n].tobytes()


def ref_shim():
    return C.CDLL(REF_SHIM) if os.path.exists(REF_SHIM) else None


def make_batch(cases):
    from nextpolish_b200.nextpolish2 import make_batch as mk
    return mk(cases, cases[0]["read_type"] if cases else 1, cases[0]["min_cov"] if cases else 4)


def first_pass_batch(call, cases):
    """call(batch_ptr, out_pos, out_base, out_qv, cap, out_off) -> total; returns per-window [(pos, base, qv)] or the code."""
    b, keep = make_batch(cases)
    cap = sum(int(c["len"]) * 3 + int(c["aln_len"].sum()) for c in cases) + 16
    pos, base, qv = np.zeros(cap, np.uint32), np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
    off = np.zeros(len(cases) + 1, np.int64)
    n = call(C.byref(b), pos.ctypes.data, base.ctypes.data, qv.ctypes.data, cap, off.ctypes.data)
    del keep
    if n < 0:
        return n
    assert off[-1] == n
    return [(pos[off[i]:off[i + 1]].copy(), base[off[i]:off[i + 1]].tobytes(), qv[off[i]:off[i + 1]].copy()) for i in range(len(cases))]


def oracle_window(O2, case):
    """(pos, base, qv) of the C restatement, or its negative code."""
    cap = int(case["len"]) * 3 + int(case["aln_len"].sum()) + 16
    pos, base, qv = np.zeros(cap, np.uint32), np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
    f = O2.np2_oracle_first_pass_qv
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    n = f(case["read_type"], len(case["aln_t_s"]), case["aln_t_s"].ctypes.data, case["aln_len"].ctypes.data, case["str_off"].ctypes.data,
          case["t_str"], case["q_str"], case["len"], case["min_cov"], pos.ctypes.data, base.ctypes.data, qv.ctypes.data, cap)
    if n < 0:
        return n
    return pos[:n].copy(), base[:n].tobytes(), qv[:n].copy()
