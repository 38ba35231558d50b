#!/usr/bin/env python
"""Mint the long-read (nextpolish2) fixtures from the reference itself.  Run in the build container, after
`make -C oracle ref ref2` (oracle/_ref/minimap2, samtools, nextpolish2.so, libnp2_refshim.so).

  lgs_td_windows.npz   two windows of the reference's own source/test_data: lreads.fasta.gz mapped onto raw.genome.fasta
                       with the vendored minimap2 (-ax map-ont, what the reference's driver runs for lgs reads,
                       source/nextPolish:199-211), alignments written out as gapped strings clipped to the window
  lgs_golden.json      (a) whole-path record: md5 / length of ctg_cns_core's output (nextpolish2.so, read type ont) for
                       both test_data contigs on that BAM — the target of the rows still to build;
                       (b) first pass (np2_ref_first_pass = get_cns_from_align_tags(fast) of the unmodified ctg_cns.c):
                       md5 of the position list and of the base string for every seeded case of tests/lgs_cases.py under
                       all four read types, and for the two test_data windows.
"""
import ctypes as C
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.realpath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
TD = "/root/reference/source/test_data"
OUT = os.path.dirname(os.path.realpath(__file__))
sys.path.insert(0, ROOT)
from tests import lgs_cases as L  # noqa: E402

WINDOWS = [("tig0000001", 10000, 16000), ("tig0000002", 30000, 35000)]


def read_fa(path):
    d, name = {}, None
    for line in open(path):
        if line.startswith(">"):
            name = line[1:].split()[0]
            d[name] = []
        elif name:
            d[name].append(line.strip())
    return {k: "".join(v) for k, v in d.items()}


def window_alignments(bam, draft, ctg, ws, we):
    """[(aln_t_s, t_str, q_str)] clipped to [ws, we), BAM order, the window against itself first."""
    alns = [(0, draft[ws:we], draft[ws:we])]
    txt = subprocess.run([os.path.join(REF, "samtools"), "view", "-F", "0x904", bam, "%s:%d-%d" % (ctg, ws + 1, we)],
                         stdout=subprocess.PIPE, check=True).stdout.decode()
    for line in txt.split("\n"):
        if not line:
            continue
        f = line.split("\t")
        pos, cigar, seq = int(f[3]) - 1, f[5], f[9]
        t, q, cols = [], [], []          # cols: target position of every column (insert columns carry the position before)
        rp, qp = pos, 0
        for n, op in re.findall(r"(\d+)([MIDNSHP=X])", cigar):
            n = int(n)
            if op in "M=X":
                for k in range(n):
                    t.append(draft[rp + k]), q.append(seq[qp + k]), cols.append(rp + k)
                rp += n
                qp += n
            elif op == "I":
                for k in range(n):
                    t.append("-"), q.append(seq[qp + k]), cols.append(rp - 1)
                qp += n
            elif op == "D":
                for k in range(n):
                    t.append(draft[rp + k]), q.append("-"), cols.append(rp + k)
                rp += n
            elif op == "S":
                qp += n
        # clip to the window; first and last column must be match columns (the reference anchors alignments on exact
        # 8-mers, get_align_shift ctg_cns.c:139)
        idx = [i for i in range(len(t)) if ws <= cols[i] < we]
        while idx and not (t[idx[0]] != "-" and q[idx[0]] != "-"):
            idx.pop(0)
        while idx and not (t[idx[-1]] != "-" and q[idx[-1]] != "-"):
            idx.pop()
        if len(idx) < 50:
            continue
        a, b = idx[0], idx[-1] + 1
        alns.append((cols[a] - ws, "".join(t[a:b]), "".join(q[a:b])))
    return alns


def md5s(res):
    pos, base = res
    return {"n": len(base), "pos_md5": hashlib.md5(pos.astype("<u4").tobytes()).hexdigest(), "base_md5": hashlib.md5(base).hexdigest()}


GEOMETRIES = [(5000000, 1000000), (20000, 4000), (9000, 1000)]


def ref_contig_windows(S, bam, ctg, seq, rt, w, ovl):
    """np2_ref_contig_windows on the reference's view of the draft -> [(start, end, n_alns, hash, pos, base)]"""
    rf = C.create_string_buffer(len(seq) + 1)
    S.np2_ref_roundtrip.argtypes = [C.c_char_p, C.c_int, C.c_char_p]
    S.np2_ref_roundtrip(seq.encode(), len(seq), rf)
    MW, cap = 256, len(seq) * 8
    ws, we, wn = np.zeros(MW, np.int32), np.zeros(MW, np.int32), np.zeros(MW, np.int32)
    wh, woff = np.zeros(MW, np.uint64), np.zeros(MW + 1, np.int64)
    pos, base = np.zeros(cap, np.uint32), np.zeros(cap, np.uint8)
    f = S.np2_ref_contig_windows
    f.restype = C.c_int
    f.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7 + [C.c_int64]
    n = f(bam.encode(), ctg.encode(), rf.value, len(seq), rt, w, ovl, 4, MW, ws.ctypes.data, we.ctypes.data, wn.ctypes.data,
          wh.ctypes.data, woff.ctypes.data, pos.ctypes.data, base.ctypes.data, cap)
    assert n > 0, n
    return [(int(ws[i]), int(we[i]), int(wn[i]), int(wh[i]), pos[woff[i]:woff[i + 1]].copy(), base[woff[i]:woff[i + 1]].tobytes()) for i in range(n)]


def ref_contig_fast(S, bam, ctg, seq, rt, w, ovl):
    """np2_ref_contig_fast -> (length, linked sequence).  Window sizes must stay well above the overlap and the overlap
    above ~100 positions: link_consensus_fast has no bounds on its walks (tiny windows make the reference spin)."""
    rf = C.create_string_buffer(len(seq) + 1)
    S.np2_ref_roundtrip.argtypes = [C.c_char_p, C.c_int, C.c_char_p]
    S.np2_ref_roundtrip(seq.encode(), len(seq), rf)
    cap = len(seq) * 2 + 1000
    out = C.create_string_buffer(cap)
    f = S.np2_ref_contig_fast
    f.restype = C.c_int64
    f.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int64]
    n = f(bam.encode(), ctg.encode(), rf.value, len(seq), rt, w, ovl, out, cap)
    assert n > 0, n
    return n, out.raw[:n]


PROD_CLEAN = {
    "hifi20": dict(L.CASES["hifi20"]),
    "hifi_exact": dict(seed=31, length=3000, depth=20, read_len=1500, sub=0.0, ins=0.0, dele=0.0, read_type=3),
    "hifi_deep": dict(seed=32, length=5000, depth=40, read_len=2000, sub=0.001, ins=0.001, dele=0.001, read_type=3),
    "clr_exact": dict(seed=33, length=2500, depth=25, read_len=900, sub=0.0, ins=0.0, dele=0.0, read_type=2),
    "ont_exact_shallow": dict(seed=34, length=2000, depth=3, read_len=700, sub=0.0, ins=0.0, dele=0.0, read_type=1),
    # noisy reads under their own read type's rules: the production stages find nothing they change here either
    "ont30/ont": dict(L.CASES["ont30"], read_type=1), "ont30/clr": dict(L.CASES["ont30"], read_type=2),
    "clr30/clr": dict(L.CASES["clr30"], read_type=2), "clr_hp/clr": dict(L.CASES["clr_hp"], read_type=2),
    "hifi20/ont": dict(L.CASES["hifi20"], read_type=1), "hifi20/clr": dict(L.CASES["hifi20"], read_type=2),
}
# ... and windows the missing stages DO change (recorded so that the gap stays visible in the tests)
PROD_CHANGED = {"clr30/ont": dict(L.CASES["clr30"], read_type=1), "ont30/hifi": dict(L.CASES["ont30"], read_type=3)}


def ref_window_prod(S, case):
    cap = case["len"] * 3 + int(case["aln_len"].sum())
    pos, base, qv = np.zeros(cap, np.uint32), np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
    f = S.np2_ref_window_prod
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    n = f(case["read_type"], len(case["aln_t_s"]), case["aln_t_s"].ctypes.data, case["aln_len"].ctypes.data, case["str_off"].ctypes.data,
          case["t_str"], case["q_str"], case["len"], case["min_cov"], pos.ctypes.data, base.ctypes.data, qv.ctypes.data, cap)
    assert n > 0, n
    return n, pos[:n].copy(), base[:n].tobytes(), qv[:n].copy()


def front_goldens(S, fa_path, bam):
    out = {}
    for ctg, seq in read_fa(fa_path).items():
        for w, ovl in GEOMETRIES:
            for rt in (1, 3):
                wins = ref_contig_windows(S, bam, ctg, seq, rt, w, ovl)
                out["%s/w%d_o%d/rt%d" % (ctg, w, ovl, rt)] = [
                    {"start": a, "end": b, "n_alns": n, "hash": "%016x" % h, "n": len(base),
                     "pos_md5": hashlib.md5(pos.astype("<u4").tobytes()).hexdigest(), "base_md5": hashlib.md5(base).hexdigest()}
                    for a, b, n, h, pos, base in wins]
    return out


def main():
    tmp = tempfile.mkdtemp(prefix="npgold_lgs")
    bam = os.path.join(tmp, "lgs.sort.bam")
    subprocess.check_call(f"{REF}/minimap2 -ax map-ont -t 1 {TD}/raw.genome.fasta {TD}/lreads.fasta.gz 2>/dev/null | "
                          f"{REF}/samtools view -b -F 0x4 - | {REF}/samtools sort -o {bam} - 2>/dev/null && {REF}/samtools index {bam}",
                          shell=True, executable="/bin/bash")
    draft = read_fa(os.path.join(TD, "raw.genome.fasta"))
    gold = {"whole_path": {}, "first_pass": {}}
    # (a) the whole path through the reference's own boundary (nextpolish2.py:54-65)
    class CT(C.Structure):
        _fields_ = [("len", C.c_uint), ("identity", C.c_float), ("seq", C.c_char_p)]

    class CTD(C.Structure):
        _fields_ = [("data", C.POINTER(CT)), ("i_m", C.c_int)]

    class REFT(C.Structure):
        _fields_ = [("n", C.c_char_p), ("s", C.POINTER(C.c_uint32)), ("qv", C.c_void_p), ("qv_l", C.c_uint32), ("length", C.c_uint32)]

    class REFS(C.Structure):
        _fields_ = [("ref", C.POINTER(REFT)), ("i", C.c_uint32), ("i_m", C.c_uint32)]
    P = C.CDLL(os.path.join(REF, "nextpolish2.so"))
    P.read_ref.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int]
    P.read_ref.restype = C.POINTER(REFS)
    P.ctg_cns_init.argtypes = [C.c_int] * 3 + [C.c_float] * 3
    P.ctg_cns_init.restype = C.c_void_p
    P.ctg_cns_core.argtypes = [C.c_void_p, C.POINTER(REFT), C.c_char_p]
    P.ctg_cns_core.restype = C.POINTER(CTD)
    names = [n.encode() for n in draft]
    lst = os.path.join(tmp, "lgs.list")
    open(lst, "w").write(bam + "\n")
    refs = P.read_ref(os.path.join(TD, "raw.genome.fasta").encode(), (C.c_char_p * len(names))(*names), len(names))
    cfg = P.ctg_cns_init(5000000, 1, 1, 0.8, 0.8, 0.8)
    for i in range(refs.contents.i):
        r = P.ctg_cns_core(cfg, refs.contents.ref[i], lst.encode())
        gold["whole_path"][refs.contents.ref[i].n.decode()] = [
            {"len": r.contents.data[k].len, "md5": hashlib.md5(r.contents.data[k].seq).hexdigest()} for k in range(r.contents.i_m)]
    # (b) the first pass
    S = L.ref_shim()
    packed = {}
    for ctg, ws, we in WINDOWS:
        alns = window_alignments(bam, draft[ctg], ctg, ws, we)
        case = L.pack_case(alns, we - ws, 1)
        key = "td_%s_%d_%d" % (ctg, ws, we)
        for k in ("aln_t_s", "aln_len", "str_off"):
            packed[key + "." + k] = case[k]
        packed[key + ".t_str"] = np.frombuffer(case["t_str"], np.uint8)
        packed[key + ".q_str"] = np.frombuffer(case["q_str"], np.uint8)
        packed[key + ".len"] = np.array([we - ws], np.int64)
        for rt in (1, 2, 3, 4):
            case["read_type"] = rt
            gold["first_pass"]["%s/rt%d" % (key, rt)] = md5s(L.first_pass_via(S.np2_ref_first_pass, case))
    np.savez_compressed(os.path.join(OUT, "lgs_td_windows.npz"), **packed)
    for name, kw in L.CASES.items():
        for rt in (1, 2, 3, 4):
            case = L.synthetic_case(**dict(kw, read_type=rt))
            gold["first_pass"]["%s/rt%d" % (name, rt)] = md5s(L.first_pass_via(S.np2_ref_first_pass, case))
    # (c) the stage in front: BAM records -> alignment strings (np2_ref_contig_windows: the reference's own functions in
    # the record loop of ctg_cns_core).  Fixture: a 30 % subsample of the BAM above (lgs_td.bam + .bai, ~1 MB) and the
    # draft with a few N / lower-case bases written into it (what read_ref's 2-bit packing does to them is part of the pin).
    import random
    import shutil
    sub = os.path.join(OUT, "lgs_td.bam")
    subprocess.check_call(f"{REF}/samtools view -b -s 7.3 -o {sub} {bam} && {REF}/samtools index {sub}", shell=True)
    rng = random.Random(5)
    fa_path = os.path.join(OUT, "lgs_td.fa")
    with open(fa_path, "w") as f:
        for name, seq in draft.items():
            d = list(seq)
            for _ in range(40):
                i = rng.randrange(len(d))
                d[i] = rng.choice(["N", "n", d[i].lower(), "R"])
            f.write(">%s\n" % name)
            for i in range(0, len(d), 70):
                f.write("".join(d[i:i + 70]) + "\n")
    gold["from_bam"] = front_goldens(S, fa_path, sub)
    # (d) the reference's fast mode end to end (np2_ref_contig_fast: first pass of every window + link_consensus_fast)
    gold["fast_mode"] = {}
    for ctg, seq in read_fa(fa_path).items():
        for w, ovl in GEOMETRIES:
            for rt in (1, 3):
                n, linked = ref_contig_fast(S, sub, ctg, seq, rt, w, ovl)
                gold["fast_mode"]["%s/w%d_o%d/rt%d" % (ctg, w, ovl, rt)] = {"len": n, "md5": hashlib.md5(linked).hexdigest()}
    # (e) the PRODUCTION window consensus (np2_ref_window_prod: get_cns_from_align_tags with fast = 0 — first pass, low-quality
    # regions, POA, second round) on accurate reads: where it finds nothing to re-polish it must equal the first pass with
    # the production letter-case rule (qv > 20), which pins that rule and the qv values
    gold["production_clean"] = {}
    for name, kw in PROD_CLEAN.items():
        case = L.synthetic_case(**kw)
        n, pos, base, qv = ref_window_prod(S, case)
        gold["production_clean"][name] = {"n": n, "base_md5": hashlib.md5(base).hexdigest()}
    gold["production_changed"] = {}
    for name, kw in PROD_CHANGED.items():
        n, pos, base, qv = ref_window_prod(S, L.synthetic_case(**kw))
        gold["production_changed"][name] = {"n": n, "base_md5": hashlib.md5(base).hexdigest()}
    json.dump(gold, open(os.path.join(OUT, "lgs_golden.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(gold["first_pass"]), "first-pass goldens;", gold["whole_path"])


if __name__ == "__main__":
    main()
