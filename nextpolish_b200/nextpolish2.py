"""Host-side handle of the long-read consensus path's first slice (include/nextpolish2_b200.h, csrc/lgs_consensus.cu;
reference boundary: source/lib/nextpolish2.py:54-65 over nextpolish2.so).  Plumbing only: all compute is in
nextpolish_b200/lib/nextpolish2.so, which has no CPU path.

    eng = LgsEngine(device=0)
    res = eng.first_pass(windows, read_type=1, min_cov=4)      # [(pos uint32[], base bytes, qv uint8[])] per window
    cw  = ContigWindows(fasta, bam, "ctg1", read_type=1)       # windows + alignment strings of a contig from an indexed BAM (host)
    res = eng.first_pass_contig(cw)
    seq = eng.polish_contig_fast(fasta, bam, "ctg1", 1)        # files -> windows -> first pass -> linked contig (the reference's fast mode)
    python -m nextpolish_b200.nextpolish2 -g genome.fa -l lgs.sort.bam.list -r ont --fast -o out.fa      # lib/nextpolish2.py's command line

A window is a dict  len, aln_t_s uint32[n], aln_len uint32[n], str_off uint64[n], t_str bytes, q_str bytes  — the gapped
alignment strings of the reads against the window, window against itself first (ctg_cns.c:3456-3468)."""
import ctypes as C
import os

import numpy as np

LIB2_PATH = os.path.join(os.path.dirname(os.path.realpath(__file__)), "lib", "nextpolish2.so")
EXPORTS2 = ["np2_engine_create", "np2_engine_destroy", "np2_last_error", "np2_first_pass", "np2_engine_launch_count",
            "np2_engine_last_stats", "np2_engine_kernel_times",
            "np2_windows_from_bam", "np2_windows_from_bams", "np2_windows_count", "np2_windows_info", "np2_windows_batch", "np2_windows_free",
            "np2_windows_starts", "np2_link_windows_fast"]                     # every symbol include/nextpolish2_b200.h declares
ERRORS = {-1: "output capacity too small", -2: "a window's last position has no node",
          -3: "an alignment is empty, starts on a gap column or leaves its window",
          -4: "backtrack through a node without links", -5: "size limit", -6: "CUDA failure"}


class WindowBatch(C.Structure):  # np2_window_batch
    _fields_ = [("n_windows", C.c_int32), ("win_len", C.c_void_p), ("win_aln0", C.c_void_p), ("read_type", C.c_int32),
                ("min_cov", C.c_int32), ("aln_t_s", C.c_void_p), ("aln_len", C.c_void_p), ("str_off", C.c_void_p),
                ("t_str", C.c_char_p), ("q_str", C.c_char_p), ("str_bytes", C.c_int64)]


class NativeError(RuntimeError):
    pass


_LIB2 = None


def lib2():
    global _LIB2
    if _LIB2 is None:
        if not os.path.exists(LIB2_PATH):
            raise OSError("%s is not built: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB2_PATH)
        L = C.CDLL(LIB2_PATH)
        L.np2_engine_create.argtypes = [C.c_int32]
        L.np2_engine_create.restype = C.c_void_p
        L.np2_engine_destroy.argtypes = [C.c_void_p]
        L.np2_engine_destroy.restype = None
        L.np2_last_error.restype = C.c_char_p
        L.np2_first_pass.argtypes = [C.c_void_p, C.POINTER(WindowBatch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.np2_first_pass.restype = C.c_int64
        L.np2_engine_launch_count.argtypes = [C.c_void_p]
        L.np2_engine_launch_count.restype = C.c_int64
        L.np2_engine_last_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.np2_engine_last_stats.restype = None
        L.np2_engine_kernel_times.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.np2_windows_from_bam.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_int32]
        L.np2_windows_from_bam.restype = C.c_void_p
        L.np2_windows_from_bams.argtypes = [C.c_char_p, C.c_void_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32, C.c_int32]
        L.np2_windows_from_bams.restype = C.c_void_p
        L.np2_windows_count.argtypes = [C.c_void_p]
        L.np2_windows_info.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.np2_windows_info.restype = None
        L.np2_windows_batch.argtypes = [C.c_void_p, C.POINTER(WindowBatch)]
        L.np2_windows_batch.restype = None
        L.np2_windows_free.argtypes = [C.c_void_p]
        L.np2_windows_free.restype = None
        L.np2_windows_starts.argtypes = [C.c_void_p, C.c_void_p]
        L.np2_windows_starts.restype = None
        L.np2_link_windows_fast.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]
        L.np2_link_windows_fast.restype = C.c_int64
        _LIB2 = L
    return _LIB2


def make_batch(windows, read_type, min_cov=4):
    """Several windows as one np2_window_batch; returns (struct, keep-alive list of the arrays it points into)."""
    win_len = np.array([w["len"] for w in windows], np.int32)
    win_aln0 = np.zeros(len(windows) + 1, np.int32)
    t_s, a_len, s_off, t_all, q_all, pos = [], [], [], [], [], 0
    for i, w in enumerate(windows):
        win_aln0[i + 1] = win_aln0[i] + len(w["aln_t_s"])
        t_s.append(np.asarray(w["aln_t_s"], np.uint32)), a_len.append(np.asarray(w["aln_len"], np.uint32))
        s_off.append(np.asarray(w["str_off"], np.uint64) + np.uint64(pos))
        t_all.append(w["t_str"]), q_all.append(w["q_str"])
        pos += len(w["t_str"])
    cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs).astype(dt)) if xs else np.zeros(0, dt)
    t_s, a_len, s_off = cat(t_s, np.uint32), cat(a_len, np.uint32), cat(s_off, np.uint64)
    t_str, q_str = b"".join(t_all), b"".join(q_all)
    b = WindowBatch(len(windows), win_len.ctypes.data, win_aln0.ctypes.data, read_type, min_cov, t_s.ctypes.data, a_len.ctypes.data,
                    s_off.ctypes.data, t_str, q_str, len(t_str))
    return b, [win_len, win_aln0, t_s, a_len, s_off, t_str, q_str]


def split_result(n, pos, base, qv, off, n_windows):
    return [(pos[off[i]:off[i + 1]].copy(), base[off[i]:off[i + 1]].tobytes(), qv[off[i]:off[i + 1]].copy()) for i in range(n_windows)]


class ContigWindows:
    """np2_windows_from_bam: the consensus windows of one contig with their alignment strings, built on the host from the
    draft FASTA and an indexed long-read BAM (the record loop of ctg_cns_core, ctg_cns.c:3444-3566)."""

    def __init__(self, fasta, bam, contig, read_type, window=5000000, overlap=1000000):
        """bam: one path, or a list of paths (merged in the reference's order, bsort.c:174-199)"""
        bams = [bam] if isinstance(bam, str) else list(bam)
        arr = (C.c_char_p * len(bams))(*[b.encode() for b in bams])
        self.h = lib2().np2_windows_from_bams(fasta.encode(), arr, len(bams), contig.encode(), read_type, window, overlap)
        if not self.h:
            raise NativeError(lib2().np2_last_error().decode(errors="replace"))
        self.read_type = read_type
        self.batch = WindowBatch()
        lib2().np2_windows_batch(self.h, C.byref(self.batch))

    def info(self):
        """[(start, end, n_alignments, hash)] per window"""
        out = []
        for i in range(lib2().np2_windows_count(self.h)):
            s, e, n, h = C.c_int32(), C.c_int32(), C.c_int32(), C.c_uint64()
            lib2().np2_windows_info(self.h, i, C.byref(s), C.byref(e), C.byref(n), C.byref(h))
            out.append((s.value, e.value, n.value, h.value))
        return out

    def windows(self):
        """The windows as dicts (the form LgsEngine.first_pass and the oracle helpers take); copies."""
        b, out = self.batch, []
        n_aln = (C.c_int32 * (b.n_windows + 1)).from_address(b.win_aln0)
        wl = (C.c_int32 * b.n_windows).from_address(b.win_len)
        tot = n_aln[b.n_windows]
        t_s = np.frombuffer((C.c_uint32 * tot).from_address(b.aln_t_s), np.uint32) if tot else np.zeros(0, np.uint32)
        a_l = np.frombuffer((C.c_uint32 * tot).from_address(b.aln_len), np.uint32) if tot else np.zeros(0, np.uint32)
        s_o = np.frombuffer((C.c_uint64 * tot).from_address(b.str_off), np.uint64) if tot else np.zeros(0, np.uint64)
        t_all = C.string_at(b.t_str, b.str_bytes)
        q_all = C.string_at(b.q_str, b.str_bytes)
        for i in range(b.n_windows):
            lo, hi = n_aln[i], n_aln[i + 1]
            b0 = int(s_o[lo]) if hi > lo else 0
            b1 = int(s_o[hi - 1] + a_l[hi - 1]) if hi > lo else 0
            out.append(dict(len=int(wl[i]), read_type=self.read_type, min_cov=b.min_cov, aln_t_s=t_s[lo:hi].copy(), aln_len=a_l[lo:hi].copy(),
                            str_off=(s_o[lo:hi] - np.uint64(b0)).astype(np.uint64), t_str=t_all[b0:b1], q_str=q_all[b0:b1]))
        return out

    def close(self):
        if self.h:
            lib2().np2_windows_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def production_case(base, qv):
    """The letter case of the reference's production first pass (generate_cns_from_best_score, ctg_cns.c:1839-1846): upper case
    only where coverage > min_cov (already in `base`, the fast variant's rule) AND the link quality qv > LQBASE_MIN_QV = 20."""
    b = np.frombuffer(base, np.uint8).copy()
    low = (np.asarray(qv) <= 20) & (b >= 65) & (b <= 90)
    b[low] |= 32
    return b.tobytes()


def link_windows_fast(starts, results, overlap):
    """np2_link_windows_fast over per-window first-pass results [(pos, base, qv)] -> the linked sequence (bytes)."""
    n = len(results)
    ws = np.asarray(starts, np.int32)
    off = np.zeros(n + 1, np.int64)
    for i, r in enumerate(results):
        off[i + 1] = off[i] + len(r[1])
    pos = np.concatenate([r[0] for r in results]).astype(np.uint32) if n else np.zeros(0, np.uint32)
    base = np.frombuffer(b"".join(r[1] for r in results), np.uint8).copy() if n else np.zeros(0, np.uint8)
    out = np.zeros(int(off[n]) + 16, np.uint8)
    m = lib2().np2_link_windows_fast(n, ws.ctypes.data, off.ctypes.data, pos.ctypes.data, base.ctypes.data, overlap, out.ctypes.data, len(out))
    if m < 0:
        raise NativeError("np2_link_windows_fast: %d: %s" % (m, lib2().np2_last_error().decode(errors="replace")))
    return out[:m].tobytes()


class LgsEngine:
    def __init__(self, device=0):
        self.h = lib2().np2_engine_create(device)
        if not self.h:
            raise NativeError(lib2().np2_last_error().decode(errors="replace"))

    def first_pass_contig(self, cw):
        """First pass of every window of a ContigWindows handle (no copies of the strings)."""
        b = cw.batch
        cap = int(b.str_bytes) + 16
        pos, base, qv = np.zeros(cap, np.uint32), np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
        off = np.zeros(b.n_windows + 1, np.int64)
        n = lib2().np2_first_pass(self.h, C.byref(b), pos.ctypes.data, base.ctypes.data, qv.ctypes.data, cap, off.ctypes.data)
        if n < 0:
            raise NativeError("np2_first_pass: %d (%s): %s" % (n, ERRORS.get(n, "?"), lib2().np2_last_error().decode(errors="replace")))
        return split_result(n, pos, base, qv, off, b.n_windows)

    def polish_contig_fast(self, fasta, bam, contig, read_type, window=5000000, overlap=1000000):
        """FASTA + indexed long-read BAM -> the contig's consensus in the reference's fast mode: windows from the BAM
        (host), first pass (GPU), link_consensus_fast (host)."""
        cw = ContigWindows(fasta, bam, contig, read_type, window, overlap)
        try:
            res = self.first_pass_contig(cw)
            return link_windows_fast([s for s, _, _, _ in cw.info()], res, overlap)
        finally:
            cw.close()

    def first_pass(self, windows, read_type, min_cov=4):
        b, keep = make_batch(windows, read_type, min_cov)
        cap = sum(int(w["len"]) * 2 + int(np.asarray(w["aln_len"]).sum()) for w in windows) + 16
        pos, base, qv = np.zeros(cap, np.uint32), np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
        off = np.zeros(len(windows) + 1, np.int64)
        n = lib2().np2_first_pass(self.h, C.byref(b), pos.ctypes.data, base.ctypes.data, qv.ctypes.data, cap, off.ctypes.data)
        del keep
        if n < 0:
            raise NativeError("np2_first_pass: %d (%s): %s" % (n, ERRORS.get(n, "?"), lib2().np2_last_error().decode(errors="replace")))
        return split_result(n, pos, base, qv, off, len(windows))

    def stats(self):
        out = (C.c_int64 * 4)()
        lib2().np2_engine_last_stats(self.h, out)
        return dict(segments=out[0], reruns=out[1], stitch_iterations=out[2], link_records=out[3], launches=lib2().np2_engine_launch_count(self.h))

    def kernel_times(self):
        """[(launch name, ms)] of the last first_pass (needs NEXTPOLISH_B200_LGS_TIMING=1 in the environment)."""
        names = (C.c_char_p * 4096)()
        ms = (C.c_float * 4096)()
        n = lib2().np2_engine_kernel_times(self.h, names, ms, 4096)
        return [(names[i].decode(), ms[i]) for i in range(n)]

    def close(self):
        if self.h:
            lib2().np2_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- host-side mirror of the reference's long-read worker (source/lib/nextpolish2.py) -------------------------------------
def _read_corrected(path, corrected):
    """read_corrected_seqs (nextpolish2.py:116-137): finished contigs of an existing output and the offset at which its
    last record (possibly partial; `_s<k>` pieces of a split contig belong together) starts."""
    last, cur, pos = "", 0, 0
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                parts = line.split()[0].split("_s")
                last = parts[0][1:]
                if len(parts) == 1 or parts[1] == "0":
                    pos += cur
                    cur = len(line)
                else:
                    cur += len(line)
                corrected.add(last)
            else:
                cur += len(line)
    if last:
        corrected.discard(last)
    return pos


def _read_uncorrected(path, index, corrected):
    """read_uncorrected_seqs (nextpolish2.py:98-114), in file order."""
    names = []
    with open(path) as f:
        for line in f:
            if index != "all":
                p = line.strip().split()
                if p and p[0] not in corrected and len(p) > 1 and p[1] == index and p[0] not in names:
                    names.append(p[0])
            elif line.startswith(">"):
                n = line.strip().split()[0][1:]
                if n not in corrected and n not in names:
                    names.append(n)
    return names


def main(argv=None):
    """python -m nextpolish_b200.nextpolish2 -g genome.fa -l lgs.sort.bam.list -r ont --fast [-b blc -i 0] [-o out] [-u] [-w 5M]

    The command line of the reference's lib/nextpolish2.py (block file, resume, `>name len` records, -u, -w) over this
    engine.  Only the reference's FAST mode exists here (first pass + link_consensus_fast), so the
    mirror insists on --fast: without it it refuses instead of printing something the reference's default run would not."""
    import argparse
    import sys
    ap = argparse.ArgumentParser(description="Long-read polish on the GPU, the reference's fast mode (mirror of lib/nextpolish2.py).")
    ap.add_argument("-g", "--genome", required=True)
    ap.add_argument("-l", "--bam_list", required=True, help="file listing the sorted, indexed long-read BAMs, one per line")
    ap.add_argument("-r", "--read_type", required=True, type=str.lower, choices=["clr", "hifi", "ont"])
    ap.add_argument("-b", "--block")
    ap.add_argument("-i", "--block_index", default="all")
    ap.add_argument("-o", "--out", default="stdout")
    ap.add_argument("-p", "--process", type=int, default=10)
    ap.add_argument("-u", "--uppercase", action="store_true")
    ap.add_argument("-w", "--window", default="5M")
    ap.add_argument("-a", "--auto", action="store_false", default=True)
    ap.add_argument("-sp", "--split", action="store_false", default=True)
    ap.add_argument("-id", "--alignment_identity_ratio", type=float, default=0.8)
    ap.add_argument("-as", "--alignment_score_ratio", type=float, default=0.8)
    ap.add_argument("--fast", action="store_true", help="the reference's fast mode (ctg_cns.c:3433,3620): the only one built")
    ap.add_argument("--plan", action="store_true", help="print the contigs this job would polish and the BAMs it would read, touch no device")
    args, _unknown = ap.parse_known_args(argv)
    if not args.fast:
        sys.stderr.write("only the reference's fast mode is built (pass --fast); its default mode re-polishes low-quality regions "
                         "(ctg_cns.c:822-1474), which this engine does not do yet: use the reference's nextpolish2.py for it\n")
        return 1
    unit = {"k": 1e3, "m": 1e6, "g": 1e9}.get(args.window[-1].lower())
    window = int(float(args.window[:-1]) * unit) if unit else int(args.window)
    window = max(window, 5000000)                      # set_window_process never goes below 5 M (nextpolish2.py:75-76)
    bams = [l.strip() for l in open(args.bam_list) if l.strip()]
    if not bams:
        sys.stderr.write("the BAM list is empty\n")
        return 1
    rt = {"ont": 1, "clr": 2, "hifi": 3}[args.read_type]
    out, corrected = sys.stdout, set()
    if args.out != "stdout":
        if os.path.exists(args.out):
            at = _read_corrected(args.out, corrected)
            out = open(args.out, "r+")
            out.seek(at)
            out.truncate()
        else:
            out = open(args.out, "w")
    block = args.genome if (args.block_index == "all" or not args.block) else args.block
    names = _read_uncorrected(block, "all" if block == args.genome else args.block_index, corrected)
    rc = 0
    if args.plan:
        sys.stdout.write("".join("bam\t%s\n" % b for b in bams) + "".join("polish\t%s\n" % n for n in names))
        names = []
    if names:
        # the windows of the next contig are built on a host thread (np2_windows_from_bams: file reads, inflate, alignment
        # strings; the ctypes call drops the GIL) while the GPU works on the current one
        from concurrent.futures import ThreadPoolExecutor
        eng = LgsEngine(int(os.environ.get("NEXTPOLISH_B200_DEVICE", "0")))
        pool = ThreadPoolExecutor(1)
        load = lambda n: ContigWindows(args.genome, bams, n, rt, window, 1000000)
        pending = pool.submit(load, names[0])
        for k, name in enumerate(names):
            cw = pending.result()
            if k + 1 < len(names):
                pending = pool.submit(load, names[k + 1])
            try:
                seq = link_windows_fast([s for s, _, _, _ in cw.info()], eng.first_pass_contig(cw), 1000000).decode()
            finally:
                cw.close()
            if args.uppercase:
                seq = seq.upper()
            if len(seq) > 10:                          # nextpolish2.py:199-203
                out.write(">%s %d\n%s\n" % (name, len(seq), seq))
            else:
                sys.stderr.write("Failed to correct sequence: %s\n" % name)
                rc = 1
                break
        pool.shutdown(wait=True)
        eng.close()
    if out is not sys.stdout:
        out.close()
    return rc


if __name__ == "__main__":
    import sys
    sys.exit(main())
