#!/usr/bin/env python
"""GPU BGZF inflate throughput on a synthetic BAM of the bench shape, next to zlib on one host core.
usage: bench_inflate.py [contig_len] [n_contigs]  -> one JSON line"""
import ctypes as C
import gzip
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from nextpolish_b200 import engine as E  # noqa: E402


def main():
    L = E.lib()
    clen = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    nctg = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    with tempfile.TemporaryDirectory() as tmp:
        fa, bam = os.path.join(tmp, "x.fa"), os.path.join(tmp, "x.bam")
        p = E.synth_params(seed=20240919, n_contigs=nctg, contig_len=clen, depth=30.0)
        p.compress_level = 6
        assert L.np_synth_write(p, fa.encode(), bam.encode()) == 0
        comp = open(bam, "rb").read()
    buf = np.frombuffer(comp, dtype=np.uint8)
    n, nb, ms = C.c_int64(0), C.c_int32(0), C.c_float(0)
    assert L.np_bgzf_inflate(0, buf.ctypes.data, len(comp), None, 0, C.byref(n), C.byref(nb), None) == 0
    out = np.zeros(n.value, np.uint8)
    best, wall = 1e9, 1e9
    for _ in range(5):
        t0 = time.time()
        rc = L.np_bgzf_inflate(0, buf.ctypes.data, len(comp), out.ctypes.data, n.value, C.byref(n), C.byref(nb), C.byref(ms))
        wall = min(wall, time.time() - t0)
        assert rc == 0, E.last_error()
        best = min(best, ms.value)
    import struct
    import zlib
    t0 = time.time()
    parts, off = [], 0
    while off < len(comp):                      # host baseline: zlib on every block's raw deflate payload, one core
        xlen = struct.unpack_from("<H", comp, off + 10)[0]
        bsize = struct.unpack_from("<H", comp, off + 16)[0] + 1
        parts.append(zlib.decompress(comp[off + 12 + xlen:off + bsize - 8], -15))
        off += bsize
    ref = b"".join(parts)
    host_s = time.time() - t0
    assert out.tobytes() == ref
    print(json.dumps({"what": "BGZF inflate, one warp per block (k_bgzf_inflate)", "bam_bytes": len(comp), "inflated_bytes": n.value,
                      "blocks": nb.value, "kernel_ms": best, "kernel_out_GBps": n.value / best / 1e6, "kernel_in_GBps": len(comp) / best / 1e6,
                      "call_wall_ms_pageable_host_buffers": wall * 1e3, "host_zlib_1core_s": host_s,
                      "host_zlib_1core_out_GBps": n.value / host_s / 1e9, "identical_to_zlib": True}))


if __name__ == "__main__":
    main()
