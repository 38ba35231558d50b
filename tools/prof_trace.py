#!/usr/bin/env python
"""Phase trace (NEXTPOLISH_B200_TRACE=1) of loads running inside the pipelined from-files front end at a given depth:
what a job spends its wall time on when several jobs are in flight.  usage: prof_trace.py [depth]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nextpolish_b200 import engine as E  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 6
tmp = tempfile.mkdtemp(prefix="npfiles")
files = bench.write_inputs(tmp, 0, [1, 2])
cfg = E.default_config(b"")
cfg.contents.read_tlen = 1750
pipe = E.FilePipeline(0, depth=depth)


def run(n):
    acc = []
    t0 = time.time()
    for i in range(n):
        for t in (1, 2):
            pipe.submit(t, files[t][0], files[t][1], cfg)
            while pipe.in_flight() > depth - 1:
                acc.append(pipe.wait_oldest(want_md5=False))
    while pipe.in_flight():
        acc.append(pipe.wait_oldest(want_md5=False))
    return (time.time() - t0) / n * 1e3, acc


for rep in range(4):
    ms, acc = run(30)
    print("depth %d rep %d: %.2f ms per step; per job load %.1f polish %.1f ms" % (
        depth, rep, ms, sum(r["load_ms"] for r in acc) / len(acc), sum(r["polish_ms"] for r in acc) / len(acc)), flush=True)
os.environ["NEXTPOLISH_B200_TRACE"] = "1"
ms, acc = run(4)
os.environ.pop("NEXTPOLISH_B200_TRACE")
print("traced: %.2f ms per step" % ms)
pipe.close()
