#!/usr/bin/env python
"""Stability of the from-files pipeline: ms per 2-task step over consecutive groups of 40 steps at one depth.
usage: prof_modes.py [depth] [groups]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nextpolish_b200 import engine as E  # noqa: E402

tmp = tempfile.mkdtemp(prefix="npfiles")
files = bench.write_inputs(tmp, 0, [1, 2])
cfg = E.default_config(b"")
cfg.contents.read_tlen = 1750
for depth in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "6").split(",")]:
    pipe = E.FilePipeline(0, depth=depth)

    def run(n):
        t0 = time.time()
        for i in range(n):
            for t in (1, 2):
                pipe.submit(t, files[t][0], files[t][1], cfg)
                while pipe.in_flight() > pipe.capacity - 1:
                    pipe.wait_oldest(want_md5=False)
        while pipe.in_flight():
            pipe.wait_oldest(want_md5=False)
        return (time.time() - t0) / n * 1e3
    run(12)
    print("depth %d:" % depth, " ".join("%.1f" % run(40) for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 8)), "| 200 steps without a drain: %.1f" % run(200), flush=True)
    pipe.close()
