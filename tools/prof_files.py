#!/usr/bin/env python
"""From-files path on the bench files (5 x 1 Mb, 30x; task 1 and task 2 inputs): phase trace of one load per task
(NEXTPOLISH_B200_TRACE=1), then steady-state ms per job through the pipelined front end.  usage: prof_files.py [depth]"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nextpolish_b200 import engine as E  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tmp = tempfile.mkdtemp(prefix="npfiles")
files = bench.write_inputs(tmp, 0, [1, 2])
cfg = E.default_config(b"")
cfg.contents.read_tlen = 1750
for t in (1, 2):
    print("task", t, "BAM bytes", os.path.getsize(files[t][1]))
pipe = E.FilePipeline(0, depth=1)
for rep in range(3):
    for t in (1, 2):
        if rep == 2:
            os.environ["NEXTPOLISH_B200_TRACE"] = "1"
            sys.stderr.write("---- task %d\n" % t)
        pipe.submit(t, files[t][0], files[t][1], cfg)
        r = pipe.wait_oldest(want_md5=False)
        if rep == 2:
            sys.stderr.write("load_ms %.2f polish_ms %.2f\n" % (r["load_ms"], r["polish_ms"]))
os.environ.pop("NEXTPOLISH_B200_TRACE", None)
pipe.close()
for dep in (1, depth):
    pipe = E.FilePipeline(0, depth=dep)
    for n in (4, 20):
        acc = {1: [0.0, 0.0, 0], 2: [0.0, 0.0, 0]}
        def take(r):
            a = acc[r["task"]]; a[0] += r["load_ms"]; a[1] += r["polish_ms"]; a[2] += 1
        t0 = time.time()
        for i in range(n):
            for t in (1, 2):
                pipe.submit(t, files[t][0], files[t][1], cfg)
                while pipe.in_flight() > dep - 1:
                    take(pipe.wait_oldest(want_md5=False))
        while pipe.in_flight():
            take(pipe.wait_oldest(want_md5=False))
        dt = time.time() - t0
        print("depth %d: %d steps, %.2f ms per 2-task step = %.0f Mbp/s; per job load/polish ms: task1 %.1f/%.1f task2 %.1f/%.1f"
              % (dep, n, dt / n * 1e3, 10.0 * n / dt, acc[1][0] / acc[1][2], acc[1][1] / acc[1][2], acc[2][0] / acc[2][2], acc[2][1] / acc[2][2]))
    pipe.close()
print("host cpus", os.cpu_count())
