#!/usr/bin/env python
"""Two jobs (task 1, task 2) of the from-files path on the bench files, depth 1: for ncu launch lists."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
import bench
from nextpolish_b200 import engine as E
tmp = tempfile.mkdtemp(prefix="npfiles")
files = bench.write_inputs(tmp, 0, [1, 2])
cfg = E.default_config(b""); cfg.contents.read_tlen = 1750
pipe = E.FilePipeline(0, depth=1)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    for t in (1, 2):
        pipe.submit(t, files[t][0], files[t][1], cfg)
        pipe.wait_oldest(want_md5=False)
pipe.close()
