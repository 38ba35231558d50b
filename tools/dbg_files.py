import sys, os, faulthandler
sys.path.insert(0, '.')
import bench
from nextpolish_b200 import engine as E
E.lib()
import tempfile
tmp = tempfile.mkdtemp()
files = bench.write_inputs(tmp, 0, [1, 2])
print("inputs written", flush=True)
cfg = E.default_config(b""); cfg.contents.read_tlen = 1750
fp = E.FilePipeline(0, depth=2)
print("pipeline created", flush=True)
for it in range(3):
    for t in (1, 2):
        fp.submit(t, files[t][0], files[t][1], cfg)
        print("submitted", t, flush=True)
        while fp.in_flight() > 1:
            r = fp.wait_oldest(); print("done", r["task"], r["load_ms"], r["polish_ms"], flush=True)
while fp.in_flight():
    r = fp.wait_oldest(); print("done", r["task"], r["load_ms"], r["polish_ms"], flush=True)
fp.close()
print("closed")
