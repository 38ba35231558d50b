#!/usr/bin/env python
"""Which part of bench.py's set-up slows the from-files pipeline down?  Builds the set-up in stages and measures the
pipeline (ms per 2-task step over 40 steps) after each one."""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nextpolish_b200 import engine as E  # noqa: E402

depth = int(os.environ.get("NP_BENCH_FILES_DEPTH", "6"))
tmp = tempfile.mkdtemp(prefix="npfiles")
files = bench.write_inputs(tmp, 0, [1, 2])
cfg = E.default_config(b"")
cfg.contents.read_tlen = 1750
pipe = E.FilePipeline(0, depth=depth)


def run(n):
    t0 = time.time()
    for i in range(n):
        for t in (1, 2):
            pipe.submit(t, files[t][0], files[t][1], cfg)
            while pipe.in_flight() > depth - 1:
                pipe.wait_oldest(want_md5=False)
    while pipe.in_flight():
        pipe.wait_oldest(want_md5=False)
    return (time.time() - t0) / n * 1e3


def report(what):
    print("%-60s %6.2f %6.2f ms per step" % (what, run(40), run(40)), flush=True)


run(40)
report("stage 0: pipeline alone")
import torch  # noqa: E402
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
x = torch.zeros(1 << 20, device=dev)
report("stage 1: + torch CUDA context use")
views, keep = {}, []
for t in (1, 2):
    views[t] = []
    for k in range(bench.N_ROTATE):
        sh = E.Shard.synthetic(E.synth_params(**bench.synth_kwargs(t, bench.seed_for(0, t, k))), 0, bench.WORKLOAD["n_contigs"],
                               with_qual=(2 if t == 2 else 0), threads=os.cpu_count() or 8)
        a = sh.arrays()
        pin = {k2: torch.from_numpy(v.copy()).pin_memory() for k2, v in a.items() if k2 in ("ctg_seq", "rec_off", "rec", "qual_off", "qual")}
        res = {k2: v.to(dev) for k2, v in pin.items()}
        v = E.ShardView()
        v.n_contigs, v.n_reads = sh.view.n_contigs, sh.view.n_reads
        v.ctg_off, v.ctg_read_off = sh.view.ctg_off, sh.view.ctg_read_off
        v.ctg_seq, v.rec_off, v.rec = res["ctg_seq"].data_ptr(), res["rec_off"].data_ptr(), res["rec"].data_ptr()
        if t == 2:
            v.qual_off, v.qual = res["qual_off"].data_ptr(), res["qual"].data_ptr()
        views[t].append(v); keep.append((sh, pin, res))
report("stage 2: + 6 shards synthesised, pinned (1 GB) and resident")
slots = 8
rp = E.ResidentSlots(0, slots)
cap = int(5e6 * 1.25) + 4096
bufs = [torch.zeros(cap + 16, dtype=torch.uint8, device=dev) for _ in range(slots)]
pend, n = [], 0
for i in range(60):
    for t in (1, 2):
        while len(pend) >= slots:
            rp.wait(pend.pop(0))
        pend.append(rp.submit(t, views[t][i % bench.N_ROTATE], cfg, bufs[n % slots].data_ptr(), cap + 16))
        n += 1
while pend:
    rp.wait(pend.pop(0))
report("stage 3: + 8 resident slots created and used")
eng = E.Engine(0)
st = E.Stream(0, 2)
report("stage 4: + one more engine and the packed stream front end")
smp = bench.ClockSampler(0)
smp.start()
report("stage 5: + clock sampler thread (NVML every 5 ms)")
smp.stop_flag = True
smp.join()
report("stage 6: sampler stopped again")
rp.close()
report("stage 7: resident slots closed")
