#!/usr/bin/env python
"""Resident shards of the bench shape through np_resident with 1..8 slots, and the from-files pipeline at several depths:
ms per 2-task step.  usage: prof_slots.py [steps]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from nextpolish_b200 import engine as E  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda", 0)
cfg = E.default_config(b"")
cfg.contents.read_tlen = 1750
views, keep, cap = {}, [], 0
for t in (1, 2):
    views[t] = []
    for k in range(bench.N_ROTATE):
        sh = E.Shard.synthetic(E.synth_params(**bench.synth_kwargs(t, bench.seed_for(0, t, k))), 0, bench.WORKLOAD["n_contigs"],
                               with_qual=(2 if t == 2 else 0), threads=os.cpu_count() or 8)
        a = sh.arrays()
        res = {k2: torch.from_numpy(v.copy()).to(dev) for k2, v in a.items() if k2 in ("ctg_seq", "rec_off", "rec", "qual_off", "qual")}
        v = E.ShardView()
        v.n_contigs, v.n_reads = sh.view.n_contigs, sh.view.n_reads
        v.ctg_off, v.ctg_read_off = sh.view.ctg_off, sh.view.ctg_read_off
        v.ctg_seq, v.rec_off, v.rec = res["ctg_seq"].data_ptr(), res["rec_off"].data_ptr(), res["rec"].data_ptr()
        if t == 2:
            v.qual_off, v.qual = res["qual_off"].data_ptr(), res["qual"].data_ptr()
        views[t].append(v); keep.append((sh, res))
        cap = max(cap, int(sh.total_bases * 1.25) + 4096)
bp = 2 * bench.WORKLOAD["n_contigs"] * bench.WORKLOAD["contig_len"]
only = os.environ.get("NP_PROF_ONLY", "")
for slots, spin in (() if only == "files" else ((1, "0"), (8, "0"), (8, "1"), (12, "0"), (16, "0"))):
    os.environ["NEXTPOLISH_B200_SPIN"] = spin
    rp = E.ResidentSlots(0, slots)
    bufs = [torch.zeros(cap + 16, dtype=torch.uint8, device=dev) for _ in range(slots)]
    pend, n = [], 0

    def run(k):
        global n
        for i in range(k):
            for t in (1, 2):
                while len(pend) >= slots:
                    rp.wait(pend.pop(0))
                pend.append(rp.submit(t, views[t][i % bench.N_ROTATE], cfg, bufs[n % slots].data_ptr(), cap + 16))
                n += 1
        while pend:
            rp.wait(pend.pop(0))
    run(6)
    torch.cuda.synchronize()
    t0 = time.time()
    run(steps)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print("resident slots %d spin %s: %.3f ms per 2-task step = %.0f Mbp/s" % (slots, spin, dt / steps * 1e3, bp * steps / dt / 1e6), flush=True)
    rp.close()
del keep
tmp = tempfile.mkdtemp(prefix="npfiles")
files = bench.write_inputs(tmp, 0, [1, 2])
for dep, spin in (() if only == "resident" else ((3, "0"), (6, "0"), (3, "1"), (6, "1"), (4, "0"), (8, "0"))):
    os.environ["NEXTPOLISH_B200_SPIN"] = spin
    pipe = E.FilePipeline(0, depth=dep)
    for n in (8, 80):
        t0 = time.time()
        for i in range(n):
            for t in (1, 2):
                pipe.submit(t, files[t][0], files[t][1], cfg)
                while pipe.in_flight() > dep - 1:
                    pipe.wait_oldest(want_md5=False)
        while pipe.in_flight():
            pipe.wait_oldest(want_md5=False)
        dt = time.time() - t0
    print("files depth %d spin %s: %.2f ms per 2-task step = %.0f Mbp/s" % (dep, spin, dt / n * 1e3, bp * n / dt / 1e6), flush=True)
    pipe.close()
print("host cpus", os.cpu_count())
