#!/usr/bin/env python
"""Task 1 on the bench shard (5 x 1 Mb, 30x), resident: per-kernel CUDA-event times.  Used under ncu and with
NEXTPOLISH_B200_PHASE_CYCLES=1.  usage: prof_task1.py [runs] [task]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
from nextpolish_b200 import engine as E  # noqa: E402

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
task = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kw = dict(seed=20240917 + 2 + 100 * task, n_contigs=5, contig_len=1000000, depth=30.0, read_len=150)
if task == 2:
    kw.update(draft_snv=1e-5, draft_indel=2e-5, lowercase_frac=6.3e-4)
sh = E.Shard.synthetic(E.synth_params(**kw), 0, 5, with_qual=(2 if task == 2 else 0), threads=os.cpu_count() or 8)
cfg = E.default_config(b"")
cfg.contents.read_tlen = 1750
eng = E.Engine(0)
eng.upload(sh.view)
eng.set_timing(True)
for _ in range(runs):
    eng.run(task, cfg)
    eng.sync()
kt = eng.kernel_times()
print("launches", eng.launch_count(), "sum_ms %.4f" % sum(v for _, v in kt))
for n, v in sorted(kt, key=lambda kv: -kv[1]):
    print("  %-18s %.4f" % (n, v))
if task == 1:
    print(eng.window_stats())
