#!/usr/bin/env python
"""One-off scale check (BASELINE config 3 shape on one GPU): many contigs of log-uniform length, 30x, both tasks,
resident shard; full-size properties instead of an oracle comparison (the oracle runs at ~1 Mbp/s).
usage: bench_scale.py [total_Mb] [n_contigs]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from nextpolish_b200 import engine as E  # noqa: E402


def main():
    total_mb = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    # log-uniform lengths between 20 kb and 1 Mb scaled to the requested total (synth: min_len/max_len)
    t0 = time.time()
    p = E.synth_params(seed=20240917 + 3, n_contigs=n, contig_len=0, min_len=20000, max_len=1000000, depth=30.0, lowercase_frac=0.0003)
    sh = E.Shard.synthetic(p, 0, n, with_qual=2, threads=os.cpu_count() or 8)
    gen_s = time.time() - t0
    bp = int(sh.total_bases)
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750
    eng = E.Engine(0)
    res = {"what": "scale check, resident shard", "contigs": n, "draft_bp": bp, "reads": int(sh.n_reads), "synth_s": gen_s}
    eng.upload(sh.view)
    for task in E.TASKS:
        eng.run(task, cfg); eng.sync()
        best = 1e9
        for _ in range(3):
            t0 = time.time(); eng.run(task, cfg); eng.sync(); best = min(best, time.time() - t0)
        out, off = eng.download(sh.n_contigs)
        lens = np.diff(off)
        src = np.diff(np.array([sh.view.ctg_off[i] for i in range(n + 1)]))
        # size-independent properties: every contig present, length within the indel budget, alphabet
        assert len(lens) == n and (np.abs(lens - src) <= 0.02 * src + 50).all()
        assert set(np.unique(out)) <= set(b"ACGTNacgtn")
        res["task%d" % task] = {"ms": best * 1e3, "Mbp_per_s": bp / best / 1e6, "out_bytes": int(off[-1])}
        if task == 1:
            res["window_stats"] = eng.window_stats()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
