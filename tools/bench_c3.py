#!/usr/bin/env python
"""BASELINE config 3 through the product's own multi-GPU entry (np_multi, csrc/multi_gpu.cu): ONE synthetic 100 Mb draft
in 1000 contigs of log-uniform length (20 kb - 1 Mb) + 30x short reads, FASTA + BGZF BAM (+ .bai) on disk
(page cache) -> polished bytes on the host, STRONG scaling over the GPUs of this box.  Parity: the N-GPU output equals the
1-GPU output byte for byte, and a sample of contigs equals the oracle (the oracle runs at ~1 Mbp/s: a sample it is).
usage: bench_c3.py [total_Mb=100] [n_contigs=1000] [tasks=1,2]   -> one JSON line"""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from nextpolish_b200 import engine as E  # noqa: E402

SIM = os.path.join(ROOT, "nextpolish_b200", "lib", "np_simulate")
SAMTOOLS = os.path.join(ROOT, "oracle", "_ref", "samtools")


def main():
    total_mb = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
    n_ctg = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    tasks = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,2").split(",")]
    tmp = tempfile.mkdtemp(prefix="npc3")
    res = {"what": "config 3: one draft, contig blocks over the GPUs of one box (np_multi), from files", "n_contigs": n_ctg}
    files = {}
    t0 = time.time()
    for t in tasks:
        fa, bam = os.path.join(tmp, "c3.t%d.fa" % t), os.path.join(tmp, "c3.t%d.bam" % t)
        # log-uniform lengths in [20 kb, 1 Mb] have mean ~250 kb; max_len is scaled so that the set sums to ~total_mb
        mean = total_mb * 1e6 / n_ctg
        kw = dict(seed=20240917 + 3 + 100 * t, n_contigs=n_ctg, contig_len=0, min_len=int(mean / 12.5), max_len=int(mean * 4), depth=30.0, read_len=150)
        kw.update(dict(lowercase_frac=0.0) if t == 1 else dict(draft_snv=1e-5, draft_indel=2e-5, lowercase_frac=6.3e-4))
        subprocess.check_call([SIM, fa, bam] + ["%s=%r" % kv for kv in kw.items()])
        subprocess.check_call([SAMTOOLS, "index", bam])
        files[t] = (fa, bam)
    res["generate_s"] = time.time() - t0
    lens_of = {}
    for t in tasks:
        lens_of[t] = {}
        for line in open(files[t][0]):
            if line.startswith(">"):
                cur = line[1:].split()[0]; lens_of[t][cur] = 0
            else:
                lens_of[t][cur] += len(line.strip())
    bp = sum(lens_of[tasks[0]].values())
    res["draft_bp"] = bp
    res["bam_bytes"] = {str(t): os.path.getsize(files[t][1]) for t in tasks}
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750
    ngpu = torch.cuda.device_count()
    md5_by_n = {}
    res["runs"] = []
    block_sizes = [x for x in os.environ.get("C3_BLOCK_MBP", "").split(",") if x]      # tuning: block budgets to try
    for blk in (block_sizes or [None]):
        if blk:
            os.environ["NEXTPOLISH_B200_BLOCK_MBP"] = blk
        for n in [g for g in (1, 2, 4, 8) if g <= ngpu]:
          m = E.MultiGpu(n)
          for t in tasks:
              m.polish(t, files[t][0], files[t][1], cfg)                  # warm-up (memory pools, pinned buffers)
          for t in tasks:
              best, out, st = 1e9, None, None
              for _ in range(2):
                  t1 = time.time()
                  out, st = m.polish(t, files[t][0], files[t][1], cfg)
                  best = min(best, time.time() - t1)
              lens = lens_of[t]
              md5 = hashlib.md5(b"".join(out[k] for k in sorted(out))).hexdigest()
              md5_by_n.setdefault(t, {})[n] = md5
              res["runs"].append({"block_mbp": blk, "n_gpus": n, "task": t, "wall_ms": best * 1e3, "Mbp_per_s": bp / best / 1e6, "blocks_ms": st["blocks_ms"], "gather_download_ms": st["gather_download_ms"],
                                  "h2d_bytes": st["h2d_bytes"], "d2h_bytes": st["d2h_bytes"], "md5": md5})
              if n == 1:
                  keep = out
                  # size-independent properties + a bit-exact sample against the oracle (the five shortest contigs)
                  assert set(out) == set(lens) and all(abs(len(out[k]) - lens[k]) <= 0.02 * lens[k] + 50 for k in lens)
                  O = C.CDLL(os.path.join(ROOT, "oracle", "libnp_oracle.so"))
                  O.np_oracle_run.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
                  names = sorted(lens, key=lambda k: lens[k])[:5]
                  sh = E.Shard.load(files[t][0], files[t][1], names=names, with_qual=True)
                  cap = int(sh.total_bases * 2) + 4096
                  buf = np.zeros(cap, np.uint8); off = np.zeros(sh.n_contigs + 1, np.int64)
                  assert O.np_oracle_run(C.addressof(sh.view), t, C.cast(cfg, C.c_void_p), buf.ctypes.data, cap, off.ctypes.data) == 0
                  raw = buf.tobytes()
                  for i, nm in enumerate(sh.names):
                      assert keep[nm] == raw[off[i]:off[i + 1]], (t, nm)
                  res.setdefault("oracle_sample_ok", {})[str(t)] = len(names)
          m.close()
    res["same_bytes_on_every_gpu_count"] = all(len(set(v.values())) == 1 for v in md5_by_n.values())
    free, tot = torch.cuda.mem_get_info(0)
    res["hbm_used_gb_dev0_after"] = (tot - free) / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
