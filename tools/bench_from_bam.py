#!/usr/bin/env python
"""T_e2e from files: FASTA + BAM on disk -> polished sequences in host memory, task 1, for
  host   np_shard_load (zlib inflate + packing on the host cores) -> upload -> kernels -> download
  gpu    np_shard_load_gpu (inflate + record unpack + packing on the GPU) -> kernels -> download
usage: bench_from_bam.py [contig_len] [n_contigs]  -> one JSON line"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
from nextpolish_b200 import engine as E  # noqa: E402


def main():
    L = E.lib()
    clen = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    nctg = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    samtools = os.path.join(ROOT, "oracle", "_ref", "samtools")
    with tempfile.TemporaryDirectory() as tmp:
        fa, bam = os.path.join(tmp, "x.fa"), os.path.join(tmp, "x.bam")
        p = E.synth_params(seed=20240920, n_contigs=nctg, contig_len=clen, depth=30.0)
        p.compress_level = 6
        assert L.np_synth_write(p, fa.encode(), bam.encode()) == 0
        subprocess.check_call([samtools, "index", bam])
        cfg = E.default_config(b"")
        eng = E.Engine(0)
        bp = nctg * clen
        res = {}

        def host_path():
            sh = E.Shard.load(fa, bam, with_qual=False, threads=os.cpu_count() or 8)
            t1 = time.time()
            got = eng.polish(sh, 1, cfg)
            sh.close()
            return got, t1

        def gpu_path():
            ds = E.DeviceShard(fa, bam, with_qual=False)
            t1 = time.time()
            eng.adopt_device(ds.view)
            eng.run(1, cfg)
            out, off = eng.download(ds.n_contigs)
            raw = out.tobytes()
            got = {nm: raw[off[i]:off[i + 1]] for i, nm in enumerate(ds.names)}
            res["gpu_load_stats"] = ds.stats()
            ds.close()                       # return the shard's buffers to the device pool before the next load
            return got, t1

        outs = {}
        for name, fn in (("host", host_path), ("gpu", gpu_path)):
            best, best_load = 1e9, 1e9
            for _ in range(4):
                t0 = time.time()
                got, t1 = fn()
                dt = time.time() - t0
                if dt < best:
                    best, best_load = dt, t1 - t0
            outs[name] = got
            res[name] = {"wall_s": best, "load_s": best_load, "Mbp_per_s": bp / best / 1e6}
        assert outs["host"] == outs["gpu"]
        res.update({"what": "FASTA+BAM files -> polished sequences (task 1), best of 4", "draft_bp": bp, "bam_bytes": os.path.getsize(bam),
                    "host_threads": os.cpu_count(), "identical": True, "speedup": res["host"]["wall_s"] / res["gpu"]["wall_s"]})
        print(json.dumps(res))


if __name__ == "__main__":
    main()
