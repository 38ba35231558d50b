#!/usr/bin/env python
"""Long-read fast mode, one contig set: Mbp/s of (a) the reference's own functions on one host core (oracle/ref2_shim.c:
np2_ref_contig_fast — record loop + first pass + link_consensus_fast; needs oracle/_ref, i.e. the build container or a box
it travelled to), (b) this engine (np2_windows_from_bam on the host, np2_first_pass on the GPU, np2_link_windows_fast),
with the per-launch device times of the last contig when NEXTPOLISH_B200_LGS_TIMING=1.  Prints one JSON line.
    python tools/bench_lgs.py [--fasta F --bam B] [--read-type 1] [--reps 3] [--cpu-only]
Defaults: the committed fixture (tests/golden/lgs_td.fa / .bam: 111 kb, ~10x ONT) — a smoke-sized input, not a benchmark."""
import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
from tests import lgs_cases as L  # noqa: E402
from tests.golden.make_golden_lgs import read_fa, ref_contig_fast  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fasta", default=os.path.join(ROOT, "tests", "golden", "lgs_td.fa"))
    ap.add_argument("--bam", default=os.path.join(ROOT, "tests", "golden", "lgs_td.bam"))
    ap.add_argument("--read-type", type=int, default=1)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-only", action="store_true")
    a = ap.parse_args()
    draft = read_fa(a.fasta)
    total = sum(len(s) for s in draft.values())
    out = {"input": {"fasta": os.path.basename(a.fasta), "bam": os.path.basename(a.bam), "contigs": len(draft), "bp": total, "read_type": a.read_type}}
    S = L.ref_shim()
    md5_ref = {}
    if S is not None:
        best = 1e9
        for _ in range(a.reps):
            t0 = time.time()
            for ctg, seq in draft.items():
                n, linked = ref_contig_fast(S, a.bam, ctg, seq, a.read_type, 5000000, 1000000)
                md5_ref[ctg] = hashlib.md5(linked).hexdigest()
            best = min(best, time.time() - t0)
        out["reference_cpu"] = {"what": "np2_ref_contig_fast: the reference's own functions, 1 host core", "seconds": round(best, 4), "Mbp_per_s": round(total / best / 1e6, 4)}
    if not a.cpu_only:
        os.environ.setdefault("NEXTPOLISH_B200_LGS_TIMING", "1")
        from nextpolish_b200 import nextpolish2 as NP2
        eng = NP2.LgsEngine(0)
        best, md5_gpu = 1e9, {}
        for _ in range(a.reps + 1):                                  # first repetition allocates
            t0 = time.time()
            for ctg in draft:
                md5_gpu[ctg] = hashlib.md5(eng.polish_contig_fast(a.fasta, a.bam, ctg, a.read_type)).hexdigest()
            best = min(best, time.time() - t0)
        agg = {}
        for nm, ms in eng.kernel_times():
            agg[nm] = round(agg.get(nm, 0.0) + ms, 4)
        out["engine"] = {"what": "np2_windows_from_bam (host) + np2_first_pass (GPU) + np2_link_windows_fast (host), wall clock", "seconds": round(best, 4),
                         "Mbp_per_s": round(total / best / 1e6, 4), "identical_to_reference": (md5_gpu == md5_ref) if md5_ref else None,
                         "last_contig_kernels_ms": dict(sorted(agg.items(), key=lambda kv: -kv[1])), "stats": eng.stats()}
        eng.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
