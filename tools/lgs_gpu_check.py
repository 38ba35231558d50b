#!/usr/bin/env python
"""Quick GPU check without torch/pytest (seconds): nextpolish2.so first pass against the oracle on seeded windows (both chain
kernels), per-kernel device times of two larger windows, and the native CLI's worker grammar on the golden td30 fixture.
Prints one line per check; exit code 0 when all agree."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from tests import lgs_cases as L  # noqa: E402
from nextpolish_b200 import nextpolish2 as NP2  # noqa: E402


def same(g, w):
    return len(g[1]) == len(w[1]) and (g[0] == w[0]).all() and g[1] == w[1] and (g[2] == w[2]).all()


def main():
    ok = True
    O2 = C.CDLL(os.path.join(ROOT, "oracle", "libnp2_oracle.so"))
    t0 = time.time()
    eng = NP2.LgsEngine(0)
    print("engine up in %.2f s" % (time.time() - t0), flush=True)
    T0 = time.time()
    DEADLINE = float(os.environ.get("LGS_CHECK_DEADLINE_S", "1e9"))
    base_cases = [L.synthetic_case(**L.CASES[n]) for n in sorted(L.CASES)]

    def small_windows(chain):
        nonlocal ok
        os.environ["NEXTPOLISH_B200_LGS_CHAIN"] = chain
        for rt in (1, 2, 3, 4):
            if time.time() - T0 > DEADLINE:
                print("deadline: skipped chain=%s rt%d" % (chain, rt), flush=True)
                continue
            cases = [dict(c, read_type=rt) for c in base_cases]
            want = [L.oracle_window(O2, c) for c in cases]
            t0 = time.time()
            got = eng.first_pass(cases, rt, 4)
            dt = time.time() - t0
            good = all(same(g, w) for g, w in zip(got, want))
            ok &= good
            print("lgs first pass chain=%s rt%d: %d windows, %s, %.1f ms, %s" % (chain, rt, len(cases), "IDENTICAL" if good else "DIFFERENT", dt * 1e3, eng.stats()), flush=True)

    small_windows("warp")
    big = {"ont_60kb_x40": dict(seed=21, length=60000, depth=40, read_len=8000, sub=0.03, ins=0.03, dele=0.03, read_type=1),
           "hifi_150kb_x20": dict(seed=22, length=150000, depth=20, read_len=12000, sub=0.002, ins=0.002, dele=0.002, read_type=3)}
    os.environ["NEXTPOLISH_B200_LGS_TIMING"] = "1"
    for name, kw in big.items():
        case = L.synthetic_case(**kw)
        t0 = time.time()
        want = L.oracle_window(O2, case)
        cpu_ms = (time.time() - t0) * 1e3
        for chain in ("warp", "thread"):
            os.environ["NEXTPOLISH_B200_LGS_CHAIN"] = chain
            for rep in range(3):
                t0 = time.time()
                got = eng.first_pass([case], kw["read_type"], 4)[0]
                dt = time.time() - t0
            good = same(got, want)
            ok &= good
            agg = {}
            for nm, ms in eng.kernel_times():
                agg[nm] = round(agg.get(nm, 0.0) + ms, 4)
            print(json.dumps({"case": name, "chain": chain, "identical": bool(good), "call_ms_host_buffers": round(dt * 1e3, 2), "device_ms": round(sum(agg.values()), 3),
                              "oracle_cpu_ms_1core": round(cpu_ms, 1), "alignment_columns": int(case["aln_len"].sum()), "window": kw["length"],
                              "stats": eng.stats(), "kernels_ms": dict(sorted(agg.items(), key=lambda kv: -kv[1]))}), flush=True)
    os.environ["NEXTPOLISH_B200_LGS_TIMING"] = "0"
    # the GPU tests of the CLI's worker grammar and of np_multi_run_names, called without pytest (tests/test_zzz_cli_worker.py)
    import pathlib
    import tempfile
    from nextpolish_b200 import engine as E
    from tests import test_zzz_cli_worker as TW
    from tests.synth_cases import CASES as SC
    tmp = pathlib.Path(tempfile.mkdtemp(prefix="lgs_check"))

    def synth_files(name):
        fa, bam = str(tmp / (name + ".fa")), str(tmp / (name + ".bam"))
        if not os.path.exists(fa):
            assert E.lib().np_synth_write(E.synth_params(**SC[name]), fa.encode(), bam.encode()) == 0
        return fa, bam
    for label, fn in (("worker grammar step 1 (block, resume, -u)", lambda: TW.test_worker_grammar_block_resume_headers(E, tmp, 1)),
                      ("np_multi_run_names subset", lambda: TW.test_multi_run_names_subset(E, synth_files)),
                      ("worker grammar step 2 (block, resume, -u)", lambda: TW.test_worker_grammar_block_resume_headers(E, tmp, 2))):
        if time.time() - T0 > DEADLINE:
            print("deadline: skipped", label, flush=True)
            continue
        t0 = time.time()
        try:
            fn()
            print("%s: PASSED (%.1f s)" % (label, time.time() - t0), flush=True)
        except Exception as ex:                                  # noqa: BLE001
            ok = False
            print("%s: FAILED %r" % (label, ex), flush=True)
    small_windows("thread")
    eng.close()
    return 0 if ok else 1


def _unused():
    cli = os.path.join(ROOT, "nextpolish_b200", "lib", "nextpolish1")
    G = os.path.join(ROOT, "tests", "golden")
    for step in (1, 2):
        out = "/tmp/lgs_check_part%d.fa" % step
        if os.path.exists(out):
            os.remove(out)
        r = subprocess.run([cli, "-g", os.path.join(G, "td30.step%d.fa" % step), "-s", os.path.join(G, "td30.step%d.bam" % step), "-t", str(step), "-o", out],
                           capture_output=True, text=True)
        exp = open(os.path.join(G, "td30.step%d.expected.fa" % step)).read().split("\n")
        want_seqs = [l for l in exp if l and not l.startswith(">")]
        got_seqs = [l for l in open(out).read().split("\n") if l and not l.startswith(">")] if r.returncode == 0 else []
        good = r.returncode == 0 and got_seqs == want_seqs
        ok &= good
        print("cli worker grammar step %d: rc %d, %s %s" % (step, r.returncode, "IDENTICAL" if good else "DIFFERENT", r.stderr.strip()[-200:] if not good else ""), flush=True)


if __name__ == "__main__":
    sys.exit(main())
