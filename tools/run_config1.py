#!/usr/bin/env python
"""BASELINE config 1: `nextPolish test_data/run.cfg` run hermetically.

The reference's driver (source/nextPolish) imports `paralleltask`, which the reference neither vendors nor pins, and
expects its binaries under <dir of nextPolish>/bin and /lib.  This tool stages a throw-away copy of the driver in a
temporary directory (outside the repository: reference sources are never copied into it), links the binaries that
`make -C oracle ref ref2 refcfg` built from the reference's own sources (oracle/_ref/), puts compat/ (our local
`paralleltask.Task` stand-in, job_type = local) on PYTHONPATH and runs the bundled test: tasks [5, 1, 2] on the 2-contig
draft with the bundled long and short reads.

    python tools/run_config1.py [--engine reference|b200] [--task default|12|...] [--keep]

--engine reference (default): every engine is the reference's own (CPU; what config 1 names: "reference plumbing, no GPU").
--engine b200: lib/nextpolish1.so is THIS repository's library (the short-read steps then run on the GPU; needs a B200 and
  works because the structs and the eight symbols are the reference's).
Needs /root/reference (build container).  Prints a JSON line with the final FASTA's contigs (length, md5)."""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
REF_SRC = "/root/reference/source"
REF_BIN = os.path.join(ROOT, "oracle", "_ref")


def stage(engine):
    if not os.path.exists(os.path.join(REF_SRC, "nextPolish")):
        raise SystemExit("the reference's driver sources (%s) are not on this machine" % REF_SRC)
    need = ["seq_split", "seq_count", "bwa", "samtools", "minimap2", "nextpolish1.so", "nextpolish2.so", "calgs.so"]
    missing = [n for n in need if not os.path.exists(os.path.join(REF_BIN, n))]
    if missing:
        raise SystemExit("missing %s under oracle/_ref: run make -C oracle ref ref2 refcfg" % missing)
    top = tempfile.mkdtemp(prefix="np_config1_")
    st = os.path.join(top, "NextPolish")
    os.makedirs(os.path.join(st, "bin"))
    os.makedirs(os.path.join(st, "lib"))
    shutil.copy(os.path.join(REF_SRC, "nextPolish"), st)
    for f in os.listdir(os.path.join(REF_SRC, "lib")):
        if f.endswith(".py"):
            shutil.copy(os.path.join(REF_SRC, "lib", f), os.path.join(st, "lib"))
    for b in ("seq_split", "seq_count", "bwa", "samtools", "minimap2"):
        os.symlink(os.path.join(REF_BIN, b), os.path.join(st, "bin", b))
    for so in ("nextpolish2.so", "calgs.so"):
        os.symlink(os.path.join(REF_BIN, so), os.path.join(st, "lib", so))
    np1 = os.path.join(REF_BIN, "nextpolish1.so") if engine == "reference" else os.path.join(ROOT, "nextpolish_b200", "lib", "nextpolish1.so")
    os.symlink(np1, os.path.join(st, "lib", "nextpolish1.so"))
    work = os.path.join(top, "test_data")
    shutil.copytree(os.path.join(REF_SRC, "test_data"), work)
    os.chmod(work, 0o755)
    for f in os.listdir(work):
        os.chmod(os.path.join(work, f), 0o644)
    return top, st, work


def read_fa(path):
    d, name = {}, None
    for line in open(path):
        if line.startswith(">"):
            name = line[1:].split()[0]
            d[name] = []
        elif name:
            d[name].append(line.strip())
    return {k: "".join(v) for k, v in d.items()}


def run(engine="reference", task=None, keep=False, timeout=3600):
    top, st, work = stage(engine)
    cfg = os.path.join(work, "run.cfg")
    if task:
        txt = open(cfg).read().replace("task = default", "task = %s" % task)
        open(cfg, "w").write(txt)
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "compat") + os.pathsep + os.environ.get("PYTHONPATH", ""))
    t0 = time.time()
    r = subprocess.run([sys.executable, os.path.join(st, "nextPolish"), "run.cfg"], cwd=work, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=timeout)
    out = {"engine": engine, "task": task or "default", "rc": r.returncode, "seconds": round(time.time() - t0, 1)}
    final = os.path.join(work, "01_rundir", "genome.nextpolish.fasta")
    if r.returncode == 0 and os.path.exists(final):
        out["contigs"] = {n: {"len": len(s), "md5": hashlib.md5(s.encode()).hexdigest()} for n, s in sorted(read_fa(final).items())}
    else:
        out["log_tail"] = r.stdout[-3000:]
    if keep:
        out["dir"] = top
    else:
        shutil.rmtree(top, ignore_errors=True)
    return out, r.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", choices=["reference", "b200"], default="reference")
    ap.add_argument("--task", default=None)
    ap.add_argument("--keep", action="store_true")
    a = ap.parse_args()
    out, log = run(a.engine, a.task, a.keep)
    print(json.dumps(out))
    return 0 if out["rc"] == 0 and "contigs" in out else 1


if __name__ == "__main__":
    sys.exit(main())
