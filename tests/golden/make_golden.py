#!/usr/bin/env python
"""Mint the committed golden fixtures from the reference itself (run in the build container,
where /root/reference exists and `make -C oracle ref` has produced oracle/_ref/).

Fixture A (td30): contig 1 region of the reference's own source/test_data, short reads mapped with
  the vendored bwa mem -> samtools fixmate/sort/markdup exactly as source/nextPolish:199-226,119-156
  does, then subsampled to ~30x so the BAM stays small.  Expected outputs: the reference binary's
  `nextpolish1 scorechain` FASTA, and `nextpolish1 kmercount` on that FASTA re-mapped.
Fixture B (synth md5s): per-contig md5 of the reference's output on seeded synthetic sets produced by
  our generator (the sets are regenerated at test time; only the md5s are committed).
"""
import hashlib, json, os, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.realpath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
TD = "/root/reference/source/test_data"
OUT = os.path.dirname(os.path.realpath(__file__))
sys.path.insert(0, ROOT)


def sh(cmd):
    subprocess.check_call(cmd, shell=True, executable="/bin/bash")


def map_reads(genome, prefix, frac):
    sh(f"{REF}/bwa index {genome} 2>/dev/null")
    sh(f"{REF}/bwa mem -p -t 1 {genome} <(paste -d'\\n' <(zcat {TD}/sreads.R1.fastq.gz | paste - - - -) "
       f"<(zcat {TD}/sreads.R2.fastq.gz | paste - - - -) | tr '\\t' '\\n') 2>/dev/null | "
       f"{REF}/samtools view -F 0x4 -b - | {REF}/samtools fixmate -m - - | {REF}/samtools sort -o {prefix}.s0.bam - 2>/dev/null")
    sh(f"{REF}/samtools markdup -r -s {prefix}.s0.bam {prefix}.s1.bam 2>/dev/null")
    sh(f"{REF}/samtools view -b -s 11.{frac} -o {prefix}.bam {prefix}.s1.bam && {REF}/samtools index {prefix}.bam")
    os.remove(f"{prefix}.s0.bam"); os.remove(f"{prefix}.s1.bam")


def read_fa(path):
    d, name = {}, None
    for line in open(path):
        if line.startswith(">"):
            name = line[1:].split()[0]
        elif name:
            d[name] = d.get(name, "") + line.strip()
    return d


def mint_snpvalid():
    """Adds the snp_valid goldens (task 4, `nextpolish1 snpvalid`) without touching the committed BAMs: md5s of the six
    synthetic cases under key "4" of synth_md5.json, and the expected FASTA of the td30 step-1 fixture.  The reference's
    second pass reads past a list on some inputs (snpvalid.c:37-66 emits an odd number of cut points when a window starts
    on a flagged column); the cases pinned here are ones where that does not happen (the oracle restatement reports it)."""
    from nextpolish_b200 import engine as E
    from tests.synth_cases import CASES
    tmp = tempfile.mkdtemp(prefix="npgold4")
    path = os.path.join(OUT, "synth_md5.json")
    md5s = json.load(open(path))
    for name, kw in CASES.items():
        fa, bam = os.path.join(tmp, name + ".fa"), os.path.join(tmp, name + ".bam")
        assert E.lib().np_synth_write(E.synth_params(**kw), fa.encode(), bam.encode()) == 0
        sh(f"{REF}/samtools index {bam}")
        o = os.path.join(tmp, f"{name}.4.fa")
        sh(f"{REF}/nextpolish1 snpvalid {fa} {bam} > {o} 2>/dev/null")
        md5s[name]["4"] = {k: hashlib.md5(v.encode()).hexdigest() for k, v in sorted(read_fa(o).items())}
    json.dump(md5s, open(path, "w"), indent=1, sort_keys=True)
    sh(f"{REF}/nextpolish1 snpvalid {OUT}/td30.step1.fa {OUT}/td30.step1.bam > {OUT}/td30.step1.snpvalid.expected.fa 2>/dev/null")
    for f in os.listdir(OUT):
        if f.endswith(".fai") and "expected" not in f and not os.path.exists(os.path.join(OUT, f[:-4])):
            pass
    print("snp_valid goldens written to", OUT)


def main():
    if "--snpvalid-only" in sys.argv:
        return mint_snpvalid()
    tmp = tempfile.mkdtemp(prefix="npgold")
    # ---- fixture A
    g1 = os.path.join(OUT, "td30.step1.fa")
    sh(f"cp {TD}/raw.genome.fasta {g1} && chmod u+w {g1}")
    map_reads(g1, os.path.join(OUT, "td30.step1"), "20")
    sh(f"{REF}/nextpolish1 scorechain {g1} {OUT}/td30.step1.bam > {OUT}/td30.step1.expected.fa 2>/dev/null")
    g2 = os.path.join(OUT, "td30.step2.fa")
    sh(f"cp {OUT}/td30.step1.expected.fa {g2}")
    map_reads(g2, os.path.join(OUT, "td30.step2"), "20")
    sh(f"{REF}/nextpolish1 kmercount {g2} {OUT}/td30.step2.bam > {OUT}/td30.step2.expected.fa 2>/dev/null")
    for f in os.listdir(OUT):
        if f.endswith((".amb", ".ann", ".bwt", ".pac", ".sa", ".fai")):
            os.remove(os.path.join(OUT, f))
    # ---- fixture B
    from nextpolish_b200 import engine as E
    from tests.synth_cases import CASES
    md5s = {}
    for name, kw in CASES.items():
        fa, bam = os.path.join(tmp, name + ".fa"), os.path.join(tmp, name + ".bam")
        p = E.synth_params(**kw)
        assert E.lib().np_synth_write(p, fa.encode(), bam.encode()) == 0
        sh(f"{REF}/samtools index {bam}")
        md5s[name] = {}
        for step, cmd in ((1, "scorechain"), (2, "kmercount")):
            o = os.path.join(tmp, f"{name}.{step}.fa")
            sh(f"{REF}/nextpolish1 {cmd} {fa} {bam} > {o} 2>/dev/null")
            md5s[name][str(step)] = {k: hashlib.md5(v.encode()).hexdigest() for k, v in sorted(read_fa(o).items())}
    json.dump(md5s, open(os.path.join(OUT, "synth_md5.json"), "w"), indent=1, sort_keys=True)
    print("golden fixtures written to", OUT)
    mint_snpvalid()


if __name__ == "__main__":
    main()
