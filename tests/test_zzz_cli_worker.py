"""GPU: the native CLI's worker grammar (the command line of the reference's lib/nextpolish1.py: -g/-s/-t/-b/-i/-o/-u/-debug,
block file, resume, ">name_np<task> <len>" records — nextpolish1.py:148-179,226-229) against the goldens minted from the
reference binary, and np_multi_run_names against the unrestricted run."""
import os
import subprocess

import pytest

from tests.conftest import GOLDEN, REF_SAMTOOLS, read_fasta

pytestmark = pytest.mark.gpu


def cli(E):
    return os.path.join(os.path.dirname(E.binding.LIB_PATH), "nextpolish1")


@pytest.mark.parametrize("step", [1, 2])
def test_worker_grammar_block_resume_headers(E, tmp_path, step):
    fa = os.path.join(GOLDEN, "td30.step%d.fa" % step)
    bam = os.path.join(GOLDEN, "td30.step%d.bam" % step)
    exp = read_fasta(os.path.join(GOLDEN, "td30.step%d.expected.fa" % step))
    names = [n[:-2] for n in exp]                           # expected headers are name_<step>
    tag = "_np%d" % step
    blc = str(tmp_path / "g.blc")
    open(blc, "w").write("".join("%s\t%d\n" % (n, i % 2) for i, n in enumerate(names)))
    out = str(tmp_path / "part000.fasta")
    base = [cli(E), "-g", fa, "-s", bam, "-t", str(step), "-p", "3"]
    r = subprocess.run(base + ["-b", blc, "-i", "0", "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = read_fasta(out)
    want0 = {n + tag: exp["%s_%d" % (n, step)] for i, n in enumerate(names) if i % 2 == 0}
    assert got == want0
    for line in open(out):
        if line.startswith(">"):
            name, length = line[1:].split()
            assert int(length) == len(got[name])
    # resume: finished contigs are skipped, a partial last record is re-done; index "all" through a one-block file
    blc_all = str(tmp_path / "all.blc")
    open(blc_all, "w").write("".join("%s\t0\n" % n for n in names))
    with open(out, "a") as f:
        f.write(">" + names[1] + tag + " 10\nACGT")
    r = subprocess.run(base + ["-b", blc_all, "-i", "0", "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert read_fasta(out) == {n + tag: exp["%s_%d" % (n, step)] for n in names}
    # stdout, no block file, -u
    r = subprocess.run(base + ["-u"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    recs = {}
    for line in r.stdout.split("\n"):
        if line.startswith(">"):
            cur = line[1:].split()[0]
        elif line:
            recs[cur] = line.encode()
    assert recs == {n + tag: exp["%s_%d" % (n, step)].upper() for n in names}


def test_worker_grammar_debug_trace_equals_mirror(E, tmp_path, capfd):
    """-debug: the PolishPoint lines on stderr (nextpolish1.py:230-231) equal the Python mirror's, sequences unchanged."""
    from nextpolish_b200 import nextpolish1
    fa = os.path.join(GOLDEN, "td30.step1.fa")
    bam = os.path.join(GOLDEN, "td30.step1.bam")
    exp = read_fasta(os.path.join(GOLDEN, "td30.step1.expected.fa"))
    out1, out2 = str(tmp_path / "cli.fa"), str(tmp_path / "mirror.fa")
    r = subprocess.run([cli(E), "-g", fa, "-s", bam, "-t", "1", "-o", out1, "-debug"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    capfd.readouterr()
    assert nextpolish1.main(["-g", fa, "-s", bam, "-t", "1", "-o", out2, "-debug"]) == 0
    mirror_err = capfd.readouterr().err
    pts = lambda text: sorted(l for l in text.split("\n") if len(l.split()) == 5 and l.split()[1].lstrip("-").isdigit())
    assert len(pts(r.stderr)) > 10 and pts(r.stderr) == pts(mirror_err)
    assert read_fasta(out1) == read_fasta(out2) == {n[:-2] + "_np1": s for n, s in exp.items()}


def test_multi_run_names_subset(E, synth_files):
    """np_multi_run_names on a subset of the contigs = those contigs of the unrestricted run, in FASTA order."""
    fa, bam = synth_files("ragged")
    if not os.path.exists(bam + ".bai"):
        subprocess.check_call([REF_SAMTOOLS, "index", bam])
    cfg = E.default_config(fa, bam)
    m = E.MultiGpu(1)
    try:
        for task in (1, 2):
            full, _ = m.polish(task, fa, bam, cfg)
            order = list(full)
            assert len(order) >= 5
            pick = [order[4], order[0], order[3]]           # given out of order: the result is in FASTA order
            got, _ = m.polish(task, fa, bam, cfg, names=pick)
            assert list(got) == [order[0], order[3], order[4]]
            assert got == {n: full[n] for n in got}
            assert m.polish(task, fa, bam, cfg, names=[])[0] == {}
        with pytest.raises(E.NativeError, match="not in the draft"):
            m.polish(1, fa, bam, cfg, names=["no_such_contig"])
    finally:
        m.close()
