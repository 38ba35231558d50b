"""BASELINE config 1 — `nextPolish test_data/run.cfg` ("reference plumbing, no GPU") — run hermetically: the reference's
unmodified driver, staged outside the repository by tools/run_config1.py with the binaries built from the reference's own
sources (oracle/_ref) and OUR local `paralleltask` stand-in (compat/paralleltask; the reference neither vendors nor pins
that dependency, source/nextPolish:11).  Tasks [5, 1, 2]: long-read polish, score_chain, kmer_count on the bundled 2-contig
draft.  Build container only (needs /root/reference); the output is not reproducible to the byte (bwa / minimap2 thread
batching and the driver's job order change tie-breaks), so the check is the reference's own pass
criterion — it finishes — plus names and lengths."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

NEED = ["seq_split", "seq_count", "bwa", "samtools", "minimap2", "nextpolish1.so", "nextpolish2.so", "calgs.so"]
have = os.path.exists("/root/reference/source/nextPolish") and all(os.path.exists(os.path.join(ROOT, "oracle", "_ref", n)) for n in NEED)


@pytest.fixture(scope="module")
def config1_run():
    import shutil
    import run_config1
    out, log = run_config1.run("reference", keep=True)
    yield out, log
    shutil.rmtree(out["dir"], ignore_errors=True)


@pytest.mark.skipif(not have, reason="needs /root/reference and make -C oracle ref ref2 refcfg")
def test_reference_driver_runs_run_cfg_on_the_paralleltask_standin(config1_run):
    out, log = config1_run
    assert out["rc"] == 0, log[-3000:]
    # default task string = 5, 1, 2 (6 dropped: no hifi_fofn; config_parser.py:86-87,104-108): names carry the sgs steps
    assert sorted(out["contigs"]) == ["tig0000001_np12", "tig0000002_np12"]
    # the reference's bundled sample output (test_data/genome.nextpolish.fa: 51009 / 60401, an older release) within 1 %
    assert abs(out["contigs"]["tig0000001_np12"]["len"] - 51009) < 500
    assert abs(out["contigs"]["tig0000002_np12"]["len"] - 60401) < 600
    assert "nextPolish has finished" in log or "N50" in log


@pytest.mark.skipif(not have, reason="needs /root/reference and make -C oracle ref ref2 refcfg")
def test_native_cli_accepts_the_drivers_job_lines(config1_run, tmp_path):
    """The job lines the driver wrote for its polish steps (source/nextPolish:87-90), handed to OUR native CLI with the
    interpreter + worker script replaced by the binary: every option parses, and the plan (--plan: no device touched) is
    the block-file selection / resume behaviour of the reference's worker (nextpolish1.py:148-179)."""
    import glob
    import shlex
    import subprocess
    from tests.test_part_writer import ref_block_names, ref_scan_output
    out, _ = config1_run
    cli = os.path.join(ROOT, "nextpolish_b200", "lib", "nextpolish1")
    scripts = sorted(glob.glob(os.path.join(out["dir"], "test_data", "01_rundir", "0[12].*", "0*.polish.ref.sh")))
    assert len(scripts) == 2                                     # score_chain and kmer_count steps
    seen = 0
    for sc in scripts:
        for line in open(sc):
            argv = shlex.split(line)
            if not argv:
                continue
            assert argv[1].endswith("lib/nextpolish1.py")
            args = argv[2:]
            opt = dict(zip(args[::2], args[1::2]))
            r = subprocess.run([cli] + args + ["--plan"], cwd=tmp_path, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            plan = [l.split("\t") for l in r.stdout.strip().split("\n")]
            assert plan[0] == ["task", opt["-t"]] and plan[1] == ["finished", "0"] and plan[2] == ["resume_offset", "0"]
            want = ref_block_names(opt["-b"], opt["-i"], set())
            assert [p[1] for p in plan[3:]] == want
            seen += len(want)
            if want:                                             # resume: a part with one finished and one partial record
                part = tmp_path / opt["-o"]
                part.write_text(">%s_np%s 4\nACGT\n>%s_np%s 9\nAC" % (want[0], opt["-t"], want[-1], opt["-t"]))
                r = subprocess.run([cli] + args + ["--plan"], cwd=tmp_path, capture_output=True, text=True)
                done, off = ref_scan_output(str(part))
                lines = r.stdout.strip().split("\n")
                assert lines[1] == "finished\t%d" % len(done) and lines[2] == "resume_offset\t%d" % off
                assert [l.split("\t")[1] for l in lines[3:]] == ref_block_names(opt["-b"], opt["-i"], done)
                part.unlink()
    assert seen == 4                                             # 2 contigs x 2 steps


@pytest.mark.skipif(not have, reason="needs /root/reference and make -C oracle ref ref2 refcfg")
def test_long_read_worker_mirror_accepts_the_drivers_job_lines(config1_run, capsys):
    """The long-read step's job lines (python lib/nextpolish2.py -sp -p N -g G -b blc -i i -l lgs.sort.bam.list -r ont -o part,
    source/nextPolish:70-84) parse in our mirror of that worker; --fast --plan prints what it would do without a device."""
    import glob
    import shlex
    from nextpolish_b200 import nextpolish2 as NP2
    from tests.test_part_writer import ref_block_names
    out, _ = config1_run
    scripts = sorted(glob.glob(os.path.join(out["dir"], "test_data", "01_rundir", "00.lgs_polish", "0*.polish.ref.sh")))
    assert len(scripts) == 1
    seen = []
    for line in open(scripts[0]):
        argv = shlex.split(line)
        if not argv:
            continue
        assert argv[1].endswith("lib/nextpolish2.py")
        args = argv[2:]
        opt = {a: args[i + 1] for i, a in enumerate(args[:-1]) if a in ("-g", "-b", "-i", "-l", "-r", "-o", "-p")}
        capsys.readouterr()
        assert NP2.main(args + ["--fast", "--plan", "-o", "stdout"]) == 0
        printed = [l.split("\t") for l in capsys.readouterr().out.strip().split("\n") if l]
        assert [p[1] for p in printed if p[0] == "bam"] == [l.strip() for l in open(opt["-l"]) if l.strip()]
        assert [p[1] for p in printed if p[0] == "polish"] == ref_block_names(opt["-b"], opt["-i"], set())
        seen += [p[1] for p in printed if p[0] == "polish"]
        assert NP2.main(args) == 1                                  # without --fast: refused (the default mode is not built)
    assert sorted(seen) == ["tig0000001", "tig0000002"]
