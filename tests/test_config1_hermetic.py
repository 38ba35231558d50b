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


@pytest.mark.skipif(not have, reason="needs /root/reference and make -C oracle ref ref2 refcfg")
def test_reference_driver_runs_run_cfg_on_the_paralleltask_standin():
    import run_config1
    out, log = run_config1.run("reference")
    assert out["rc"] == 0, log[-3000:]
    # default task string = 5, 1, 2 (6 dropped: no hifi_fofn; config_parser.py:86-87,104-108): names carry the sgs steps
    assert sorted(out["contigs"]) == ["tig0000001_np12", "tig0000002_np12"]
    # the reference's bundled sample output (test_data/genome.nextpolish.fa: 51009 / 60401, an older release) within 1 %
    assert abs(out["contigs"]["tig0000001_np12"]["len"] - 51009) < 500
    assert abs(out["contigs"]["tig0000002_np12"]["len"] - 60401) < 600
    assert "nextPolish has finished" in log or "N50" in log
