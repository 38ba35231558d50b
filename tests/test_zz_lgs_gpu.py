"""GPU: first pass of the long-read consensus window (nextpolish2.so, include/nextpolish2_b200.h) through the C ABI against
the goldens minted from the reference (tests/golden/lgs_golden.json) and against the oracle restatement
(oracle/np2_oracle.c) — bit-exact: positions, bases and link qualities are integer work."""
import ctypes as C
import hashlib
import json
import os
import random
import subprocess

import numpy as np
import pytest

from tests import lgs_cases as L
from tests.conftest import GOLDEN, ROOT
from tests.test_lgs_first_pass import all_golden_cases, td_windows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def O2():
    path = os.path.join(ROOT, "oracle", "libnp2_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    return C.CDLL(path)


@pytest.fixture(scope="module")
def lgs():
    from nextpolish_b200 import nextpolish2 as NP2
    e = NP2.LgsEngine(0)
    yield e
    e.close()


def same(got, want):
    return len(got[1]) == len(want[1]) and (got[0] == want[0]).all() and got[1] == want[1] and (got[2] == want[2]).all()


def test_gpu_first_pass_matches_reference_goldens(lgs):
    gold = json.load(open(os.path.join(GOLDEN, "lgs_golden.json")))["first_pass"]
    seen = 0
    for key, make in all_golden_cases():
        case = make()
        pos, base, _ = lgs.first_pass([case], case["read_type"], case["min_cov"])[0]
        got = {"n": len(base), "pos_md5": hashlib.md5(pos.astype("<u4").tobytes()).hexdigest(), "base_md5": hashlib.md5(base).hexdigest()}
        assert got == gold[key], key
        seen += 1
    assert seen == len(gold)
    assert lgs.stats()["launches"] > 0


def test_gpu_first_pass_matches_oracle_with_qualities_and_batches(lgs, O2):
    for rt in (1, 2, 3, 4):
        cases = [dict(w, read_type=rt) for w in td_windows().values()]
        cases += [L.synthetic_case(**dict(L.CASES[n], read_type=rt)) for n in sorted(L.CASES)]
        want = [L.oracle_window(O2, c) for c in cases]
        got = lgs.first_pass(cases, rt, 4)                               # one batch of 17 windows
        for g, w in zip(got, want):
            assert same(g, w), rt
        st = lgs.stats()
        assert st["segments"] >= len(cases) and st["reruns"] > 0         # the "zones" windows force re-runs
        for c, w in zip(cases[:4], want[:4]):                             # and alone
            assert same(lgs.first_pass([c], rt, 4)[0], w)


def test_gpu_first_pass_fuzz_and_rejections(lgs, O2):
    from nextpolish_b200 import nextpolish2 as NP2
    rng = random.Random(99)
    for it in range(40):
        kw = dict(seed=rng.randrange(1 << 30), length=rng.choice([1, 5, 60, 300, 2000]), depth=rng.choice([1, 2, 5, 15, 40]),
                  read_len=rng.choice([10, 30, 100, 400]), sub=rng.choice([0.0, 0.02, 0.08, 0.4]), ins=rng.choice([0.0, 0.03, 0.1, 0.3]),
                  dele=rng.choice([0.0, 0.03, 0.1, 0.3]), read_type=rng.choice([1, 2, 3, 4]), min_cov=rng.choice([0, 4, 10]),
                  long_ins=rng.choice([0, 0.003]), masked=rng.choice([0, 0, 0.02]), homopolymer=rng.random() < 0.3,
                  odd_chars=rng.random() < 0.2, zones=rng.choice([0, 0, (4, 20), (6, 60)]))
        case = L.synthetic_case(**kw)
        want = L.oracle_window(O2, case)
        if isinstance(want, int):
            with pytest.raises(NP2.NativeError):
                lgs.first_pass([case], case["read_type"], case["min_cov"])
        else:
            assert same(lgs.first_pass([case], case["read_type"], case["min_cov"])[0], want), kw
    case = L.synthetic_case(**L.CASES["ont30"])
    for bad, code in ((dict(case, len=case["len"] + 50), "-2"), (dict(case, len=case["len"] - 50), "-3")):
        with pytest.raises(NP2.NativeError, match=code):
            lgs.first_pass([bad], 1, 4)
    assert same(lgs.first_pass([case], 1, 4)[0], L.oracle_window(O2, case))        # the engine is usable after a rejection
    assert lgs.first_pass([], 1, 4) == []


def test_gpu_first_pass_larger_window_deep(lgs, O2):
    """A 60 kb window at 40x (2.6 M alignment columns): the shape of a real window in miniature."""
    case = L.synthetic_case(seed=21, length=60000, depth=40, read_len=8000, sub=0.03, ins=0.03, dele=0.03, read_type=1)
    want = L.oracle_window(O2, case)
    got = lgs.first_pass([case], 1, 4)[0]
    assert same(got, want)
    assert abs(len(got[1]) - 60000) < 600


def test_gpu_bam_to_first_pass_consensus_matches_the_reference(lgs):
    """Draft FASTA + indexed long-read BAM -> np2_windows_from_bam (host) -> np2_first_pass (GPU): the reference's own
    first-pass consensus of every window (goldens minted through oracle/ref2_shim.c), several window geometries."""
    from nextpolish_b200 import nextpolish2 as NP2
    gold = json.load(open(os.path.join(GOLDEN, "lgs_golden.json")))["from_bam"]
    fa, bam = os.path.join(GOLDEN, "lgs_td.fa"), os.path.join(GOLDEN, "lgs_td.bam")
    for key in sorted(gold):
        ctg, geo, rt = key.split("/")
        w, o = geo[1:].split("_o")
        cw = NP2.ContigWindows(fa, bam, ctg, int(rt[2:]), int(w), int(o))
        res = lgs.first_pass_contig(cw)
        assert len(res) == len(gold[key])
        for (pos, base, _), x in zip(res, gold[key]):
            got = {"n": len(base), "pos_md5": hashlib.md5(pos.astype("<u4").tobytes()).hexdigest(), "base_md5": hashlib.md5(base).hexdigest()}
            assert got == {k: x[k] for k in got}, key
        cw.close()


def test_gpu_fast_mode_end_to_end_matches_the_reference(lgs):
    """LgsEngine.polish_contig_fast: FASTA + BAM -> linked contig consensus, the reference's fast mode byte for byte."""
    gold = json.load(open(os.path.join(GOLDEN, "lgs_golden.json")))["fast_mode"]
    fa, bam = os.path.join(GOLDEN, "lgs_td.fa"), os.path.join(GOLDEN, "lgs_td.bam")
    for key in sorted(gold):
        ctg, geo, rt = key.split("/")
        w, o = (int(x) for x in geo[1:].split("_o"))
        seq = lgs.polish_contig_fast(fa, bam, ctg, int(rt[2:]), w, o)
        assert {"len": len(seq), "md5": hashlib.md5(seq).hexdigest()} == gold[key], key


def test_gpu_worker_mirror_writes_the_fast_mode_contigs(tmp_path):
    """python -m nextpolish_b200.nextpolish2 -g -l -r --fast -o: `>name len` records of both contigs (nextpolish2.py:196-200),
    resume of a cut output, -u."""
    from nextpolish_b200 import nextpolish2 as NP2
    from tests.conftest import read_fasta
    gold = json.load(open(os.path.join(GOLDEN, "lgs_golden.json")))["fast_mode"]
    fa, bam = os.path.join(GOLDEN, "lgs_td.fa"), os.path.join(GOLDEN, "lgs_td.bam")
    lst = tmp_path / "lgs.list"
    lst.write_text(bam + "\n")
    out = str(tmp_path / "lgs.part000.fasta")
    assert NP2.main(["-g", fa, "-l", str(lst), "-r", "ont", "--fast", "-o", out]) == 0
    got = read_fasta(out)
    for ctg in ("tig0000001", "tig0000002"):
        want = gold["%s/w5000000_o1000000/rt1" % ctg]
        assert {"len": len(got[ctg]), "md5": hashlib.md5(got[ctg]).hexdigest()} == want
    headers = [l.split() for l in open(out) if l.startswith(">")]
    assert all(int(h[1]) == len(got[h[0][1:]]) for h in headers)
    full = open(out).read()
    cut = full[:len(full) - 2000]                                    # the last record loses its tail: redone on resume
    open(out, "w").write(cut)
    assert NP2.main(["-g", fa, "-l", str(lst), "-r", "ont", "--fast", "-o", out]) == 0
    assert open(out).read() == full
