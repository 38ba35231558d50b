"""The stage in front of the long-read first pass: BAM records of a contig -> the alignment strings of its
consensus windows, np2_windows_from_bam (csrc/lgs_hostio.cpp, host code in nextpolish2.so) against the reference's own
functions run in the record loop of ctg_cns_core (oracle/ref2_shim.c: np2_ref_contig_windows) — goldens in
tests/golden/lgs_golden.json["from_bam"] (window ranges, alignment counts, a hash over every alignment string, and the
reference's first-pass consensus from that BAM), minted by make_golden_lgs.py on a 30 % subsample of the reference's
test_data long reads mapped with the vendored minimap2 (tests/golden/lgs_td.bam) and a draft with N / lower-case / IUPAC
bases written into it (read_ref's 2-bit packing is part of the pin).  The loader's batch then goes through the kernel
bodies of the GPU path in the host test build: BAM -> first-pass consensus equals the reference's."""
import ctypes as C
import hashlib
import json
import os
import shutil

import numpy as np
import pytest

from tests import lgs_cases as L
from tests.conftest import GOLDEN
from tests.test_lgs_first_pass import _emu2_build

FA = os.path.join(GOLDEN, "lgs_td.fa")
BAM = os.path.join(GOLDEN, "lgs_td.bam")


@pytest.fixture(scope="module")
def NP2(E):                                                   # E: makes sure the libraries are built
    from nextpolish_b200 import nextpolish2
    nextpolish2.lib2()
    return nextpolish2


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(GOLDEN, "lgs_golden.json")))["from_bam"]


def keys(gold):
    for key in sorted(gold):
        ctg, geo, rt = key.split("/")
        w, o = geo[1:].split("_o")
        yield key, ctg, int(w), int(o), int(rt[2:])


def test_windows_and_alignment_strings_match_the_reference(NP2, gold):
    seen = 0
    for key, ctg, w, o, rt in keys(gold):
        cw = NP2.ContigWindows(FA, BAM, ctg, rt, w, o)
        got = [(s, e, n, "%016x" % h) for s, e, n, h in cw.info()]
        want = [(x["start"], x["end"], x["n_alns"], x["hash"]) for x in gold[key]]
        assert got == want, key
        seen += len(want)
        cw.close()
    assert seen > 30                                           # 2 contigs x (1 + 3-4 + 6-7 windows) x 2 read types


def test_bam_to_first_pass_consensus_matches_the_reference(NP2, gold):
    E2 = _emu2_build("libnp2_emu.so")
    for key, ctg, w, o, rt in keys(gold):
        cw = NP2.ContigWindows(FA, BAM, ctg, rt, w, o)
        b = cw.batch
        cap = int(b.str_bytes) + 16
        pos, base, qv = np.zeros(cap, np.uint32), np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
        off = np.zeros(b.n_windows + 1, np.int64)
        n = E2.np2_emu_first_pass_batch(C.byref(b), pos.ctypes.data, base.ctypes.data, qv.ctypes.data, cap, off.ctypes.data, 9, None)
        assert n > 0, (key, n)
        for i, x in enumerate(gold[key]):
            p, s = pos[off[i]:off[i + 1]], base[off[i]:off[i + 1]].tobytes()
            got = {"n": len(s), "pos_md5": hashlib.md5(p.astype("<u4").tobytes()).hexdigest(), "base_md5": hashlib.md5(s).hexdigest()}
            assert got == {k: x[k] for k in got}, (key, i)
        cw.close()


def test_windows_as_dicts_round_trip_through_the_oracle(NP2):
    """ContigWindows.windows(): the same windows as stand-alone inputs (what LgsEngine.first_pass takes)."""
    import subprocess
    from tests.conftest import ROOT
    path = os.path.join(ROOT, "oracle", "libnp2_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    O2 = C.CDLL(path)
    cw = NP2.ContigWindows(FA, BAM, "tig0000002", 1, 20000, 4000)
    wins = cw.windows()
    assert [w["len"] for w in wins] == [e - s for s, e, _, _ in cw.info()]
    E2 = _emu2_build("libnp2_emu.so")
    whole = L.first_pass_batch(lambda b, p, ba, q, cap, off: E2.np2_emu_first_pass_batch(b, p, ba, q, cap, off, 0, None), wins)
    for w, got in zip(wins, whole):
        want = L.oracle_window(O2, w)
        assert (got[0] == want[0]).all() and got[1] == want[1] and (got[2] == want[2]).all()
    cw.close()


@pytest.mark.skipif(L.ref_shim() is None, reason="oracle/_ref/libnp2_refshim.so not built (needs /root/reference)")
def test_live_reference_other_read_types_and_geometry(NP2):
    from tests.golden.make_golden_lgs import read_fa, ref_contig_windows
    S = L.ref_shim()
    draft = read_fa(FA)
    for ctg in draft:
        for rt, w, o in ((2, 30000, 5000), (4, 12345, 678), (1, 1200, 100), (3, 501, 0)):      # tiny windows: clip_aln's short branch, the unsigned length test
            want = [(a, b, n, h) for a, b, n, h, _, _ in ref_contig_windows(S, BAM, ctg, draft[ctg], rt, w, o)]
            cw = NP2.ContigWindows(FA, BAM, ctg, rt, w, o)
            assert cw.info() == want, (ctg, rt, w, o)
            cw.close()


def test_loader_edges(NP2, tmp_path):
    with pytest.raises(NP2.NativeError, match="not in the FASTA|contig"):
        NP2.ContigWindows(FA, BAM, "no_such_contig", 1)
    # a contig the BAM does not know: one window, only the window itself
    fa2 = str(tmp_path / "two.fa")
    open(fa2, "w").write(open(FA).read() + ">lonely\n" + "ACGTTGCA" * 50 + "\n")
    cw = NP2.ContigWindows(fa2, BAM, "lonely", 1)
    assert [(s, e, n) for s, e, n, _ in cw.info()] == [(0, 400, 1)]
    cw.close()
    # without the index the loader says so
    bam2 = str(tmp_path / "noindex.bam")
    shutil.copy(BAM, bam2)
    with pytest.raises(NP2.NativeError, match="bai"):
        NP2.ContigWindows(FA, bam2, "tig0000001", 1)
    with pytest.raises(NP2.NativeError, match="bad arguments"):
        NP2.ContigWindows(FA, BAM, "tig0000001", 1, window=1000, overlap=1000)


def test_fast_mode_end_to_end_matches_the_reference(NP2):
    """FASTA + BAM -> windows (host) -> first pass (kernel bodies, host test build) -> np2_link_windows_fast (host) against the
    reference's fast mode (first pass of every window + link_consensus_fast, ctg_cns.c:3053-3119): the linked contig,
    byte for byte, for one window and for 3-8 linked windows; the result does not depend on the window geometry."""
    gold = json.load(open(os.path.join(GOLDEN, "lgs_golden.json")))["fast_mode"]
    E2 = _emu2_build("libnp2_emu.so")
    per_contig = {}
    for key in sorted(gold):
        ctg, geo, rt = key.split("/")
        w, o = (int(x) for x in geo[1:].split("_o"))
        cw = NP2.ContigWindows(FA, BAM, ctg, int(rt[2:]), w, o)
        b = cw.batch
        cap = int(b.str_bytes) + 16
        pos, base, qv = np.zeros(cap, np.uint32), np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
        off = np.zeros(b.n_windows + 1, np.int64)
        n = E2.np2_emu_first_pass_batch(C.byref(b), pos.ctypes.data, base.ctypes.data, qv.ctypes.data, cap, off.ctypes.data, 0, None)
        assert n > 0
        linked = NP2.link_windows_fast([s for s, _, _, _ in cw.info()], NP2.split_result(n, pos, base, qv, off, b.n_windows), o)
        assert {"len": len(linked), "md5": hashlib.md5(linked).hexdigest()} == gold[key], key
        per_contig.setdefault((ctg, rt), set()).add(linked)
        cw.close()
    assert all(len(v) == 1 for v in per_contig.values())
    # windows that cannot be linked are reported, not walked off (the reference has no bounds there)
    pos = np.arange(100, dtype=np.uint32)
    res = [(pos, b"A" * 100, None), (pos, b"C" * 100, None)]
    with pytest.raises(NP2.NativeError, match="-7"):
        NP2.link_windows_fast([0, 60], res, 40)


def test_worker_mirror_command_line_without_gpu_work(NP2, tmp_path, capsys):
    """python -m nextpolish_b200.nextpolish2: refuses the production mode, one BAM per list, resume / block rules of
    nextpolish2.py:98-137 (a finished part leaves nothing to polish: no device is touched)."""
    lst = tmp_path / "lgs.list"
    lst.write_text(BAM + "\n")
    assert NP2.main(["-g", FA, "-l", str(lst), "-r", "ont"]) == 1
    assert "fast mode" in capsys.readouterr().err
    empty = tmp_path / "empty.list"
    empty.write_text("\n")
    assert NP2.main(["-g", FA, "-l", str(empty), "-r", "ont", "--fast"]) == 1
    # resume scan: tig0000001 finished, tig0000002 split into two pieces of which the second is partial -> redone
    out = tmp_path / "part.fasta"
    out.write_text(">tig0000001 4\nACGT\n>tig0000002_s0 8\nACGTACGT\n>tig0000002_s1 9\nAC")
    done = set()
    at = NP2._read_corrected(str(out), done)
    assert done == {"tig0000001"} and at == len(">tig0000001 4\nACGT\n")
    assert NP2._read_uncorrected(FA, "all", done) == ["tig0000002"]
    blc = tmp_path / "g.blc"
    blc.write_text("tig0000001\t0\ntig0000002\t1\n\n")
    assert NP2._read_uncorrected(str(blc), "0", set()) == ["tig0000001"]
    assert NP2._read_uncorrected(str(blc), "1", {"tig0000002"}) == []
    # a block whose contigs are all finished: exit 0, the partial tail is cut, nothing else happens
    out.write_text(">tig0000001 4\nACGT\n>tig0000002 8\nACGTACGT\n>tig0000002 3\nAC")
    assert NP2.main(["-g", FA, "-l", str(lst), "-r", "hifi", "--fast", "-b", str(blc), "-i", "0", "-o", str(out), "-w", "5M"]) == 0
    assert out.read_text() == ">tig0000001 4\nACGT\n>tig0000002 8\nACGTACGT\n"


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "samtools")), reason="needs oracle/_ref/samtools")
def test_loader_refuses_what_is_not_built(NP2, tmp_path):
    """(-12) a CIGAR operation the reference's bam2aln rejects ("bamaln error", ctg_cns.c:3527-3530); (-10) a split-read gap
    on a supplementary record of a contig longer than 100 kb, which would switch the reference's large-indel path on."""
    import subprocess
    from tests.conftest import REF_SAMTOOLS
    sam = subprocess.run([REF_SAMTOOLS, "view", "-h", BAM], stdout=subprocess.PIPE, check=True).stdout.decode().split("\n")

    def write_bam(lines, path):
        p = subprocess.run([REF_SAMTOOLS, "view", "-b", "-o", path, "-"], input="\n".join(lines).encode(), check=True)
        subprocess.check_call([REF_SAMTOOLS, "index", path])

    # -12: the first primary record's leading M run becomes '=' (valid BAM, but not an operation bam2aln handles)
    lines, done = [], False
    for l in sam:
        f = l.split("\t")
        if not done and not l.startswith("@") and len(f) > 5 and not int(f[1]) & 0x904 and f[5][0].isdigit():
            import re
            f[5] = re.sub(r"^(\d+S)?(\d+)M", lambda m: (m.group(1) or "") + m.group(2) + "=", f[5], count=1)
            l, done = "\t".join(f), True
        lines.append(l)
    bam12 = str(tmp_path / "eq.bam")
    write_bam(lines, bam12)
    with pytest.raises(NP2.NativeError, match="-12"):
        NP2.ContigWindows(FA, bam12, "tig0000001", 1)
    # -10: the same alignments on a contig declared (and padded to) 150 kb: short contigs take split reads in their stride,
    # long ones hand them to the large-indel path
    has_split = any("\tSA:Z:" in l and not l.startswith("@") and int(l.split("\t")[1]) & 0x800 for l in sam)
    assert has_split
    fa150 = str(tmp_path / "long.fa")
    from tests.golden.make_golden_lgs import read_fa
    d = read_fa(FA)
    with open(fa150, "w") as f:
        for n, s in d.items():
            f.write(">%s\n%s\n" % (n, s + "ACGT" * ((150000 - len(s)) // 4 + 1)))
    lines = [l.replace("LN:%d" % len(d["tig0000001"]), "LN:%d" % (len(d["tig0000001"]) + 4 * ((150000 - len(d["tig0000001"])) // 4 + 1)))
             .replace("LN:%d" % len(d["tig0000002"]), "LN:%d" % (len(d["tig0000002"]) + 4 * ((150000 - len(d["tig0000002"])) // 4 + 1))) if l.startswith("@SQ") else l for l in sam]
    # a split read on tig0000001: primary 5000M5000S at 1001, supplementary 5000S5000M at 6201 (a 200-base gap on the contig),
    # each naming the other in SA — check_indel (ctg_cns.c:2463) gives it a gap score, and the record is supplementary
    seq = d["tig0000001"].upper().replace("N", "A").replace("R", "A")
    rd = seq[1000:6000] + seq[6200:11200]
    lines = [l for l in lines if l]
    lines.append("\t".join(["split1", "0", "tig0000001", "1001", "60", "5000M5000S", "*", "0", "0", rd, "*", "SA:Z:tig0000001,6201,+,5000S5000M,60,0;"]))
    lines.append("\t".join(["split1", "2048", "tig0000001", "6201", "60", "5000S5000M", "*", "0", "0", rd, "*", "SA:Z:tig0000001,1001,+,5000M5000S,60,0;"]))
    unsorted = str(tmp_path / "long.unsorted.bam")
    subprocess.run([REF_SAMTOOLS, "view", "-b", "-o", unsorted, "-"], input="\n".join(lines).encode(), check=True)
    bam10 = str(tmp_path / "long.bam")
    subprocess.check_call([REF_SAMTOOLS, "sort", "-o", bam10, unsorted], stderr=subprocess.DEVNULL)
    subprocess.check_call([REF_SAMTOOLS, "index", bam10])
    outcomes = []
    for ctg in d:
        try:
            NP2.ContigWindows(fa150, bam10, ctg, 1).close()
            outcomes.append("ok")
        except NP2.NativeError as ex:
            assert "-10" in str(ex)
            outcomes.append("refused")
    assert outcomes == ["refused", "ok"]
    # the same split read on the 51 kb contig is taken in its stride (no large-indel path below 100 kb, ctg_cns.c:3450)
    short_lines = [l for l in sam if l] + lines[-2:]
    short_unsorted, short_bam = str(tmp_path / "short.unsorted.bam"), str(tmp_path / "short.bam")
    subprocess.run([REF_SAMTOOLS, "view", "-b", "-o", short_unsorted, "-"], input="\n".join(short_lines).encode(), check=True)
    subprocess.check_call([REF_SAMTOOLS, "sort", "-o", short_bam, short_unsorted], stderr=subprocess.DEVNULL)
    subprocess.check_call([REF_SAMTOOLS, "index", short_bam])
    cw = NP2.ContigWindows(FA, short_bam, "tig0000001", 1)
    base_cw = NP2.ContigWindows(FA, BAM, "tig0000001", 1)
    assert cw.info()[0][2] == base_cw.info()[0][2] + 1           # one more alignment: the primary half (the supplementary one is filtered)
    if L.ref_shim() is not None:
        from tests.golden.make_golden_lgs import ref_contig_windows as rcw
        want = [(a, b, n, h) for a, b, n, h, _, _ in rcw(L.ref_shim(), short_bam, "tig0000001", d["tig0000001"], 1, 5000000, 1000000)]
        assert cw.info() == want
    cw.close(), base_cw.close()
    # whether a contig is refused depends on its reads (a supplementary record whose SA partner passes check_indel); the
    # reference's own loop (shim) must agree contig by contig when it is available
    if L.ref_shim() is not None:
        from tests.golden.make_golden_lgs import ref_contig_windows
        d150 = read_fa(fa150)
        for ctg, got in zip(d, outcomes):
            try:
                ref_contig_windows(L.ref_shim(), bam10, ctg, d150[ctg], 1, 5000000, 1000000)
                want = "ok"
            except AssertionError as ex:
                want = "refused" if "-10" in str(ex) else "error"
            assert got == want, ctg


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "samtools")), reason="needs oracle/_ref/samtools")
def test_several_bams_are_merged_in_the_reference_order(NP2, tmp_path):
    """The driver maps long reads in parts and lists the part BAMs (nextpolish2.py -l): the loader takes the records in the
    order of the reference's merge iterator (bsort.c:174-199,1428-1461: position, forward strand first, list order).  The
    fixture BAM split in two by read name; against the reference's own bam_merge_iter (shim) when it is available."""
    import subprocess
    from tests.conftest import REF_SAMTOOLS
    a, b = str(tmp_path / "part_a.bam"), str(tmp_path / "part_b.bam")
    subprocess.check_call([REF_SAMTOOLS, "view", "-b", "-s", "3.5", "-o", a, "-U", b, BAM])
    subprocess.check_call([REF_SAMTOOLS, "index", a])
    subprocess.check_call([REF_SAMTOOLS, "index", b])
    lst = tmp_path / "parts.list"
    lst.write_text(a + "\n" + b + "\n")
    single = {}
    for ctg in ("tig0000001", "tig0000002"):
        for order in ([a, b], [b, a]):
            cw = NP2.ContigWindows(FA, order, ctg, 1, 20000, 4000)
            one = NP2.ContigWindows(FA, BAM, ctg, 1, 20000, 4000)
            assert [x[:3] for x in cw.info()] == [x[:3] for x in one.info()]          # same windows and alignment counts as the unsplit BAM
            single[(ctg, tuple(order))] = cw.info()
            cw.close(), one.close()
    # a list whose second BAM has nothing for the contig is the first BAM alone
    lonely = NP2.ContigWindows(FA, [a, a], "tig0000001", 1, 20000, 4000)
    assert lonely.info()[0][2] > NP2.ContigWindows(FA, a, "tig0000001", 1, 20000, 4000).info()[0][2]   # (the same reads twice: more alignments)
    lonely.close()
    if L.ref_shim() is not None:
        from tests.golden.make_golden_lgs import read_fa, ref_contig_windows
        draft = read_fa(FA)
        for ctg in draft:
            for rt, w, o in ((1, 20000, 4000), (3, 5000000, 1000000), (2, 7000, 900)):
                want = [(x[0], x[1], x[2], x[3]) for x in ref_contig_windows(L.ref_shim(), str(lst), ctg, draft[ctg], rt, w, o)]
                cw = NP2.ContigWindows(FA, [a, b], ctg, rt, w, o)
                assert cw.info() == want, (ctg, rt, w, o)
                cw.close()
