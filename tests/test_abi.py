"""CPU-side checks of the drop-in boundary: the shared object loads, exports every symbol the
header declares, and the struct layouts are the reference's."""
import ctypes as C
import os
import re

import pytest

from tests.conftest import ROOT, _have_gpu


def test_exports_every_declared_symbol(E):
    from nextpolish_b200 import binding
    L = C.CDLL(binding.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "nextpolish_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b((?:np_[a-z_0-9]+|config_init|config_destory|score_chain|kmer_count|snp_phase|"
                              r"snp_valid|lgspolish|polishresult_init|polishresult_destory))\s*\(", hdr))
    declared -= {"np_shard_view", "np_synth_params"}
    assert declared == set(binding.EXPORTS), declared ^ set(binding.EXPORTS)
    for sym in declared:
        assert hasattr(L, sym), sym


def test_struct_layouts_match_reference(E):
    from nextpolish_b200.binding import Configure, PolishPoint, PolishResult
    assert C.sizeof(Configure) == 152          # config.h:25-67 on x86-64
    assert Configure.ploidy.offset == 32
    assert Configure.region_count.offset == 72
    assert Configure.read_tlen.offset == 116
    assert Configure.fastafn.offset == 128
    assert C.sizeof(PolishPoint) == 8
    assert C.sizeof(PolishResult) == 24


def test_config_init_defaults_and_insert_estimate(E):
    fa = os.path.join(ROOT, "tests", "golden", "td30.step1.fa")
    bam = os.path.join(ROOT, "tests", "golden", "td30.step1.bam")
    cfg = E.default_config(fa, bam)
    c = cfg.contents
    assert (c.trim_len_edge, c.ext_len_edge, c.min_map_quality) == (2, 2, 0)       # config.c:11-13
    assert (c.indel_balance_factor_sgs, c.min_count_ratio_skip) == (0.5, 0.8)
    assert (c.min_len_ldr, c.min_len_inter_kmer, c.max_len_kmer, c.max_count_kmer) == (3, 5, 50, 50)
    assert (c.max_clip_ratio_sgs, c.max_clip_ratio_lgs) == (0.15, 0.4)
    assert c.read_len == 150 and 1000 < c.read_tlen < 3000                         # config.c:45-50,80-101
    assert c.fastafn == fa.encode() and c.bamfn == bam.encode() and c.thirdbamfn is None
    E.lib().config_destory(cfg)
    cfg = E.default_config(fa, "/nonexistent.bam")
    assert cfg.contents.bamfn is None and cfg.contents.read_tlen == 0              # config.c:43-50
    E.lib().config_destory(cfg)


def test_shard_loader_matches_bam(E):
    fa = os.path.join(ROOT, "tests", "golden", "td30.step1.fa")
    bam = os.path.join(ROOT, "tests", "golden", "td30.step1.bam")
    sh = E.Shard.load(fa, bam, with_qual=True)
    assert sh.n_contigs == 2 and sh.n_reads > 15000
    a = sh.arrays()
    assert a["ctg_off"][-1] == sh.total_bases == 111129
    assert a["ctg_read_off"][-1] == sh.n_reads
    # one contig through the .bai path gives the same records as the whole-file path
    one = E.Shard.load(fa, bam, names=[sh.names[1]], with_qual=True)
    lo, hi = a["ctg_read_off"][1], a["ctg_read_off"][2]
    b = one.arrays()
    assert one.n_reads == hi - lo
    assert bytes(b["rec"]) == bytes(a["rec"][int(a["rec_off"][lo]) * 16:int(a["rec_off"][hi]) * 16])
    assert bytes(b["qual"]) == bytes(a["qual"][int(a["qual_off"][lo]) * 16:int(a["qual_off"][hi]) * 16])


@pytest.mark.skipif(_have_gpu(), reason="only meaningful without a GPU")
def test_compute_fails_loudly_without_gpu(E):
    with pytest.raises(E.NativeError) as ei:
        E.Engine(0)
    assert "no CPU path" in str(ei.value) or "CUDA" in str(ei.value)


def test_worker_mirror_bookkeeping(tmp_path):
    """Block-file selection and resume scan of nextpolish_b200/nextpolish1.py (nextpolish1.py:148-179)."""
    from nextpolish_b200 import nextpolish1 as n1
    out = tmp_path / "o.fa"
    out.write_text(">a_np1 4\nACGT\n>b_np1 4\nAC")          # last record is partial
    polished = set()
    assert n1.scan_output(str(out), polished) == 14 and polished == {"a"}
    blc = tmp_path / "b.blc"
    blc.write_text("a\t0\nb\t0\nc\t1\n")
    assert n1.read_block(str(blc), "0", polished) == ["b"]
    fa = tmp_path / "g.fa"
    fa.write_text(">a desc\nAC\n>b\nGT\n")
    assert n1.read_block(str(fa), "all", set()) == ["a", "b"]


def test_partition_contiguous_blocks_are_contiguous_and_balanced():
    """np_partition_contiguous (csrc/multi_gpu.cu): what source/nextPolish:93-117 (blc_genome) does for the worker jobs."""
    import random
    from nextpolish_b200 import engine as E
    rng = random.Random(5)
    for n_parts in (1, 2, 4, 8):
        for n in (1, 3, 8, 200, 1000):
            lens = [int(20000 * (50 ** rng.random())) for _ in range(n)]
            part = E.partition_contiguous(lens, n_parts)
            assert part == sorted(part) and part[0] == 0 and max(part) < n_parts
            if n >= 100:
                loads = [sum(l for l, p in zip(lens, part) if p == b) for b in range(n_parts)]
                assert max(loads) - min(loads) <= 2 * max(lens)
