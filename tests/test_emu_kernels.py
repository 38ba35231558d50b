"""Runs the engine's kernel bodies (the same source nvcc compiles) on the CPU through the loop
backend of tests/emu and checks them against the oracle.  This exercises device_logic.h and
engine_impl.h without a GPU; the product library never contains this path."""
import ctypes as C

import pytest

from tests.conftest import run_checker
from tests.synth_cases import CASES

TASKS_IMPLEMENTED = [1, 2, 4]      # 4 = snp_valid (snpvalid.c:3-35)


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("task", TASKS_IMPLEMENTED)
def test_emulated_kernels_match_oracle(E, oracle, emu, case, task):
    sh = E.Shard.synthetic(E.synth_params(**CASES[case]), 0, CASES[case]["n_contigs"], with_qual=True)
    cfg = E.default_config(b"")
    want = run_checker(oracle.np_oracle_run, sh, task, cfg)
    stats = (C.c_int32 * 4)()
    got = run_checker(emu.np_emu_run, sh, task, cfg, (C.cast(stats, C.c_void_p),))
    for n in want:
        assert got[n] == want[n], n


@pytest.mark.parametrize("task", TASKS_IMPLEMENTED)
def test_emulated_kernels_non_default_thresholds(E, oracle, emu, task):
    sh = E.Shard.synthetic(E.synth_params(seed=77, n_contigs=2, contig_len=30000, depth=20.0,
                                          draft_indel=0.01, lowercase_frac=0.02), 0, 2, with_qual=True)
    cfg = E.default_config(b"")
    cfg.contents.trim_len_edge = 0
    cfg.contents.indel_balance_factor_sgs = 0.25
    cfg.contents.min_count_ratio_skip = 0.95
    want = run_checker(oracle.np_oracle_run, sh, task, cfg)
    got = run_checker(emu.np_emu_run, sh, task, cfg, (None,))
    assert got == want


def test_synthetic_shard_equals_bam_roundtrip(E, synth_files):
    """np_synth_shard (direct packing) and np_synth_write -> BAM -> np_shard_load agree byte for byte."""
    fa, bam = synth_files("c30")
    sa = E.Shard.load(fa, bam, with_qual=True)
    sb = E.Shard.synthetic(E.synth_params(**CASES["c30"]), 0, 3, with_qual=True)
    a, b = sa.arrays(), sb.arrays()      # views into sa / sb: keep both alive
    for k in a:
        assert bytes(a[k]) == bytes(b[k]), k


@pytest.mark.parametrize("case", sorted(CASES))
def test_emulated_fused_window_kernel_matches_oracle(E, oracle, emu, case):
    """Task 1 through the fused shared-memory window kernel (window_kernel.h) + fallback, emulated."""
    emu.np_emu_run_impl.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    emu.np_emu_run_impl.restype = C.c_int
    sh = E.Shard.synthetic(E.synth_params(**CASES[case]), 0, CASES[case]["n_contigs"])
    cfg = E.default_config(b"")
    want = run_checker(oracle.np_oracle_run, sh, 1, cfg)
    stats = (C.c_int32 * 8)()
    got = run_checker(emu.np_emu_run_impl, sh, 1, cfg, (C.cast(stats, C.c_void_p), 2))
    assert got == want
    assert stats[5] > 0                      # windows were planned
    if case == "noisy":
        assert stats[1] > 0                  # exercises the fallback path (more than WK distinct 3-mers)
    if case == "c30":
        assert stats[1] == 0                 # plain 30x data is resolved entirely in shared memory


@pytest.mark.parametrize("seed", range(8))
def test_emulated_fused_window_kernel_fuzz(E, oracle, emu, seed):
    import random
    emu.np_emu_run_impl.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    emu.np_emu_run_impl.restype = C.c_int
    rng = random.Random(1000 + seed)
    kw = dict(seed=rng.randrange(1 << 30), n_contigs=rng.choice([1, 2, 5]), contig_len=rng.choice([700, 3000, 8000, 20000]),
              depth=rng.choice([2, 5, 12, 30, 60, 120]), draft_snv=rng.choice([0.001, 0.01]), draft_indel=rng.choice([0.003, 0.02]),
              read_sub=rng.choice([0.002, 0.02]), read_indel=rng.choice([0.0001, 0.003]))
    sh = E.Shard.synthetic(E.synth_params(**kw), 0, kw["n_contigs"])
    cfg = E.default_config(b"")
    cfg.contents.trim_len_edge = rng.choice([0, 1, 2, 4])
    cfg.contents.indel_balance_factor_sgs = rng.choice([0.5, 0.25])
    cfg.contents.min_count_ratio_skip = rng.choice([0.8, 0.95])
    want = run_checker(oracle.np_oracle_run, sh, 1, cfg)
    got = run_checker(emu.np_emu_run_impl, sh, 1, cfg, (None, 2))
    assert got == want


@pytest.mark.parametrize("case", ["lower", "lower_shallow", "noisy"])
def test_emulated_task2_with_sparse_quality_stream(E, oracle, emu, case):
    dense = E.Shard.synthetic(E.synth_params(**CASES[case]), 0, CASES[case]["n_contigs"], with_qual=1)
    sparse = E.Shard.synthetic(E.synth_params(**CASES[case]), 0, CASES[case]["n_contigs"], with_qual=2)
    assert len(sparse.arrays()["qual"]) < len(dense.arrays()["qual"])
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750
    want = run_checker(oracle.np_oracle_run, dense, 2, cfg)
    assert run_checker(emu.np_emu_run, sparse, 2, cfg, (None,)) == want


@pytest.mark.parametrize("seed", range(12))
def test_emulated_kernels_fuzz_both_tasks(E, oracle, emu, seed):
    """Random data shapes x random thresholds, both tasks (this fuzzer found the window-vote flag-clearing
    subtlety: candidates span a window by alignment, but their trimmed usable intervals need not overlap)."""
    import random
    rng = random.Random(4100 + seed)
    kw = dict(seed=rng.randrange(1 << 30), n_contigs=rng.choice([1, 2, 5]), contig_len=rng.choice([3000, 8000, 20000]),
              depth=rng.choice([2, 5, 12, 30, 60]), draft_snv=rng.choice([0.001, 0.01]), draft_indel=rng.choice([0.003, 0.02]),
              read_sub=rng.choice([0.002, 0.02]), read_indel=rng.choice([0.0001, 0.003]), lowercase_frac=rng.choice([0, 0.01, 0.05, 0.15]))
    sh = E.Shard.synthetic(E.synth_params(**kw), 0, kw["n_contigs"], with_qual=True)
    cfg = E.default_config(b"")
    c = cfg.contents
    c.read_tlen = rng.choice([0, 2000])
    c.trim_len_edge, c.ext_len_edge = rng.choice([0, 1, 2, 4]), rng.choice([0, 1, 2, 3])
    c.min_len_ldr, c.min_len_inter_kmer = rng.choice([1, 3, 6]), rng.choice([0, 2, 5, 9])
    c.max_len_kmer, c.max_count_kmer, c.min_map_quality = rng.choice([10, 50, 120]), rng.choice([3, 50]), rng.choice([0, 30])
    c.indel_balance_factor_sgs, c.min_count_ratio_skip = rng.choice([0.5, 0.25, 0.75]), rng.choice([0.8, 0.6, 0.95])
    for task in (1, 2):
        want = run_checker(oracle.np_oracle_run, sh, task, cfg)
        assert run_checker(emu.np_emu_run, sh, task, cfg, (None,)) == want, (task, kw)


@pytest.mark.parametrize("rate", [0.33, 0.7, 0.1])
def test_emulated_non_dyadic_rate_uses_sequential_chain(E, oracle, emu, rate):
    """-indel_balance_factor_sgs values that are not k/1024: score sums are not exact, so the engine runs the
    chain strictly left to right over whole contigs and must still match the oracle bit for bit."""
    emu.np_emu_run_impl.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    emu.np_emu_run_impl.restype = C.c_int
    sh = E.Shard.synthetic(E.synth_params(seed=91, n_contigs=3, contig_len=30000, depth=40.0, draft_indel=0.01, read_sub=0.01), 0, 3)
    cfg = E.default_config(b"")
    cfg.contents.indel_balance_factor_sgs = rate
    want = run_checker(oracle.np_oracle_run, sh, 1, cfg)
    stats = (C.c_int32 * 8)()
    assert run_checker(emu.np_emu_run_impl, sh, 1, cfg, (C.cast(stats, C.c_void_p), 2)) == want
    assert stats[1] == stats[0]          # every column went through the tables


@pytest.mark.parametrize("kw", [
    dict(seed=5, n_contigs=3, contig_len=2000, depth=0.0),                       # no reads at all
    dict(seed=6, n_contigs=4, contig_len=0, min_len=40, max_len=160, depth=30.0),  # contigs shorter than a read
    dict(seed=7, n_contigs=1, contig_len=1, depth=0.0),                           # single-base contig
    dict(seed=8, n_contigs=2, contig_len=700, depth=400.0),                       # very deep, tiny
], ids=["noreads", "tiny_contigs", "one_base", "deep_tiny"])
def test_emulated_edge_shapes(E, oracle, emu, kw):
    emu.np_emu_run_impl.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    emu.np_emu_run_impl.restype = C.c_int
    sh = E.Shard.synthetic(E.synth_params(lowercase_frac=0.05, **kw), 0, kw["n_contigs"], with_qual=True)
    cfg = E.default_config(b"")
    for task in (1, 2):
        want = run_checker(oracle.np_oracle_run, sh, task, cfg)
        assert run_checker(emu.np_emu_run, sh, task, cfg, (None,)) == want, task
    assert run_checker(emu.np_emu_run_impl, sh, 1, cfg, (None, 2)) == run_checker(oracle.np_oracle_run, sh, 1, cfg)


def test_two_bit_and_four_bit_records_polish_identically(E, oracle, emu, monkeypatch):
    """The packer ships A/C/G/T-only reads with 2 bits per base (include/nextpolish_b200.h); forcing the 4-bit
    form must change the record stream but not a single polished base (oracle, general kernels, window kernel)."""
    emu.np_emu_run_impl.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    emu.np_emu_run_impl.restype = C.c_int
    kw = dict(seed=123, n_contigs=2, contig_len=30000, depth=30.0, draft_indel=0.01, lowercase_frac=0.02)
    two = E.Shard.synthetic(E.synth_params(**kw), 0, 2, with_qual=True)
    monkeypatch.setenv("NEXTPOLISH_B200_4BIT", "1")
    four = E.Shard.synthetic(E.synth_params(**kw), 0, 2, with_qual=True)
    monkeypatch.delenv("NEXTPOLISH_B200_4BIT")
    assert len(two.arrays()["rec"]) < 0.75 * len(four.arrays()["rec"])
    assert two.algorithmic_bytes(1) == four.algorithmic_bytes(1)          # defined on the 4-bit form
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750
    for task in (1, 2):
        want = run_checker(oracle.np_oracle_run, four, task, cfg)
        assert run_checker(oracle.np_oracle_run, two, task, cfg) == want
        assert run_checker(emu.np_emu_run, two, task, cfg, (None,)) == want
        assert run_checker(emu.np_emu_run, four, task, cfg, (None,)) == want
    want = run_checker(oracle.np_oracle_run, four, 1, cfg)
    assert run_checker(emu.np_emu_run_impl, two, 1, cfg, (None, 2)) == want
    assert run_checker(emu.np_emu_run_impl, four, 1, cfg, (None, 2)) == want


@pytest.mark.parametrize("case", ["c30", "noisy", "ragged", "shallow"])
def test_emulated_column_pass_general_slices(E, oracle, emu_general_slices, case):
    """column_pass.h handles a thread slice of up to 62 columns with 64-bit bit tricks and longer ones (dense
    insertion sub-columns) with per-column loops: a build with the limit lowered to 9 runs those loops on ordinary data."""
    L = emu_general_slices
    L.np_emu_run_impl.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    L.np_emu_run_impl.restype = C.c_int
    sh = E.Shard.synthetic(E.synth_params(**CASES[case]), 0, CASES[case]["n_contigs"])
    cfg = E.default_config(b"")
    want = run_checker(oracle.np_oracle_run, sh, 1, cfg)
    assert run_checker(L.np_emu_run_impl, sh, 1, cfg, (None, 2)) == want


@pytest.mark.parametrize("case", ["c30", "lower", "ragged", "lower_shallow"])
@pytest.mark.parametrize("task", [2, 4])
def test_emulated_region_lists_literal_walk(E, oracle, emu_general_slices, case, task):
    """Task 2 / 4 merge their region lists with flat kernels when every list has the usual shape and with the literal
    per-contig walk of contig_merge_region otherwise; this build always takes the walk."""
    if case not in CASES:
        pytest.skip("no such synthetic case")
    sh = E.Shard.synthetic(E.synth_params(**CASES[case]), 0, CASES[case]["n_contigs"], with_qual=1)
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750
    want = run_checker(oracle.np_oracle_run, sh, task, cfg)
    assert run_checker(emu_general_slices.np_emu_run, sh, task, cfg, (None,)) == want


@pytest.mark.parametrize("seed", range(12))
def test_emulated_region_merge_mutated_thresholds(E, oracle, emu, seed):
    """Task 2 / 4 with the thresholds that shape the region lists mutated (ext_len_edge = 0 makes one-base regions whose
    first pair the literal contig_merge_region compares with itself: the flat merge must detect the shape and hand over
    to the per-contig walk) and lowercase densities from sparse to 40 %: emulated kernels vs the oracle."""
    import random
    rng = random.Random(1000 + seed)
    p = dict(seed=900 + seed, n_contigs=rng.choice([1, 2, 5]), contig_len=rng.choice([3000, 20000]), depth=rng.choice([8.0, 30.0]),
             lowercase_frac=rng.choice([0.002, 0.02, 0.1, 0.4]), draft_indel=rng.choice([0, 0.01]))
    sh = E.Shard.synthetic(E.synth_params(**p), 0, p["n_contigs"], with_qual=1)
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750
    cfg.contents.ext_len_edge = rng.choice([0, 0, 1, 2, 5])
    cfg.contents.min_len_inter_kmer = rng.choice([0, 1, 5, 20])
    cfg.contents.min_len_ldr = rng.choice([0, 1, 3, 10])
    for task in (2, 4):
        try:
            want = run_checker(oracle.np_oracle_run, sh, task, cfg)
        except Exception:
            assert task == 4          # snp_valid on cut-point lists the reference itself leaves undefined (oracle returns -2)
            continue
        assert run_checker(emu.np_emu_run, sh, task, cfg, (None,)) == want, (task, p)
