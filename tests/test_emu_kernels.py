"""Runs the engine's kernel bodies (the same source nvcc compiles) on the CPU through the loop
backend of tests/emu and checks them against the oracle.  This exercises device_logic.h and
engine_impl.h without a GPU; the product library never contains this path."""
import ctypes as C

import pytest

from tests.conftest import run_checker
from tests.synth_cases import CASES

TASKS_IMPLEMENTED = [1, 2]


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
