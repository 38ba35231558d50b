"""Seeded synthetic parity cases shared by the tests and tests/golden/make_golden.py.
Keyword arguments of nextpolish_b200.engine.synth_params."""

CASES = {
    # plain 30x, the BASELINE shape in miniature
    "c30": dict(seed=11, n_contigs=3, contig_len=60000, depth=30.0),
    # many small contigs of ragged length, some shorter than a read pair
    "ragged": dict(seed=12, n_contigs=24, contig_len=0, min_len=400, max_len=30000, depth=25.0),
    # deep coverage, noisy reads and draft: long non-anchor stretches, many insertion columns
    "noisy": dict(seed=13, n_contigs=2, contig_len=40000, depth=80.0, draft_snv=0.01, draft_indel=0.02,
                  read_sub=0.01, read_indel=0.004),
    # shallow coverage: zero-depth columns (FLAG_ZERO), low-support columns
    "shallow": dict(seed=14, n_contigs=2, contig_len=50000, depth=3.0),
    # lowercase draft marks: task-2 regions, windows, low-depth re-scoring
    "lower": dict(seed=15, n_contigs=3, contig_len=50000, depth=30.0, lowercase_frac=0.02),
    "lower_shallow": dict(seed=16, n_contigs=2, contig_len=40000, depth=6.0, lowercase_frac=0.05,
                          draft_indel=0.01),
}
