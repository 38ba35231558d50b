/*
 * np_oracle.h — TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
 *
 * CPU restatement of the NextPolish short-read hot path (tasks 1 and 2) over the packed shard
 * format of include/nextpolish_b200.h.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.
 *
 * Parity pinning: the reference ships NO golden vectors for this path (SURVEY.md section 4 /
 * 8c), so this restatement is pinned against the reference itself compiled from source
 * (oracle/_ref/nextpolish1, built by oracle/Makefile) on source/test_data and on seeded
 * synthetic sets: tests/test_oracle_vs_ref.py and tests/golden/.
 */
#ifndef NP_ORACLE_H
#define NP_ORACLE_H
#include "../include/nextpolish_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Runs `task` (1 = score_chain, 2 = kmer_count) on every contig of the shard, sequentially,
 * single thread. out_seq receives the concatenated polished sequences, out_off[n_contigs+1]
 * their offsets. Returns 0, or -1 if out_cap is too small / arguments are invalid. */
int np_oracle_run(const np_shard_view* shard, int task, const Configure* cfg,
                  uint8_t* out_seq, int64_t out_cap, int64_t* out_off);

/* Same for one contig of the shard. */
int np_oracle_run_contig(const np_shard_view* shard, int contig, int task, const Configure* cfg,
                         uint8_t* out_seq, int64_t out_cap, int64_t* out_len);

/* One contig with the PolishPoint trace of contig.c:743-799 (what Configure.trace_polish_open makes the reference return). */
int np_oracle_run_contig_points(const np_shard_view* shard, int contig, int task, const Configure* cfg,
                                uint8_t* out_seq, int64_t out_cap, int64_t* out_len,
                                PolishPoint* pts, int64_t pts_cap, int64_t* n_pts);

/* Default thresholds exactly as config.c:11-40 (file names left NULL, read_tlen 0). */
void np_oracle_default_config(Configure* cfg);

#ifdef __cplusplus
}
#endif
#endif
