/* nextpolish2_b200.h — C ABI of the long-read consensus path (reference: nextpolish2.so, built from
 * source/lib/ctg_cns.c, bound by source/lib/nextpolish2.py:54-65).
 *
 * STATUS: first slice — a contig can be polished end to end in the reference's FAST mode (np2_windows_from_bam ->
 * np2_first_pass -> np2_link_windows_fast), not yet in its production mode.  What runs on the GPU is the FIRST PASS of a consensus window — tags, link tally, score chain and
 * backtrack, i.e. what get_cns_from_align_tags (ctg_cns.c:1876) does up to and including the backtrack of
 * generate_cns_from_best_score{,_fast} (:1475-1509, :1839-1857): np2_first_pass below, bit-identical to the reference
 * (tests/test_lgs_first_pass.py, tests/test_zz_lgs_gpu.py).  The reference's own six entry points (read_ref, ctg_cns_init,
 * ctg_cns_core, free_consensus_trimed_data, ctg_cns_destroy, refs_destroy — ctg_cns.c:2269,3355,3399,2151,3384,2195) are
 * NOT exported yet: ctg_cns_core also needs the low-quality-region / POA stage behind the first pass (ctg_cns.c:822-1474,
 * dag.c, align.c), the large-indel path and the window linking (:3053-3330).  The stage in front — BAM records to alignment
 * strings — exists on the host (np2_windows_from_bam / np2_windows_from_bams below; several BAMs are merged in the order
 * of the reference's bam_merge_iter, bsort.c).  INTEGRATION.md shows where these calls sit inside ctg_cns_core.
 *
 * Library: nextpolish_b200/lib/nextpolish2.so (sm_100a only, no CPU path: np2_engine_create fails without a GPU). */
#ifndef NEXTPOLISH2_B200_H
#define NEXTPOLISH2_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct np2_engine np2_engine;
np2_engine* np2_engine_create(int32_t device);
void        np2_engine_destroy(np2_engine* e);
const char* np2_last_error(void);

/* A batch of consensus windows (HOST pointers).  Window w is win_len[w] target positions long and owns the alignments
 * [win_aln0[w], win_aln0[w + 1]), in the order of the reference's tags_list (BAM order; ctg_cns_core puts the window
 * against itself first, ctg_cns.c:3456-3468).  Alignment i is the pair of gapped strings t_str / q_str
 * [str_off[i], str_off[i] + aln_len[i]) — the alignment.t_aln_str / q_aln_str of ctg_cns.h:150-162 from aln->shift on:
 * '-' = gap, 'M' = masked column — whose first column sits on window position aln_t_s[i] and is not a gap. */
typedef struct {
    int32_t         n_windows;
    const int32_t*  win_len;       /* [n_windows]                                            */
    const int32_t*  win_aln0;      /* [n_windows + 1]                                        */
    int32_t         read_type;     /* 1 ont, 2 clr, 3 hifi, 4 rs (ctg_cns.c:23-26, -r of nextpolish2.py) */
    int32_t         min_cov;       /* lower-case threshold; ctg_cns_core passes 4 (:3587)     */
    const uint32_t* aln_t_s;       /* [n_alignments]                                         */
    const uint32_t* aln_len;       /* [n_alignments] columns                                 */
    const uint64_t* str_off;       /* [n_alignments]                                         */
    const char*     t_str;         /* target (window) side of the alignments                 */
    const char*     q_str;         /* read side                                              */
    int64_t         str_bytes;
} np2_window_batch;

/* First pass of every window of the batch.  Outputs (host, capacity cap entries each): consensus bases in window order and,
 * inside a window, in forward order: out_pos = window position of the base (sub-column bases repeat the position),
 * out_base = the base, lower case where coverage <= min_cov (the rule of generate_cns_from_best_score_fast, :1492),
 * out_qv (may be NULL) = 100 * links / coverage of the chosen link (:1840); out_off[n_windows + 1] = window offsets.
 * Device memory: about 60 bytes per alignment column of the batch (the two strings, one 24-byte link record and one 32-byte
 * entry slot per column) plus 30 bytes per window position: a 5 Mb window at 30x is ~10 GB, so size batches accordingly.
 * Returns the total number of bases or: -1 cap too small, -2 a window whose last position has no node, -3 an alignment
 * that is empty, starts on a gap column or leaves its window, -4 the backtrack ran into a node without links (the
 * reference reads unallocated memory there), -5 a size limit (2^31 columns or alignment columns per batch, a 65535-long
 * insertion), -6 CUDA failure (np2_last_error). */
int64_t np2_first_pass(np2_engine* e, const np2_window_batch* batch, uint32_t* out_pos, char* out_base, uint8_t* out_qv,
                       int64_t cap, int64_t* out_off);

/* kernels launched by this engine so far / figures of the last call: [0] chain segments, [1] segments run again with their
 * true cut score, [2] stitch iterations, [3] link records */
int64_t np2_engine_launch_count(np2_engine* e);
void    np2_engine_last_stats(np2_engine* e, int64_t out[4]);
/* device time (ms, CUDA events on the engine's stream) of every launch of the last np2_first_pass, in launch order; only
 * recorded when the environment has NEXTPOLISH_B200_LGS_TIMING=1.  names[i] stay valid until the next call.  Returns the
 * number of entries written.  (NEXTPOLISH_B200_LGS_CHAIN=thread selects the one-thread-per-segment chain kernel instead
 * of the warp-per-segment one: an A/B switch for measurements.) */
int32_t np2_engine_kernel_times(np2_engine* e, const char** names, float* ms, int32_t cap);

/* ---- the stage in front of the first pass, on the host: the BAM records of one contig -> the alignment strings of its
 * consensus windows.  Restates the record loop of ctg_cns_core (ctg_cns.c:3444-3566: window geometry with `window` /
 * `overlap` positions — the reference uses >= 5 Mb / 1 Mb —, record filters, bam2aln, clip_aln, get_align_shift, the
 * coverage caps, and the draft as read_ref's 2-bit packing leaves it) for everything that does not need the large-indel
 * machinery; a contig longer than 100 kb with a split-read gap on a supplementary record is refused (NULL,
 * np2_last_error: code -10).  Needs <bam>.bai.  np2_windows_batch fills a batch that points into the handle (valid until
 * np2_windows_free) and can go straight into np2_first_pass. */
typedef struct np2_windows np2_windows;
np2_windows* np2_windows_from_bam(const char* fasta, const char* bam, const char* contig, int32_t read_type, int32_t window, int32_t overlap);
/* the same for a list of sorted, indexed BAMs (nextpolish2.py -l: the driver lists the part BAMs of the mapping step): records
 * in the order of the reference's merge iterator (bsort.c:174-199: position, forward strand first, then list order) */
np2_windows* np2_windows_from_bams(const char* fasta, const char* const* bams, int32_t n_bams, const char* contig, int32_t read_type,
                                   int32_t window, int32_t overlap);
int32_t      np2_windows_count(const np2_windows* w);
/* window i: [start, end) on the contig, its number of alignments as the reference counts them (the window itself and the
 * rare alignments left empty by the anchoring included; the empty ones are not in the batch) and an FNV-1a hash over
 * (start, length, target string, read string) of every alignment in order (diagnostics / tests) */
void         np2_windows_info(const np2_windows* w, int32_t i, int32_t* start, int32_t* end, int32_t* n_alignments, uint64_t* hash);
void         np2_windows_batch(const np2_windows* w, np2_window_batch* out);
void         np2_windows_free(np2_windows* w);
/* contig start of window i (what ctg_cns_core keeps as uncorrected_len) for all windows: out[n_windows] */
void         np2_windows_starts(const np2_windows* w, int32_t* out);

/* ---- behind the first pass: the windows' consensus joined into one sequence the way the reference's FAST mode does
 * (link_consensus_fast, ctg_cns.c:3053-3119; ctg_cns_core returns this when its local `fast` is set, :3620 — the shipped
 * reference never sets it: its production mode re-polishes low-quality regions first, which is not built).  Host code.
 * win_start[i] = contig start of window i, win_off / pos / base = np2_first_pass's outputs, overlap as given to
 * np2_windows_from_bam.  Returns the length written to out_seq, -1 cap too small, -7 windows that cannot be linked. */
int64_t      np2_link_windows_fast(int32_t n_windows, const int32_t* win_start, const int64_t* win_off, const uint32_t* pos,
                                   const char* base, int32_t overlap, char* out_seq, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* NEXTPOLISH2_B200_H */
