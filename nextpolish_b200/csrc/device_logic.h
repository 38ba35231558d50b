// device_logic.h — per-read / per-column logic of the polishing engine, written once as
// __host__ __device__ code.  The product compiles it with nvcc into the sm_100a kernels of
// engine.cu; tests/emu compiles the very same functions with g++ and drives them with plain
// loops so the kernel bodies can be unit-tested on a box without a GPU (test build only —
// the product library contains no CPU execution path).
//
// Coordinates: all contigs of a shard are concatenated into one global position space
// (gpos = ctg_goff[k] + pos) and one global column space (a column is a reference position or
// an insertion sub-column behind it, in contig_data_next order, contig.c:385-399).
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NP_HD __host__ __device__ __forceinline__
#else
#define NP_HD inline
#endif

namespace npd {

enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };
enum { SYM_GAP = 3 };                       // BASE_DEL, config.h:19
enum { FLAG_ZERO = 1, FLAG_COVERAGE = 2,    // base.h:10-11
       CF_FIRST = 64, CF_LAST = 128 };      // engine-private: first / last column of a contig

// thresholds copied out of Configure (config.h:25-67) for the device
struct Params {
    int32_t trim_len_edge, ext_len_edge, min_map_quality;
    double  rate;                  // indel_balance_factor_sgs
    double  min_count_ratio_skip;
    int32_t min_len_ldr, min_len_inter_kmer, max_len_kmer, max_count_kmer;
    double  max_clip_ratio_sgs;
    int32_t read_tlen;
    int32_t trace;                 // Configure.trace_polish_open: also produce the PolishPoint trace
};

// packed record header (include/nextpolish_b200.h)
struct Rec {
    int32_t pos; uint32_t flag, mapq; int32_t isize; int32_t l_qseq, n_cigar;
    const uint32_t* cigar; const uint8_t* seq;
    uint32_t enc;                  // 0: 4-bit nt16 codes, 1: 2 bits per base (include/nextpolish_b200.h)
};
struct alignas(16) RecHead { uint32_t w0, w1, w2, w3; };   // records are 16-byte aligned: one 128-bit load
NP_HD Rec load_rec(const uint8_t* rec, const uint32_t* rec_off, int64_t r) {
    const uint8_t* p = rec + (size_t)rec_off[r] * 16;
    const uint32_t* w = (const uint32_t*)p;
    Rec o;
    const RecHead hd = *(const RecHead*)p;
    const uint32_t w0 = hd.w0, w1 = hd.w1, w2 = hd.w2, w3 = hd.w3;
    o.pos = (int32_t)w0;
    o.flag = w1 & 0xffffu;
    o.mapq = (w1 >> 16) & 0xffu;
    o.enc = w1 >> 24;
    o.isize = (int32_t)w2;
    o.l_qseq = (int32_t)(w3 & 0xffffu);
    o.n_cigar = (int32_t)(w3 >> 16);
    o.cigar = w + 4;
    o.seq = p + 16 + 4 * (size_t)o.n_cigar;
    return o;
}
NP_HD uint32_t seqi(const uint8_t* s, int32_t i) { return (s[i >> 1] >> ((~i & 1) << 2)) & 0xfu; }   // bam_seqi
NP_HD uint32_t rseq(const Rec& r, int32_t i) {                                                        // nt16 code of base i
    return r.enc ? 1u << ((r.seq[i >> 2] >> (6 - 2 * (i & 3))) & 3u) : seqi(r.seq, i);
}
NP_HD int cig_op(uint32_t c) { return (int)(c & 0xf); }
NP_HD int32_t cig_len(uint32_t c) { return (int32_t)(c >> 4); }

// base.c:6-15 (strtobase): ASCII -> nt16 code, everything unknown -> 15
NP_HD uint32_t base_code(uint32_t c) {
    // nibble LUT over 'A'..'P' and 'Q'..'Z': A1 B14 C2 D13 G4 H11 K12 M3 | R5 S6 T8 V7 W9 Y10, every other letter 15
    const unsigned long long lo = 0xfff3fcffb4ffd2e1ull, hi = 0xfffffffaf97f865full;
    const uint32_t i = c - 65u;
    if (i < 16u) return (uint32_t)(lo >> (4u * i)) & 0xfu;
    if (i < 26u) return (uint32_t)(hi >> (4u * (i - 16u))) & 0xfu;
    return c == 61u ? 0u : 15u;        // '='
}
NP_HD uint8_t code_char(uint32_t b) { return (uint8_t)("=ACMGRSVTWYHKDBN"[b & 15]); }   // base.c:5

// contig.c:632-646
NP_HD double clip_rate(const Rec& r) {
    int32_t add = 0;
    if (cig_op(r.cigar[0]) == OP_S) add += cig_len(r.cigar[0]);
    if (cig_op(r.cigar[r.n_cigar - 1]) == OP_S) add += cig_len(r.cigar[r.n_cigar - 1]);
    return r.l_qseq > 0 ? add / (double)r.l_qseq : 0;
}
// contig.c:667-677 (task 1) / contig.c:648-665 (task 2): admissibility level 0/1/2
NP_HD int filter_level(const Rec& r, int task, const Params& P) {
    if ((r.flag & 0xC04u) != 0) return 0;
    if (task == 1) return 1;
    int32_t tl = r.isize >= 0 ? r.isize : -r.isize;
    double cr = clip_rate(r);
    int res = 0;
    if ((tl > 0 && tl < P.read_tlen) || cr < P.max_clip_ratio_sgs) {
        res = 1;
        if ((int32_t)r.mapq >= P.min_map_quality && cr < P.max_clip_ratio_sgs + 0.05) res = 2;
    }
    return res;
}
// contig.c:333-358: usable query interval (homopolymer loops bounded to the read)
NP_HD void cut_read(const Rec& r, int32_t trim, int32_t* qstart, int32_t* qend) {
    int32_t add = 0;
    if (cig_op(r.cigar[0]) == OP_S) add = cig_len(r.cigar[0]);
    int32_t qs = trim + add;
    add = 0;
    if (cig_op(r.cigar[r.n_cigar - 1]) == OP_S) add = cig_len(r.cigar[r.n_cigar - 1]);
    int32_t qe = r.l_qseq - trim - add - 1;
    if (trim > 0) {
        while (qs < r.l_qseq && qs >= 1 && rseq(r, qs) == rseq(r, qs - 1)) qs++;
        while (qe >= 0 && qe + 1 < r.l_qseq && rseq(r, qe) == rseq(r, qe + 1)) qe--;
    }
    *qstart = qs; *qend = qe;
}
// reference span: walk end (M,D only — what contig_parse_read advances on) and htslib's
// bam_endpos (M,D,N,=,X; sam.c:391-397)
NP_HD void ref_spans(const Rec& r, int32_t* walk_len, int32_t* hts_len) {
    int32_t w = 0, h = 0;
    for (int i = 0; i < r.n_cigar; i++) {
        int op = cig_op(r.cigar[i]); int32_t len = cig_len(r.cigar[i]);
        if (op == OP_M || op == OP_D) { w += len; h += len; }
        else if (op == OP_N || op == OP_EQ || op == OP_X) h += len;
    }
    *walk_len = w; *hts_len = h;
}

// The per-read walk of contig_parse_read (contig.c:247-331) / ss_parse_read_kmer
// (kmercount.c:365-465) over region [start,end] (global positions), delivering every symbol the
// read casts to the visitor: v.sym(column, symbol, query_pos_or_-1, is_subcolumn_gap).
// gshift = global offset of the read's contig (gpos = gshift + pos); contig-local pos 0 is
// tested for the "insertion before the first base" rule (contig.c:300,315-319).
template <class V>
NP_HD void walk_read(const Rec& r, int32_t gshift, int32_t start, int32_t end,
                     int32_t qstart, int32_t qend, const int32_t* colbase, V& v) {
    int32_t pos = gshift + r.pos, qpos = 0;
    int last = OP_I;
    for (int i = 0; i < r.n_cigar; i++) {
        int32_t len = cig_len(r.cigar[i]);
        int cur = cig_op(r.cigar[i]);
        if (cur == OP_M || cur == OP_D) {
            // only bases with start <= pos <= end and qstart <= qpos <= qend cast anything: jump
            // straight to that sub-run [ja, jb] of the op instead of stepping base by base
            int32_t ja = pos < start ? start - pos : 0, jb = pos + len - 1 > end ? end - pos : len - 1;
            if (cur == OP_M) {
                if (qstart - qpos > ja) ja = qstart - qpos;
                if (qend - qpos < jb) jb = qend - qpos;
            } else if (qpos < qstart || qpos > qend) jb = ja - 1;
            for (int32_t j = ja; j <= jb; j++) {
                int32_t p = pos + j, q = cur == OP_M ? qpos + j : qpos;
                int lastj = j == 0 ? last : cur;
                if (lastj != OP_I && p > start && (q > qstart || (q == qstart && lastj == OP_D))) {
                    int32_t cb = colbase[p - 1], n = colbase[p] - cb - 1;
                    for (int32_t k = 0; k < n; k++) v.sym(cb + 1 + k, (uint32_t)SYM_GAP, -1, true);
                }
                if (cur == OP_D) v.sym(colbase[p], (uint32_t)SYM_GAP, -1, false);
                else v.sym(colbase[p], rseq(r, q), q, false);
            }
            pos += len;
            if (cur == OP_M) qpos += len;
            if (len > 0) last = cur;
        } else if (cur == OP_I) {
            if (pos != gshift) {
                bool in_reg = pos > start && pos <= end;
                int32_t cb = 0, n = 0;
                if (in_reg) { cb = colbase[pos - 1]; n = colbase[pos] - cb - 1; }
                int32_t j = 0;
                for (; j < len; j++, qpos++) {
                    if (in_reg && qpos >= qstart && qpos <= qend) {
                        if (j < n) v.sym(cb + 1 + j, rseq(r, qpos), qpos, false);
                        else v.overflow();
                    }
                }
                if (in_reg && qpos > qstart && qpos <= qend + 1)
                    for (; j < n; j++) v.sym(cb + 1 + j, (uint32_t)SYM_GAP, -1, true);
                last = cur;
            } else {
                qpos += len; qstart += len; last = cur;
            }
        } else if (cur == OP_S || cur == OP_H) {
            qpos += len;
        }
        if (pos > end) break;
    }
}

// ---- small search helpers ----------------------------------------------------------------
// first index in [lo,hi) with a[idx] > key (a non-decreasing)
NP_HD int64_t upper_bound_i32(const int32_t* a, int64_t lo, int64_t hi, int32_t key) {
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (a[m] <= key) lo = m + 1; else hi = m; }
    return lo;
}
// first index in [lo,hi) with a[idx] >= key
NP_HD int64_t lower_bound_i32(const int32_t* a, int64_t lo, int64_t hi, int32_t key) {
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (a[m] < key) lo = m + 1; else hi = m; }
    return lo;
}
NP_HD int32_t find_contig_i64(const int64_t* off, int32_t n, int64_t key) {   // off[k] <= key < off[k+1]
    int32_t lo = 0, hi = n;
    while (lo + 1 < hi) { int32_t m = (lo + hi) >> 1; if (off[m] <= key) lo = m; else hi = m; }
    return lo;
}
NP_HD int32_t find_contig_i32(const int32_t* off, int32_t n, int32_t key) {
    int32_t lo = 0, hi = n;
    while (lo + 1 < hi) { int32_t m = (lo + hi) >> 1; if (off[m] <= key) lo = m; else hi = m; }
    return lo;
}

NP_HD uint32_t sym_get(const uint32_t* words, int32_t i) { return (words[i >> 3] >> ((i & 7) << 2)) & 0xfu; }

}  // namespace npd
