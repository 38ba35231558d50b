// lgs_first_pass.h — first pass of the long-read consensus window (reference: nextpolish2.so, source/lib/ctg_cns.c;
// SURVEY.md 8f-2), written once as __host__ __device__ functors: nvcc turns them into the sm_100a kernels of
// lgs_consensus.cu, tests/emu/emu_lgs.cpp compiles the very same bodies with g++ and drives them with plain loops (test
// build only: the product library has no CPU path).
//
// What the pass computes (ctg_cns.c line numbers):
//   tags       every alignment column becomes a tag (target position, sub-column, base)      get_align_tags :1213, get_align_tag :303
//   tally      per node (position, sub-column, base): the distinct (previous, pre-previous)   update_msa :324
//              tag pairs with their link counts, in first-seen order
//   chain      integer scores along the links, one rule set per read type                    get_cns_from_align_tags :1876-2128
//   backtrack  best node of the last column back to the first                                :1475-1509, qv :1839-1846
//
// How it is laid out for the GPU.  All windows of a batch share one global column space (col = win_col0[w] + position).
//   * alignments are cut into stretches of 256 columns; a scan over the target-column counts of the stretches gives every
//     stretch its starting position, so tags are produced by one thread per stretch (count pass, then emit pass);
//   * every countable tag becomes a 24-byte record in its column's bucket (counts -> exclusive scan -> atomically filled);
//     one thread per column tallies its bucket into entries and orders them by (sub-column, base, first read seen): the
//     reference's first-seen order is the order of the alignments, and a node is visited at most once per alignment;
//   * the chain is sequential in the column order, but a column that holds exactly ONE entry and no sub-columns is a cut:
//     every later score is that entry's score plus something local.  Segments between cuts run in parallel, each with the
//     cut's score replaced by the symbol BIG; scores are only ever compared, and a comparison between a score that
//     descends from the cut (>= BIG/2) and one that does not (chains started by reads that begin inside the segment)
//     is decided for the cut's side while recording the smallest true cut score T for which that is right.  A
//     sequential stitch over the cuts then computes the true scores and re-runs, with its true score, every segment whose
//     assumption fails (typically only near a window's start); repeated until nothing changes, the result is exactly the
//     sequential chain's.  The backtrack passes through every cut it reaches, so it runs per segment as well.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NP2_HD __host__ __device__ __forceinline__
#else
#define NP2_HD inline
#endif

namespace np2 {

#ifndef NP2_STRETCH
#define NP2_STRETCH 256          // alignment columns per tag-walk thread
#endif
#ifndef NP2_CUT_BLOCK
#define NP2_CUT_BLOCK 128        // at most one chain segment starts per this many window columns
#endif
enum { STRETCH = NP2_STRETCH, CUT_BLOCK = NP2_CUT_BLOCK };   // (the test build also compiles tiny values to reach every seam)
enum { ERR_RANGE = 1, ERR_LAST = 2, ERR_EMPTY = 4, ERR_LIMIT = 8 };
enum { READS_ONT = 1, READS_CLR = 2, READS_HIFI = 3, READS_RS = 4 };           // ctg_cns.c:23-26
enum { MODE_SPEC = 0, MODE_RERUN = 1, MODE_EXACT = 2 };
static const int64_t BIG = (int64_t)1 << 50;
static const int64_t NEG_INF = INT64_MIN;

struct Rec {                       // one countable tag with its two predecessors (columns are global, -1 = head)
    int32_t r; uint16_t d; uint8_t b, pp_b; int32_t pp_t, ppp_t; uint16_t pp_d, ppp_d; uint8_t ppp_b, pad[3];
};
struct Ent {                       // one (pp, ppp) pair of a node; aux: first read seen (tally), then best index (first entry of a node)
    int32_t pp_t, ppp_t; uint16_t pp_d, ppp_d, d, link; uint8_t b, pp_b, ppp_b, pad; int32_t aux; int64_t score;
};
static_assert(sizeof(Rec) == 24 && sizeof(Ent) == 32, "record layout");

struct Dev {
    int32_t n_win, n_aln, Ctot, n_blk; int64_t n_stretch;
    int32_t read_type, min_cov;
    const int32_t* win_col0; const int32_t* win_aln0; const int32_t* win_blk0;      // [n_win + 1]
    const uint32_t* aln_t_s; const uint32_t* aln_len; const uint64_t* str_off; const int64_t* aln_st0;   // [n_aln] / [n_aln + 1]
    const char* t_str; const char* q_str;
    int32_t* st_nt; int32_t* st_ex;                                                  // [n_stretch + 1]
    uint32_t* cov; uint32_t* msz; int32_t* cnt; int32_t* rec_off; int32_t* fill; int32_t* n_ent;   // [Ctot + 1]
    Rec* rec; Ent* ent;
    int32_t* blk_col; int32_t* blk_has; int32_t* blk_idx;                            // [n_blk + 1]
    int32_t* seg_col; int32_t* seg_win; int32_t* seg_mode; int32_t* seg_n; int32_t* seg_out; int32_t* seg_reach;   // [n_blk + 1]
    int64_t* seg_S; int64_t* seg_T; int64_t* seg_abs;
    int32_t* win_gb;                                                                 // [n_win * 3]
    int32_t* n_invalid; uint32_t* err;
    uint32_t* out_pos; uint8_t* out_base; uint8_t* out_qv; int64_t* win_out;        // outputs; win_out [n_win + 1]
};

NP2_HD uint8_t base_to_int(unsigned char c) {                                        // ctg_cns.c:58-67
    switch (c) {
    case 'A': case 'a': return 0;
    case 'T': case 't': return 1;
    case 'G': case 'g': return 2;
    case 'C': case 'c': return 3;
    case 'N': return 5;
    case 'M': return 6;
    default: return 4;
    }
}
NP2_HD uint8_t int_to_base(int b) { return b == 0 ? 'A' : b == 1 ? 'T' : b == 2 ? 'G' : b == 3 ? 'C' : b == 4 ? '-' : b == 5 ? 'N' : 'M'; }

// largest i in [0, n) with a[i] <= v (a ascending, a[0] <= v)
template <class T, class V> NP2_HD int32_t owner(const T* a, int32_t n, V v) {
    int32_t lo = 0, hi = n - 1;
    while (lo < hi) { const int32_t mid = (lo + hi + 1) >> 1; if ((V)a[mid] <= v) lo = mid; else hi = mid - 1; }
    return lo;
}

struct Stretch { int32_t r, w; int64_t i0, i1; const char* t; const char* q; };
NP2_HD Stretch stretch_of(const Dev& d, int64_t s) {
    Stretch x;
    x.r = owner(d.aln_st0, d.n_aln, s);
    x.w = owner(d.win_aln0, d.n_win, x.r);
    x.i0 = (s - d.aln_st0[x.r]) * STRETCH;
    x.i1 = x.i0 + STRETCH < (int64_t)d.aln_len[x.r] ? x.i0 + STRETCH : (int64_t)d.aln_len[x.r];
    x.t = d.t_str + d.str_off[x.r]; x.q = d.q_str + d.str_off[x.r];
    return x;
}
// length of the run of '-' target columns ending at column j (0 when column j is a target column)
NP2_HD int32_t gap_run(const char* t, int64_t j) { int32_t n = 0; while (j >= 0 && t[j] == '-') { n++; j--; } return n; }

struct StretchCount {               // target columns of every stretch; an alignment must start on a target column
    Dev d;
    template <class Ops> NP2_HD void operator()(int64_t s, Ops& ops) const {
        const Stretch x = stretch_of(d, s);
        if (x.i0 == 0 && x.t[0] == '-') ops.atomic_or(d.err, ERR_RANGE);
        int32_t n = 0;
        for (int64_t j = x.i0; j < x.i1; j++) n += x.t[j] != '-';
        d.st_nt[s] = n;
    }
};

template <bool EMIT> struct TagWalk {   // count pass: coverage, sub-column count, bucket sizes; emit pass: the records
    Dev d;
    template <class Ops> NP2_HD void operator()(int64_t s, Ops& ops) const {
        const Stretch x = stretch_of(d, s);
        const int32_t col0 = d.win_col0[x.w], wlen = d.win_col0[x.w + 1] - col0;
        int64_t te = (int64_t)d.aln_t_s[x.r] - 1 + (d.st_ex[s] - d.st_ex[d.aln_st0[x.r]]);   // position of the last target column before i0
        int32_t delta = x.i0 > 0 ? gap_run(x.t, x.i0 - 1) : 0;
        // the two tags before the stretch (align_tag_head before the alignment's first tag, ctg_cns.c:52-56)
        int32_t pp_t = -1, ppp_t = -1; uint16_t pp_d = 0, ppp_d = 0; uint8_t pp_b = 0, ppp_b = 0;
        if (x.i0 > 0) {
            pp_t = col0 + (int32_t)te; pp_d = (uint16_t)delta; pp_b = base_to_int((unsigned char)x.q[x.i0 - 1]);
            if (x.i0 > 1) {
                ppp_t = col0 + (int32_t)(te - (x.t[x.i0 - 1] != '-' ? 1 : 0));
                ppp_d = (uint16_t)gap_run(x.t, x.i0 - 2); ppp_b = base_to_int((unsigned char)x.q[x.i0 - 2]);
            }
        }
        for (int64_t j = x.i0; j < x.i1; j++) {
            if (x.t[j] == '-') delta++; else { te++; delta = 0; }
            if (te < 0 || te >= wlen) { ops.atomic_or(d.err, ERR_RANGE); return; }
            if (delta > 65534) { ops.atomic_or(d.err, ERR_LIMIT); return; }
            const int32_t col = col0 + (int32_t)te;
            const uint8_t b = base_to_int((unsigned char)x.q[j]);
            const bool countable = b != 6 && pp_b != 6;                              // update_msa :331-335
            if (!EMIT) {
                if (delta == 0 && b != 6) ops.atomic_add_u32(d.cov + col, 1);       // :1232-1234
                ops.atomic_max_u32(d.msz + col, (uint32_t)delta + 1);                // :1236-1238
                if (countable) ops.atomic_add(d.cnt + col, 1);
            } else if (countable) {
                const int32_t slot = d.rec_off[col] + ops.atomic_add_ret(d.fill + col, 1);
                Rec rc;
                rc.r = x.r - d.win_aln0[x.w]; rc.d = (uint16_t)delta; rc.b = b; rc.pp_b = pp_b; rc.pp_t = pp_t; rc.ppp_t = ppp_t;
                rc.pp_d = pp_d; rc.ppp_d = ppp_d; rc.ppp_b = ppp_b; rc.pad[0] = rc.pad[1] = rc.pad[2] = 0;
                d.rec[slot] = rc;
            }
            ppp_t = pp_t; ppp_d = pp_d; ppp_b = pp_b;
            pp_t = col; pp_d = (uint16_t)delta; pp_b = b;
        }
    }
};

struct ColumnLinks {                // one column's bucket -> entries, ordered by (sub-column, base, first read seen)
    Dev d;
    template <class Ops> NP2_HD void operator()(int64_t col, Ops&) const {
        const int32_t o = d.rec_off[col], n = d.fill[col];
        Ent* E = d.ent + o;
        int32_t ne = 0;
        for (int32_t i = 0; i < n; i++) {
            const Rec rc = d.rec[o + i];
            int32_t k = 0;
            for (; k < ne; k++) {
                const Ent& e = E[k];
                if (e.d == rc.d && e.b == rc.b && e.pp_t == rc.pp_t && e.pp_d == rc.pp_d && e.pp_b == rc.pp_b &&
                    e.ppp_t == rc.ppp_t && e.ppp_d == rc.ppp_d && e.ppp_b == rc.ppp_b) break;
            }
            if (k < ne) { E[k].link = (uint16_t)(E[k].link + 1); if (rc.r < E[k].aux) E[k].aux = rc.r; }
            else {
                Ent e;
                e.pp_t = rc.pp_t; e.ppp_t = rc.ppp_t; e.pp_d = rc.pp_d; e.ppp_d = rc.ppp_d; e.d = rc.d; e.link = 1;
                e.b = rc.b; e.pp_b = rc.pp_b; e.ppp_b = rc.ppp_b; e.pad = 0; e.aux = rc.r; e.score = 0;
                E[ne++] = e;
            }
        }
        for (int32_t i = 1; i < ne; i++) {                                           // insertion sort (a handful of entries)
            const Ent e = E[i];
            int32_t j = i - 1;
            while (j >= 0 && (E[j].d > e.d || (E[j].d == e.d && (E[j].b > e.b || (E[j].b == e.b && E[j].aux > e.aux))))) { E[j + 1] = E[j]; j--; }
            E[j + 1] = e;
        }
        d.n_ent[col] = ne;
    }
};

struct CutBlocks {                  // first cut column of every block of 128 columns (a window's first block: its first column)
    Dev d;
    template <class Ops> NP2_HD void operator()(int64_t j, Ops&) const {
        const int32_t w = owner(d.win_blk0, d.n_win, (int32_t)j);
        const int32_t kb = (int32_t)j - d.win_blk0[w];
        const int32_t c0 = d.win_col0[w] + kb * CUT_BLOCK;
        const int32_t c1 = c0 + CUT_BLOCK < d.win_col0[w + 1] ? c0 + CUT_BLOCK : d.win_col0[w + 1];
        int32_t cut = -1;
        if (kb == 0) cut = c0;
        else for (int32_t c = c0; c < c1 && c < d.win_col0[w + 1] - 1; c++)          // (the window's last column ends a segment, it never starts one)
            if (d.n_ent[c] == 1 && d.msz[c] == 1) { cut = c; break; }
        d.blk_col[j] = cut; d.blk_has[j] = cut >= 0 ? 1 : 0;
    }
};
struct SegScatter {
    Dev d;
    template <class Ops> NP2_HD void operator()(int64_t j, Ops&) const {
        if (!d.blk_has[j]) return;
        const int32_t k = d.blk_idx[j];
        d.seg_col[k] = d.blk_col[j]; d.seg_win[k] = owner(d.win_blk0, d.n_win, (int32_t)j);
        d.seg_mode[k] = MODE_SPEC; d.seg_S[k] = 0; d.seg_T[k] = NEG_INF; d.seg_abs[k] = 0; d.seg_n[k] = 0; d.seg_reach[k] = 0;
    }
};

// comparisons of scores inside a segment that runs with the cut's score replaced by BIG (see the header)
struct Cmp {
    int64_t T; bool spec;
    NP2_HD bool mixed(int64_t a, int64_t b) {
        if (!spec) return false;
        const bool ma = a >= BIG / 2, mb = b >= BIG / 2;
        if (ma == mb) return false;
        const int64_t f = ma ? b : a, r = (ma ? a : b) - BIG;
        if (f > NEG_INF / 2) { const int64_t need = f - r + 1; if (need > T) T = need; }
        return true;
    }
    NP2_HD bool gt(int64_t a, int64_t b) { if (mixed(a, b)) return a >= BIG / 2; return a > b; }
    NP2_HD bool eq(int64_t a, int64_t b) { if (mixed(a, b)) return false; return a == b; }
};

struct SegRange { int32_t w, col0, cend, c, hi; bool start; };
NP2_HD SegRange seg_range(const Dev& d, int32_t k, int32_t ns) {
    SegRange s;
    s.w = d.seg_win[k]; s.col0 = d.win_col0[s.w]; s.cend = d.win_col0[s.w + 1] - 1; s.c = d.seg_col[k];
    s.start = s.c == s.col0;
    s.hi = (k + 1 < ns && d.seg_win[k + 1] == s.w) ? d.seg_col[k + 1] : s.cend;
    return s;
}
// first entry of node (dd, bb) in column col, its run length in *len (0 when the node has no entries)
NP2_HD int32_t find_node(const Dev& d, int32_t col, uint16_t dd, uint8_t bb, int32_t* len) {
    const int32_t o = d.rec_off[col], ne = d.n_ent[col];
    int32_t i = 0;
    while (i < ne && !(d.ent[o + i].d == dd && d.ent[o + i].b == bb)) i++;
    int32_t j = i;
    while (j < ne && d.ent[o + j].d == dd && d.ent[o + j].b == bb) j++;
    *len = j - i;
    return o + i;
}

struct Chain {                      // the score chain of one segment (get_cns_from_align_tags :1890-2128)
    Dev d; int32_t rerun_only;
    template <class Ops> NP2_HD void operator()(int64_t kk, Ops&) const {
        const int32_t ns = d.blk_idx[d.n_blk], k = (int32_t)kk;
        if (k >= ns) return;
        const int32_t mode = d.seg_mode[k];
        if (rerun_only && mode != MODE_RERUN) return;
        const SegRange sr = seg_range(d, k, ns);
        const bool exact = sr.start || mode == MODE_RERUN;
        const int64_t base = sr.start ? 0 : (mode == MODE_RERUN ? d.seg_S[k] : BIG);
        Cmp cmp; cmp.T = NEG_INF; cmp.spec = !exact;
        const int64_t pen = d.read_type == READS_HIFI ? 4 : 3;
        const int32_t rt = d.read_type;
        int64_t gbs = NEG_INF; int32_t gb_c = -1, gb_d = 0, gb_b = 0;
        for (int32_t col = sr.start ? sr.c : sr.c + 1; col <= sr.hi; col++) {
            const int32_t o = d.rec_off[col], ne = d.n_ent[col];
            const int64_t cov = (int64_t)(uint16_t)d.cov[col];
            int32_t i = 0;
            while (i < ne) {
                Ent* N = d.ent + o + i;                                              // the node's entries, first-seen order
                int32_t len = 1;
                while (i + len < ne && N[len].d == N[0].d && N[len].b == N[0].b) len++;
                const int32_t b = N[0].b;
                int32_t best = 0, tmp = 0;
                int64_t p_pp = NEG_INF, p_pp_ = NEG_INF;
                for (int32_t m = 0; m < len; m++) if (N[m].link > tmp) tmp = N[m].link;
                for (int32_t m = 0; m < len; m++) {
                    Ent& em = N[m];
                    em.score = 0;
                    if (em.pp_t == -1) em.score = 10 * (int64_t)em.link - pen * cov;
                    else {
                        int32_t plen;
                        const int32_t pn = find_node(d, em.pp_t, em.pp_d, em.pp_b, &plen);
                        const bool at_cut = !sr.start && em.pp_t == sr.c;
                        for (int32_t n = 0; n < plen; n++) {
                            const Ent& en = d.ent[pn + n];
                            if (!(en.pp_t == em.ppp_t && en.pp_d == em.ppp_d && en.pp_b == em.ppp_b)) continue;
                            const int64_t ns_ = at_cut ? base : en.score;
                            const int64_t s = ns_ + 10 * (int64_t)em.link - pen * cov;
                            if (cmp.gt(s, em.score)) { em.score = s; p_pp_ = ns_; }
                            if (rt == READS_CLR || rt == READS_HIFI) {                // :1958-1963, :2023-2028
                                if (cmp.gt(ns_, p_pp) || (cmp.eq(ns_, p_pp) && em.pp_b != 4)) { best = m; p_pp = ns_; }
                            } else if (rt != READS_RS) {                              // ONT :2086-2094
                                if (((em.ppp_d > 1 || em.pp_d > 0) && ((double)em.link > (double)cov * 0.2 || (int32_t)em.link > tmp / 2)) ||
                                    ((int32_t)em.link > (int32_t)N[best].link / 2 && cmp.gt(ns_, p_pp) &&
                                     (em.pp_b == 4 || em.pp_b == b || em.ppp_b == b || em.pp_b == em.ppp_b))) { best = m; p_pp = ns_; }
                            }
                        }
                    }
                    if (rt == READS_RS) { if (!cmp.gt(N[best].score, em.score)) { best = m; p_pp = p_pp_; } }       // :1918-1921
                    else if (cmp.gt(em.score, N[best].score) || (cmp.eq(em.score, N[best].score) && em.pp_b != 4)) { best = m; p_pp = p_pp_; }
                }
                (void)p_pp;
                N[0].aux = best;
                if (col == sr.cend && !cmp.gt(gbs, N[best].score)) {                  // :1923-1930 (last column only)
                    gb_c = col; gb_d = N[0].d; gb_b = b;
                    if (cmp.gt(N[best].score, gbs)) gbs = N[best].score;
                }
                i += len;
            }
        }
        d.seg_T[k] = exact ? NEG_INF : cmp.T;
        d.seg_mode[k] = exact ? MODE_EXACT : MODE_SPEC;
        if (sr.hi == sr.cend) { d.win_gb[sr.w * 3] = gb_c; d.win_gb[sr.w * 3 + 1] = gb_d; d.win_gb[sr.w * 3 + 2] = gb_b; }
    }
};

// ---- the same chain with a warp per segment --------------------------------------------------------------------------------
// What is slow in the chain is not arithmetic but the look-ups: for every entry, find the node of its predecessor tag and
// the entries of that node that continue the same pre-predecessor (dependent loads from L2).  The look-ups of all entries
// that share a sub-column are independent of each other, so the lanes of a warp do them at once (gather: the matching
// predecessor scores, in order, into the warp's shared scratch); lane 0 then replays the reference's sequential rules over
// the gathered values, which costs no memory latency.  Groups that do not fit the scratch (more than GMAX entries, or an
// entry with more than MAXM matches) are done by lane 0 the literal way.
#ifndef NP2_GMAX
#define NP2_GMAX 64
#endif
#ifndef NP2_MAXM
#define NP2_MAXM 6
#endif
enum { GMAX = NP2_GMAX, MAXM = NP2_MAXM };
struct Gath { int32_t cnt; int32_t pad; int64_t v[MAXM]; };          // cnt < 0: more matches than MAXM

NP2_HD void gather_entry(const Dev& d, const SegRange& sr, int64_t base, const Ent& em, Gath& g) {
    g.cnt = 0;
    if (em.pp_t == -1) return;
    int32_t plen;
    const int32_t pn = find_node(d, em.pp_t, em.pp_d, em.pp_b, &plen);
    const bool at_cut = !sr.start && em.pp_t == sr.c;
    for (int32_t n = 0; n < plen; n++) {
        const Ent& en = d.ent[pn + n];
        if (!(en.pp_t == em.ppp_t && en.pp_d == em.ppp_d && en.pp_b == em.ppp_b)) continue;
        if (g.cnt == MAXM) { g.cnt = -1; return; }
        g.v[g.cnt++] = at_cut ? base : en.score;
    }
}

struct NodeState { int32_t best, tmp, b, rt; int64_t p_pp, p_pp_, pen, cov; };
// one matching predecessor score ns_ for entry m of node N (the body of the reference's inner loop)
NP2_HD void chain_step(Cmp& cmp, NodeState& st, Ent* N, int32_t m, int64_t ns_) {
    Ent& em = N[m];
    const int64_t s = ns_ + 10 * (int64_t)em.link - st.pen * st.cov;
    if (cmp.gt(s, em.score)) { em.score = s; st.p_pp_ = ns_; }
    if (st.rt == READS_CLR || st.rt == READS_HIFI) {
        if (cmp.gt(ns_, st.p_pp) || (cmp.eq(ns_, st.p_pp) && em.pp_b != 4)) { st.best = m; st.p_pp = ns_; }
    } else if (st.rt != READS_RS) {
        if (((em.ppp_d > 1 || em.pp_d > 0) && ((double)em.link > (double)st.cov * 0.2 || (int32_t)em.link > st.tmp / 2)) ||
            ((int32_t)em.link > (int32_t)N[st.best].link / 2 && cmp.gt(ns_, st.p_pp) &&
             (em.pp_b == 4 || em.pp_b == st.b || em.ppp_b == st.b || em.pp_b == em.ppp_b))) { st.best = m; st.p_pp = ns_; }
    }
}

// W: lane(), lanes(), sync(), scratch() -> Gath[GMAX] private to the warp (one lane and a local array in the test build)
struct ChainWarp {
    Dev d; int32_t rerun_only;
    template <class W> NP2_HD void operator()(int64_t kk, W& w) const {
        const int32_t ns = d.blk_idx[d.n_blk], k = (int32_t)kk;
        if (k >= ns) return;
        const int32_t mode = d.seg_mode[k];
        if (rerun_only && mode != MODE_RERUN) return;
        const SegRange sr = seg_range(d, k, ns);
        const bool exact = sr.start || mode == MODE_RERUN;
        const int64_t base = sr.start ? 0 : (mode == MODE_RERUN ? d.seg_S[k] : BIG);
        Cmp cmp; cmp.T = NEG_INF; cmp.spec = !exact;
        NodeState st;
        st.pen = d.read_type == READS_HIFI ? 4 : 3; st.rt = d.read_type;
        int64_t gbs = NEG_INF; int32_t gb_c = -1, gb_d = 0, gb_b = 0;
        Gath* G = w.scratch();
        for (int32_t col = sr.start ? sr.c : sr.c + 1; col <= sr.hi; col++) {
            const int32_t o = d.rec_off[col], ne = d.n_ent[col];
            st.cov = (int64_t)(uint16_t)d.cov[col];
            int32_t g0 = 0;
            while (g0 < ne) {                                                        // group: the entries of one sub-column
                int32_t glen = 1;
                while (g0 + glen < ne && d.ent[o + g0 + glen].d == d.ent[o + g0].d) glen++;
                const bool fits = glen <= GMAX;
                if (fits) for (int32_t e = w.lane(); e < glen; e += w.lanes()) gather_entry(d, sr, base, d.ent[o + g0 + e], G[e]);
                w.sync();
                if (w.lane() == 0) {
                    int32_t i = g0;
                    while (i < g0 + glen) {
                        Ent* N = d.ent + o + i;
                        int32_t len = 1;
                        while (i + len < g0 + glen && N[len].b == N[0].b) len++;
                        st.b = N[0].b; st.best = 0; st.tmp = 0; st.p_pp = NEG_INF; st.p_pp_ = NEG_INF;
                        for (int32_t m = 0; m < len; m++) if (N[m].link > st.tmp) st.tmp = N[m].link;
                        for (int32_t m = 0; m < len; m++) {
                            Ent& em = N[m];
                            em.score = 0;
                            if (em.pp_t == -1) em.score = 10 * (int64_t)em.link - st.pen * st.cov;
                            else if (fits && G[i - g0 + m].cnt >= 0) {
                                const Gath& g = G[i - g0 + m];
                                for (int32_t j = 0; j < g.cnt; j++) chain_step(cmp, st, N, m, g.v[j]);
                            } else {                                                 // the literal look-up
                                int32_t plen;
                                const int32_t pn = find_node(d, em.pp_t, em.pp_d, em.pp_b, &plen);
                                const bool at_cut = !sr.start && em.pp_t == sr.c;
                                for (int32_t n = 0; n < plen; n++) {
                                    const Ent& en = d.ent[pn + n];
                                    if (!(en.pp_t == em.ppp_t && en.pp_d == em.ppp_d && en.pp_b == em.ppp_b)) continue;
                                    chain_step(cmp, st, N, m, at_cut ? base : en.score);
                                }
                            }
                            if (st.rt == READS_RS) { if (!cmp.gt(N[st.best].score, em.score)) { st.best = m; st.p_pp = st.p_pp_; } }
                            else if (cmp.gt(em.score, N[st.best].score) || (cmp.eq(em.score, N[st.best].score) && em.pp_b != 4)) { st.best = m; st.p_pp = st.p_pp_; }
                        }
                        N[0].aux = st.best;
                        if (col == sr.cend && !cmp.gt(gbs, N[st.best].score)) {
                            gb_c = col; gb_d = N[0].d; gb_b = st.b;
                            if (cmp.gt(N[st.best].score, gbs)) gbs = N[st.best].score;
                        }
                        i += len;
                    }
                }
                w.sync();
                g0 += glen;
            }
        }
        if (w.lane() == 0) {
            d.seg_T[k] = exact ? NEG_INF : cmp.T;
            d.seg_mode[k] = exact ? MODE_EXACT : MODE_SPEC;
            if (sr.hi == sr.cend) { d.win_gb[sr.w * 3] = gb_c; d.win_gb[sr.w * 3 + 1] = gb_d; d.win_gb[sr.w * 3 + 2] = gb_b; }
        }
    }
};

struct Stitch {                     // true cut scores of one window, in order; marks the segments to run again
    Dev d;
    template <class Ops> NP2_HD void operator()(int64_t w, Ops& ops) const {
        const int32_t k0 = d.blk_idx[d.win_blk0[w]], k1 = d.blk_idx[d.win_blk0[w + 1]];
        int32_t bad = 0;
        for (int32_t k = k0 + 1; k < k1; k++) {
            const int64_t v = d.ent[d.rec_off[d.seg_col[k]]].score;                  // the cut's only entry, written by segment k - 1
            const int32_t km = k - 1;
            int64_t S;
            if (km == k0 || d.seg_mode[km] == MODE_EXACT) S = v;                     // absolute
            else S = v >= BIG / 2 ? d.seg_abs[km] + (v - BIG) : v;
            d.seg_abs[k] = S;
            if (d.seg_mode[k] == MODE_EXACT) { if (d.seg_S[k] != S) { d.seg_S[k] = S; d.seg_mode[k] = MODE_RERUN; bad++; } }
            else if (d.seg_mode[k] == MODE_RERUN || S < d.seg_T[k]) { d.seg_S[k] = S; d.seg_mode[k] = MODE_RERUN; bad++; }
        }
        if (bad) ops.atomic_add(d.n_invalid, bad);
    }
};

struct Backtrack {                  // one segment's part of the path; fill == 0: count, fill == 1: write (forward order)
    Dev d; int32_t fill;
    template <class Ops> NP2_HD void operator()(int64_t kk, Ops& ops) const {
        const int32_t ns = d.blk_idx[d.n_blk], k = (int32_t)kk;
        if (k >= ns) return;
        if (fill && d.seg_n[k] == 0) return;
        const SegRange sr = seg_range(d, k, ns);
        int32_t ct, cd, cb;
        if (sr.hi == sr.cend) {
            ct = d.win_gb[sr.w * 3]; cd = d.win_gb[sr.w * 3 + 1]; cb = d.win_gb[sr.w * 3 + 2];
            if (ct < 0) { ops.atomic_or(d.err, ERR_LAST); if (!fill) { d.seg_n[k] = 0; d.seg_reach[k] = 0; } return; }
        } else {
            const Ent& x = d.ent[d.rec_off[sr.hi]];                                   // the next cut's only entry
            ct = x.pp_t; cd = x.pp_d; cb = x.pp_b;
        }
        int32_t n = 0, reach = 0;
        const int64_t end = (int64_t)d.seg_out[k] + d.seg_n[k];
        while (ct != -1) {
            int32_t len;
            const int32_t nd = find_node(d, ct, (uint16_t)cd, (uint8_t)cb, &len);
            if (len == 0) { ops.atomic_or(d.err, ERR_EMPTY); break; }
            const Ent& be = d.ent[nd + d.ent[nd].aux];
            if (cb != 4) {
                if (fill) {
                    const int64_t at = end - 1 - n;
                    const uint32_t cov = (uint16_t)d.cov[ct];
                    d.out_pos[at] = (uint32_t)(ct - sr.col0);
                    const uint8_t ch = int_to_base(cb);
                    d.out_base[at] = (int64_t)cov > (int64_t)d.min_cov ? ch : (uint8_t)(ch >= 'A' && ch <= 'Z' ? ch + 32 : ch);
                    d.out_qv[at] = cov ? (uint8_t)(100 * (uint32_t)be.link / cov) : 0;
                }
                n++;
            }
            if (!sr.start && ct == sr.c) { reach = 1; break; }                         // the segment's own cut: the rest belongs to the segment before
            ct = be.pp_t; cd = be.pp_d; cb = be.pp_b;
        }
        if (sr.start && ct == -1) reach = 1;
        if (!fill) { d.seg_n[k] = n; d.seg_reach[k] = reach; }
    }
};

struct Prune {                      // segments before the one in which the path ends (a chain started by a read) emit nothing
    Dev d;
    template <class Ops> NP2_HD void operator()(int64_t w, Ops&) const {
        const int32_t k0 = d.blk_idx[d.win_blk0[w]], k1 = d.blk_idx[d.win_blk0[w + 1]];
        bool active = true;
        for (int32_t k = k1 - 1; k >= k0; k--) {
            if (!active) d.seg_n[k] = 0;
            else if (!d.seg_reach[k]) active = false;
        }
    }
};
struct WinOut {
    Dev d;
    template <class Ops> NP2_HD void operator()(int64_t w, Ops&) const {
        d.win_out[w] = w < d.n_win ? d.seg_out[d.blk_idx[d.win_blk0[w]]] : d.seg_out[d.blk_idx[d.n_blk]];
    }
};

// ---- host side: one batch of windows (host pointers), through a backend that owns the device (or emulated) memory -------
struct Batch {
    int32_t n_win; const int32_t* win_len; const int32_t* win_aln0;                  // alignments of window w: [win_aln0[w], win_aln0[w + 1])
    int32_t read_type, min_cov;
    const uint32_t* aln_t_s; const uint32_t* aln_len; const uint64_t* str_off; const char* t_str; const char* q_str; int64_t str_bytes;
};
struct Stats { int32_t n_seg, reruns, iterations; int64_t n_rec; };

// Returns the number of consensus bases (all windows), or a negative code: -1 cap, -2 no node in a window's last column,
// -3 an alignment outside its window / starting on a gap column / empty, -4 backtrack through a node without entries,
// -5 size limits (columns, records or a sub-column run beyond the index types), -6 backend failure.
template <class BE>
int64_t run_first_pass(BE& be, const Batch& hb, uint32_t* out_pos, uint8_t* out_base, uint8_t* out_qv, int64_t cap, int64_t* out_off, Stats* st) {
    if (hb.n_win < 1) { if (out_off) out_off[0] = 0; return 0; }
    if (hb.win_aln0[0] != 0) return -3;
    for (int32_t w = 0; w < hb.n_win; w++) if (hb.win_aln0[w + 1] < hb.win_aln0[w]) return -3;
    const int32_t n_aln = hb.win_aln0[hb.n_win];
    // host-side tables: column / block offsets of the windows, stretch offsets of the alignments
    int64_t ctot = 0, nblk = 0, nst = 0, total_cols = 0;
    int32_t* win_col0 = be.template host<int32_t>("h_win_col0", (size_t)hb.n_win + 1);
    int32_t* win_blk0 = be.template host<int32_t>("h_win_blk0", (size_t)hb.n_win + 1);
    int64_t* aln_st0 = be.template host<int64_t>("h_aln_st0", (size_t)n_aln + 1);
    for (int32_t w = 0; w < hb.n_win; w++) {
        if (hb.win_len[w] < 1) return -2;
        win_col0[w] = (int32_t)ctot; win_blk0[w] = (int32_t)nblk;
        ctot += hb.win_len[w]; nblk += (hb.win_len[w] + CUT_BLOCK - 1) / CUT_BLOCK;
        if (ctot > 0x7ffffff0LL) return -5;
    }
    win_col0[hb.n_win] = (int32_t)ctot; win_blk0[hb.n_win] = (int32_t)nblk;
    for (int32_t r = 0; r < n_aln; r++) {
        if (hb.aln_len[r] == 0) return -3;
        aln_st0[r] = nst; nst += ((int64_t)hb.aln_len[r] + STRETCH - 1) / STRETCH; total_cols += hb.aln_len[r];
        if (hb.str_off[r] + hb.aln_len[r] > (uint64_t)hb.str_bytes) return -3;
    }
    aln_st0[n_aln] = nst;
    if (total_cols > 0x7ffffff0LL || nst > 0x7ffffff0LL) return -5;

    Dev d;
    d.n_win = hb.n_win; d.n_aln = n_aln; d.Ctot = (int32_t)ctot; d.n_blk = (int32_t)nblk; d.n_stretch = nst;
    d.read_type = hb.read_type; d.min_cov = hb.min_cov;
    d.win_col0 = be.upload("win_col0", win_col0, (size_t)hb.n_win + 1);
    d.win_blk0 = be.upload("win_blk0", win_blk0, (size_t)hb.n_win + 1);
    d.win_aln0 = be.upload("win_aln0", hb.win_aln0, (size_t)hb.n_win + 1);
    d.aln_t_s = be.upload("aln_t_s", hb.aln_t_s, (size_t)n_aln);
    d.aln_len = be.upload("aln_len", hb.aln_len, (size_t)n_aln);
    d.str_off = be.upload("str_off", hb.str_off, (size_t)n_aln);
    d.aln_st0 = be.upload("aln_st0", aln_st0, (size_t)n_aln + 1);
    d.t_str = be.upload("t_str", hb.t_str, (size_t)hb.str_bytes);
    d.q_str = be.upload("q_str", hb.q_str, (size_t)hb.str_bytes);
    const size_t C1 = (size_t)ctot + 1, B1 = (size_t)nblk + 1, S1 = (size_t)nst + 1;
    d.st_nt = be.template buf<int32_t>("st_nt", S1); d.st_ex = be.template buf<int32_t>("st_ex", S1);
    d.cov = be.template buf<uint32_t>("cov", C1); d.msz = be.template buf<uint32_t>("msz", C1);
    d.cnt = be.template buf<int32_t>("cnt", C1); d.rec_off = be.template buf<int32_t>("rec_off", C1);
    d.fill = be.template buf<int32_t>("fill", C1); d.n_ent = be.template buf<int32_t>("n_ent", C1);
    d.rec = be.template buf<Rec>("rec", (size_t)total_cols + 1); d.ent = be.template buf<Ent>("ent", (size_t)total_cols + 1);
    d.blk_col = be.template buf<int32_t>("blk_col", B1); d.blk_has = be.template buf<int32_t>("blk_has", B1); d.blk_idx = be.template buf<int32_t>("blk_idx", B1);
    d.seg_col = be.template buf<int32_t>("seg_col", B1); d.seg_win = be.template buf<int32_t>("seg_win", B1); d.seg_mode = be.template buf<int32_t>("seg_mode", B1);
    d.seg_n = be.template buf<int32_t>("seg_n", B1); d.seg_out = be.template buf<int32_t>("seg_out", B1); d.seg_reach = be.template buf<int32_t>("seg_reach", B1);
    d.seg_S = be.template buf<int64_t>("seg_S", B1); d.seg_T = be.template buf<int64_t>("seg_T", B1); d.seg_abs = be.template buf<int64_t>("seg_abs", B1);
    d.win_gb = be.template buf<int32_t>("win_gb", (size_t)hb.n_win * 3);
    d.n_invalid = be.template buf<int32_t>("n_invalid", 4); d.err = (uint32_t*)(d.n_invalid + 1);
    d.win_out = be.template buf<int64_t>("win_out", (size_t)hb.n_win + 1);
    d.out_pos = nullptr; d.out_base = nullptr; d.out_qv = nullptr;
    be.zero(d.st_nt, S1 * 4); be.zero(d.cov, C1 * 4); be.zero(d.msz, C1 * 4); be.zero(d.cnt, C1 * 4); be.zero(d.fill, C1 * 4); be.zero(d.n_ent, C1 * 4);
    be.zero(d.blk_has, B1 * 4); be.zero(d.seg_n, B1 * 4); be.zero(d.n_invalid, 16);
    be.fill_ff(d.win_gb, (size_t)hb.n_win * 12);

    be.launch("lgs_stretch_count", nst, StretchCount{d});
    be.exscan_i32(d.st_nt, d.st_ex, (int64_t)S1);
    be.launch("lgs_tag_count", nst, TagWalk<false>{d});
    be.exscan_i32(d.cnt, d.rec_off, (int64_t)C1);
    be.launch("lgs_tag_emit", nst, TagWalk<true>{d});
    be.launch("lgs_column_links", ctot, ColumnLinks{d});
    be.launch("lgs_cut_blocks", nblk, CutBlocks{d});
    be.exscan_i32(d.blk_has, d.blk_idx, (int64_t)B1);
    be.launch("lgs_seg_scatter", nblk, SegScatter{d});
    const bool warp_chain = be.warp_chain();
    if (warp_chain) be.launch_warps("lgs_chain", nblk, ChainWarp{d, 0}); else be.launch("lgs_chain", nblk, Chain{d, 0});
    int32_t iterations = 0, reruns = 0;
    for (;;) {
        be.zero(d.n_invalid, 4);
        be.launch("lgs_stitch", hb.n_win, Stitch{d});
        const int32_t bad = be.read_i32(d.n_invalid);
        if (!be.good()) return -6;
        if (bad <= 0) break;
        reruns += bad;
        if (++iterations > d.n_blk + 2) return -6;                                  // cannot happen: every pass fixes at least one segment
        if (warp_chain) be.launch_warps("lgs_chain_rerun", nblk, ChainWarp{d, 1}); else be.launch("lgs_chain_rerun", nblk, Chain{d, 1});
    }
    be.launch("lgs_backtrack_count", nblk, Backtrack{d, 0});
    be.launch("lgs_prune", hb.n_win, Prune{d});
    be.exscan_i32(d.seg_n, d.seg_out, (int64_t)B1);
    be.launch("lgs_win_out", (int64_t)hb.n_win + 1, WinOut{d});
    int32_t tail[2];
    be.download(tail, d.n_invalid, 8);                                               // [1] = error bits
    const int32_t n_seg = be.read_i32(d.blk_idx + d.n_blk);
    if (!be.good()) return -6;
    const uint32_t err = (uint32_t)tail[1];
    if (err & ERR_RANGE) return -3;
    if (err & ERR_LIMIT) return -5;
    if (err & ERR_LAST) return -2;
    if (err & ERR_EMPTY) return -4;
    be.download(out_off, d.win_out, ((size_t)hb.n_win + 1) * 8);
    if (!be.good()) return -6;
    const int64_t total = out_off[hb.n_win];
    if (st) { st->n_seg = n_seg; st->reruns = reruns; st->iterations = iterations; st->n_rec = be.read_i32(d.rec_off + d.Ctot); }
    if (total > cap) return -1;
    d.out_pos = be.template buf<uint32_t>("out_pos", (size_t)total + 1);
    d.out_base = be.template buf<uint8_t>("out_base", (size_t)total + 1);
    d.out_qv = be.template buf<uint8_t>("out_qv", (size_t)total + 1);
    be.launch("lgs_backtrack_fill", nblk, Backtrack{d, 1});
    if (total) {
        be.download(out_pos, d.out_pos, (size_t)total * 4);
        be.download(out_base, d.out_base, (size_t)total);
        if (out_qv) be.download(out_qv, d.out_qv, (size_t)total);
    }
    if (!be.good()) return -6;
    tail[1] = 0;
    be.download(tail, d.n_invalid, 8);
    if ((uint32_t)tail[1] & ERR_EMPTY) return -4;
    return total;
}

}  // namespace np2
