// engine_impl.h — orchestration of the polishing step over a Backend (device allocation,
// per-item launches, scans).  The product instantiates it with the CUDA backend (engine.cu);
// tests/emu instantiates the same code with a loop backend to unit-test the kernel bodies on
// CPU-only machines.  Nothing here is a CPU fallback of the product.
//
// Task 1 (score_chain, scorechain.c:3-15 -> contig_score_correct, contig.c:706-734) as a
// sequence of data-parallel passes over one packed shard (all contigs at once):
//
//   read_prep    per read   : filter level, usable query interval, reference span, and the
//                             insertion-length maxima ins[p]            (contig.c:170-245,333-358,648-686)
//   scan         per pos    : colbase = exclusive_sum(1 + ins)          (column layout, contig.c:385-399)
//   col_init     per pos    : reference symbol / flags of every column  (contig.c:81-102,373-383)
//   sym_bound    per read   : upper bound of the read's column string
//   expand       per read   : the read's column string (one 4-bit symbol per consecutive
//                             column: base, or 3 for D / unfilled sub-columns) (contig.c:247-331)
//   pileup_scan  per column : votes[c] and "some read disagrees with the draft symbol"
//   table        per table column : first-seen-ordered 3-mer tallies      (base.c:60-71)
//   chain_dp     per stretch: forward score chain + backtrack over maximal runs of
//                             non-anchor columns                          (contig.c:424-496)
//   emit         per column : deletion skipping, lowercase flags, compaction (contig.c:736-799)
//
// Anchor decomposition: a column whose votes all carry the draft's own symbol has a
// single score entry, so every successor lookup resolves to it and the DP is invariant under a
// common offset of the incoming scores (exact: scores are dyadic rationals for rate = k/2^m).
// Only maximal runs of non-anchor columns (+ their right neighbour) need the k-mer tables and the
// sequential chain; all runs are independent.
#pragma once
#include <stdint.h>
#include <string.h>
#include "device_logic.h"

namespace npe {
using namespace npd;

// Device-visible state of one run. All pointers are device pointers of the backend.
struct Dev {
    // resident shard
    int32_t n_ctg; int64_t n_reads; int32_t G;
    const uint8_t* ctg_seq; const int32_t* ctg_goff; const int64_t* ctg_read_off;
    const uint32_t* rec_off; const uint8_t* rec; const uint32_t* qual_off; const uint8_t* qual;
    Params P; int task;
    // per read
    int32_t *r_ctg, *r_gpos, *r_qstart, *r_qend, *r_wend, *r_hend, *r_pm, *r_c0, *r_n, *r_symoff, *r_bound;
    uint8_t* r_level;
    uint32_t* sym;
    // per position
    int32_t *ins, *ncol, *colbase;
    // per column
    int32_t C;
    uint8_t *refsym, *cflag, *mism, *obase, *oflag, *need, *keep;
    int32_t *colpos, *tidx, *keepidx, *needi, *keepi;
    uint32_t* votes;
    // table columns
    int32_t T; int32_t *tcols, *tcap, *toff, *tnk;
    uint32_t* ktab;          // (kmer | count << 16) entries, first-seen order
    uint16_t* bpk;           // [T*16] winning k-mer per base code
    uint8_t*  amax;          // [T] argmax base of the column's score list
    // output
    uint8_t* out; int64_t* out_off;
    // optional change trace (contig_get_contig's PolishPoint list, contig.c:743-799)
    int32_t *pcnt, *poff; struct TracePoint* pts; int64_t* pts_off; int32_t n_pts;
    int32_t* err;            // device error word (0 = ok)
};

struct TracePoint { int32_t pos; int16_t index; char curbase, base; };   // PolishPoint, contig.h:10-15

enum { ERR_INS_OVERFLOW = 1, ERR_DEPTH = 2, ERR_MISSING_SCORE = 4, ERR_SYM_BOUND = 8 };

// ---------------------------------------------------------------------------------------------
struct ReadPrep {
    Dev d;
    template <class B> NP_HD void operator()(int64_t r, B& be) const {
        Rec rc = load_rec(d.rec, d.rec_off, r);
        int32_t k = find_contig_i64(d.ctg_read_off, d.n_ctg, r);
        int32_t gs = d.ctg_goff[k], ge = d.ctg_goff[k + 1] - 1;
        int lvl = filter_level(rc, d.task, d.P);
        int32_t qs, qe, wl, hl;
        cut_read(rc, d.P.trim_len_edge, &qs, &qe);
        ref_spans(rc, &wl, &hl);
        d.r_ctg[r] = k;
        d.r_gpos[r] = gs + rc.pos;
        d.r_qstart[r] = qs; d.r_qend[r] = qe;
        d.r_wend[r] = gs + rc.pos + wl;
        d.r_hend[r] = gs + rc.pos + (hl > 0 && !(rc.flag & 4u) ? hl : 1);
        d.r_level[r] = (uint8_t)lvl;
        if (d.task == 1 && lvl >= 1) {
            // contig_parse_read_insert (contig.c:202-245), region = whole contig
            int32_t pos = gs + rc.pos;
            for (int i = 0; i < rc.n_cigar; i++) {
                int op = cig_op(rc.cigar[i]); int32_t len = cig_len(rc.cigar[i]);
                if (op == OP_M || op == OP_D) pos += len;
                else if (op == OP_I && pos > gs && pos <= ge) be.atomic_max(&d.ins[pos - 1], len);
            }
        }
    }
};

struct NcolFromIns {   // ncol[p] = 1 + ins[p]; ncol[G] = 0 (scan pad)
    Dev d;
    template <class B> NP_HD void operator()(int64_t p, B&) const { d.ncol[p] = p < d.G ? 1 + d.ins[p] : 0; }
};

struct ColInit {
    Dev d;
    template <class B> NP_HD void operator()(int64_t p, B&) const {
        int32_t c = d.colbase[p], n = d.colbase[p + 1] - c;
        uint32_t ch = d.ctg_seq[p];
        uint8_t fl = 0;
        if (ch >= 97 && ch <= 122) { ch -= 32; fl = FLAG_ZERO; }   // contig.c:94-97
        d.refsym[c] = (uint8_t)base_code(ch);
        d.cflag[c] = fl; d.colpos[c] = (int32_t)p;
        for (int32_t j = 1; j < n; j++) { d.refsym[c + j] = SYM_GAP; d.cflag[c + j] = fl; d.colpos[c + j] = (int32_t)p; }
    }
};
struct ColEnds {   // mark first / last column of every contig
    Dev d;
    template <class B> NP_HD void operator()(int64_t k, B&) const {
        int32_t gs = d.ctg_goff[k], ge = d.ctg_goff[k + 1] - 1;
        if (ge < gs) return;
        d.cflag[d.colbase[gs]] |= CF_FIRST;
        d.cflag[d.colbase[ge]] |= CF_LAST;
    }
};

struct SymBound {  // words needed for the read's column string (+1 pad word)
    Dev d; const uint8_t* need = nullptr;   // optional: only reads with need[r] != 0
    template <class B> NP_HD void operator()(int64_t r, B&) const {
        int32_t words = 0;
        if (r < d.n_reads && d.r_level[r] >= 1 && (!need || need[r])) {
            int32_t k = d.r_ctg[r];
            int32_t gs = d.ctg_goff[k], ge = d.ctg_goff[k + 1] - 1;
            int32_t a = d.r_gpos[r], b = d.r_wend[r] - 1;
            if (a < gs) a = gs;
            if (b > ge) b = ge;
            if (b >= a) {
                int32_t ncolumns = d.colbase[b + 1] - d.colbase[a];
                words = (ncolumns + 7) / 8 + 1;
            }
        }
        d.r_bound[r] = words;
    }
};

struct ExpandVisitor {
    uint32_t* w; int32_t cap_syms; int32_t c0, n; int32_t* err;
    NP_HD void sym(int32_t col, uint32_t s, int32_t, bool) {
        if (n == 0) c0 = col;
        int32_t i = col - c0;
        if (i != n || i >= cap_syms) { *err |= ERR_SYM_BOUND; return; }
        w[i >> 3] |= s << ((i & 7) << 2);
        n++;
    }
    NP_HD void overflow() { *err |= ERR_INS_OVERFLOW; }
};
struct Expand {
    Dev d; int level_eq;   // expand reads whose level == level_eq (task 1: 1)
    template <class B> NP_HD void operator()(int64_t r, B&) const {
        d.r_c0[r] = 0; d.r_n[r] = 0;
        int32_t words = d.r_symoff[r + 1] - d.r_symoff[r];
        if (words == 0 || d.r_level[r] != level_eq) return;
        uint32_t* w = d.sym + d.r_symoff[r];
        for (int32_t i = 0; i < words; i++) w[i] = 0;
        Rec rc = load_rec(d.rec, d.rec_off, r);
        int32_t k = d.r_ctg[r];
        int32_t gs = d.ctg_goff[k], ge = d.ctg_goff[k + 1] - 1;
        ExpandVisitor v{w, (words - 1) * 8, 0, 0, d.err};
        walk_read(rc, gs, gs, ge, d.r_qstart[r], d.r_qend[r], d.colbase, v);
        d.r_c0[r] = v.c0; d.r_n[r] = v.n;
    }
};

// candidate reads of reference position p: [lo,hi) = first read whose prefix-max walk end
// exceeds p .. first read starting after p
NP_HD void cand_range(const Dev& d, int32_t p, int64_t* lo, int64_t* hi) {
    *hi = upper_bound_i32(d.r_gpos, 0, d.n_reads, p);
    *lo = upper_bound_i32(d.r_pm, 0, *hi, p);
}

struct PileupScan {   // per column: votes and draft-disagreement
    Dev d;
    template <class B> NP_HD void operator()(int64_t c, B&) const {
        int32_t p = d.colpos[c];
        uint32_t ref = d.refsym[c];
        int64_t lo, hi; cand_range(d, p, &lo, &hi);
        uint32_t cnt = 1, mis = 0;
        for (int64_t r = lo; r < hi; r++) {
            int32_t i = (int32_t)c - d.r_c0[r];
            if (i < 0 || i >= d.r_n[r]) continue;
            uint32_t s = sym_get(d.sym + d.r_symoff[r], i);
            cnt++;
            mis |= (uint32_t)(s != ref);
        }
        d.votes[c] = cnt;
        d.mism[c] = (uint8_t)mis;
        if (cnt >= 65535u) *d.err |= ERR_DEPTH;   // uint16 counters of the reference would wrap (base.h:28-31,45)
    }
};

struct NeedTable {   // table columns: disagreeing columns and their right neighbours
    Dev d; int all = 0;   // all != 0: every column (strictly sequential chain over whole contigs)
    template <class B> NP_HD void operator()(int64_t c, B&) const {
        int32_t nd = 0;
        if (c < d.C) {
            nd = all ? 1 : d.mism[c];
            if (!nd && !(d.cflag[c] & CF_FIRST) && c > 0) nd = d.mism[c - 1];
        }
        d.needi[c] = nd;
    }
};
struct TableCols {
    Dev d;
    template <class B> NP_HD void operator()(int64_t c, B&) const {
        if (d.needi[c]) { int32_t t = d.tidx[c]; d.tcols[t] = (int32_t)c; d.tcap[t] = (int32_t)d.votes[c]; }
        if (c == 0) d.tcap[d.T] = 0;
    }
};

NP_HD uint32_t ref_kmer(const Dev& d, int32_t c) {   // contig_as_read, contig.c:373-383
    uint32_t k = d.refsym[c];
    if (!(d.cflag[c] & CF_FIRST)) {
        k |= (uint32_t)d.refsym[c - 1] << 4;
        if (!(d.cflag[c - 1] & CF_FIRST)) k |= (uint32_t)d.refsym[c - 2] << 8;
    }
    return k;
}

struct BuildTable {   // first-seen-ordered k-mer tally of one column (base.c:60-71)
    Dev d;
    template <class B> NP_HD void operator()(int64_t t, B&) const {
        int32_t c = d.tcols[t];
        uint32_t* tab = d.ktab + d.toff[t];
        int32_t nk = 0;
        tab[nk++] = ref_kmer(d, c) | (1u << 16);
        int64_t lo, hi; cand_range(d, d.colpos[c], &lo, &hi);
        for (int64_t r = lo; r < hi; r++) {
            int32_t i = c - d.r_c0[r];
            if (i < 0 || i >= d.r_n[r]) continue;
            const uint32_t* w = d.sym + d.r_symoff[r];
            uint32_t k = sym_get(w, i);
            if (i >= 1) k |= sym_get(w, i - 1) << 4;
            if (i >= 2) k |= sym_get(w, i - 2) << 8;
            int32_t j = 0;
            for (; j < nk; j++) if ((tab[j] & 0xffffu) == k) { tab[j] += 1u << 16; break; }
            if (j == nk) tab[nk++] = k | (1u << 16);
        }
        d.tnk[t] = nk;
    }
};

struct ChainDP {   // one thread per stretch of consecutive table columns (contig.c:424-496)
    Dev d;
    template <class B> NP_HD void operator()(int64_t t0, B&) const {
        int32_t c0 = d.tcols[t0];
        bool start = t0 == 0 || d.tcols[t0 - 1] != c0 - 1 || (d.cflag[c0] & CF_FIRST);
        if (!start) return;
        const double rate = d.P.rate;
        double sp[16]; bool hp[16]; double spmax = 0;   // previous column's scores by base code
        bool zero_prev = true;                          // predecessor resolves every lookup to 0
        for (int b = 0; b < 16; b++) { sp[b] = 0; hp[b] = false; }
        int64_t t = t0;
        for (;; t++) {
            int32_t c = d.tcols[t];
            const uint32_t* tab = d.ktab + d.toff[t];
            int32_t nk = d.tnk[t];
            uint32_t total = d.votes[c], refk = tab[0] & 0xffffu;
            uint32_t tot = total > 1 ? total - 1 : total;
            double sc[16]; bool hc[16]; uint16_t kc[16]; uint8_t order[16]; int no = 0;
            for (int b = 0; b < 16; b++) { hc[b] = false; sc[b] = 0; kc[b] = 0; }
            for (int32_t j = 0; j < nk; j++) {
                uint32_t k = tab[j] & 0xffffu, cnt = tab[j] >> 16;
                uint32_t pv = (k >> 4) & 0xfu;
                double s;
                if (zero_prev) s = 0;
                else if (pv == 0) s = spmax;
                else { if (!hp[pv]) { *d.err |= ERR_MISSING_SCORE; } s = sp[pv]; }
                if (k == refk && total > 1) cnt--;
                s = s + ((double)cnt - (double)tot * rate);            // contig.c:448
                uint32_t b = k & 0xfu;
                if (!hc[b]) { hc[b] = true; sc[b] = s; kc[b] = (uint16_t)k; order[no++] = (uint8_t)b; }
                else if (sc[b] < s) { sc[b] = s; kc[b] = (uint16_t)k; }   // contig.c:450
            }
            // base_max_score: first strictly greater in first-seen order (base.c:185-197)
            int am = order[0]; double mx = sc[am];
            for (int q = 1; q < no; q++) if (sc[order[q]] > mx) { mx = sc[order[q]]; am = order[q]; }
            d.amax[t] = (uint8_t)am;
            for (int b = 0; b < 16; b++) { d.bpk[t * 16 + b] = kc[b]; sp[b] = sc[b]; hp[b] = hc[b]; }
            spmax = mx; zero_prev = false;
            bool last = (d.cflag[c] & CF_LAST) || t + 1 >= d.T || d.tcols[t + 1] != c + 1;
            if (last) break;
        }
        // backtrack (contig.c:473-496)
        uint32_t chosen = d.amax[t];
        for (;; t--) {
            int32_t c = d.tcols[t];
            const uint32_t* tab = d.ktab + d.toff[t];
            int32_t nk = d.tnk[t];
            uint32_t total = d.votes[c], support = 0;
            for (int32_t j = 0; j < nk; j++) if ((tab[j] & 0xfu) == chosen) support += tab[j] >> 16;
            uint8_t fl = d.cflag[c] & (uint8_t)(CF_FIRST | CF_LAST);
            if (total == 1) fl |= FLAG_ZERO;
            if (support / (double)total < d.P.min_count_ratio_skip) fl |= FLAG_COVERAGE;   // base.c:79-89
            d.obase[c] = (uint8_t)chosen; d.oflag[c] = fl;
            if (t == t0) break;
            uint32_t k = d.bpk[t * 16 + chosen], pv = (k >> 4) & 0xfu;
            chosen = (pv == 0) ? d.amax[t - 1] : pv;
        }
    }
};

struct AnchorCols {   // columns outside every stretch keep the draft symbol
    Dev d;
    template <class B> NP_HD void operator()(int64_t c, B&) const {
        if (d.needi[c]) return;
        uint8_t fl = d.cflag[c] & (uint8_t)(CF_FIRST | CF_LAST);
        if (d.votes[c] == 1) fl |= FLAG_ZERO;
        if (1.0 < d.P.min_count_ratio_skip) fl |= FLAG_COVERAGE;
        d.obase[c] = d.refsym[c]; d.oflag[c] = fl;
    }
};

struct KeepFlag {
    Dev d;
    template <class B> NP_HD void operator()(int64_t c, B&) const { d.keepi[c] = c < d.C ? (d.obase[c] != SYM_GAP) : 0; }
};
struct Emit {   // contig_get_contig, contig.c:748-786
    Dev d; uint8_t lowmask;
    template <class B> NP_HD void operator()(int64_t c, B&) const {
        if (d.obase[c] == SYM_GAP) return;
        bool low = (d.oflag[c] & lowmask) != 0;
        // `sign` carry: flagged deleted columns since the previous emitted column of this contig
        for (int64_t q = c - 1; !low && q >= 0 && !(d.oflag[q + 1] & CF_FIRST) && d.obase[q] == SYM_GAP; q--)
            if (d.oflag[q] & lowmask) low = true;
        uint8_t ch = code_char(d.obase[c]);
        if (low) ch += 32;
        d.out[d.keepidx[c]] = ch;
    }
};
struct OutOffsets {
    Dev d;
    template <class B> NP_HD void operator()(int64_t k, B&) const {
        if (k == d.n_ctg) { d.out_off[k] = d.keepidx[d.C]; return; }
        int32_t gs = d.ctg_goff[k];
        d.out_off[k] = gs < d.G ? d.keepidx[d.colbase[gs]] : d.keepidx[d.C];
    }
};

// ---- PolishPoint trace (contig.c:743-799), per draft position: a deleted main column, an emitted base that differs
//      from the (upper-cased) draft character, and every emitted insertion sub-column are change points
struct TraceCount {
    Dev d; int fill;
    template <class B> NP_HD void operator()(int64_t p, B&) const {
        if (p >= d.G) { if (!fill) d.pcnt[p] = 0; return; }
        const int32_t c0 = d.colbase[p], n = d.colbase[p + 1] - c0;
        uint32_t raw = d.ctg_seq[p];
        if (raw >= 97 && raw <= 122) raw -= 32;
        int32_t k = 0;
        TracePoint* out = nullptr; int32_t rel = 0;
        if (fill) {
            out = d.pts + d.poff[p];
            rel = (int32_t)p - d.ctg_goff[find_contig_i32(d.ctg_goff, d.n_ctg, (int32_t)p)];
        }
        for (int32_t j = 0; j < n; j++) {
            const uint32_t b = d.obase[c0 + j];
            const char ch = (char)code_char(b);
            bool pt; char cur, base;
            if (b == SYM_GAP) { pt = j == 0; cur = '.'; base = (char)raw; }
            else if (j != 0) { pt = true; cur = ch; base = '.'; }
            else { pt = (uint32_t)(uint8_t)ch != raw; cur = ch; base = (char)raw; }
            if (!pt) continue;
            if (fill) out[k] = TracePoint{rel, (int16_t)j, cur, base};
            k++;
        }
        if (!fill) d.pcnt[p] = k;
    }
};
struct TraceOffsets {
    Dev d;
    template <class B> NP_HD void operator()(int64_t k, B&) const { d.pts_off[k] = k == d.n_ctg ? d.poff[d.G] : d.poff[d.ctg_goff[k]]; }
};

// after obase is final: builds the trace when Params.trace is set (every task path ends with this)
template <class BE>
void run_trace(BE& be, Dev& d) {
    d.n_pts = 0;
    if (!d.P.trace) return;
    const int32_t G = d.G;
    d.pcnt = be.template buf<int32_t>("pcnt", (size_t)G + 2);
    d.poff = be.template buf<int32_t>("poff", (size_t)G + 2);
    d.pts_off = be.template buf<int64_t>("pts_off", (size_t)d.n_ctg + 1);
    be.launch("trace_count", (int64_t)G + 1, TraceCount{d, 0});
    be.exscan_i32(d.pcnt, d.poff, (int64_t)G + 1);
    d.n_pts = be.read_i32(d.poff + G);
    d.pts = be.template buf<TracePoint>("pts", (size_t)d.n_pts + 1);
    if (d.n_pts > 0) be.launch("trace_fill", G, TraceCount{d, 1});
    be.launch("trace_offsets", (int64_t)d.n_ctg + 1, TraceOffsets{d});
}

}  // namespace npe

// =============================================================================================
// Orchestration (templated on the backend; see the header comment)
// =============================================================================================
namespace npe {

struct RunStats { int32_t C, T; int64_t sym_words, table_entries, out_bytes; };

// The anchor decomposition needs exact (order-independent) score sums: true when the indel balance
// factor is a dyadic rational k/1024 of moderate size.  For any other user-supplied rate the chain is run
// strictly left to right over whole contigs (every column gets a table, one thread per contig), which
// reproduces the reference's rounding sequence exactly.
inline bool rate_is_dyadic(double r) {
    double s = r * 1024.0;
    return r > -64.0 && r < 64.0 && s == (double)(long long)s;
}

template <class BE>
int run_score_chain(BE& be, Dev& d, RunStats* st, bool exact_sequential = false) {
    const int64_t R = d.n_reads; const int32_t G = d.G;
    d.task = 1;
    d.err = be.template buf<int32_t>("err", 1);
    be.zero(d.err, sizeof(int32_t));
    d.r_ctg = be.template buf<int32_t>("r_ctg", R + 1);
    d.r_gpos = be.template buf<int32_t>("r_gpos", R + 1);
    d.r_qstart = be.template buf<int32_t>("r_qstart", R + 1);
    d.r_qend = be.template buf<int32_t>("r_qend", R + 1);
    d.r_wend = be.template buf<int32_t>("r_wend", R + 1);
    d.r_hend = be.template buf<int32_t>("r_hend", R + 1);
    d.r_pm = be.template buf<int32_t>("r_pm", R + 1);
    d.r_c0 = be.template buf<int32_t>("r_c0", R + 1);
    d.r_n = be.template buf<int32_t>("r_n", R + 1);
    d.r_bound = be.template buf<int32_t>("r_bound", R + 1);
    d.r_symoff = be.template buf<int32_t>("r_symoff", R + 1);
    d.r_level = be.template buf<uint8_t>("r_level", R + 1);
    d.ins = be.template buf<int32_t>("ins", (size_t)G + 1);
    d.colbase = be.template buf<int32_t>("colbase", (size_t)G + 1);
    d.out_off = be.template buf<int64_t>("out_off", (size_t)d.n_ctg + 1);
    be.zero(d.ins, sizeof(int32_t) * ((size_t)G + 1));

    if (R > 0) {
        be.launch("read_prep", R, ReadPrep{d});
        be.inclmax_i32(d.r_wend, d.r_pm, R);
    }
    be.exscan_ncol(d.ins, d.colbase, (int64_t)G);
    d.C = be.read_i32(d.colbase + G);
    const int32_t C = d.C;
    d.refsym = be.template buf<uint8_t>("refsym", (size_t)C + 1);
    d.cflag = be.template buf<uint8_t>("cflag", (size_t)C + 1);
    d.mism = be.template buf<uint8_t>("mism", (size_t)C + 1);
    d.obase = be.template buf<uint8_t>("obase", (size_t)C + 1);
    d.oflag = be.template buf<uint8_t>("oflag", (size_t)C + 1);
    d.colpos = be.template buf<int32_t>("colpos", (size_t)C + 1);
    d.votes = be.template buf<uint32_t>("votes", (size_t)C + 1);
    d.needi = be.template buf<int32_t>("needi", (size_t)C + 1);
    d.tidx = be.template buf<int32_t>("tidx", (size_t)C + 1);
    d.keepidx = be.template buf<int32_t>("keepidx", (size_t)C + 1);
    if (G > 0) {
        be.launch("col_init", G, ColInit{d});
        be.launch("col_ends", d.n_ctg, ColEnds{d});
    }
    be.launch("sym_bound", R + 1, SymBound{d});
    be.exscan_i32(d.r_bound, d.r_symoff, R + 1);
    int32_t W = be.read_i32(d.r_symoff + R);
    d.sym = be.template buf<uint32_t>("sym", (size_t)W + 1);
    if (R > 0) be.launch("expand", R, Expand{d, 1});
    if (C > 0) be.launch("pileup_scan", C, PileupScan{d});
    be.launch("need_table", (int64_t)C + 1, NeedTable{d, exact_sequential ? 1 : 0});
    be.exscan_i32(d.needi, d.tidx, (int64_t)C + 1);
    d.T = be.read_i32(d.tidx + C);
    int32_t E = 0;
    if (d.T > 0) {
        const int32_t T = d.T;
        d.tcols = be.template buf<int32_t>("tcols", (size_t)T + 1);
        d.tcap = be.template buf<int32_t>("tcap", (size_t)T + 1);
        d.toff = be.template buf<int32_t>("toff", (size_t)T + 1);
        d.tnk = be.template buf<int32_t>("tnk", (size_t)T + 1);
        d.bpk = be.template buf<uint16_t>("bpk", (size_t)T * 16);
        d.amax = be.template buf<uint8_t>("amax", (size_t)T + 1);
        be.launch("table_cols", C, TableCols{d});
        be.exscan_i32(d.tcap, d.toff, (int64_t)T + 1);
        E = be.read_i32(d.toff + T);
        d.ktab = be.template buf<uint32_t>("ktab", (size_t)E + 1);
        be.launch("build_table", T, BuildTable{d});
        be.launch("chain_dp", T, ChainDP{d});
    }
    if (C > 0) be.launch("anchor_cols", C, AnchorCols{d});
    be.exscan_keep(d.obase, d.keepidx, (int64_t)C);
    int32_t total = be.read_i32(d.keepidx + C);
    d.out = be.template buf<uint8_t>("out", (size_t)total + 1);
    if (C > 0) be.launch("emit", C, Emit{d, (uint8_t)(FLAG_ZERO | FLAG_COVERAGE)});
    be.launch("out_offsets", (int64_t)d.n_ctg + 1, OutOffsets{d});
    run_trace(be, d);
    int32_t err = be.read_i32(d.err);
    if (st) { st->C = C; st->T = d.T; st->sym_words = W; st->table_entries = E; st->out_bytes = total; }
    return err;
}

}  // namespace npe
