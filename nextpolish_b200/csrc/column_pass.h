// column_pass.h — the column half of the task-1 pileup scan as flat, full-width kernels.
//
// The streaming half (diff_pass.h) reduced every read to its column extent (+1/-1 coverage marks in cov[]), its sparse
// diff entries (pool) and one bit per column where some read disagrees with the draft (disb).  What is left is column
// work, and only ~13 % of the columns (those that disagree + their right neighbours: "table columns") need more than
// a prefix sum:
//
//   tile_agg    CTA per tile of TW draft positions: sum of the tile's coverage marks, number of its table columns
//   tile_scan   one CTA: exclusive prefix of both over the tiles (the tile's carry-in)
//   col_pass    CTA per tile, one thread per 8 positions: the tile's colbase and cov slices are staged into shared
//               memory by bulk copies (cp.async.bulk / TMA, mbarrier completion); block-wide exclusive scan (warp
//               shuffles) of per-thread sums, then every thread walks its columns once: votes = 1 + reads covering the
//               column, draft symbol nibbles (refw), table bitmap + rank (tblb, tblp); anchors (columns no read
//               disagrees with, and whose left neighbour nobody disagrees with) get their final base and flags here;
//               table columns get a table: slot 0 = the draft's own 3-mer (contig_as_read, contig.c:373-383)
//   votes       thread per diff entry: the read's 3-mer differs from the draft's at the entry's column and the two
//               after it — find-or-insert (atomicCAS) in those columns' tables, carrying the smallest read index per
//               slot; thread per read: the partial 3-mers of a read start (zeros for the missing symbols)
//   chain       thread per stretch (maximal run of table columns): per column order the slots by first voter =
//               first-seen (BAM) order (base.c:60-71), give the draft's 3-mer the remaining votes, score chain in
//               registers (contig.c:424-471), backtrack through a 9-byte-per-column trail (contig.c:473-496)
//
// Tables live in global memory (L2-resident: 64 B per table column) so that no phase is limited by what fits a CTA's
// shared memory: any depth, any stretch length.  A column with more than WK distinct 3-mers marks its stretch
// unresolved; those stretches are redone by the general kernels of engine_impl.h (same results, slower).
//
// Anchor decomposition: see DESIGN.md / engine_impl.h.  Everything here is integer work except the score chain, which
// is evaluated in double with separate multiply / subtract / add exactly like contig.c:448 (-fmad=false).
#pragma once
#include "diff_pass.h"

namespace npc {
using namespace npd;
using npe::Dev;
using npw::DiffEnt;
using npw::DiffGlobals;
using npw::ReadDesc;

#ifndef NP_SLICE_FAST_MAX
#define NP_SLICE_FAST_MAX 62     // columns of a thread's slice handled with 64-bit bit tricks (tests lower it to reach the general loops)
#endif
enum { TW = 2048,          // draft positions per tile
       TT = 256,           // threads per tile CTA
       PPT = TW / TT,      // positions per thread
       WK = 8,             // table capacity (distinct 3-mers per column)
       COV_CAP = TW + TW / 2 + 16 };   // columns of a tile staged in shared memory (more: read from global memory)

struct ColGlobals {
    int32_t n_tiles; const int32_t* tile_off;        // [n_ctg+1] first tile of every contig
    const int32_t* cov; const uint32_t* disb;         // written by the diff pass
    int32_t *tile_cov, *tile_tbl, *tile_str;          // [n_tiles+1] aggregates (coverage marks, table columns, stretch starts), then
                                                      // (tile_scan) exclusive prefixes; [n_tiles] = totals
    uint32_t *refw, *tblb, *tblp;                     // [C/8+2] draft symbols (little-endian nibbles); [C/32+2] table bitmap, rank before the word
    // tables (entry-major: slot j of table t at [j * T + t])
    int32_t T;
    int32_t* tcol; uint32_t* tvotes; uint8_t *tflag, *tbad, *tunres;
    uint32_t *te, *tfs;                               // kmer | count << 16 (0: empty slot); 0xffffffff - smallest read index of the slot's voters
                                                      // (zero-initialised by one memset; slot 0 = the draft's 3-mer, written by col_pass)
    uint32_t *bt_base, *bt_pv; uint16_t* bt_am;       // chain trail per table column: bases / previous-base nibbles, argmax | entries << 4 | slots << 8
    int32_t* sstart; const int32_t* n_start;          // [T] first table of every stretch, their number (= tile_str[n_tiles])
    int32_t* n_unresolved;
};

NP_HD int32_t find_tile_contig(const int32_t* tile_off, int32_t n_ctg, int32_t w) { return find_contig_i32(tile_off, n_ctg, w); }

// position range, column range and contig of tile w
struct Tile { int32_t k, gs, ge, p0, p1, cbeg, cend, cgs; };   // cgs: the contig's first column
NP_HD Tile tile_of(const Dev& d, const ColGlobals& g, int32_t w) {
    Tile t;
    t.k = find_tile_contig(g.tile_off, d.n_ctg, w);
    t.gs = d.ctg_goff[t.k]; t.ge = d.ctg_goff[t.k + 1] - 1;
    t.p0 = t.gs + (w - g.tile_off[t.k]) * TW;
    t.p1 = t.p0 + TW; if (t.p1 > t.ge + 1) t.p1 = t.ge + 1;
    t.cbeg = d.colbase[t.p0]; t.cend = d.colbase[t.p1]; t.cgs = d.colbase[t.gs];
    return t;
}
NP_HD uint32_t get_bit(const uint32_t* b, int32_t i) { return i < 0 ? 0u : (b[i >> 5] >> (i & 31)) & 1u; }
NP_HD int32_t popc32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
// 64 bits of a bitmap starting at bit i (i >= -2; bits before 0 read as 0)
NP_HD unsigned long long bits64(const uint32_t* b, int32_t i) {
    if (i < 0) return bits64(b, 0) << -i;
    const int32_t w = i >> 5, s = i & 31;
    const unsigned long long lo = (unsigned long long)b[w] | (unsigned long long)b[w + 1] << 32;
    return s == 0 ? lo : (lo >> s) | ((unsigned long long)b[w + 2] << (64 - s));
}
// table status of column c (not the first column of its contig when `first` is false): it disagrees, or its left
// neighbour (same contig) does
NP_HD bool col_is_table(const uint32_t* disb, int32_t c, bool first) { return get_bit(disb, c) || (!first && get_bit(disb, c - 1)); }

// ---- tile aggregates: thread per 8 positions, block sums (the caller reduces) ------------------------------------
// The per-thread slice: positions [pa, pb), columns [ca, cb)
struct Slice { int32_t pa, pb, ca, cb; };
NP_HD Slice slice_of(const Dev& d, const Tile& t, int32_t tid) {
    Slice s;
    s.pa = t.p0 + tid * PPT; if (s.pa > t.p1) s.pa = t.p1;
    s.pb = s.pa + PPT; if (s.pb > t.p1) s.pb = t.p1;
    s.ca = d.colbase[s.pa]; s.cb = d.colbase[s.pb];
    return s;
}
// table / stretch-start bits of the slice's first <= 62 columns (bit i = column ca + i).  A table column starts a
// stretch when it is its contig's first column or its left neighbour is not a table column.
NP_HD void slice_bits(const ColGlobals& g, const Tile& t, const Slice& s, unsigned long long& tb, unsigned long long& sb) {
    const unsigned long long b = bits64(g.disb, s.ca - 2);           // bit i = dis(ca - 2 + i)
    const bool cfirst = s.pa == t.gs;                                // column ca is the contig's first column
    unsigned long long left = b >> 1;                                // left neighbour disagrees
    if (cfirst) left &= ~1ull;
    tb = (b >> 2) | left;
    unsigned long long tprev = 0;                                    // table status of column ca - 1
    if (!cfirst) tprev = ((b >> 1) & 1ull) | ((b & 1ull) & (s.ca - 1 == t.cgs ? 0ull : 1ull));
    sb = tb & ~((tb << 1) | tprev);
}
struct Sums { int32_t cov, tbl, str; };
NP_HD Sums slice_sums(const Dev& d, const ColGlobals& g, const Tile& t, const Slice& s, const int32_t* cov) {
    Sums r{0, 0, 0};
    const int32_t n = s.cb - s.ca;
    if (n <= 0) return r;
    for (int32_t i = 0; i < n; i++) r.cov += cov[i];
    if (n <= NP_SLICE_FAST_MAX) {
        unsigned long long tb, sb;
        slice_bits(g, t, s, tb, sb);
        const unsigned long long m = (1ull << n) - 1ull;
        tb &= m; sb &= m;
        r.tbl = popc32((uint32_t)tb) + popc32((uint32_t)(tb >> 32));
        r.str = popc32((uint32_t)sb) + popc32((uint32_t)(sb >> 32));
    } else {
        bool prev = s.pa != t.gs && col_is_table(g.disb, s.ca - 1, s.ca - 1 == t.cgs);
        for (int32_t i = 0; i < n; i++) {
            const bool first = i == 0 && s.pa == t.gs;
            const bool tab = col_is_table(g.disb, s.ca + i, first);
            if (tab) { r.tbl++; if (first || !prev) r.str++; }
            prev = tab;
        }
    }
    return r;
}

// ---- the column walk of one thread (after the block scan gave it its carries) -------------------------------------
// run: reads covering the column before ca; trank / srank: tables / stretch starts before ca
template <class B>
NP_HD void slice_walk(const Dev& d, const ColGlobals& g, const Tile& t, const Slice& s, const int32_t* cov, int32_t run, int32_t trank, int32_t srank, B& be) {
    const int32_t n = s.cb - s.ca;
    if (n <= 0) return;
    // the two symbols before the slice's first column (k-mer context), and how many columns of the contig precede it (0, 1, 2+)
    uint32_t prev1 = 0, prev2 = 0; int nctx = 0;
    if (s.pa > t.gs) {
        const int32_t p = s.pa, ins1 = d.colbase[p] - d.colbase[p - 1] - 1;
        prev1 = ins1 > 0 ? (uint32_t)SYM_GAP : npw::draft_sym(d, p - 1); nctx = 1;
        if (ins1 >= 2) { prev2 = SYM_GAP; nctx = 2; }
        else if (ins1 == 1) { prev2 = npw::draft_sym(d, p - 1); nctx = 2; }
        else if (p - 1 > t.gs) { const int32_t ins2 = d.colbase[p - 1] - d.colbase[p - 2] - 1; prev2 = ins2 > 0 ? (uint32_t)SYM_GAP : npw::draft_sym(d, p - 2); nctx = 2; }
    }
    const bool fast = n <= NP_SLICE_FAST_MAX;
    unsigned long long tb = 0, sb = 0;
    if (fast) slice_bits(g, t, s, tb, sb);
    bool prevtab = !fast && s.pa != t.gs && col_is_table(g.disb, s.ca - 1, s.ca - 1 == t.cgs);
    const uint8_t cov_flag = 1.0 < d.P.min_count_ratio_skip ? (uint8_t)FLAG_COVERAGE : (uint8_t)0;
    // the slice's colbase entries and draft bases go to registers first (independent loads, issued together)
    int32_t cbv[PPT + 1]; uint32_t chv[PPT];
    const int32_t np = s.pb - s.pa;
    #pragma unroll
    for (int u = 0; u <= PPT; u++) cbv[u] = u <= np ? d.colbase[s.pa + u] : 0;
    #pragma unroll
    for (int u = 0; u < PPT; u++) chv[u] = u < np ? d.ctg_seq[s.pa + u] : 0u;
    int32_t c = s.ca, i = 0;
    uint32_t rw = 0, tw = 0;                           // pending refw nibbles / table bits of the current word
    #pragma unroll
    for (int u = 0; u < PPT; u++) {
        if (u >= np) break;
        const int32_t p = s.pa + u;
        const int32_t ncol = cbv[u + 1] - cbv[u];
        uint32_t ch = chv[u];
        if (ch >= 97u && ch <= 122u) ch -= 32u;
        const uint32_t psym = base_code(ch);
        for (int32_t j = 0; j < ncol; j++, c++, i++) {
            const uint32_t sym = j == 0 ? psym : (uint32_t)SYM_GAP;
            run += cov[i];
            const bool first = p == t.gs && j == 0, last = p == t.ge && j == 0;
            const bool tab = fast ? ((tb >> i) & 1ull) != 0ull : col_is_table(g.disb, c, first);
            if ((c & 31) == 0) { g.tblp[c >> 5] = (uint32_t)trank; }
            uint8_t fl = (uint8_t)((first ? CF_FIRST : 0) | (last ? CF_LAST : 0));
            if (run >= 65534) *d.err |= npe::ERR_DEPTH;              // uint16 counters of the reference would wrap (base.h:28-31,45)
            if (!tab) {
                if (run == 0) fl |= FLAG_ZERO;                       // only the draft's own vote
                d.obase[c] = (uint8_t)sym; d.oflag[c] = fl | cov_flag;
            } else {
                const int32_t tt = trank++;
                uint32_t k = sym;
                if (nctx >= 1) k |= prev1 << 4;
                if (nctx >= 2) k |= prev2 << 8;
                g.tcol[tt] = c; g.tvotes[tt] = 1u + (uint32_t)run; g.tflag[tt] = fl; g.te[tt] = k;
                tw |= 1u << (c & 31);
                if (fast ? ((sb >> i) & 1ull) != 0ull : (first || !prevtab)) g.sstart[srank++] = tt;
            }
            prevtab = tab;
            rw |= sym << ((c & 7) << 2);
            if ((c & 7) == 7) { be.atomic_or(&g.refw[c >> 3], rw); rw = 0; }
            if ((c & 31) == 31) { if (tw) be.atomic_or(&g.tblb[c >> 5], tw); tw = 0; }
            prev2 = prev1; prev1 = sym; if (nctx < 2) nctx++;
        }
    }
    if (rw) be.atomic_or(&g.refw[(c - 1) >> 3], rw);
    if (tw) be.atomic_or(&g.tblb[(c - 1) >> 5], tw);
}

// ---- votes -----------------------------------------------------------------------------------------------------
NP_HD int32_t table_of(const ColGlobals& g, int32_t c) {              // table index of column c, or -1
    const uint32_t w = g.tblb[c >> 5];
    if (!((w >> (c & 31)) & 1u)) return -1;
    return (int32_t)g.tblp[c >> 5] + popc32(w & ((1u << (c & 31)) - 1u));
}
template <class B>
NP_HD void tab_vote(const ColGlobals& g, int32_t t, uint32_t kmer, uint32_t ridx, B& be) {
    if (kmer == (g.te[t] & 0xffffu)) return;
    for (int j = 1; j < WK; j++) {
        uint32_t* e = &g.te[(size_t)j * g.T + t];
        const uint32_t old = be.atomic_cas_u32(e, 0u, kmer | (1u << 16));
        if (old == 0u) { be.atomic_max_u32(&g.tfs[(size_t)j * g.T + t], 0xffffffffu - ridx); return; }
        if ((old & 0xffffu) == kmer) { be.atomic_add_u32(e, 1u << 16); be.atomic_max_u32(&g.tfs[(size_t)j * g.T + t], 0xffffffffu - ridx); return; }
    }
    g.tbad[t] = 1;
}
NP_HD uint32_t refsym(const ColGlobals& g, int32_t c) { return (g.refw[c >> 3] >> ((c & 7) << 2)) & 0xfu; }

struct EntryVotes {     // per diff entry: the read's 3-mer differs from the draft's at the entry's column and the two after it
    Dev d; DiffGlobals dg; ColGlobals g;
    template <class B> NP_HD void operator()(int64_t i, B& be) const {
        const DiffEnt en = dg.pool[i];
        const uint32_t ridx = en.sr >> 4;
        const ReadDesc rd = dg.rdesc[ridx];
        const int64_t lo = rd.doff, hi = (int64_t)rd.doff + rd.dcnt;          // the read's entries (ascending columns)
        const int32_t cs = rd.cs, ce = rd.cs + rd.n, col = en.col;
        const int32_t nlc = i + 1 < hi ? dg.pool[i + 1].col : 0x7fffffff;      // a column a LATER entry also reaches is left to it
        // the read's symbols on columns col-2 .. col+2 (nibble j = column col-2+j): the draft's, overlaid with this entry
        // and the read's two previous entries (later entries lie beyond every column looked at here)
        const int32_t a = col - 2;
        uint32_t W;
        if (a >= 0) {
            const unsigned long long w64 = (unsigned long long)g.refw[a >> 3] | (unsigned long long)g.refw[(a >> 3) + 1] << 32;
            W = (uint32_t)(w64 >> (4 * (a & 7))) & 0xfffffu;
        } else W = (g.refw[0] << (4 * -a)) & 0xfffffu;
        W = (W & ~0xf00u) | (en.sr & 0xfu) << 8;
        if (i - 1 >= lo) {
            const DiffEnt e1 = dg.pool[i - 1];
            if (e1.col >= a) W = (W & ~(0xfu << (4 * (e1.col - a)))) | (e1.sr & 0xfu) << (4 * (e1.col - a));
            if (i - 2 >= lo) {
                const DiffEnt e2 = dg.pool[i - 2];
                if (e2.col >= a) W = (W & ~(0xfu << (4 * (e2.col - a)))) | (e2.sr & 0xfu) << (4 * (e2.col - a));
            }
        }
        for (int32_t k = 0; k < 3; k++) {
            const int32_t c = col + k;
            if (c >= ce || c >= nlc) break;
            if (c - cs < 2) continue;                                          // the first two symbols of a read: StartVotes
            const int32_t t = table_of(g, c);
            if (t < 0) continue;
            const uint32_t v = (W >> (4 * k)) & 0xfffu;                        // nibbles: s(c-2), s(c-1), s(c)
            tab_vote(g, t, (v & 0xfu) << 8 | (v & 0xf0u) | v >> 8, ridx, be);
        }
    }
};
struct StartVotes {     // per read: the first two symbols of a string cast partial 3-mers (zeros for the missing symbols)
    Dev d; DiffGlobals dg; ColGlobals g;
    template <class B> NP_HD void operator()(int64_t r, B& be) const {
        if (r >= d.n_reads) return;
        const ReadDesc rd = dg.rdesc[r];
        if (rd.cs < 0 || rd.n <= 0) return;
        const int32_t t0 = table_of(g, rd.cs), t1 = rd.n >= 2 ? table_of(g, rd.cs + 1) : -1;
        if (t0 < 0 && t1 < 0) return;
        uint32_t s0 = refsym(g, rd.cs), s1 = t1 >= 0 ? refsym(g, rd.cs + 1) : 0u;
        for (uint32_t j = 0; j < rd.dcnt && j < 2u; j++) {
            const DiffEnt en = dg.pool[rd.doff + j];
            if (en.col == rd.cs) s0 = en.sr & 0xfu;
            else if (en.col == rd.cs + 1) s1 = en.sr & 0xfu;
        }
        if (t0 >= 0) tab_vote(g, t0, s0, (uint32_t)r, be);
        if (t1 >= 0) tab_vote(g, t1, s0 << 4 | s1, (uint32_t)r, be);
    }
};

struct Votes {          // one launch for all votes: 64 threads per read group (two region slots each), one per overflow entry, one per read
    EntryVotes ev; StartVotes sv; int64_t n_ovf;
    NP_HD int64_t items() const { return (int64_t)ev.dg.n_groups * 64 + n_ovf + ev.d.n_reads; }
    template <class B> NP_HD void operator()(int64_t i, B& be) const {
        const int64_t nreg = (int64_t)ev.dg.n_groups * 64;
        if (i < nreg) {
            const int64_t grp = i >> 6; const int32_t s = (int32_t)(i & 63), cnt = ev.dg.gcnt[grp];
            if (s < cnt) ev(grp * npw::DIFF_GROUP_SLOTS + s, be);
            if (s + 64 < cnt) ev(grp * npw::DIFF_GROUP_SLOTS + s + 64, be);
        } else if (i < nreg + n_ovf) ev((int64_t)ev.dg.n_groups * npw::DIFF_GROUP_SLOTS + (i - nreg), be);
        else sv(i - nreg - n_ovf, be);
    }
};

// ---- chain -----------------------------------------------------------------------------------------------------
// compare-exchange of two (key, value) slots
#define NP_CE(a, b) do { if (k##b > k##a) { uint32_t tk_ = k##a; k##a = k##b; k##b = tk_; uint32_t tv_ = e##a; e##a = e##b; e##b = tv_; } } while (0)

// One column of the forward chain.  Slots e0..e7 (kmer | count << 16) in first-seen order, nk of them valid.  Previous
// column: pb = base nibbles of its score entries (first-seen order), pn of them, ps[] their scores, pam = argmax entry.
// Everything is kept in scalars / fully unrolled so that nothing is indexed dynamically (registers, not local memory).
struct ChainCol {
    double s0, s1, s2, s3, s4, s5, s6, s7;    // scores of the score entries
    uint32_t base, pv;                         // nibble q: base code / previous-base nibble of the winning 3-mer of entry q
    int32_t n, am;                             // entries, argmax
};
NP_HD double cc_get(const ChainCol& c, int q) {
    return q == 0 ? c.s0 : q == 1 ? c.s1 : q == 2 ? c.s2 : q == 3 ? c.s3 : q == 4 ? c.s4 : q == 5 ? c.s5 : q == 6 ? c.s6 : c.s7;
}
NP_HD void cc_set(ChainCol& c, int q, double v) {
    if (q == 0) c.s0 = v; else if (q == 1) c.s1 = v; else if (q == 2) c.s2 = v; else if (q == 3) c.s3 = v;
    else if (q == 4) c.s4 = v; else if (q == 5) c.s5 = v; else if (q == 6) c.s6 = v; else c.s7 = v;
}
NP_HD int nib_find(uint32_t nibs, int n, uint32_t v) {      // first q < n with nibble q == v, else n
    for (int q = 0; q < WK; q++) if (q < n && ((nibs >> (4 * q)) & 0xfu) == v) return q;
    return n;
}

struct Chain {          // per stretch
    Dev d; ColGlobals g;
    template <class B> NP_HD void operator()(int64_t si, B& be) const {
        if (si >= *g.n_start) return;
        const int32_t T = g.T;
        const int64_t t0 = g.sstart[si];
        const double rate = d.P.rate;
        // extent of the stretch, and whether every table of it is usable
        int32_t t1 = (int32_t)t0; bool ok = true;
        for (;;) {
            if (g.tbad[t1]) ok = false;
            if ((g.tflag[t1] & CF_LAST) || t1 + 1 >= T || g.tcol[t1 + 1] != g.tcol[t1] + 1) break;
            t1++;
        }
        if (!ok) {
            for (int32_t t = (int32_t)t0; t <= t1; t++) g.tunres[t] = 1;
            be.atomic_add(g.n_unresolved, 1);
            return;
        }
        ChainCol P; P.s0 = P.s1 = P.s2 = P.s3 = P.s4 = P.s5 = P.s6 = P.s7 = 0; P.base = P.pv = 0; P.n = 0; P.am = 0;
        for (int32_t t = (int32_t)t0; t <= t1; t++) {
            // slots in first-seen order: slot 0 is the draft's 3-mer (the reference's own vote comes first), the others
            // by smallest voter = largest key; empty slots (key 0) sink to the end.  Slots fill from 1 upwards.
            uint32_t e0 = g.te[t], e1 = g.te[(size_t)1 * T + t], e2 = g.te[(size_t)2 * T + t], e3 = g.te[(size_t)3 * T + t];
            uint32_t k1 = g.tfs[(size_t)1 * T + t], k2 = g.tfs[(size_t)2 * T + t], k3 = g.tfs[(size_t)3 * T + t];
            uint32_t e4 = 0, e5 = 0, e6 = 0, e7 = 0, k4 = 0, k5 = 0, k6 = 0, k7 = 0;
            if (e3 != 0u) {
                e4 = g.te[(size_t)4 * T + t]; e5 = g.te[(size_t)5 * T + t]; e6 = g.te[(size_t)6 * T + t]; e7 = g.te[(size_t)7 * T + t];
                k4 = g.tfs[(size_t)4 * T + t]; k5 = g.tfs[(size_t)5 * T + t]; k6 = g.tfs[(size_t)6 * T + t]; k7 = g.tfs[(size_t)7 * T + t];
            }
            int nk = 1 + (e1 != 0u) + (e2 != 0u) + (e3 != 0u) + (e4 != 0u) + (e5 != 0u) + (e6 != 0u) + (e7 != 0u);
            if (nk == 3) { NP_CE(1, 2); }
            else if (nk > 3) {                  // odd-even transposition sort of slots 1..7 (7 rounds sort 7 keys)
                NP_CE(1, 2); NP_CE(3, 4); NP_CE(5, 6);
                NP_CE(2, 3); NP_CE(4, 5); NP_CE(6, 7);
                NP_CE(1, 2); NP_CE(3, 4); NP_CE(5, 6);
                NP_CE(2, 3); NP_CE(4, 5); NP_CE(6, 7);
                NP_CE(1, 2); NP_CE(3, 4); NP_CE(5, 6);
                NP_CE(2, 3); NP_CE(4, 5); NP_CE(6, 7);
                NP_CE(1, 2); NP_CE(3, 4); NP_CE(5, 6);
            }
            const uint32_t total = g.tvotes[t];
            const uint32_t nd = (e1 >> 16) + (e2 >> 16) + (e3 >> 16) + (e4 >> 16) + (e5 >> 16) + (e6 >> 16) + (e7 >> 16);
            e0 = (e0 & 0xffffu) | ((total - nd) << 16);                 // the draft's 3-mer: its own vote + every read that agrees
            g.te[t] = e0;                                               // the backtrack sums counts per base
            const uint32_t refk = e0 & 0xffffu, tot = total > 1 ? total - 1 : total;
            const double dec = (double)tot * rate;
            const bool hasP = t > (int32_t)t0;                           // first column: every lookup resolves to 0
            ChainCol Q; Q.s0 = Q.s1 = Q.s2 = Q.s3 = Q.s4 = Q.s5 = Q.s6 = Q.s7 = 0; Q.base = Q.pv = 0; Q.n = 0; Q.am = 0;
#define NP_CHAIN_SLOT(J, E)                                                                                         \
            if (J < nk) {                                                                                           \
                const uint32_t k = (E) & 0xffffu, pvn = (k >> 4) & 0xfu, b = k & 0xfu;                               \
                uint32_t cnt = (E) >> 16;                                                                           \
                double s = 0;                                                                                       \
                if (hasP) {                                                                                         \
                    int q = P.am;                                                                                   \
                    if (pvn != 0u) { q = nib_find(P.base, P.n, pvn); if (q == P.n) { *d.err |= npe::ERR_MISSING_SCORE; q = 0; } } \
                    s = cc_get(P, q);                                                                               \
                }                                                                                                   \
                if (k == refk && total > 1) cnt--;                                                                  \
                s = s + ((double)cnt - dec);                                                                        \
                const int q = nib_find(Q.base, Q.n, b);                                                             \
                if (q == Q.n) { Q.base |= b << (4 * q); Q.pv |= pvn << (4 * q); cc_set(Q, q, s); Q.n++; }            \
                else if (cc_get(Q, q) < s) { cc_set(Q, q, s); Q.pv = (Q.pv & ~(0xfu << (4 * q))) | pvn << (4 * q); } \
            }
            NP_CHAIN_SLOT(0, e0) NP_CHAIN_SLOT(1, e1) NP_CHAIN_SLOT(2, e2) NP_CHAIN_SLOT(3, e3)
            NP_CHAIN_SLOT(4, e4) NP_CHAIN_SLOT(5, e5) NP_CHAIN_SLOT(6, e6) NP_CHAIN_SLOT(7, e7)
#undef NP_CHAIN_SLOT
            // base_max_score: first strictly greater in first-seen order (base.c:185-197)
            int am = 0; double mx = Q.s0;
            for (int q = 1; q < WK; q++) if (q < Q.n && cc_get(Q, q) > mx) { mx = cc_get(Q, q); am = q; }
            Q.am = am;
            g.bt_base[t] = Q.base; g.bt_pv[t] = Q.pv; g.bt_am[t] = (uint16_t)(am | Q.n << 4 | nk << 8);
            P = Q;
        }
        // backtrack (contig.c:473-496)
        int ent = g.bt_am[t1] & 0xf;
        for (int32_t t = t1;; t--) {
            const uint32_t chosen = (g.bt_base[t] >> (4 * ent)) & 0xfu;
            uint32_t support = 0;
            const int nkb = g.bt_am[t] >> 8;
            for (int j = 0; j < WK; j++) if (j < nkb) { const uint32_t e = g.te[(size_t)j * T + t]; if ((e & 0xfu) == chosen) support += e >> 16; }
            const uint32_t total = g.tvotes[t];
            uint8_t fl = g.tflag[t];
            if (total == 1) fl |= FLAG_ZERO;
            if (support / (double)total < d.P.min_count_ratio_skip) fl |= FLAG_COVERAGE;      // base.c:79-89
            const int32_t c = g.tcol[t];
            d.obase[c] = (uint8_t)chosen; d.oflag[c] = fl;
            if (t == (int32_t)t0) break;
            const uint32_t pvn = (g.bt_pv[t] >> (4 * ent)) & 0xfu;
            const uint32_t pam = g.bt_am[t - 1];
            if (pvn == 0u) ent = (int)(pam & 0xfu);
            else { const int pn = (int)((pam >> 4) & 0xfu); int q = nib_find(g.bt_base[t - 1], pn, pvn); if (q == pn) { *d.err |= npe::ERR_MISSING_SCORE; q = 0; } ent = q; }
        }
    }
};
#undef NP_CE

// unresolved stretches -> the general kernels' inputs: needi marks, votes of the marked columns
struct MarkUnresolved {
    Dev d; ColGlobals g;
    template <class B> NP_HD void operator()(int64_t t, B&) const {
        if (!g.tunres[t]) return;
        const int32_t c = g.tcol[t];
        d.needi[c] = 1; d.votes[c] = g.tvotes[t];
    }
};
struct ReadNeed {       // reads that cast a symbol on a marked column (tidx = exclusive scan of needi)
    Dev d; DiffGlobals dg; uint8_t* r_need;
    template <class B> NP_HD void operator()(int64_t r, B&) const {
        const ReadDesc rd = dg.rdesc[r];
        r_need[r] = (rd.cs >= 0 && rd.n > 0 && d.tidx[rd.cs + rd.n] - d.tidx[rd.cs] > 0) ? 1 : 0;
    }
};

}  // namespace npc
