// window_kernel.h — the fused pileup-scan kernel of task 1 ("one CTA per pileup window").
//
// A window owns W consecutive draft positions of one contig and works on an extended range
// [e0,e1) = [p0-HL, p1+HR) so that k-mer context (2 columns) and short non-anchor stretches that
// cross its right edge are available locally.  Per window:
//
//   stage    the contiguous byte range of the packed records of every read overlapping [e0,e1)
//            is copied into shared memory with ONE bulk copy (cp.async.bulk / TMA, mbarrier
//            completion), together with their rec_off entries;
//   compare  one thread per read: CIGAR walk at op granularity (contig.c:247-358); the read's column
//            string (4-bit symbols) is never materialised: it is produced 8 symbols at a time, aligned
//            to the draft's 8-column symbol words, and XORed against them.  A read leaves behind only
//            (a) +1/-1 at the ends of its (contiguous) voted column range, (b) a flag on every column
//            where it disagrees with the draft and (c) one EVENT (column, read index, 3-mer) for every
//            column within two columns after a disagreement — the only columns where its 3-mer can
//            differ from the draft's — plus its first two symbols (partial 3-mers of a read start);
//   scan     block-wide prefix sum of (a): reads voting on every column (votes = 1 + that);
//   tally    one thread per event: find-or-insert of the 3-mer in the table of its column (columns that
//            disagree + right neighbours) with shared-memory atomics, carrying the smallest read index
//            per entry; one thread per table then orders the entries by that index = first-seen (BAM)
//            order (base.c:60-71); the draft's own 3-mer gets all remaining votes;
//   chain    one thread per stretch that starts in the owned range: score chain + backtrack
//            (contig.c:424-496), result bases/flags written for every column of the stretch;
//   anchors  owned columns outside stretches keep the draft symbol.
//
// Anything the window cannot finish locally (a stretch running past the extended range, more than
// WK distinct 3-mers or base codes in a column, table pool exhausted) is marked `needi` and handled
// by the general global-memory kernels of engine_impl.h afterwards (same results, slower).
//
// The phase functions are plain NP_HD code over a context of shared-memory pointers so that the
// CPU test build (tests/emu) can run a window with a loop per phase.
#pragma once
#include "engine_impl.h"

namespace npw {
using namespace npd;
using npe::Dev;

enum { WK = 8,            // table capacity per column (distinct 3-mers) handled in shared memory
       HL = 2, HR = 32 }; // halo positions left / right of the owned range

struct WinGlobals {       // extra global arrays of the fused path
    int32_t W, n_win;
    const int32_t *win_ctg, *win_p0;          // [n_win]
    int32_t *win_rlo, *win_rhi, *win_need;    // [n_win] plan: staged reads, smem bytes
    int32_t* maxneed;                         // [1]
    uint8_t* r_need;                          // [n_reads] reads the fallback path must expand
    int32_t* n_unresolved;                    // [1]
    unsigned long long* phase_cycles;         // [16] optional per-phase cycle sums over all CTAs (tuning), or null
};

// Table columns live in a structure-of-arrays pool (entry-major) so that the threads of a warp, which
// work on consecutive tables, touch consecutive shared-memory words.
enum { TAB_BYTES = 128 };  // pool bytes per table column (8*8 scores + 8*4 k-mers + 8*2 + 8 + 2 + 4, rounded)
struct TabPool {
    double*   sc;          // [WK][tmax] score per score entry (chain phase); the tally phase keeps the
                           //            smallest read index per k-mer entry in the same bytes (uint32 [WK][tmax])
    uint32_t* e;           // [WK][tmax] kmer | count << 16; entry 0 = the draft's 3-mer, then first-seen order
    uint16_t* ekmer;       // [WK][tmax] winning k-mer per score entry
    uint16_t* votes;       // [tmax]
    uint8_t*  ebase;       // [WK][tmax] base code per score entry, first-seen order
    uint8_t  *nk, *nent, *amax, *bad;   // [tmax]
    int32_t   tmax;
};
struct TabEntry {          // accessor of one table column
    const TabPool* p; int32_t t;
    NP_HD double&   sc(int j) const { return p->sc[j * p->tmax + t]; }
    NP_HD uint32_t& fs(int j) const { return ((uint32_t*)p->sc)[j * p->tmax + t]; }
    NP_HD uint32_t& e(int j) const { return p->e[j * p->tmax + t]; }
    NP_HD uint16_t& ekmer(int j) const { return p->ekmer[j * p->tmax + t]; }
    NP_HD uint8_t&  ebase(int j) const { return p->ebase[j * p->tmax + t]; }
    NP_HD uint16_t& votes() const { return p->votes[t]; }
    NP_HD uint8_t&  nk() const { return p->nk[t]; }
    NP_HD uint8_t&  nent() const { return p->nent[t]; }
    NP_HD uint8_t&  amax() const { return p->amax[t]; }
    NP_HD uint8_t&  bad() const { return p->bad[t]; }
};

struct alignas(16) Quad { uint32_t a, b, c, d; };
struct alignas(8) ReadStart {
    int32_t cs;            // local column of the read's first stored symbol, -1: the read casts nothing here
    uint32_t info;         // sym0 << 4 | sym1 | min(stored symbols, 2) << 8
};

enum { RUNCAP = 8 };      // runs per read kept in shared memory (longer lists continue with the generator)
#ifndef __CUDACC__
struct uint2 { unsigned int x, y; };
#endif
enum { CTR_EVENTS = 0, CTR_TABLES = 1, CTR_ERROR = 2, CTR_UNRESOLVED = 3, CTR_TICKET = 4, N_CTR = 6,
       SCAN_SCRATCH_BYTES = 160 };

struct WCtx {             // per-window context: globals + carved shared memory
    Dev d; WinGlobals g;
    int32_t win, k, gs, ge, p0, p1, e0, e1;   // contig, owned [p0,p1), extended [e0,e1)
    int32_t cb0, ncols, cown0, cown1;         // first ext column, #ext columns, owned local range [cown0,cown1)
    int32_t chr_end;                          // local column bound that a prev-window stretch may reach (HR rule)
    int32_t lc_first, lc_last;                // local columns of the contig's first / last column
    int32_t npos;                             // ext positions
    int32_t rlo, nr;                          // staged reads [rlo, rlo+nr)
    int32_t evcap, tmax;                      // event list capacity, table pool entries
    // shared memory
    uint8_t* rec; const uint32_t* recoff;     // staged records and their offsets (16-byte units, global)
    ReadStart* rs;                            // per read
    uint32_t *ev_s, *ev_m; uint16_t* ev_r;    // parked chunks (see ph_compare): symbols, meta word, read index
    uint2* runs;                              // [nr][RUNCAP] pre-walked runs: .x = lc | len << 16, .y = first query index or -1
    uint32_t* refw;                           // draft symbol words (8 columns per word, big-endian nibbles)
    uint32_t* acc;                            // same layout as refw: bit 0 of a nibble set = some read disagrees at that column
    int32_t* cov;                             // [ncols+1] +1/-1 at string ends, then (scan) reads voting on each column
    uint8_t* colinfo;                         // per local column: 1 mism, 2 covered
    int16_t* tabidx;                          // per local column: table index or -1
    int16_t* tabcol;                          // dense list: table index -> local column
    uint16_t* lcb;                            // [npos+1] local column of every ext position (staged colbase)
    TabPool tab;                              // aliases the record area (records are dead after the compare phase)
    int32_t* ctr;                             // counters (CTR_*)
    int32_t* scan_scratch;                    // block scan scratch
};

NP_HD uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }
NP_HD int32_t win_evcap(int32_t nr) { int32_t c = (6 * nr + 7) & ~7; return c < 1024 ? 1024 : c; }   // parked chunks per window (~170 on 30x data, outliers 3-4x)
// shared-memory bytes of a window with nr reads, recbytes of records, ncols ext columns
NP_HD uint32_t win_smem_bytes(int32_t nr, uint32_t recbytes, int32_t ncols, int32_t npos) {
    uint32_t b = 64 + SCAN_SCRATCH_BYTES;              // mbarrier + counters, scan scratch
    uint32_t recarea = align16(recbytes);              // later re-used as the table pool
    uint32_t mintab = (uint32_t)(ncols / 4 + 16) * (uint32_t)TAB_BYTES   /* ~13 % of columns need a table at 30x; 2x headroom */;
    if (recarea < mintab) recarea = mintab;
    b += recarea + align16(4u * (uint32_t)(nr + 1));
    b += align16(2u * (uint32_t)(npos + 2));
    b += align16(8u * (uint32_t)(nr + 1));
    b += 10u * (uint32_t)win_evcap(nr);
    b += 8u * (uint32_t)RUNCAP * (uint32_t)(nr + 1);
    b += 2 * align16(4u * (uint32_t)(ncols / 8 + 3));
    b += align16(4u * (uint32_t)(ncols + 2));
    b += align16((uint32_t)ncols + 16);
    b += align16(2u * (uint32_t)(ncols + 8)) + align16(2u * (uint32_t)(ncols / 2 + 16));   // tabidx, tabcol
    return b;
}

// ---- plan (global kernel, one thread per window) -------------------------------------------------
struct WinPlan {
    Dev d; WinGlobals g;
    template <class B> NP_HD void operator()(int64_t w, B& be) const {
        int32_t k = g.win_ctg[w], p0 = g.win_p0[w];
        int32_t gs = d.ctg_goff[k], ge = d.ctg_goff[k + 1] - 1;
        int32_t p1 = p0 + g.W; if (p1 > ge + 1) p1 = ge + 1;
        int32_t e0 = p0 - HL < gs ? gs : p0 - HL, e1 = p1 + HR > ge + 1 ? ge + 1 : p1 + HR;
        int64_t r0 = d.ctg_read_off[k], r1 = d.ctg_read_off[k + 1];
        int64_t hi = lower_bound_i32(d.r_gpos, r0, r1, e1);          // first read starting at/after e1
        int64_t lo = upper_bound_i32(d.r_pm, 0, hi, e0);             // first read whose prefix-max end > e0
        if (lo < r0) lo = r0;
        if (lo > hi) lo = hi;
        int32_t ncols = d.colbase[e1] - d.colbase[e0];
        uint32_t recbytes = (d.rec_off[hi] - d.rec_off[lo]) * 16u;
        g.win_rlo[w] = (int32_t)lo; g.win_rhi[w] = (int32_t)hi;
        uint32_t need = win_smem_bytes((int32_t)(hi - lo), recbytes, ncols, e1 - e0);
        g.win_need[w] = (int32_t)need;
        be.atomic_max(g.maxneed, (int32_t)need);
    }
};

// ---- context setup (all threads compute the same values) ----------------------------------------
NP_HD void win_setup(WCtx& x, int32_t w, uint8_t* smem) {
    const Dev& d = x.d;
    x.win = w; x.k = x.g.win_ctg[w]; x.p0 = x.g.win_p0[w];
    x.gs = d.ctg_goff[x.k]; x.ge = d.ctg_goff[x.k + 1] - 1;
    x.p1 = x.p0 + x.g.W; if (x.p1 > x.ge + 1) x.p1 = x.ge + 1;
    x.e0 = x.p0 - HL < x.gs ? x.gs : x.p0 - HL;
    x.e1 = x.p1 + HR > x.ge + 1 ? x.ge + 1 : x.p1 + HR;
    x.cb0 = d.colbase[x.e0]; x.ncols = d.colbase[x.e1] - x.cb0;
    x.cown0 = d.colbase[x.p0] - x.cb0; x.cown1 = d.colbase[x.p1] - x.cb0;
    // a stretch handed over from the previous window must END before that window's extended range:
    int32_t pe = x.p0 + HR > x.ge + 1 ? x.ge + 1 : x.p0 + HR;
    x.chr_end = d.colbase[pe] - x.cb0;
    x.lc_first = d.colbase[x.gs] - x.cb0; x.lc_last = d.colbase[x.ge] - x.cb0;
    x.npos = x.e1 - x.e0;
    x.rlo = x.g.win_rlo[w]; x.nr = x.g.win_rhi[w] - x.rlo;
    x.evcap = win_evcap(x.nr);
    uint32_t recbytes = (d.rec_off[x.rlo + x.nr] - d.rec_off[x.rlo]) * 16u;
    uint32_t recarea = align16(recbytes), mintab = (uint32_t)(x.ncols / 4 + 16) * (uint32_t)TAB_BYTES;
    if (recarea < mintab) recarea = mintab;
    x.tmax = (int32_t)(recarea / TAB_BYTES) & ~7;      // multiple of 8 keeps every array 8-byte aligned
    if (x.tmax > ((x.ncols / 2 + 8) & ~7)) x.tmax = (x.ncols / 2 + 8) & ~7;   // the dense list (tabcol) holds ncols/2+16 entries
    x.ctr = (int32_t*)(smem + 16);
    x.scan_scratch = (int32_t*)(smem + 64);
    uint8_t* p = smem + 64 + SCAN_SCRATCH_BYTES;
    x.rec = p;
    {   // table pool carved from the same bytes
        uint8_t* q = p; size_t tm = (size_t)x.tmax;
        x.tab.tmax = x.tmax;
        x.tab.sc = (double*)q; q += 8 * WK * tm;
        x.tab.e = (uint32_t*)q; q += 4 * WK * tm;
        x.tab.ekmer = (uint16_t*)q; q += 2 * WK * tm;
        x.tab.votes = (uint16_t*)q; q += 2 * tm;
        x.tab.ebase = q; q += WK * tm;
        x.tab.nk = q; q += tm; x.tab.nent = q; q += tm; x.tab.amax = q; q += tm; x.tab.bad = q;
    }
    p += recarea;
    x.recoff = (const uint32_t*)p; p += align16(4u * (uint32_t)(x.nr + 1));
    x.lcb = (uint16_t*)p; p += align16(2u * (uint32_t)(x.npos + 2));
    x.rs = (ReadStart*)p; p += align16(8u * (uint32_t)(x.nr + 1));
    x.ev_s = (uint32_t*)p; p += 4u * (uint32_t)x.evcap;
    x.ev_m = (uint32_t*)p; p += 4u * (uint32_t)x.evcap;
    x.ev_r = (uint16_t*)p; p += 2u * (uint32_t)x.evcap;
    x.runs = (uint2*)p; p += 8u * (uint32_t)RUNCAP * (uint32_t)(x.nr + 1);
    x.refw = (uint32_t*)p; p += align16(4u * (uint32_t)(x.ncols / 8 + 3));
    x.acc = (uint32_t*)p; p += align16(4u * (uint32_t)(x.ncols / 8 + 3));
    x.cov = (int32_t*)p; p += align16(4u * (uint32_t)(x.ncols + 2));
    x.colinfo = p; p += align16((uint32_t)x.ncols + 16);
    x.tabidx = (int16_t*)p; p += align16(2u * (uint32_t)(x.ncols + 8));
    x.tabcol = (int16_t*)p;
}

// ---- big-endian nibble words ---------------------------------------------------------------------
// nibble i of a symbol array lives in word i>>3 at bits [28-4*(i&7), +4)
NP_HD uint32_t be_get(const uint32_t* w, int32_t i) { return (w[i >> 3] >> (28 - ((i & 7) << 2))) & 0xfu; }
NP_HD uint32_t bswap32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __byte_perm(v, 0u, 0x0123u);
#else
    return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24);
#endif
}
NP_HD uint32_t fsl(uint32_t hi, uint32_t lo, uint32_t nib) {   // 8 nibbles starting at nibble `nib` (0..7) of (hi:lo)
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, nib * 4u);
#else
    return nib == 0 ? hi : (hi << (nib * 4)) | (lo >> (32 - nib * 4));
#endif
}

// 8 bases of a 2-bit read (16 bits, first base in the two highest bits) -> 8 nt16 nibbles (1 << v), first base
// in the highest nibble
NP_HD uint32_t expand2(uint32_t v16) {
    uint32_t t = v16 & 0xffffu;
    t = (t | (t << 8)) & 0x00ff00ffu;
    t = (t | (t << 4)) & 0x0f0f0f0fu;
    t = (t | (t << 2)) & 0x33333333u;                       // one 2-bit value per nibble
    const uint32_t hi = (t >> 1) & 0x11111111u, mh = hi * 0xfu;
    const uint32_t r = 0x11111111u + (t & 0x11111111u);     // 1 or 2
    return (r & ~mh) | ((r << 2) & mh);                     // x4 where the high bit was set
}

// ---- phase 0: clear + draft symbols --------------------------------------------------------------
// Global loads of the window's colbase slice and draft bases are ISSUED first (up to PF per thread, held
// in registers) so that their latency overlaps the shared-memory clears and the bulk copy.
enum { PF = 4 };
struct Prefetch { int32_t cb[PF]; uint32_t ch[PF]; };
NP_HD void ph_prefetch(const WCtx& x, int32_t tid, int32_t nt, Prefetch& pf) {
    for (int k = 0; k < PF; k++) {
        int32_t i = tid + k * nt;
        pf.cb[k] = i <= x.npos ? x.d.colbase[x.e0 + i] : 0;
        pf.ch[k] = i < x.npos ? x.d.ctg_seq[x.e0 + i] : 0;
    }
}
NP_HD void ph_clear(WCtx& x, int32_t tid, int32_t nt, const Prefetch& pf) {
    for (int32_t i = tid; i < x.ncols / 8 + 3; i += nt) { x.refw[i] = 0; x.acc[i] = 0; }
    for (int32_t i = tid; i < x.ncols + 2; i += nt) x.cov[i] = 0;
    for (int32_t i = tid; i < x.ncols + 8; i += nt) x.tabidx[i] = -1;
    if (tid == 0) for (int k = 0; k < N_CTR; k++) x.ctr[k] = 0;
    for (int k = 0; k < PF; k++) { int32_t i = tid + k * nt; if (i <= x.npos) x.lcb[i] = (uint16_t)(pf.cb[k] - x.cb0); }
    for (int32_t i = tid + PF * nt; i <= x.npos; i += nt) x.lcb[i] = (uint16_t)(x.d.colbase[x.e0 + i] - x.cb0);
}
template <class B>
NP_HD void ph_ref(WCtx& x, int32_t tid, int32_t nt, B& be, const Prefetch& pf) {   // one thread per ext position
    const Dev& d = x.d;
    int k = 0;
    for (int32_t p = x.e0 + tid; p < x.e1; p += nt, k++) {
        int32_t lc = x.lcb[p - x.e0], n = (int32_t)x.lcb[p - x.e0 + 1] - lc;
        uint32_t ch = k < PF ? pf.ch[k] : d.ctg_seq[p];
        if (ch >= 97 && ch <= 122) ch -= 32;
        be.atomic_or(&x.refw[lc >> 3], base_code(ch) << (28 - ((lc & 7) << 2)));
        for (int32_t j = 1; j < n; j++) {
            int32_t c = lc + j;
            be.atomic_or(&x.refw[c >> 3], (uint32_t)SYM_GAP << (28 - ((c & 7) << 2)));
        }
    }
}

// ---- phase 1: compare one read's column string against the draft ------------------------------------
// Emits the symbols of contig_parse_read (contig.c:247-331) at CIGAR-op granularity; only columns
// inside the extended range are looked at.
NP_HD int32_t lcol(const WCtx& x, int32_t p) {     // local column of position p; virtual outside the range
    int32_t i = p - x.e0;
    if (i < 0) return i;
    if (i > x.npos) return (int32_t)x.lcb[x.npos] + (i - x.npos);
    return (int32_t)x.lcb[i];
}
// The CIGAR walk is a resumable generator of RUNS (a run = consecutive local columns that take their symbols
// either from consecutive read bases or from the gap symbol), so that the compare loop below is ONE loop over
// 8-column chunks shared by all lanes of a warp, whatever CIGAR op each lane is in.
struct Walk {
    const uint32_t* cigar; int32_t n_cigar, ci;
    int32_t pos, qpos, qstart, qend; int last;
    int32_t mj, mjb, mlen; bool in_m;         // M op in progress: next base index, last in-range index, op length
    int32_t plc, plen, pq;                    // pending second run of the current op (plen > 0)
    int32_t next; bool started, done;         // contiguity check, walk finished
};
template <class X>
NP_HD void walk_op_done(const X& x, Walk& w) {
    w.ci++;
    if (w.pos > x.ge || w.pos > x.e1 + 1) w.done = true;
}
// X: a context with gs, ge, e0, e1, ncols, ctr[], d.err and an lcol(x, p) overload (the window's WCtx, or the
// whole-contig context of engine_v3.h)
template <class X, class B>
NP_HD bool next_run(X& x, Walk& w, int32_t& lc, int32_t& len, int32_t& q, B& be) {
    (void)be;
    const int32_t start = x.gs, end = x.ge;
    for (;;) {
        int32_t rl = 0, rn = 0, rq = -1;      // raw run: first local column, length, first query index or -1 (gaps)
        if (w.plen > 0) { rl = w.plc; rn = w.plen; rq = w.pq; w.plen = 0; }
        else if (w.in_m) {
            // compare segments between positions that carry sub-columns
            int32_t pj = w.pos + w.mj, cj = lcol(x, pj);
            int32_t kmax = w.mjb - w.mj + 1, run = 1;
            if (lcol(x, pj + kmax - 1) - cj == kmax - 1) run = kmax;          // no sub-columns inside
            else {
                // largest run with lcol(pj + run - 1) - cj == run - 1 (the excess is monotone): binary search
                int32_t lo = 1, hi = kmax - 1;                                // answer in [lo, hi]
                while (lo < hi) { int32_t mid = (lo + hi + 1) >> 1; if (lcol(x, pj + mid - 1) - cj == mid - 1) lo = mid; else hi = mid - 1; }
                run = lo;
            }
            rl = cj; rn = run; rq = w.qpos + w.mj;
            w.mj += run;
            if (w.mj <= w.mjb) {                                              // sub-columns behind pos+mj-1
                int32_t cb = lcol(x, w.pos + w.mj - 1);
                w.plc = cb + 1; w.plen = lcol(x, w.pos + w.mj) - cb - 1; w.pq = -1;
            } else {
                w.in_m = false; w.pos += w.mlen; w.qpos += w.mlen; w.last = OP_M;
                walk_op_done(x, w);
            }
        } else {
            if (w.done || w.ci >= w.n_cigar) return false;
            const uint32_t cg = w.cigar[w.ci];
            const int32_t oplen = cig_len(cg); const int cur = cig_op(cg);
            const int32_t pos = w.pos, qpos = w.qpos, qstart = w.qstart, qend = w.qend;
            if (cur == OP_M) {
                // in-range bases: query [max(qpos,qstart), min(qpos+len-1,qend)], pos in [start,end]
                int32_t ja = qstart > qpos ? qstart - qpos : 0;
                if (pos + ja < start) ja = start - pos;
                int32_t jb = qend - qpos < oplen - 1 ? qend - qpos : oplen - 1;
                if (pos + jb > end) jb = end - pos;
                if (pos + jb > x.e1) jb = x.e1 - pos;               // nothing beyond the extended range is looked at
                if (ja <= jb) {
                    // first in-range base: sub-columns behind pos-1 are filled only under the rule of
                    // contig.c:273; every later base of the op fills unconditionally
                    int lastj = ja == 0 ? w.last : OP_M;
                    int32_t q0 = qpos + ja, p_a = pos + ja;
                    bool fill0 = lastj != OP_I && p_a > start && (q0 > qstart || (q0 == qstart && lastj == OP_D));
                    w.in_m = true; w.mj = ja; w.mjb = jb; w.mlen = oplen;
                    if (fill0) { int32_t cb = lcol(x, p_a - 1); rl = cb + 1; rn = lcol(x, p_a) - cb - 1; rq = -1; }
                } else {
                    w.pos += oplen; w.qpos += oplen; w.last = OP_M;
                    walk_op_done(x, w);
                }
            } else if (cur == OP_D) {
                if (qpos >= qstart && qpos <= qend) {
                    int32_t ja = pos < start ? start - pos : 0, jb = pos + oplen - 1 > end ? end - pos : oplen - 1;
                    if (ja <= jb) {
                        int lastj = ja == 0 ? w.last : OP_D;
                        int32_t p_a = pos + ja, p_b = pos + jb;
                        bool fill0 = lastj != OP_I && p_a > start && (qpos > qstart || (qpos == qstart && lastj == OP_D));
                        int32_t c_from = fill0 ? lcol(x, p_a - 1) + 1 : lcol(x, p_a);
                        if (p_b > x.e1) p_b = x.e1;                    // clip far-right work
                        if (p_b >= p_a) { rl = c_from; rn = lcol(x, p_b) - c_from + 1; rq = -1; }
                    }
                }
                w.pos += oplen; w.last = OP_D;
                walk_op_done(x, w);
            } else if (cur == OP_I) {
                if (pos != x.gs) {
                    // sub-columns behind a position left of the range are not looked at (virtual columns)
                    bool in_reg = pos > start && pos <= end && pos - 1 >= x.e0 && pos <= x.e1;
                    if (in_reg) {
                        int32_t cb = lcol(x, pos - 1), nsub = lcol(x, pos) - cb - 1;
                        int32_t ja = qstart > qpos ? qstart - qpos : 0;
                        int32_t jb = qend - qpos < oplen - 1 ? qend - qpos : oplen - 1;
                        if (ja <= jb) {
                            if (jb >= nsub) { *x.d.err |= npe::ERR_INS_OVERFLOW; jb = nsub - 1; }
                            if (ja <= jb) { rl = cb + 1 + ja; rn = jb - ja + 1; rq = qpos + ja; }
                        }
                        int32_t qa = qpos + oplen;
                        if (qa > qstart && qa <= qend + 1 && nsub > oplen) {
                            if (rn > 0) { w.plc = cb + 1 + oplen; w.plen = nsub - oplen; w.pq = -1; }
                            else { rl = cb + 1 + oplen; rn = nsub - oplen; rq = -1; }
                        }
                    }
                    w.qpos += oplen; w.last = OP_I;
                } else { w.qpos += oplen; w.qstart += oplen; w.last = OP_I; }
                walk_op_done(x, w);
            } else {
                if (cur == OP_S || cur == OP_H) w.qpos += oplen;
                walk_op_done(x, w);
            }
        }
        if (rn <= 0) continue;
        if (w.started && rl != w.next) x.ctr[CTR_ERROR] = 1;   // cannot happen (votes are contiguous)
        w.next = rl + rn; w.started = true;
        int32_t skip = rl < 0 ? -rl : 0;                  // clip to the extended range
        if (skip >= rn) continue;
        rl += skip; rn -= skip; if (rq >= 0) rq += skip;
        if (rl + rn > x.ncols) rn = x.ncols - rl;
        if (rn <= 0) continue;
        lc = rl; len = rn; q = rq;
        return true;
    }
}

// A chunk that disagrees with the draft (or still owes events for a disagreement just before it) is parked as a
// 16-byte descriptor; its votes are cast by ph_votes once the table columns are known.
//   symbols (left-justified, zero below); meta = lc | take << 16 | pend_in << 20 | min(idx0, 2) << 22 | hist << 24;
//   read index (uint16: a window with 65 536 reads does not fit shared memory)
template <class B>
NP_HD void ph_compare(WCtx& x, int32_t tid, int32_t nt, B& be) {
    const Dev& d = x.d;
    // Static assignment and a loop whose only exit is its condition: the lanes of a warp enter and leave the
    // chunk loop together and reconverge at its latch after every divergent `if`.
    for (int32_t i = tid; i < x.nr; i += nt) {
        // filter level, usable query interval and reference span were computed per read by ReadPrep
        // (coalesced global arrays); only the record header, CIGAR and bases come from shared memory
        const int64_t r = (int64_t)x.rlo + i;
        const uint8_t* p = x.rec + (size_t)(x.recoff[i] - x.recoff[0]) * 16;
        const uint32_t hd = ((const uint32_t*)p)[3];
        const int32_t n_cigar = (int32_t)(hd >> 16);
        const bool two = (((const uint32_t*)p)[1] >> 24) != 0u;           // 2 bits per base
        const uint32_t* sw = (const uint32_t*)(p + 16 + 4 * (size_t)n_cigar);   // bases, 4-byte aligned
        Walk w;
        w.cigar = (const uint32_t*)p + 4; w.n_cigar = n_cigar; w.ci = 0;
        w.pos = d.r_gpos[r]; w.qpos = 0; w.qstart = d.r_qstart[r]; w.qend = d.r_qend[r]; w.last = OP_I;
        w.mj = w.mjb = w.mlen = 0; w.in_m = false; w.plc = w.plen = 0; w.pq = -1;
        w.next = 0; w.started = false; w.done = false;
        int32_t lc = 0, len = 0, q = -1;
        int32_t cs = 0, n = 0;                            // first stored local column, stored symbols
        uint32_t hist = 0, pend = 0, s01 = 0;             // last two symbols; columns still owing an event; first two symbols
        uint32_t cur = 0; int32_t cur_si = -1;            // rolling big-endian word of the bases
        // pass A: the first RUNCAP runs of every read are generated with all lanes in step (the generator is
        // long, divergent code: called from inside the chunk loop it would run for one lane at a time)
        uint2* rl = x.runs + (size_t)i * RUNCAP;
        int32_t nrun = 0, krun = 1;
        bool more = d.r_level[r] == 1;
        while (more && nrun < RUNCAP) {
            more = next_run(x, w, lc, len, q, be);
            if (more) { rl[nrun] = uint2{(uint32_t)lc | (uint32_t)len << 16, (uint32_t)q}; nrun++; }
        }
        bool alive = nrun > 0;
        if (alive) { const uint2 u = rl[0]; lc = (int32_t)(u.x & 0xffffu); len = (int32_t)(u.x >> 16); q = (int32_t)u.y; cs = lc; }
        while (alive) {
            // one chunk, aligned to the draft's symbol words
            const int32_t dn = lc & 7;
            const int32_t take = 8 - dn < len ? 8 - dn : len;
            const uint32_t m = 0xffffffffu << (32 - 4 * take);
            uint32_t S = 0x33333333u;                     // SYM_GAP x 8
            if (q >= 0) {
                // `cur` = big-endian word holding base q (8 bases per word, or 16 for 2-bit reads)
                const int32_t sh = two ? 4 : 3, per = 1 << sh;
                const int32_t si = q >> sh, sn = q & (per - 1);
                if (si != cur_si) { cur = bswap32(sw[si]); cur_si = si; }
                uint32_t lo = 0u;
                if (sn + take > per) lo = bswap32(sw[si + 1]);
                if (two) {
#ifdef __CUDA_ARCH__
                    S = expand2(__funnelshift_l(lo, cur, 2u * (uint32_t)sn) >> 16);
#else
                    S = expand2((uint32_t)(((((unsigned long long)cur) << 32 | lo) << (2 * sn)) >> 48));
#endif
                } else S = fsl(cur, lo, (uint32_t)sn);
                if (sn + take > per) { cur = lo; cur_si = si + 1; }
                q += take;
            }
            S &= m;
            const uint32_t R = (x.refw[lc >> 3] << (4 * dn)) & m;
            const uint32_t D = S ^ R;
            if (n < 2 && n + take >= 2) s01 = n == 0 ? S >> 24 : ((hist << 4) | (S >> 28)) & 0xffu;
            if (D | pend) {
                uint32_t Dn = D | (D >> 1); Dn |= Dn >> 2; Dn &= 0x11111111u;
                if (Dn) be.atomic_or(&x.acc[lc >> 3], Dn >> (4 * dn));
                int32_t slot = be.atomic_add_ret(&x.ctr[CTR_EVENTS], 1);
                if (slot < x.evcap) {
                    x.ev_s[slot] = S;
                    x.ev_m[slot] = (uint32_t)lc | (uint32_t)take << 16 | pend << 20 | (uint32_t)(n < 2 ? n : 2) << 22 | hist << 24;
                    x.ev_r[slot] = (uint16_t)i;
                }
                // events spill two columns past a disagreement: what the next chunk still owes
                const unsigned long long M = (unsigned long long)Dn << 32;
                const unsigned long long E = M | (M >> 4) | (M >> 8) |
                                             ((unsigned long long)((pend >= 1u ? 0x10000000u : 0u) | (pend >= 2u ? 0x01000000u : 0u)) << 32);
                pend = ((E >> (56 - 4 * take)) & 1u) ? 2u : ((E >> (60 - 4 * take)) & 1u) ? 1u : 0u;
            }
            hist = (uint32_t)(((((unsigned long long)hist) << 32) | S) >> (32 - 4 * take)) & 0xffu;
            n += take; lc += take; len -= take;
            if (len == 0) {
                if (krun < nrun) { const uint2 u = rl[krun++]; lc = (int32_t)(u.x & 0xffffu); len = (int32_t)(u.x >> 16); q = (int32_t)u.y; }
                else alive = more && next_run(x, w, lc, len, q, be);
                if (alive && lc != cs + n) { x.ctr[CTR_ERROR] = 1; alive = false; }
            }
        }
        ReadStart rs{-1, 0u};
        if (n > 0) {
            if (n == 1) s01 = (hist & 0xfu) << 4;
            be.atomic_add(&x.cov[cs], 1);
            be.atomic_add(&x.cov[cs + n], -1);
            rs = ReadStart{cs, s01 | (uint32_t)(n < 2 ? n : 2) << 8};
        }
        x.rs[i] = rs;
    }
}

// ---- phase 2: votes per column (block-wide prefix sum of the +1/-1 marks) + per-column info ------------
template <class B>
NP_HD void ph_scan(WCtx& x, int32_t tid, int32_t nt, B& be) {
    const int32_t n = x.ncols + 1, per = (n + nt - 1) / nt;
    int32_t a = tid * per, b = a + per;
    if (a > n) a = n;
    if (b > n) b = n;
    int32_t s = 0;
    for (int32_t i = a; i < b; i++) s += x.cov[i];
    int32_t run = be.block_exscan(s, x.scan_scratch);          // barrier inside: every thread calls it
    for (int32_t i = a; i < b; i++) {
        run += x.cov[i];
        x.cov[i] = run;                                        // reads voting on column i
        x.colinfo[i] = (uint8_t)(((x.acc[i >> 3] >> (28 - 4 * (i & 7))) & 1u) | (run > 0 ? 2u : 0u));
        if (run >= 65534) *x.d.err |= npe::ERR_DEPTH;          // uint16 counters of the reference would wrap (base.h:28-31,45)
    }
}

// ---- phase 3: table columns ---------------------------------------------------------------------------
// table status: the column disagrees, or its left neighbour (same contig) does
NP_HD bool col_first(const WCtx& x, int32_t lc) { return lc == x.lc_first; }
NP_HD bool col_last(const WCtx& x, int32_t lc) { return lc == x.lc_last; }
NP_HD bool is_table(const WCtx& x, int32_t lc) {
    if (lc < 0 || lc >= x.ncols) return false;
    if (x.colinfo[lc] & 1) return true;
    return lc > 0 && !col_first(x, lc) && (x.colinfo[lc - 1] & 1);
}
template <class B>
NP_HD void ph_mark_tables(WCtx& x, int32_t tid, int32_t nt, B& be) {
    // columns needed by stretches this window may resolve: from the owned start (and two columns of
    // k-mer context are available because HL >= 2 positions) to the end of the extended range
    for (int32_t lc = x.cown0 + tid; lc < x.ncols; lc += nt) {
        if (!is_table(x, lc)) continue;
        int32_t t = be.atomic_add_ret(&x.ctr[CTR_TABLES], 1);
        x.tabidx[lc] = t < x.tmax ? (int16_t)t : (int16_t)-2;        // -2: pool exhausted
        if (t >= x.tmax) continue;
        x.tabcol[t] = (int16_t)lc;
        TabEntry T{&x.tab, t};
        // entry 0 = the draft's own 3-mer (contig_as_read, contig.c:373-383); the others start empty
        uint32_t k = be_get(x.refw, lc);
        if (!col_first(x, lc)) {
            k |= be_get(x.refw, lc - 1) << 4;
            if (!col_first(x, lc - 1)) k |= be_get(x.refw, lc - 2) << 8;
        }
        T.e(0) = k;
        for (int j = 1; j < WK; j++) { T.e(j) = 0; T.fs(j) = 0xffffffffu; }
        T.bad() = 0; T.nent() = 0; T.amax() = 0;
    }
}

// one vote of read `ridx` for 3-mer `kmer` at local column lc (no-op unless lc has a table and the 3-mer is
// not the draft's): find-or-insert with atomics; entries fill slots 1.. in claim order, the smallest read
// index per entry restores the first-seen order afterwards
template <class B>
NP_HD void tab_vote(WCtx& x, int32_t lc, uint32_t kmer, uint32_t ridx, B& be) {
    if (lc < 0 || lc >= x.ncols) return;
    int32_t t = x.tabidx[lc];
    if (t < 0) return;
    TabEntry T{&x.tab, t};
    if (kmer == (T.e(0) & 0xffffu)) return;
    for (int j = 1; j < WK; j++) {
        uint32_t old = be.atomic_cas_u32(&T.e(j), 0u, kmer | (1u << 16));
        if (old == 0u) { be.atomic_min_u32(&T.fs(j), ridx); return; }
        if ((old & 0xffffu) == kmer) { be.atomic_add_u32(&T.e(j), 1u << 16); be.atomic_min_u32(&T.fs(j), ridx); return; }
    }
    T.bad() = 1;
}
NP_HD int32_t clz32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}
template <class B>
NP_HD void ph_votes(WCtx& x, int32_t tid, int32_t nt, B& be) {
    // parked chunks: every column that disagrees, and the two after it, casts the read's 3-mer there
    int32_t nev = x.ctr[CTR_EVENTS] < x.evcap ? x.ctr[CTR_EVENTS] : x.evcap;
    for (int32_t i = tid; i < nev; i += nt) {
        const uint32_t S = x.ev_s[i], eb = x.ev_m[i], hist = eb >> 24, pin = (eb >> 20) & 3u, ridx = x.ev_r[i];
        const int32_t lc = (int32_t)(eb & 0xffffu), take = (int32_t)((eb >> 16) & 0xfu), idx0 = (int32_t)((eb >> 22) & 3u);
        const uint32_t m = 0xffffffffu << (32 - 4 * take);
        const uint32_t R = (x.refw[lc >> 3] << (4 * (lc & 7))) & m;
        uint32_t Dn = S ^ R; Dn |= Dn >> 1; Dn |= Dn >> 2; Dn &= 0x11111111u;
        uint32_t Ev = (Dn | (Dn >> 4) | (Dn >> 8) | (pin >= 1u ? 0x10000000u : 0u) | (pin >= 2u ? 0x01000000u : 0u)) & m & 0x11111111u;
        const unsigned long long W = ((unsigned long long)hist << 32) | S;
        while (Ev) {
            const int32_t k = clz32(Ev) >> 2;
            Ev &= ~(0x10000000u >> (4 * k));
            if (idx0 + k >= 2 && lc + k >= x.cown0) tab_vote(x, lc + k, (uint32_t)(W >> (28 - 4 * k)) & 0xfffu, ridx, be);
        }
    }
    // read starts: the first two symbols of a string cast partial 3-mers (zeros for the missing symbols)
    for (int32_t i = tid; i < x.nr; i += nt) {
        const ReadStart rs = x.rs[i];
        if (rs.cs < 0) continue;
        tab_vote(x, rs.cs, (rs.info >> 4) & 0xfu, (uint32_t)i, be);
        if ((rs.info >> 8) >= 2u) tab_vote(x, rs.cs + 1, rs.info & 0xffu, (uint32_t)i, be);
    }
}
NP_HD void ph_tally(WCtx& x, int32_t tid, int32_t nt) {
    int32_t ntab = x.ctr[CTR_TABLES] < x.tmax ? x.ctr[CTR_TABLES] : x.tmax;
    const bool overflow = x.ctr[CTR_EVENTS] > x.evcap;       // lost events: nothing of this window is trusted
    for (int32_t ti = tid; ti < ntab; ti += nt) {
        int32_t lc = x.tabcol[ti];
        TabEntry T{&x.tab, ti};
        if (overflow) T.bad() = 1;
        int32_t nk = 1; uint32_t nd = 0;
        while (nk < WK && T.e(nk) != 0u) { nd += T.e(nk) >> 16; nk++; }
        for (int a = 1; a < nk; a++) {                        // order by first voter = first-seen order
            int m = a;
            for (int b = a + 1; b < nk; b++) if (T.fs(b) < T.fs(m)) m = b;
            if (m != a) { uint32_t te = T.e(a), tf = T.fs(a); T.e(a) = T.e(m); T.fs(a) = T.fs(m); T.e(m) = te; T.fs(m) = tf; }
        }
        uint32_t votes = 1u + (uint32_t)x.cov[lc];
        T.e(0) = (T.e(0) & 0xffffu) | ((votes - nd) << 16);
        T.nk() = (uint8_t)nk; T.votes() = (uint16_t)votes;
    }
}

// ---- phase 4: stretches ---------------------------------------------------------------------------------
// Marks every column of an unfinished stretch that lies inside this window's extended range (also
// columns owned by the right neighbour: that window skips a handed-over leading run whenever the run
// closes before this window's extended range ends, so both rules agree on who marks what).
NP_HD void mark_unresolved(WCtx& x, int32_t lc_from, int32_t lc_to) {
    const Dev& d = x.d;
    for (int32_t lc = lc_from; lc <= lc_to && lc < x.ncols; lc++) d.needi[x.cb0 + lc] = 1;
    x.ctr[CTR_UNRESOLVED] = 1;
}

NP_HD void ph_chain(WCtx& x, int32_t tid, int32_t nt) {
    const Dev& d = x.d;
    const double rate = d.P.rate;
    int32_t ntab_all = x.ctr[CTR_TABLES] < x.tmax ? x.ctr[CTR_TABLES] : x.tmax;
    if (x.ctr[CTR_TABLES] > x.tmax) {
        // table pool exhausted: some table columns are not in the dense list; walk the columns instead
        ntab_all = -1;
    }
    int32_t nitems = ntab_all >= 0 ? ntab_all : x.cown1 - x.cown0;
    for (int32_t it = tid; it < nitems; it += nt) {
        int32_t lc0 = ntab_all >= 0 ? (int32_t)x.tabcol[it] : x.cown0 + it;
        if (lc0 < x.cown0 || lc0 >= x.cown1 || x.tabidx[lc0] == -1) continue;
        bool prev_table = lc0 > 0 && !col_first(x, lc0) && is_table(x, lc0 - 1);
        if (prev_table && lc0 != x.cown0) continue;              // not a stretch start
        // extent of the run of table columns starting here
        int32_t lend = lc0;
        bool closed = false;
        for (;;) {
            if (col_last(x, lend)) { closed = true; break; }
            if (lend + 1 >= x.ncols) break;                      // runs past the extended range
            if (!is_table(x, lend + 1)) { closed = true; break; }
            lend++;
        }
        if (prev_table) {
            // leading run handed over from the previous window: it resolved it iff the run closes
            // before ITS extended range ends; otherwise every window marks its own part
            if (!(closed && lend < x.chr_end)) mark_unresolved(x, lc0, lend);
            continue;
        }
        bool ok = closed;
        for (int32_t lc = lc0; ok && lc <= lend; lc++) { int32_t ti = x.tabidx[lc]; if (ti < 0 || x.tab.bad[ti]) ok = false; }
        if (!ok) { mark_unresolved(x, lc0, lend); continue; }
        // forward score chain (contig.c:424-471); scores live in the table entries (shared memory)
        for (int32_t lc = lc0; lc <= lend; lc++) {
            TabEntry T{&x.tab, x.tabidx[lc]};
            const bool hasP = lc > lc0;                                          // first column: every lookup resolves to 0
            TabEntry P{&x.tab, hasP ? x.tabidx[lc - 1] : 0};
            uint32_t total = T.votes(), refk = T.e(0) & 0xffffu, tot = total > 1 ? total - 1 : total;
            const double dec = (double)tot * rate;
            int no = 0;
            for (int j = 0; j < T.nk(); j++) {
                uint32_t k = T.e(j) & 0xffffu, cnt = T.e(j) >> 16, pv = (k >> 4) & 0xfu;
                double s = 0;
                if (hasP) {
                    int q = P.amax();
                    if (pv != 0) { for (q = 0; q < P.nent(); q++) if (P.ebase(q) == pv) break; if (q == P.nent()) { *d.err |= npe::ERR_MISSING_SCORE; q = 0; } }
                    s = P.sc(q);
                }
                if (k == refk && total > 1) cnt--;
                s = s + ((double)cnt - dec);
                uint32_t b = k & 0xfu;
                int q = 0; for (; q < no; q++) if (T.ebase(q) == b) break;
                if (q == no) { T.ebase(no) = (uint8_t)b; T.ekmer(no) = (uint16_t)k; T.sc(no) = s; no++; }
                else if (T.sc(q) < s) { T.sc(q) = s; T.ekmer(q) = (uint16_t)k; }
            }
            int am = 0; double mx = T.sc(0);
            for (int q = 1; q < no; q++) if (T.sc(q) > mx) { mx = T.sc(q); am = q; }
            T.nent() = (uint8_t)no; T.amax() = (uint8_t)am;
        }
        // backtrack (contig.c:473-496)
        int32_t ent = x.tab.amax[x.tabidx[lend]];
        for (int32_t lc = lend;; lc--) {
            TabEntry T{&x.tab, x.tabidx[lc]};
            uint32_t chosen = T.ebase(ent), support = 0;
            for (int j = 0; j < T.nk(); j++) if ((T.e(j) & 0xfu) == chosen) support += T.e(j) >> 16;
            int32_t c = x.cb0 + lc;
            uint8_t fl = 0;
            if (col_first(x, lc)) fl |= CF_FIRST;
            if (col_last(x, lc)) fl |= CF_LAST;
            if (T.votes() == 1) fl |= FLAG_ZERO;
            if (support / (double)T.votes() < d.P.min_count_ratio_skip) fl |= FLAG_COVERAGE;
            d.obase[c] = (uint8_t)chosen; d.oflag[c] = fl;
            if (lc == lc0) break;
            uint32_t k = T.ekmer(ent), pv = (k >> 4) & 0xfu;
            TabEntry P{&x.tab, x.tabidx[lc - 1]};
            if (pv == 0) ent = P.amax();
            else { int q = 0; for (; q < P.nent(); q++) if (P.ebase(q) == pv) break; if (q == P.nent()) { *d.err |= npe::ERR_MISSING_SCORE; q = 0; } ent = q; }
        }
    }
}

// ---- phase 5: anchors + bookkeeping for the fallback path -------------------------------------------------
NP_HD void ph_anchors(WCtx& x, int32_t tid, int32_t nt) {
    const Dev& d = x.d;
    for (int32_t lc = x.cown0 + tid; lc < x.cown1; lc += nt) {
        int32_t c = x.cb0 + lc;
        uint8_t ci = x.colinfo[lc];
        if (x.tabidx[lc] != -1) {
            // votes of table columns are needed by the fallback tables (capacity)
            d.votes[c] = 1u + (uint32_t)x.cov[lc];
            continue;
        }
        uint8_t fl = 0;
        if (col_first(x, lc)) fl |= CF_FIRST;
        if (col_last(x, lc)) fl |= CF_LAST;
        if (!(ci & 2)) fl |= FLAG_ZERO;                      // only the draft's own vote
        if (1.0 < d.P.min_count_ratio_skip) fl |= FLAG_COVERAGE;
        d.obase[c] = (uint8_t)be_get(x.refw, lc); d.oflag[c] = fl;      // needi stays 0 (cleared before the launch)
    }
}
template <class B>
NP_HD void ph_finish(WCtx& x, int32_t tid, int32_t nt, B& be) {
    if (x.ctr[CTR_ERROR]) { if (tid == 0) *x.d.err |= npe::ERR_SYM_BOUND; }
#if !defined(__CUDA_ARCH__) && defined(NP_DEBUG_UNRESOLVED)
    { static long long se = 0, sw = 0, st = 0, sr = 0; se += x.ctr[CTR_EVENTS]; st += x.ctr[CTR_TABLES]; sr += x.nr; sw++;
      if (sw % 500 == 0) printf("avg events %.1f tables %.1f reads %.1f\n", (double)se / sw, (double)st / sw, (double)sr / sw); }
    if (x.ctr[CTR_UNRESOLVED]) printf("unresolved window %d: nr %d events %d/%d tables %d/%d\n", x.win, x.nr, x.ctr[CTR_EVENTS], x.evcap, x.ctr[CTR_TABLES], x.tmax);
#endif
    if (x.ctr[CTR_UNRESOLVED]) {
        for (int32_t i = tid; i < x.nr; i += nt) x.g.r_need[x.rlo + i] = 1;
        if (tid == 0) be.atomic_add(x.g.n_unresolved, 1);
    }
}

}  // namespace npw
