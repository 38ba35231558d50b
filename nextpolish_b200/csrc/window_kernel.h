// window_kernel.h — the fused pileup-scan kernel of task 1 ("one CTA per pileup window").
//
// A window owns W consecutive draft positions of one contig and works on an extended range
// [e0,e1) = [p0-HL, p1+HR) so that k-mer context (2 columns) and short non-anchor stretches that
// cross its right edge are available locally.  Per window:
//
//   stage    the contiguous byte range of the packed records of every read overlapping [e0,e1)
//            is copied into shared memory with ONE bulk copy (cp.async.bulk / TMA, mbarrier
//            completion), together with their rec_off entries;
//   expand   one thread per read: filter level, contig_cut_read, CIGAR walk at op granularity;
//            the read's column string (4-bit symbols, big-endian nibble order inside 32-bit words)
//            is written to shared memory with word-parallel nibble copies (contig.c:247-358);
//   compare  fused into expand: strings are stored aligned to the draft's 8-column symbol words, so each
//            string word is XORed against one draft word -> per column "covered" / "some read disagrees"
//            (OR-ed into two shared-memory accumulators per word);
//   tally    one thread per table column (disagreeing columns + right neighbours): 3-mer tallies in
//            first-seen (BAM) order from the shared-memory strings (base.c:60-71);
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
    int32_t *win_rlo, *win_rhi, *win_strw, *win_need;   // [n_win] plan: staged reads, string words, smem bytes
    int32_t* maxneed;                         // [1]
    uint8_t* r_need;                          // [n_reads] reads the fallback path must expand
    int32_t* n_unresolved;                    // [1]
    unsigned long long* phase_cycles;         // [16] optional per-phase cycle sums over all CTAs (tuning), or null
};

// Table columns live in a structure-of-arrays pool (entry-major) so that the threads of a warp, which
// work on consecutive tables, touch consecutive shared-memory words (no bank conflicts in the tally).
enum { TAB_BYTES = 128 };  // pool bytes per table column (8*8 scores + 8*4 k-mers + 8*2 + 8 + 2 + 4, rounded)
struct TabPool {
    double*   sc;          // [WK][tmax] score per score entry (chain phase)
    uint32_t* e;           // [WK][tmax] kmer | count << 16, first-seen order
    uint16_t* ekmer;       // [WK][tmax] winning k-mer per score entry
    uint16_t* votes;       // [tmax]
    uint8_t*  ebase;       // [WK][tmax] base code per score entry, first-seen order
    uint8_t  *nk, *nent, *amax, *bad;   // [tmax]
    int32_t   tmax;
};
struct TabEntry {          // accessor of one table column
    const TabPool* p; int32_t t;
    NP_HD double&   sc(int j) const { return p->sc[j * p->tmax + t]; }
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
struct alignas(16) ReadMeta {
    int32_t cs;            // local column of the first stored symbol
    int32_t cn;            // stored symbols
    int32_t so;            // word offset of the string in the pool
    uint32_t mm;           // bit k set = the k-th string word disagrees with the draft somewhere
};

struct WCtx {             // per-window context: globals + carved shared memory
    Dev d; WinGlobals g;
    int32_t win, k, gs, ge, p0, p1, e0, e1;   // contig, owned [p0,p1), extended [e0,e1)
    int32_t cb0, ncols, cown0, cown1;         // first ext column, #ext columns, owned local range [cown0,cown1)
    int32_t chr_end;                          // local column bound that a prev-window stretch may reach (HR rule)
    int32_t lc_first, lc_last;                // local columns of the contig's first / last column
    int32_t npos, nblk;                       // ext positions, 32-column blocks
    int32_t rlo, nr;                          // staged reads [rlo, rlo+nr)
    int32_t strw, tmax;                       // string pool words, table pool entries
    // shared memory
    uint8_t* rec; const uint32_t* recoff;     // staged records and their offsets (16-byte units, global)
    ReadMeta* rd;                             // per read: one 16-byte record (a single 128-bit shared-memory load)
    uint32_t *str, *refw, *acc;               // string pool, draft symbol words, compare accumulators
    uint8_t* colinfo;                         // per local column: 1 mism, 2 covered, 8 sub-column
    int16_t* tabidx;                          // per local column: table index or -1
    int16_t* tabcol;                          // dense list: table index -> local column
    uint16_t* lcb;                            // [npos+1] local column of every ext position (staged colbase)
    int32_t* blk;                             // [2*nblk] first / last+1 staged read overlapping each 32-column block
    TabPool tab;                              // aliases the record area (records are dead after expand)
    int32_t* ctr;                             // [0] string words used, [1] tables used, [2] internal error, [3] unresolved
};

NP_HD uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }
// shared-memory bytes of a window with nr reads, recbytes of records, ncols ext columns
NP_HD uint32_t win_smem_bytes(int32_t nr, uint32_t recbytes, int32_t ncols, int32_t strw, int32_t npos) {
    uint32_t b = 64;                                   // mbarrier + counters
    uint32_t recarea = align16(recbytes);              // later re-used as the table pool
    uint32_t mintab = (uint32_t)(ncols / 4 + 16) * (uint32_t)TAB_BYTES   /* ~13 % of columns need a table at 30x; 2x headroom */;
    if (recarea < mintab) recarea = mintab;
    b += recarea + align16(4u * (uint32_t)(nr + 1));
    b += align16(2u * (uint32_t)(npos + 2)) + align16(8u * (uint32_t)(ncols / 32 + 2));
    b += 16u * (uint32_t)nr;
    b += align16(4u * (uint32_t)(strw + 4));
    b += 3 * align16(4u * (uint32_t)(ncols / 8 + 3));
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
        int32_t extra = ncols - (e1 - e0);                           // insertion sub-columns in range
        int32_t strw = 0;
        for (int64_t r = lo; r < hi; r++) {
            int32_t a = d.r_gpos[r] < e0 ? e0 : d.r_gpos[r], b = d.r_wend[r] > e1 ? e1 : d.r_wend[r];
            int32_t span = b > a ? b - a : 0;
            strw += (span + extra + 14) / 8 + 2;
        }
        uint32_t recbytes = (d.rec_off[hi] - d.rec_off[lo]) * 16u;
        g.win_rlo[w] = (int32_t)lo; g.win_rhi[w] = (int32_t)hi; g.win_strw[w] = strw;
        uint32_t need = win_smem_bytes((int32_t)(hi - lo), recbytes, ncols, strw, e1 - e0);
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
    x.npos = x.e1 - x.e0; x.nblk = x.ncols / 32 + 1;
    x.rlo = x.g.win_rlo[w]; x.nr = x.g.win_rhi[w] - x.rlo;
    x.strw = x.g.win_strw[w];
    uint32_t recbytes = (d.rec_off[x.rlo + x.nr] - d.rec_off[x.rlo]) * 16u;
    uint32_t recarea = align16(recbytes), mintab = (uint32_t)(x.ncols / 4 + 16) * (uint32_t)TAB_BYTES;
    if (recarea < mintab) recarea = mintab;
    x.tmax = (int32_t)(recarea / TAB_BYTES) & ~7;      // multiple of 8 keeps every array 8-byte aligned
    if (x.tmax > ((x.ncols / 2 + 8) & ~7)) x.tmax = (x.ncols / 2 + 8) & ~7;   // the dense list (tabcol) holds ncols/2+16 entries
    uint8_t* p = smem + 64;
    x.ctr = (int32_t*)(smem + 16);
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
    x.blk = (int32_t*)p; p += align16(8u * (uint32_t)(x.ncols / 32 + 2));
    x.rd = (ReadMeta*)p; p += 16u * (uint32_t)x.nr;
    x.str = (uint32_t*)p; p += align16(4u * (uint32_t)(x.strw + 4));
    x.refw = (uint32_t*)p; p += align16(4u * (uint32_t)(x.ncols / 8 + 3));
    x.acc = (uint32_t*)p; p += 2 * align16(4u * (uint32_t)(x.ncols / 8 + 3));
    x.colinfo = p; p += align16((uint32_t)x.ncols + 16);
    x.tabidx = (int16_t*)p; p += align16(2u * (uint32_t)(x.ncols + 8));
    x.tabcol = (int16_t*)p;
}

// ---- big-endian nibble strings -------------------------------------------------------------------
// nibble i of a string lives in word i>>3 at bits [28-4*(i&7), +4)
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
// mask with nibbles [a, b) set (0 <= a <= b <= 8), nibble 0 = top
NP_HD uint32_t nib_mask(int32_t a, int32_t b) {
    a = a < 0 ? 0 : (a > 8 ? 8 : a);
    b = b < 0 ? 0 : (b > 8 ? 8 : b);
    uint32_t hi = (uint32_t)(0xffffffffull >> (a * 4));         // nibbles a..7
    uint32_t lo = (uint32_t)(0xffffffffull >> (b * 4));         // nibbles b..7
    return hi & ~lo;
}
// copy `len` nibbles of BAM seq (bytes, high nibble first == big-endian nibble order) starting at
// query index q into dst string at nibble index di
NP_HD void put_seq(uint32_t* dst, int32_t di, const uint8_t* seq, int32_t q, int32_t len) {
    const uint32_t* sw = (const uint32_t*)seq;   // seq is 4-byte aligned inside a record
    int32_t dn = di & 7;
    if (dn) {                                    // head: finish the partially filled destination word
        int32_t take = 8 - dn < len ? 8 - dn : len;
        int32_t si = q >> 3, sn = q & 7;
        uint32_t hi = bswap32(sw[si]), lo = (sn + take > 8) ? bswap32(sw[si + 1]) : 0u;
        uint32_t v = fsl(hi, lo, (uint32_t)sn) >> (dn * 4);
        uint32_t m = nib_mask(dn, dn + take);
        dst[di >> 3] = (dst[di >> 3] & ~m) | (v & m);
        di += take; q += take; len -= take;
    }
    int32_t dw = di >> 3, si = q >> 3;
    uint32_t sn = (uint32_t)(q & 7);
    if (len >= 8) {                              // body: whole destination words, one source load each
        uint32_t cur = bswap32(sw[si]);
        if (sn == 0) {
            for (;;) { dst[dw++] = cur; len -= 8; si++; if (len < 8) break; cur = bswap32(sw[si]); }
        } else {
            do { uint32_t nxt = bswap32(sw[si + 1]); dst[dw++] = fsl(cur, nxt, sn); cur = nxt; si++; len -= 8; } while (len >= 8);
        }
    }
    if (len > 0) {                               // tail
        uint32_t hi = bswap32(sw[si]), lo = (sn + (uint32_t)len > 8u) ? bswap32(sw[si + 1]) : 0u;
        uint32_t m = nib_mask(0, len);
        dst[dw] = (dst[dw] & ~m) | (fsl(hi, lo, sn) & m);
    }
}
NP_HD void put_const(uint32_t* dst, int32_t di, int32_t len, uint32_t sym) {
    uint32_t pat = sym * 0x11111111u;
    while (len > 0) {
        int32_t dw = di >> 3, dn = di & 7;
        int32_t take = 8 - dn < len ? 8 - dn : len;
        uint32_t m = nib_mask(dn, dn + take);
        dst[dw] = (dst[dw] & ~m) | (pat & m);
        di += take; len -= take;
    }
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
    {   // the string pool is 16-byte aligned: clear it with 128-bit stores
        Quad* q4 = (Quad*)x.str; const Quad z{0u, 0u, 0u, 0u};
        for (int32_t i = tid; i < (x.strw + 4 + 3) / 4; i += nt) q4[i] = z;
    }
    for (int32_t i = tid; i < x.ncols / 8 + 3; i += nt) { x.refw[i] = 0; x.acc[2 * i] = 0; x.acc[2 * i + 1] = 0; }
    for (int32_t i = tid; i < x.ncols + 16; i += nt) x.colinfo[i] = 0;
    for (int32_t i = tid; i < x.ncols + 8; i += nt) x.tabidx[i] = -1;
    for (int32_t i = tid; i < x.nblk; i += nt) { x.blk[2 * i] = 0x7fffffff; x.blk[2 * i + 1] = 0; }
    if (tid == 0) { x.ctr[0] = 0; x.ctr[1] = 0; x.ctr[2] = 0; x.ctr[3] = 0; x.ctr[4] = 0; }
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
            x.colinfo[c] = 8;
        }
    }
}

// ---- phase 1: expand one read into its column string ----------------------------------------------
// Emits the symbols of contig_parse_read (contig.c:247-331) at CIGAR-op granularity; only columns
// inside the extended range are stored.
NP_HD int32_t lcol(const WCtx& x, int32_t p) {     // local column of position p; virtual outside the range
    int32_t i = p - x.e0;
    if (i < 0) return i;
    if (i > x.npos) return (int32_t)x.lcb[x.npos] + (i - x.npos);
    return (int32_t)x.lcb[i];
}
struct StrWriter {
    WCtx* x; uint32_t* w; int32_t cap;        // string words of this read, capacity in nibbles
    int32_t cs, n;                            // local column of the first stored symbol, stored count
    int32_t next;                             // next expected local column (contiguity)
    bool started;
    const uint8_t* seq;
    // a run of `len` columns starting at local column lc: from seq[q..] (q >= 0) or constant gaps
    NP_HD void run(int32_t lc, int32_t len, int32_t q) {
        if (len <= 0) return;
        if (started && lc != next) x->ctr[2] = 1;         // cannot happen (votes are contiguous)
        next = lc + len; started = true;
        int32_t skip = lc < 0 ? -lc : 0;                  // clip to the extended range
        if (skip >= len) return;
        lc += skip; len -= skip; if (q >= 0) q += skip;
        if (lc + len > x->ncols) len = x->ncols - lc;
        if (len <= 0) return;
        if (n == 0) cs = lc;
        int32_t di = lc - (cs & ~7);                      // strings are aligned to 8-column words of the window
        if (lc != cs + n || di + len > cap) { x->ctr[2] = 1; return; }
        if (q >= 0) put_seq(w, di, seq, q, len); else put_const(w, di, len, SYM_GAP);
        n += len;
    }
};

template <class B>
NP_HD void ph_expand(WCtx& x, int32_t tid, int32_t nt, B& be) {
    const Dev& d = x.d;
    (void)tid; (void)nt;
    for (;;) {                                            // reads are handed out dynamically: no straggler round
        int32_t i = be.atomic_add_ret(&x.ctr[4], 1);
        if (i >= x.nr) break;
        x.rd[i] = ReadMeta{0, 0, 0, 0u};
        // filter level, usable query interval and reference span were computed per read by ReadPrep
        // (coalesced global arrays); only the record header, CIGAR and bases come from shared memory
        const int64_t r = (int64_t)x.rlo + i;
        if (d.r_level[r] != 1) continue;
        int32_t qstart = d.r_qstart[r], qend = d.r_qend[r];
        const int32_t gpos = d.r_gpos[r], wl = d.r_wend[r] - gpos;
        const uint8_t* p = x.rec + (size_t)(x.recoff[i] - x.recoff[0]) * 16;
        const Quad h = *(const Quad*)p;                       // one 128-bit load of the header
        Rec rc;
        rc.pos = (int32_t)h.a; rc.flag = h.b & 0xffffu; rc.mapq = (h.b >> 16) & 0xffu; rc.isize = (int32_t)h.c;
        rc.l_qseq = (int32_t)(h.d & 0xffffu); rc.n_cigar = (int32_t)(h.d >> 16);
        rc.cigar = (const uint32_t*)p + 4; rc.seq = p + 16 + 4 * (size_t)rc.n_cigar;
        // string capacity: same bound as the plan kernel
        int32_t a = gpos < x.e0 ? x.e0 : gpos, b = gpos + wl > x.e1 ? x.e1 : gpos + wl;
        int32_t span = b > a ? b - a : 0, extra = x.ncols - (x.e1 - x.e0);
        int32_t words = (span + extra + 14) / 8 + 2;
        int32_t off = be.atomic_add_ret(&x.ctr[0], words);
        if (off + words > x.strw + 4) { x.ctr[2] = 1; continue; }
        StrWriter sw{&x, x.str + off, (words - 1) * 8, 0, 0, 0, false, rc.seq};
        const int32_t start = x.gs, end = x.ge;
        int32_t pos = gpos, qpos = 0;
        int last = OP_I;
        for (int ci = 0; ci < rc.n_cigar; ci++) {
            int32_t len = cig_len(rc.cigar[ci]); int cur = cig_op(rc.cigar[ci]);
            if (cur == OP_M) {
                // in-range bases: query [max(qpos,qstart), min(qpos+len-1,qend)], pos in [start,end]
                int32_t ja = qstart > qpos ? qstart - qpos : 0;
                if (pos + ja < start) ja = start - pos;
                int32_t jb = qend - qpos < len - 1 ? qend - qpos : len - 1;
                if (pos + jb > end) jb = end - pos;
                if (pos + jb > x.e1) jb = x.e1 - pos;               // nothing beyond the extended range is stored
                if (ja <= jb) {
                    // first in-range base: sub-columns behind pos-1 are filled only under the rule of
                    // contig.c:273; every later base of the op fills unconditionally
                    int lastj = ja == 0 ? last : OP_M;
                    int32_t q0 = qpos + ja, p_a = pos + ja;
                    bool fill0 = lastj != OP_I && p_a > start && (q0 > qstart || (q0 == qstart && lastj == OP_D));
                    if (fill0) { int32_t cb = lcol(x, p_a - 1); sw.run(cb + 1, lcol(x, p_a) - cb - 1, -1); }
                    // copy segments between positions that carry sub-columns
                    int32_t j = ja;
                    while (j <= jb) {
                        int32_t pj = pos + j, cj = lcol(x, pj);
                        int32_t kmax = jb - j + 1, run = 1;
                        if (lcol(x, pj + kmax - 1) - cj == kmax - 1) run = kmax;          // no sub-columns inside
                        else while (run < kmax && lcol(x, pj + run) - cj == run) run++;
                        sw.run(cj, run, qpos + j);
                        j += run;
                        if (j <= jb) {                                                    // sub-columns behind pos+j-1
                            int32_t cb = lcol(x, pos + j - 1);
                            sw.run(cb + 1, lcol(x, pos + j) - cb - 1, -1);
                        }
                    }
                }
                pos += len; qpos += len; last = OP_M;
            } else if (cur == OP_D) {
                if (qpos >= qstart && qpos <= qend) {
                    int32_t ja = pos < start ? start - pos : 0, jb = pos + len - 1 > end ? end - pos : len - 1;
                    if (ja <= jb) {
                        int lastj = ja == 0 ? last : OP_D;
                        int32_t p_a = pos + ja, p_b = pos + jb;
                        bool fill0 = lastj != OP_I && p_a > start && (qpos > qstart || (qpos == qstart && lastj == OP_D));
                        int32_t c_from = fill0 ? lcol(x, p_a - 1) + 1 : lcol(x, p_a);
                        if (p_b > x.e1) p_b = x.e1;                    // clip far-right work
                        if (p_b >= p_a) sw.run(c_from, lcol(x, p_b) - c_from + 1, -1);
                    }
                }
                pos += len; last = OP_D;
            } else if (cur == OP_I) {
                if (pos != x.gs) {
                    // sub-columns behind a position left of the range are not stored (virtual columns)
                    bool in_reg = pos > start && pos <= end && pos - 1 >= x.e0 && pos <= x.e1;
                    if (in_reg) {
                        int32_t cb = lcol(x, pos - 1), nsub = lcol(x, pos) - cb - 1;
                        int32_t ja = qstart > qpos ? qstart - qpos : 0;
                        int32_t jb = qend - qpos < len - 1 ? qend - qpos : len - 1;
                        if (ja <= jb) {
                            if (jb >= nsub) { *d.err |= npe::ERR_INS_OVERFLOW; jb = nsub - 1; }
                            if (ja <= jb) sw.run(cb + 1 + ja, jb - ja + 1, qpos + ja);
                        }
                        int32_t qa = qpos + len;
                        if (qa > qstart && qa <= qend + 1 && nsub > len) sw.run(cb + 1 + len, nsub - len, -1);
                    }
                    qpos += len; last = OP_I;
                } else { qpos += len; qstart += len; last = OP_I; }
            } else if (cur == OP_S || cur == OP_H) {
                qpos += len;
            }
            if (pos > end || pos > x.e1 + 1) break;
        }
        uint32_t mmask = 0;
        if (sw.n > 0) {
            for (int32_t bq = sw.cs >> 5; bq <= (sw.cs + sw.n - 1) >> 5; bq++) {
                be.atomic_min(&x.blk[2 * bq], i);
                be.atomic_max(&x.blk[2 * bq + 1], i + 1);
            }
            // compare against the draft's symbol words (same alignment): per column "covered" / "disagrees"
            const uint32_t* sp = x.str + off;
            int32_t base = sw.cs & ~7, cw0 = base >> 3, nwd = (sw.cs + sw.n - base + 7) >> 3;
            mmask = nwd > 32 ? 0xffffffffu : 0u;                // very long strings: no fast path in the tally
            for (int32_t kq = 0; kq < nwd; kq++) {
                uint32_t m = nib_mask(sw.cs - base - 8 * kq, sw.cs + sw.n - base - 8 * kq);
                uint32_t df = (sp[kq] ^ x.refw[cw0 + kq]) & m;
                df |= df >> 1; df |= df >> 2; df &= 0x11111111u;
                if (df) { be.atomic_or(&x.acc[2 * (cw0 + kq)], df); if (kq < 32) mmask |= 1u << kq; }
                be.atomic_or(&x.acc[2 * (cw0 + kq) + 1], m & 0x11111111u);
            }
        }
        x.rd[i] = ReadMeta{sw.cs, sw.n, off, mmask};
    }
}

// ---- phase 2: per-column info from the accumulators the expand phase OR-ed together ---------------------
NP_HD void ph_colinfo(WCtx& x, int32_t tid, int32_t nt) {
    for (int32_t lc = tid; lc < x.ncols; lc += nt) {
        uint32_t bit = 28 - 4 * (lc & 7);
        uint8_t f = x.colinfo[lc] & 8;
        if ((x.acc[2 * (lc >> 3)] >> bit) & 1u) f |= 1;
        if ((x.acc[2 * (lc >> 3) + 1] >> bit) & 1u) f |= 2;
        x.colinfo[lc] = f;
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
        int32_t t = be.atomic_add_ret(&x.ctr[1], 1);
        x.tabidx[lc] = t < x.tmax ? (int16_t)t : (int16_t)-2;        // -2: pool exhausted
        if (t < x.tmax) x.tabcol[t] = (int16_t)lc;
    }
}
NP_HD void ph_tally(WCtx& x, int32_t tid, int32_t nt) {
    int32_t ntab = x.ctr[1] < x.tmax ? x.ctr[1] : x.tmax;
    for (int32_t ti = tid; ti < ntab; ti += nt) {
        int32_t lc = x.tabcol[ti];
        TabEntry T{&x.tab, ti};
        T.bad() = 0; T.nent() = 0; T.amax() = 0;
        // reference vote first (contig_as_read, contig.c:373-383)
        uint32_t k = be_get(x.refw, lc);
        if (!col_first(x, lc)) {
            k |= be_get(x.refw, lc - 1) << 4;
            if (!col_first(x, lc - 1)) k |= be_get(x.refw, lc - 2) << 8;
        }
        int32_t nk = 1; uint32_t votes = 1, e0 = 1;            // entry 0 = the draft's 3-mer, its count kept in a register
        int32_t blo = x.blk[2 * (lc >> 5)], bhi = x.blk[2 * (lc >> 5) + 1];
        for (int32_t r = blo; r < bhi; r++) {
            const ReadMeta m = x.rd[r];                     // one 128-bit load
            int32_t i = lc - m.cs;
            if (i < 0 || i >= m.cn) continue;
            int32_t al = m.cs & 7;                          // the string starts at nibble `al` of its first word
            votes++;
            // fast path: the read agrees with the draft in the words holding columns lc-2..lc and has cast at
            // least two symbols before lc -> it votes the draft's own 3-mer (entry 0)
            if (i >= 2) {
                uint32_t mmr = m.mm;
                int32_t k0 = (i - 2 + al) >> 3, k1 = (i + al) >> 3;
                if (k1 < 32 && !(((mmr >> k0) | (mmr >> k1)) & 1u)) { e0++; continue; }
            }
            const uint32_t* s = x.str + m.so;
            // symbols i-2..i of the read's string as one funnel-shifted extract
            uint32_t kk;
            if (i >= 2) {
                int32_t a = i - 2 + al;
                uint32_t v = fsl(s[a >> 3], (a & 7) > 5 ? s[(a >> 3) + 1] : 0u, (uint32_t)(a & 7));
                kk = v >> 20;
            } else {
                kk = be_get(s, i + al);
                if (i >= 1) kk |= be_get(s, i - 1 + al) << 4;
            }
            if (kk == k) { e0++; continue; }
            int32_t j = 1;
            for (; j < nk; j++) if ((T.e(j) & 0xffffu) == kk) { T.e(j) += 1u << 16; break; }
            if (j == nk) { if (nk < WK) T.e(nk++) = kk | (1u << 16); else T.bad() = 1; }
        }
        T.e(0) = k | (e0 << 16);
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
    x.ctr[3] = 1;
}

NP_HD void ph_chain(WCtx& x, int32_t tid, int32_t nt) {
    const Dev& d = x.d;
    const double rate = d.P.rate;
    int32_t ntab_all = x.ctr[1] < x.tmax ? x.ctr[1] : x.tmax;
    if (x.ctr[1] > x.tmax) {
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
            int32_t ti = x.tabidx[lc];
            // votes of table columns are needed by the fallback tables (capacity); recount if no table
            uint32_t v = 1;
            if (ti >= 0) v = x.tab.votes[ti];
            else for (int32_t r = 0; r < x.nr; r++) { int32_t i = lc - x.rd[r].cs; if (i >= 0 && i < x.rd[r].cn) v++; }
            d.votes[c] = v;
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
    if (x.ctr[2]) { if (tid == 0) *x.d.err |= npe::ERR_SYM_BOUND; }
    if (x.ctr[3]) {
        for (int32_t i = tid; i < x.nr; i += nt) x.g.r_need[x.rlo + i] = 1;
        if (tid == 0) be.atomic_add(x.g.n_unresolved, 1);
    }
}

}  // namespace npw
