// engine_v3.h — task 1 as a chain of full-GPU data-parallel kernels without windows ("v3").
//
// EXPERIMENTAL alternative to the fused window kernel (engine_v2.h), selected with NEXTPOLISH_B200_V3=1 for
// A/B runs; bit-identical results (tests/test_emu_kernels.py, tests/test_gpu_parity.py).  Measured on the
// 5 x 1 Mb / 30x bench shard (B200, round 1): compare_chunks 0.47 ms + walk 0.19 ms + chain_dp 0.23 ms + votes/
// tally 0.11 ms + ~0.25 ms of per-column passes = ~1.3 ms per task-1 step against ~0.9 ms for the window path,
// so the window path stays the default.  What v3 buys: no halo (every read is visited once), no shared-memory
// limits (any depth without the general fallback), no CTA barriers.  What it still pays: one thread per
// 8-column word re-reads the read's metadata and run list (400 warp instructions per word against ~100 in the
// window kernel's serial loop), a same-address atomic per parked chunk, and one-item-per-thread column passes
// that are latency- rather than bandwidth-bound.
//
// Same semantics as engine_v2.h / window_kernel.h (and the same building blocks: the run generator
// npw::next_run, the parked-chunk descriptors, the find-or-insert tally ordered by first voter), but every
// phase is its own kernel over ALL reads / chunks / columns / tables of the shard, so nothing waits at a CTA
// barrier and no read is visited twice:
//
//   read_prep, layout scan, col_init                (engine_impl.h)
//   pack_ref        per 8 columns : draft symbols as nibble words (refw)
//   walk_runs x2    per read      : CIGAR -> runs (count, scan, write), string extent, +1/-1 coverage marks,
//                                   first two symbols (partial 3-mers of a read start)
//   compare_chunks  per (read, 8-column word) : the pileup scan proper — gathers the read's symbols of the word
//                                   (+2 columns of look-back) from its runs and XORs them against the draft word;
//                                   disagreeing columns are flagged, words that disagree (or follow a
//                                   disagreement within 2 columns) are parked as 16-byte descriptors
//   votes scan, need_table, table layout (capacity = 1 + events of the column)
//   votes           per parked chunk / per read start : atomic find-or-insert of the 3-mer in its column's table
//   tally           per table     : first-seen order (by smallest read index), draft 3-mer gets the rest
//   chain_dp, anchor_cols, emit   (engine_impl.h)
//
// Reference semantics: contig.c:247-331 (walk), base.c:60-71 (tally), contig.c:424-496 (chain).
#pragma once
#include "window_kernel.h"

namespace npv {
using namespace npd;
using npe::Dev;

struct Run { int32_t lc, len, q; };            // first GLOBAL column, columns, first query index or -1 (gap symbols)

struct V3 {
    int32_t *r_nrun, *r_runoff;                // [R+1] runs per read, exclusive scan
    Run* runs;
    int32_t *r_cs, *r_n;                       // [R] global column of the first symbol (-1: none), string length
    uint32_t *r_info, *r_seq;                  // [R] sym0 << 4 | sym1 | min(n, 2) << 8;  word offset of the bases in rec | enc << 31
    uint32_t* refw;                            // [C/8 + 2] draft symbols, 8 columns per word, first column in the highest nibble
    int32_t *cov, *covs;                       // [C+2] +1/-1 marks; inclusive scan = reads voting on the column
    int32_t* evcnt;                            // [C+2] upper bound of non-draft votes per column (table capacity)
    uint32_t *ev_s, *ev_m; int32_t *ev_c, *ev_r;   // parked chunks: symbols, meta, first column, read
    int32_t* ev_n; int32_t ev_cap;
    uint32_t* kfirst;                          // parallel to Dev::ktab: smallest read index per entry
};

enum { SLOTS = 24 };                           // compare threads per read (one 8-column word each; longer reads loop)

// whole-contig context for npw::next_run
struct GCtx {
    const Dev& d;
    int32_t gs, ge, e0, e1, ncols, cb0, npos;
    int32_t ctr[npw::N_CTR];
};
NP_HD int32_t lcol(const GCtx& x, int32_t p) {
    int32_t i = p - x.e0;
    if (i < 0) return i;
    if (i > x.npos) return x.ncols + (i - x.npos);
    return x.d.colbase[p] - x.cb0;
}

struct PackRef {   // per symbol word
    Dev d; V3 v;
    template <class B> NP_HD void operator()(int64_t w, B&) const {
        uint32_t x = 0;
        for (int k = 0; k < 8; k++) { int64_t c = w * 8 + k; if (c < d.C) x |= (uint32_t)d.refsym[c] << (28 - 4 * k); }
        v.refw[w] = x;
    }
};

struct WalkRuns {  // per read; write == 0: count runs only
    Dev d; V3 v; int write;
    template <class B> NP_HD void operator()(int64_t r, B& be) const {
        if (r >= d.n_reads) { if (!write) { v.r_nrun[r] = 0; v.r_nrun[r + 1] = 0; } return; }
        int32_t nrun = 0, cs = -1, n = 0;
        uint32_t info = 0, sq = 0;
        if (d.r_level[r] == 1) {
            const Rec rc = load_rec(d.rec, d.rec_off, r);
            const int32_t k = d.r_ctg[r];
            const int32_t gs = d.ctg_goff[k], ge = d.ctg_goff[k + 1] - 1;
            const int32_t cb0 = d.colbase[gs];
            GCtx x{d, gs, ge, gs, ge + 1, d.colbase[ge + 1] - cb0, cb0, ge + 1 - gs, {0, 0, 0, 0, 0, 0}};
            npw::Walk w;
            w.cigar = rc.cigar; w.n_cigar = rc.n_cigar; w.ci = 0;
            w.pos = d.r_gpos[r]; w.qpos = 0; w.qstart = d.r_qstart[r]; w.qend = d.r_qend[r]; w.last = OP_I;
            w.mj = w.mjb = w.mlen = 0; w.in_m = false; w.plc = w.plen = 0; w.pq = -1;
            w.next = 0; w.started = false; w.done = false;
            Run* out = write ? v.runs + v.r_runoff[r] : nullptr;
            int32_t lc = 0, len = 0, q = -1;
            uint32_t s0 = 0, s1 = 0;
            while (npw::next_run(x, w, lc, len, q, be)) {
                if (n == 0) cs = cb0 + lc;
                if (cb0 + lc != cs + n) { x.ctr[npw::CTR_ERROR] = 1; break; }
                if (write) {
                    out[nrun] = Run{cb0 + lc, len, q};
                    if (n == 0) { s0 = q >= 0 ? rseq(rc, q) : (uint32_t)SYM_GAP; if (len >= 2) s1 = q >= 0 ? rseq(rc, q + 1) : (uint32_t)SYM_GAP; }
                    else if (n == 1) s1 = q >= 0 ? rseq(rc, q) : (uint32_t)SYM_GAP;
                }
                nrun++; n += len;
            }
            if (x.ctr[npw::CTR_ERROR]) *d.err |= npe::ERR_SYM_BOUND;
            info = (s0 << 4) | s1 | (uint32_t)(n < 2 ? n : 2) << 8;
            sq = (uint32_t)((rc.seq - d.rec) >> 2) | (rc.enc ? 0x80000000u : 0u);
        }
        if (!write) { v.r_nrun[r] = nrun; return; }
        v.r_cs[r] = n > 0 ? cs : -1; v.r_n[r] = n; v.r_info[r] = info; v.r_seq[r] = sq;
        if (n > 0) {
            be.atomic_add(&v.cov[cs], 1);
            be.atomic_add(&v.cov[cs + n], -1);
            // read starts cast partial 3-mers at their first two columns: room in those columns' tables
            be.atomic_add(&v.evcnt[cs], 1);
            if (n >= 2) be.atomic_add(&v.evcnt[cs + 1], 1);
        }
    }
};

// 16 symbols of a read starting at base qq, first one in the highest nibble; only the words holding the first
// `need` symbols are loaded
NP_HD unsigned long long fetch16(const uint32_t* sw, bool two, int32_t qq, int32_t need) {
    using namespace npw;
    if (two) {
        const int32_t si = qq >> 4, sn = qq & 15;
        const uint32_t x0 = bswap32(sw[si]), x1 = sn + need > 16 ? bswap32(sw[si + 1]) : 0u;
        const unsigned long long v = ((unsigned long long)x0 << 32 | x1) << (2 * sn);
        return (unsigned long long)expand2((uint32_t)(v >> 48)) << 32 | expand2((uint32_t)(v >> 32) & 0xffffu);
    }
    const int32_t si = qq >> 3, sn = qq & 7;
    const uint32_t b0 = bswap32(sw[si]), b1 = sn + need > 8 ? bswap32(sw[si + 1]) : 0u, b2 = sn + need > 16 ? bswap32(sw[si + 2]) : 0u;
    return (unsigned long long)fsl(b0, b1, (uint32_t)sn) << 32 | fsl(b1, b2, (uint32_t)sn);
}

struct CompareChunks {   // per (read, slot): the pileup scan
    Dev d; V3 v;
    template <class B> NP_HD void operator()(int64_t t, B& be) const {
        const int64_t r = t / SLOTS; const int32_t slot = (int32_t)(t % SLOTS);
        const int32_t cs = v.r_cs[r];
        if (cs < 0) return;
        const int32_t n = v.r_n[r];
        const int32_t nch = ((cs & 7) + n + 7) >> 3;
        if (slot >= nch) return;
        const int32_t ro = v.r_runoff[r], nrun = v.r_runoff[r + 1] - ro;
        const Run* runs = v.runs + ro;
        const uint32_t sq = v.r_seq[r];
        const uint32_t* sw = (const uint32_t*)d.rec + (sq & 0x7fffffffu);
        const bool two = (sq >> 31) != 0u;
        for (int32_t k = slot; k < nch; k += SLOTS) {
            const int32_t w0 = (cs >> 3) + k;
            const int32_t a = cs > 8 * w0 ? cs : 8 * w0, b = cs + n < 8 * w0 + 8 ? cs + n : 8 * w0 + 8;
            const int32_t take = b - a, dn = a & 7, idx0 = a - cs;
            const int32_t lo = idx0 >= 2 ? a - 2 : cs;               // first column gathered (look-back)
            // symbols of columns [a-2, b): nibble 0,1 = look-back, 2.. = the word's columns
            unsigned long long V = 0;
            for (int32_t j = 0; j < nrun; j++) {
                const Run ru = runs[j];
                const int32_t re = ru.lc + ru.len;
                if (re <= lo) continue;
                if (ru.lc >= b) break;
                const int32_t pa = ru.lc > lo ? ru.lc : lo, pb = re < b ? re : b, L = pb - pa;
                unsigned long long val = ru.q >= 0 ? fetch16(sw, two, ru.q + (pa - ru.lc), L) : 0x3333333333333333ull;
                val &= ~0ull << (64 - 4 * L);
                V |= val >> (4 * (pa - (a - 2)));
            }
            const uint32_t m = 0xffffffffu << (32 - 4 * take);
            const uint32_t S = (uint32_t)((V << 8) >> 32) & m, hist = (uint32_t)(V >> 56);
            const uint32_t R = (v.refw[w0] << (4 * dn)) & m;
            uint32_t Dn = S ^ R; Dn |= Dn >> 1; Dn |= Dn >> 2; Dn &= 0x11111111u;
            // a disagreement one / two columns back still owes events here (contig.c:360-363: 3-mers)
            uint32_t pin = 0;
            if (idx0 >= 1 && (hist & 0xfu) != npw::be_get(v.refw, a - 1)) pin = 2;
            else if (idx0 >= 2 && (hist >> 4) != npw::be_get(v.refw, a - 2)) pin = 1;
            uint32_t E = (Dn | (Dn >> 4) | (Dn >> 8) | (pin >= 1u ? 0x10000000u : 0u) | (pin >= 2u ? 0x01000000u : 0u)) & m & 0x11111111u;
            if (!E) continue;
            for (uint32_t x = Dn; x;) { const int32_t q = npw::clz32(x) >> 2; x &= ~(0x10000000u >> (4 * q)); d.mism[a + q] = 1; }
            bool any = false;
            for (uint32_t x = E; x;) {
                const int32_t q = npw::clz32(x) >> 2; x &= ~(0x10000000u >> (4 * q));
                if (idx0 + q >= 2) { be.atomic_add(&v.evcnt[a + q], 1); any = true; }
            }
            if (!any) continue;
            const int32_t slotp = be.atomic_add_ret(v.ev_n, 1);
            if (slotp < v.ev_cap) {
                v.ev_s[slotp] = S;
                v.ev_m[slotp] = (uint32_t)take << 16 | pin << 20 | (uint32_t)(idx0 < 2 ? idx0 : 2) << 22 | hist << 24;
                v.ev_c[slotp] = a; v.ev_r[slotp] = (int32_t)r;
            }
        }
    }
};

struct NeedTableV3 {   // per column: votes and table status (disagreeing columns + right neighbours)
    Dev d; V3 v;
    template <class B> NP_HD void operator()(int64_t c, B&) const {
        int32_t nd = 0;
        if (c < d.C) {
            const uint32_t votes = 1u + (uint32_t)v.covs[c];
            d.votes[c] = votes;
            if (votes >= 65535u) *d.err |= npe::ERR_DEPTH;     // uint16 counters of the reference would wrap (base.h:28-31,45)
            nd = d.mism[c];
            if (!nd && !(d.cflag[c] & CF_FIRST) && c > 0) nd = d.mism[c - 1];
        }
        d.needi[c] = nd;
    }
};
struct TableColsV3 {
    Dev d; V3 v;
    template <class B> NP_HD void operator()(int64_t c, B&) const {
        if (d.needi[c]) { int32_t t = d.tidx[c]; d.tcols[t] = (int32_t)c; d.tcap[t] = 1 + v.evcnt[c]; }
        if (c == 0) d.tcap[d.T] = 0;
    }
};
struct TabInit {       // entry 0 = the draft's own 3-mer (contig_as_read, contig.c:373-383)
    Dev d; V3 v;
    template <class B> NP_HD void operator()(int64_t t, B&) const { d.ktab[d.toff[t]] = npe::ref_kmer(d, d.tcols[t]); }
};

template <class B>
NP_HD void vote(const Dev& d, const V3& v, int32_t c, uint32_t kmer, uint32_t ridx, B& be) {
    if (!d.needi[c]) return;
    const int32_t t = d.tidx[c], o = d.toff[t], cap = d.toff[t + 1] - o;
    if (kmer == (d.ktab[o] & 0xffffu)) return;
    for (int32_t j = 1; j < cap; j++) {
        const uint32_t old = be.atomic_cas_u32(&d.ktab[o + j], 0u, kmer | (1u << 16));
        if (old == 0u) { be.atomic_min_u32(&v.kfirst[o + j], ridx); return; }
        if ((old & 0xffffu) == kmer) { be.atomic_add_u32(&d.ktab[o + j], 1u << 16); be.atomic_min_u32(&v.kfirst[o + j], ridx); return; }
    }
    *d.err |= npe::ERR_MISSING_SCORE;          // cannot happen: capacity = 1 + votes that may differ from the draft
}
struct ChunkVotes {    // per parked chunk
    Dev d; V3 v;
    template <class B> NP_HD void operator()(int64_t i, B& be) const {
        const uint32_t S = v.ev_s[i], eb = v.ev_m[i], hist = eb >> 24, pin = (eb >> 20) & 3u;
        const int32_t a = v.ev_c[i], take = (int32_t)((eb >> 16) & 0xfu), idx0 = (int32_t)((eb >> 22) & 3u);
        const uint32_t m = 0xffffffffu << (32 - 4 * take);
        const uint32_t R = (v.refw[a >> 3] << (4 * (a & 7))) & m;
        uint32_t Dn = S ^ R; Dn |= Dn >> 1; Dn |= Dn >> 2; Dn &= 0x11111111u;
        uint32_t Ev = (Dn | (Dn >> 4) | (Dn >> 8) | (pin >= 1u ? 0x10000000u : 0u) | (pin >= 2u ? 0x01000000u : 0u)) & m & 0x11111111u;
        const unsigned long long W = ((unsigned long long)hist << 32) | S;
        while (Ev) {
            const int32_t k = npw::clz32(Ev) >> 2;
            Ev &= ~(0x10000000u >> (4 * k));
            if (idx0 + k >= 2) vote(d, v, a + k, (uint32_t)(W >> (28 - 4 * k)) & 0xfffu, (uint32_t)v.ev_r[i], be);
        }
    }
};
struct StartVotes {    // per read: the first two symbols cast partial 3-mers (zeros for the missing symbols)
    Dev d; V3 v;
    template <class B> NP_HD void operator()(int64_t r, B& be) const {
        const int32_t cs = v.r_cs[r];
        if (cs < 0) return;
        const uint32_t info = v.r_info[r];
        vote(d, v, cs, (info >> 4) & 0xfu, (uint32_t)r, be);
        if ((info >> 8) >= 2u) vote(d, v, cs + 1, info & 0xffu, (uint32_t)r, be);
    }
};
struct TallyV3 {       // per table: first-seen order, draft 3-mer gets the remaining votes
    Dev d; V3 v;
    template <class B> NP_HD void operator()(int64_t t, B&) const {
        const int32_t o = d.toff[t], cap = d.toff[t + 1] - o;
        uint32_t* e = d.ktab + o; uint32_t* f = v.kfirst + o;
        int32_t nk = 1; uint32_t nd = 0;
        while (nk < cap && e[nk] != 0u) { nd += e[nk] >> 16; nk++; }
        for (int32_t a = 1; a < nk; a++) {
            int32_t mi = a;
            for (int32_t b = a + 1; b < nk; b++) if (f[b] < f[mi]) mi = b;
            if (mi != a) { uint32_t te = e[a], tf = f[a]; e[a] = e[mi]; f[a] = f[mi]; e[mi] = te; f[mi] = tf; }
        }
        e[0] = (e[0] & 0xffffu) | ((d.votes[d.tcols[t]] - nd) << 16);
        d.tnk[t] = nk;
    }
};

}  // namespace npv

namespace npe {

template <class BE>
int run_score_chain_v3(BE& be, Dev& d, RunStats* st) {
    if (!rate_is_dyadic(d.P.rate)) return run_score_chain(be, d, st, true);
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
    d.r_level = be.template buf<uint8_t>("r_level", R + 1);
    d.ins = be.template buf<int32_t>("ins", (size_t)G + 1);
    d.colbase = be.template buf<int32_t>("colbase", (size_t)G + 1);
    d.out_off = be.template buf<int64_t>("out_off", (size_t)d.n_ctg + 1);
    be.zero(d.ins, sizeof(int32_t) * ((size_t)G + 1));
    if (R > 0) be.launch("read_prep", R, ReadPrep{d});
    be.exscan_ncol(d.ins, d.colbase, (int64_t)G);
    d.C = be.read_i32(d.colbase + G);
    const int32_t C = d.C;
    d.refsym = be.template buf<uint8_t>("refsym", (size_t)C + 1);
    d.cflag = be.template buf<uint8_t>("cflag", (size_t)C + 1);
    d.mism = be.template buf<uint8_t>("mism", (size_t)C + 16);
    d.obase = be.template buf<uint8_t>("obase", (size_t)C + 1);
    d.oflag = be.template buf<uint8_t>("oflag", (size_t)C + 1);
    d.colpos = be.template buf<int32_t>("colpos", (size_t)C + 1);
    d.votes = be.template buf<uint32_t>("votes", (size_t)C + 1);
    d.needi = be.template buf<int32_t>("needi", (size_t)C + 1);
    d.tidx = be.template buf<int32_t>("tidx", (size_t)C + 1);
    d.keepidx = be.template buf<int32_t>("keepidx", (size_t)C + 1);
    npv::V3 v; memset(&v, 0, sizeof(v));
    v.r_nrun = be.template buf<int32_t>("v3_nrun", R + 2);
    v.r_runoff = be.template buf<int32_t>("v3_runoff", R + 2);
    v.r_cs = be.template buf<int32_t>("v3_cs", R + 1);
    v.r_n = be.template buf<int32_t>("v3_n", R + 1);
    v.r_info = be.template buf<uint32_t>("v3_info", R + 1);
    v.r_seq = be.template buf<uint32_t>("v3_seq", R + 1);
    v.refw = be.template buf<uint32_t>("v3_refw", (size_t)C / 8 + 4);
    v.cov = be.template buf<int32_t>("v3_cov", (size_t)C + 4);
    v.covs = be.template buf<int32_t>("v3_covs", (size_t)C + 4);
    v.evcnt = be.template buf<int32_t>("v3_evcnt", (size_t)C + 4);
    v.ev_n = be.template buf<int32_t>("v3_evn", 2);
    v.ev_cap = (int32_t)((R * 4 + 4096 < 0x7ffffff0ll) ? R * 4 + 4096 : 0x7ffffff0ll);
    v.ev_s = be.template buf<uint32_t>("v3_evs", (size_t)v.ev_cap);
    v.ev_m = be.template buf<uint32_t>("v3_evm", (size_t)v.ev_cap);
    v.ev_c = be.template buf<int32_t>("v3_evc", (size_t)v.ev_cap);
    v.ev_r = be.template buf<int32_t>("v3_evr", (size_t)v.ev_cap);
    be.zero(d.mism, (size_t)C + 16);
    be.zero(v.cov, sizeof(int32_t) * ((size_t)C + 4));
    be.zero(v.evcnt, sizeof(int32_t) * ((size_t)C + 4));
    be.zero(v.ev_n, 2 * sizeof(int32_t));
    if (G > 0) {
        be.launch("col_init", G, ColInit{d});
        be.launch("col_ends", d.n_ctg, ColEnds{d});
        be.launch("pack_ref", (int64_t)C / 8 + 1, npv::PackRef{d, v});
    }
    be.launch("walk_count", R + 1, npv::WalkRuns{d, v, 0});
    be.exscan_i32(v.r_nrun, v.r_runoff, R + 2);
    const int32_t n_runs = be.read_i32(v.r_runoff + R + 1);
    v.runs = be.template buf<npv::Run>("v3_runs", (size_t)n_runs + 1);
    if (R > 0) {
        be.launch("walk_write", R, npv::WalkRuns{d, v, 1});
        be.launch("pileup_scan", R * npv::SLOTS, npv::CompareChunks{d, v});
    }
    be.inclsum_i32(v.cov, v.covs, (int64_t)C + 1);
    be.launch("need_table", (int64_t)C + 1, npv::NeedTableV3{d, v});
    be.exscan_i32(d.needi, d.tidx, (int64_t)C + 1);
    int32_t n_ev = 0;
    {
        const int32_t* ptrs[2] = {d.tidx + C, v.ev_n};
        int32_t vals[2];
        be.read_many(ptrs, 2, vals);
        d.T = vals[0]; n_ev = vals[1];
    }
    if (n_ev > v.ev_cap) return run_score_chain(be, d, st);     // pathologically noisy shard: general kernels
    int32_t E = 0;
    if (d.T > 0) {
        const int32_t T = d.T;
        d.tcols = be.template buf<int32_t>("tcols", (size_t)T + 1);
        d.tcap = be.template buf<int32_t>("tcap", (size_t)T + 1);
        d.toff = be.template buf<int32_t>("toff", (size_t)T + 1);
        d.tnk = be.template buf<int32_t>("tnk", (size_t)T + 1);
        d.bpk = be.template buf<uint16_t>("bpk", (size_t)T * 16);
        d.amax = be.template buf<uint8_t>("amax", (size_t)T + 1);
        be.launch("table_cols", C, npv::TableColsV3{d, v});
        be.exscan_i32(d.tcap, d.toff, (int64_t)T + 1);
        E = be.read_i32(d.toff + T);
        d.ktab = be.template buf<uint32_t>("ktab", (size_t)E + 1);
        v.kfirst = be.template buf<uint32_t>("v3_kfirst", (size_t)E + 1);
        be.zero(d.ktab, sizeof(uint32_t) * ((size_t)E + 1));
        be.fill_ff(v.kfirst, sizeof(uint32_t) * ((size_t)E + 1));
        be.launch("tab_init", T, npv::TabInit{d, v});
        if (n_ev > 0) be.launch("chunk_votes", n_ev, npv::ChunkVotes{d, v});
        be.launch("start_votes", R, npv::StartVotes{d, v});
        be.launch("tally", T, npv::TallyV3{d, v});
        be.launch("chain_dp", T, ChainDP{d});
    }
    if (C > 0) be.launch("anchor_cols", C, AnchorCols{d});
    be.exscan_keep(d.obase, d.keepidx, (int64_t)C);
    int32_t total = 0, err = 0;
    {
        const int32_t* ptrs[2] = {d.keepidx + C, d.err};
        int32_t vals[2];
        be.read_many(ptrs, 2, vals);
        total = vals[0]; err = vals[1];
    }
    d.out = be.template buf<uint8_t>("out", (size_t)total + 1);
    if (C > 0) be.launch("emit", C, Emit{d, (uint8_t)(FLAG_ZERO | FLAG_COVERAGE)});
    be.launch("out_offsets", (int64_t)d.n_ctg + 1, OutOffsets{d});
    run_trace(be, d);
    if (st) { st->C = C; st->T = d.T; st->sym_words = n_runs; st->table_entries = E; st->out_bytes = total; }
    return err;
}

}  // namespace npe
