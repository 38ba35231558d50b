// diff_pass.h — the streaming half of the task-1 pileup scan: every read is compared against the draft ONCE, in a
// global kernel (one thread per read), and leaves behind only its sparse DIFF against the draft's column string.
//
// A read votes on a contiguous run of columns [cs, cs+n) with one symbol per column (contig_parse_read,
// contig.c:247-331: its base, or 3 for a deletion / an insertion sub-column it does not fill).  Wherever that symbol
// equals the draft's own symbol of the column (the draft base; 3 on sub-columns) the vote cannot change any decision
// except through the per-column vote count, which follows from (cs, n) alone.  So a read is reduced to
//
//   ReadDesc { cs, n, doff, dcnt }      16 bytes: first global column, columns, its entries in the diff pool
//   DiffEnt  { col, sym } x dcnt        one entry per column where its symbol differs from the draft's, ascending
//
// (+1/-1 coverage marks at the ends of its extent, one bit per column it disagrees on), and the column kernels
// (column_pass.h) rebuild every 3-mer they need as "entry if present, else draft symbol".
//
// Compare in POSITION space: the bases of an M run are consecutive draft positions, so 2-bit reads are XORed against
// a 2-bit copy of the draft (PackDraft2) 16 bases per 32-bit operation; sub-columns between two voted positions get
// the gap symbol = the draft's symbol there and never produce an entry; colbase is touched only at the two ends of a
// read op and at its mismatches.  The CIGAR is walked at op granularity in a straight-line loop, so the threads of a
// warp (one read each) step through their ops together and meet at the single M-run compare.
#pragma once
#include "engine_impl.h"

namespace npw {
using namespace npd;
using npe::Dev;

struct alignas(16) ReadDesc { int32_t cs, n; uint32_t doff, dcnt; };   // cs < 0: the read casts nothing
struct alignas(8) DiffEnt { int32_t col; uint32_t sr; };               // sr = sym | read index << 4
enum { DIFF_MAX_READS = 1 << 28, DIFF_GROUP = 32, DIFF_GROUP_SLOTS = 128 };

NP_HD uint32_t bswap32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __byte_perm(v, 0u, 0x0123u);
#else
    return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24);
#endif
}
NP_HD int32_t clz32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}
// high 32 bits of (hi:lo) << sh, sh in [0, 31]
NP_HD uint32_t funnel_l(uint32_t hi, uint32_t lo, uint32_t sh) {
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, sh);
#else
    return sh == 0 ? hi : (hi << sh) | (lo >> (32 - sh));
#endif
}

// ---- 2-bit draft ------------------------------------------------------------------------------------------------
// d2[k] = positions 16k .. 16k+15, two bits per position (A 0, C 1, G 2, T 3 = log2 of the nt16 code), position 16k
// in the two highest bits; dn[k] = 0b11 at every position whose (upper-cased) draft character is not A/C/G/T.
struct alignas(8) Draft2 { uint32_t x, y; };
struct DiffGlobals {
    Draft2* dd;                         // [-4 .. G/16 + 2] .x = d2 word, .y = dn word (4 words of front padding: reads near position 0)
    ReadDesc* rdesc;                   // [R]
    // entry pool: group w of 32 consecutive reads owns slots [w * DIFF_GROUP_SLOTS, +gcnt[w]) (no atomics); reads that do
    // not fit their group's region take space from the overflow area [n_groups * DIFF_GROUP_SLOTS, pool_cap) instead
    DiffEnt* pool; int32_t* pool_n; int32_t pool_cap;    // pool_n: entries in the overflow area
    int32_t* gcnt; int32_t n_groups;
    int32_t* cov;                      // [C+2] +1 at a read's first column, -1 one past its last (prefix sum = reads voting on a column)
    uint32_t* disb;                    // [C/32+2] bit c: some read disagrees with the draft at column c
};
enum { ERR_DIFF_POOL = 128 };          // device error bit: the diff pool overflowed (the caller falls back to the general kernels)

struct PackDraft2 {                    // per 16 positions
    Dev d; DiffGlobals g;
    template <class B> NP_HD void operator()(int64_t k, B&) const {
        uint32_t w = 0, nmask = 0;
        for (int j = 0; j < 16; j++) {
            const int64_t p = k * 16 + j;
            uint32_t v = 0, bad = 0;
            if (p < d.G) {
                uint32_t ch = d.ctg_seq[p];
                if (ch >= 97 && ch <= 122) ch -= 32;
                switch (ch) { case 'A': v = 0; break; case 'C': v = 1; break; case 'G': v = 2; break; case 'T': v = 3; break; default: bad = 3; }
            }
            w |= v << (30 - 2 * j); nmask |= bad << (30 - 2 * j);
        }
        g.dd[k] = Draft2{w, nmask};
    }
};

NP_HD uint32_t draft_sym(const Dev& d, int32_t p) {
    uint32_t ch = d.ctg_seq[p];
    if (ch >= 97 && ch <= 122) ch -= 32;
    return base_code(ch);
}

// Entries of one read, buffered in local memory; a read with more than LB entries is walked a second time
// in write mode.  Every entry also sets its column's bit in the disagreement bitmap.
enum { DIFF_LB = 12 };
struct DiffSink {
    DiffEnt* out;                      // write mode: destination (null: buffer + count)
    DiffEnt buf[DIFF_LB];
    int32_t n; uint32_t rtag;          // entries so far; read index << 4
    uint32_t* disb;
    template <class B> NP_HD void put(int32_t col, uint32_t sym, B& be) {
        if (out) out[n] = DiffEnt{col, sym | rtag};
        else {
            if (n < DIFF_LB) buf[n] = DiffEnt{col, sym | rtag};
            be.atomic_or(&disb[col >> 5], 1u << (col & 31));
        }
        n++;
    }
};

// M run: read bases [q0, q0+len) against draft positions [p0, p0+len).  col0 >= 0: the columns are consecutive from
// col0; col0 < 0: column = colbase[p] (sub-columns lie inside the run).
// 2-bit reads: one iteration per 16-base word of the read.  Read base q sits at draft position q + delta, so the draft
// side is a sliding 64-bit window over dd[] shifted by a loop-invariant amount: one 8-byte load per iteration.
template <class B>
NP_HD void diff_m_run(const Dev& d, const Draft2* dd, const Rec& rc, int32_t q0, int32_t p0, int32_t len, int32_t col0, DiffSink& s, B& be) {
    if (!rc.enc) {                     // 4-bit reads (a base other than A/C/G/T somewhere): base by base
        for (int32_t j = 0; j < len; j++) {
            const uint32_t sy = seqi(rc.seq, q0 + j);
            if (sy != draft_sym(d, p0 + j)) s.put(col0 >= 0 ? col0 + j : d.colbase[p0 + j], sy, be);
        }
        return;
    }
    const uint32_t* sw = (const uint32_t*)rc.seq;
    const int32_t delta = p0 - q0, q1 = q0 + len - 1;
    const int32_t wi0 = q0 >> 4, wi1 = q1 >> 4;
    const int32_t P = 16 * wi0 + delta;                                      // draft position of the first word's base 0 (may be < 0)
    int32_t pi = P >> 4; const uint32_t ps = 2u * (uint32_t)(P & 15);
    const uint32_t mfirst = 0xffffffffu >> (2 * (q0 & 15));
    const uint32_t mlast = (q1 & 15) == 15 ? 0xffffffffu : ~(0xffffffffu >> (2 * ((q1 & 15) + 1)));
    Draft2 lo = dd[pi];
    for (int32_t wi = wi0; wi <= wi1; wi++, pi++) {
        const Draft2 hi = dd[pi + 1];
        const uint32_t rw = bswap32(sw[wi]);
        const uint32_t X = rw ^ funnel_l(lo.x, hi.x, ps);
        uint32_t D = ((X | (X >> 1)) | funnel_l(lo.y, hi.y, ps)) & 0x55555555u;
        if (wi == wi0) D &= mfirst;
        if (wi == wi1) D &= mlast;
        while (D) {
            const int32_t b = clz32(D) >> 1;
            D &= ~(0x40000000u >> (2 * b));
            const uint32_t sy = 1u << ((rw >> (30 - 2 * b)) & 3u);
            const int32_t q = 16 * wi + b;
            s.put(col0 >= 0 ? col0 + (q - q0) : d.colbase[q + delta], sy, be);
        }
        lo = hi;
    }
}

// One read -> its column extent and diff entries (count / buffer when s.out == null, write otherwise).
// The CIGAR walk of contig_parse_read (contig.c:247-331) at op granularity over the whole contig — the same
// conditions as the run generator above, straight-line: the threads of a warp step through their ops together and
// meet at the single M-run compare below.
enum { DK_NONE = 0, DK_M = 1, DK_D = 2, DK_I = 3 };
template <class B>
NP_HD void diff_read(const Dev& d, const Draft2* dd, int64_t r, const Rec& rc, int32_t& cs_out, int32_t& n_out, DiffSink& s, B& be) {
    const int32_t k = d.r_ctg[r];
    const int32_t gs = d.ctg_goff[k], ge = d.ctg_goff[k + 1] - 1;
    int32_t pos = d.r_gpos[r], qpos = 0, qstart = d.r_qstart[r];
    const int32_t qend = d.r_qend[r];
    int last = OP_I;
    int32_t cs = -1, next = 0;              // first column cast, one past the last column cast so far
    bool bad = false;
    for (int32_t ci = 0; ci < rc.n_cigar; ci++) {
        const uint32_t cg = rc.cigar[ci];
        const int32_t oplen = cig_len(cg); const int cur = cig_op(cg);
        int kind = DK_NONE; int32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        int32_t first = 0, lastc = -1;      // columns [first, lastc] cast by this op (lastc < first: none)
        if (cur == OP_M) {
            int32_t ja = qstart > qpos ? qstart - qpos : 0;
            if (pos + ja < gs) ja = gs - pos;
            int32_t jb = qend - qpos < oplen - 1 ? qend - qpos : oplen - 1;
            if (pos + jb > ge) jb = ge - pos;
            if (ja <= jb) {
                // first in-range base: sub-columns behind pos-1 are filled only under the rule of contig.c:273;
                // every later base of the op fills unconditionally
                const int lastj = ja == 0 ? last : OP_M;
                const int32_t q0 = qpos + ja, p_a = pos + ja, p_b = pos + jb;
                const bool fill0 = lastj != OP_I && p_a > gs && (q0 > qstart || (q0 == qstart && lastj == OP_D));
                const int32_t ca = d.colbase[p_a], cb = d.colbase[p_b];
                first = fill0 ? d.colbase[p_a - 1] + 1 : ca; lastc = cb;
                kind = DK_M; a0 = q0; a1 = p_a; a2 = jb - ja + 1; a3 = cb - ca == p_b - p_a ? ca : -1;
            }
            pos += oplen; qpos += oplen; last = OP_M;
        } else if (cur == OP_D) {
            if (qpos >= qstart && qpos <= qend) {
                const int32_t ja = pos < gs ? gs - pos : 0, jb = pos + oplen - 1 > ge ? ge - pos : oplen - 1;
                if (ja <= jb) {
                    const int lastj = ja == 0 ? last : OP_D;
                    const int32_t p_a = pos + ja, p_b = pos + jb;
                    const bool fill0 = lastj != OP_I && p_a > gs && (qpos > qstart || (qpos == qstart && lastj == OP_D));
                    first = fill0 ? d.colbase[p_a - 1] + 1 : d.colbase[p_a]; lastc = d.colbase[p_b];
                    kind = DK_D; a1 = p_a; a2 = p_b;
                }
            }
            pos += oplen; last = OP_D;
        } else if (cur == OP_I) {
            if (pos != gs) {
                if (pos > gs && pos <= ge) {
                    const int32_t cb = d.colbase[pos - 1], nsub = d.colbase[pos] - cb - 1;
                    const int32_t ja = qstart > qpos ? qstart - qpos : 0;
                    int32_t jb = qend - qpos < oplen - 1 ? qend - qpos : oplen - 1;
                    if (ja <= jb) {
                        if (jb >= nsub) { *d.err |= npe::ERR_INS_OVERFLOW; jb = nsub - 1; }
                        if (ja <= jb) { first = cb + 1 + ja; lastc = cb + 1 + jb; kind = DK_I; a0 = qpos + ja; a1 = first; a2 = jb - ja + 1; }
                    }
                    const int32_t qa = qpos + oplen;
                    if (qa > qstart && qa <= qend + 1 && nsub > oplen) {           // remaining sub-columns: gap votes
                        if (lastc < first) first = cb + 1 + oplen;
                        else if (lastc + 1 != cb + 1 + oplen) bad = true;
                        lastc = cb + nsub;
                    }
                }
                qpos += oplen; last = OP_I;
            } else { qpos += oplen; qstart += oplen; last = OP_I; }
        } else if (cur == OP_S || cur == OP_H) qpos += oplen;
        if (lastc >= first) {
            if (cs < 0) cs = first; else if (first != next) bad = true;   // cannot happen (votes are contiguous)
            next = lastc + 1;
        }
        if (kind == DK_M) diff_m_run(d, dd, rc, a0, a1, a2, a3, s, be);
        else if (kind == DK_D) {
            for (int32_t p = a1; p <= a2; p++) if (draft_sym(d, p) != (uint32_t)SYM_GAP) s.put(d.colbase[p], (uint32_t)SYM_GAP, be);
        } else if (kind == DK_I) {
            for (int32_t j = 0; j < a2; j++) { const uint32_t sy = rseq(rc, a0 + j); if (sy != (uint32_t)SYM_GAP) s.put(a1 + j, sy, be); }
        }
        if (pos > ge) break;
    }
    if (bad) *d.err |= npe::ERR_SYM_BOUND;
    cs_out = cs; n_out = cs >= 0 ? next - cs : 0;
}

struct DiffPass {
    Dev d; DiffGlobals g;
    // part 1 (per read): walk, entries buffered in the sink; returns the number of entries
    // rec / dd: where the records / the 2-bit draft are read from (global memory, or a staged copy)
    template <class B> NP_HD int32_t walk(int64_t r, const uint8_t* rec, const Draft2* dd, DiffSink& s, Rec& rc, int32_t& cs, int32_t& n, B& be) const {
        s.out = nullptr; s.n = 0; s.rtag = (uint32_t)r << 4; s.disb = g.disb;
        cs = -1; n = 0;
        if (r < d.n_reads && d.r_level[r] == 1) { rc = load_rec(rec, d.rec_off, r); diff_read(d, dd, r, rc, cs, n, s, be); }
        return s.n;
    }
    // part 2 (per read, r < n_reads): entries to the pool at `base`, coverage marks, descriptor
    template <class B> NP_HD void commit(int64_t r, int32_t base, DiffSink& s, const Rec& rc, int32_t cs, int32_t n, B& be) const {
        const int32_t cnt = s.n;
        ReadDesc rd{cs, n, (uint32_t)base, (uint32_t)cnt};
        if (cs >= 0) { be.atomic_add(&g.cov[cs], 1); be.atomic_add(&g.cov[cs + n], -1); }
        if (cnt > 0) {
            if (base < 0 || (int64_t)base + cnt > (int64_t)g.pool_cap) { *d.err |= ERR_DIFF_POOL; rd.dcnt = 0; rd.doff = 0; }
            else if (cnt <= DIFF_LB) { for (int32_t i = 0; i < cnt; i++) g.pool[base + i] = s.buf[i]; }
            else {
                DiffSink s2; s2.out = g.pool + base; s2.n = 0; s2.rtag = s.rtag; s2.disb = g.disb;
                int32_t cs2, n2;
                diff_read(d, g.dd, r, rc, cs2, n2, s2, be);
            }
        }
        g.rdesc[r] = rd;
    }
};
// slot of a read inside its group's region given the inclusive prefix of the group's counts, or -1: overflow area
NP_HD int32_t diff_group_slot(int32_t group, int32_t inc, int32_t cnt) {
    return inc <= DIFF_GROUP_SLOTS ? group * DIFF_GROUP_SLOTS + inc - cnt : -1;
}
}  // namespace npw
