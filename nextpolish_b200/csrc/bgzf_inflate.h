// bgzf_inflate.h — DEFLATE (RFC 1951) decoder for BGZF blocks, one WARP per block.
//
// the step immediately before the polishing path.  The reference inflates the BAM on
// the host through htslib (bgzf.c -> zlib inflate), ~67 MB/s of compressed input per core, twice per
// contig (contig.c:170-180 and :688-704); BGZF blocks are independent <= 64 KiB deflate members, so a
// whole BAM region inflates as thousands of independent streams.
//
// Division of labour inside a warp: DECODE and RESOLVE are decoupled.  Lane 0 owns the bit reader and decodes a
// batch of up to 32 symbols (a strictly sequential dependency chain) into shared memory without touching the output;
// then the warp resolves the batch: lane i takes symbol i, an exclusive prefix sum of the lengths gives every symbol
// its output position, literals are stored at once and LZ77 matches are copied by their own lanes as soon as all the
// bytes they read are final (round by round; a match that reads output of the same batch waits for the lanes before
// it).  BAM payloads are match-dominated (read bases repeat in the overlapping reads of the 32 KiB window, names and
// qualities likewise: ~9 bytes per symbol), and the old scheme — every match copied by the whole warp before the next
// symbol is decoded — put one L2 round trip per symbol on the critical path.
// Code tables live in shared memory, one small region per warp:
//   literal/length and distance symbols ordered by code (canonical Huffman, RFC 1951 3.2.2), code-length
//   histograms, and lookup tables indexed by the next 10 (literal/length) / 8 (distance) stream bits whose 32-bit
//   entries carry code length, extra-bit count and base value (one shared-memory load per code; longer codes
//   fall back to the canonical walk).
//
// The decoder is NP_HD code over a warp backend (lane id, broadcast, barrier) so that tests/emu can run
// it on the CPU with a 1-lane backend and compare with zlib; the product compiles it into
// k_bgzf_inflate (bgzf_inflate.cu).  Nothing here is a CPU fallback of the product.
#pragma once
#include <stdint.h>
#include "device_logic.h"

#ifndef NP_UNLIKELY
#define NP_UNLIKELY(x) __builtin_expect(!!(x), 0)
#endif
#if defined(__CUDACC__)
#define NP_NO_UNROLL _Pragma("unroll 1")
#else
#define NP_NO_UNROLL
#endif

namespace npz {

enum { OK = 0, ERR_BTYPE = 1, ERR_STORED = 2, ERR_LENGTHS = 3, ERR_CODE = 4, ERR_DIST = 5, ERR_OUTPUT = 6,
       ERR_INPUT = 7, ERR_SIZE = 8 };

struct Block {             // one BGZF block: where its raw deflate payload lives, where its output goes
    uint64_t in_off;       // offset of the deflate payload (after the gzip header) in the compressed buffer
    uint32_t in_len;       // payload bytes (BSIZE + 1 - header - 8 trailer bytes)
    uint32_t out_len;      // ISIZE
    uint64_t out_off;      // offset of the block's bytes in the output buffer
};

enum { MAXBITS = 15, MAXL = 288, MAXD = 30, FASTBITS = 10, DFASTBITS = 8, BATCH = 32 };

// Lookup-table entries (literal/length and distance alike), one 32-bit word:
//   bits 0-3   code length in bits
//   bit 4      literal
//   bit 5      not a symbol with a value (E_SPECIAL): end of block (E_END), a code longer than the table's index — the
//              canonical walk finds it (E_LONG; takes nothing out of the bit buffer) — or an invalid symbol (neither)
//   bits 8-15  code length + extra bits (RFC 1951 3.2.5): what the symbol takes out of the bit buffer
//   bits 16-31 literal byte, base length or base distance
// so that one shared-memory load yields everything a symbol needs, code + extra bits leave the bit buffer together, and
// every rare case hides behind ONE flag test.
enum { E_LIT = 0x10, E_SPECIAL = 0x20, E_END = 0x40, E_LONG = 0x80 };
NP_HD uint32_t entry_bits(int l, int x) { return (uint32_t)l | (uint32_t)(l + x) << 8; }
NP_HD uint32_t entry_take(uint32_t e) {                     // bits 8-15
#ifdef __CUDA_ARCH__
    return __byte_perm(e, 0u, 0x4441u);
#else
    return (e >> 8) & 0xffu;
#endif
}
NP_HD uint32_t low_mask(uint32_t n) {                       // n ones (n < 32)
#ifdef __CUDA_ARCH__
    uint32_t m; asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(m) : "r"(n)); return m;
#else
    return ~(0xffffffffu << n);
#endif
}
NP_HD uint32_t entry_value(uint32_t e, uint32_t lo) {       // base + the extra bits that follow the code
    const uint32_t l = e & 15u;
    return (e >> 16) + ((lo >> l) & low_mask(entry_take(e) - l));
}

struct Tables {            // per-warp shared memory (5.9 KB)
    uint32_t batch[BATCH];                 // decoded symbols: literal = 0x8000 | byte; match = length | distance << 16
    int32_t lfirst, lindex, dfirst, dindex; // state of the canonical walk after FASTBITS / DFASTBITS steps (the slow path resumes there)
    uint16_t lcount[MAXBITS + 1], lsym[MAXL];
    uint16_t dcount[MAXBITS + 1], dsym[MAXD + 2];
    uint32_t fast[1 << FASTBITS];          // literal/length entries by the next FASTBITS stream bits
    uint32_t dfast[1 << DFASTBITS];        // distance entries by the next DFASTBITS stream bits
    // scratch while a block header is read (code lengths of both alphabets): the lookup table is built afterwards
    NP_HD uint16_t* lens() { return (uint16_t*)fast; }
};
static_assert((MAXL + MAXD + 2) * 2 <= (1 << FASTBITS) * 4, "the code-length scratch borrows the lookup table");

struct Bits {              // LSB-first bit reader over the compressed bytes (the decoding lane only)
    // The payload is consumed as ALIGNED 32-bit words (one load per refill, no end test on the way): the word holding the
    // first payload byte is shifted into place at the start, and reading runs at most 7 bytes past the payload — inside the
    // next block's header, or the slack every compressed buffer of this engine carries.  Bits that were never there are
    // noticed afterwards by comparing the consumed count with the payload size (past_end).
    const uint8_t* in; uint32_t in_len;
    uint32_t wi;           // next word to load, in words from the aligned start
    uint32_t skew;         // bits of the first word that precede the payload
    uint64_t buf; int32_t cnt;
    const uint32_t* wbase; // the aligned word that holds the first payload byte
    NP_HD uint32_t word(uint32_t i) const {
#ifdef __CUDA_ARCH__
        return wbase[i];
#else
        uint32_t v = 0;                                   // test build: byte loads, nothing outside [in, in + in_len)
        for (int j = 0; j < 4; j++) {
            const int64_t o = (int64_t)i * 4 + j - (int64_t)(skew >> 3);
            if (o >= 0 && o < (int64_t)in_len) v |= (uint32_t)in[o] << (8 * j);
        }
        return v;
#endif
    }
    NP_HD void seek(const uint8_t* base, uint32_t len, uint32_t byte_off) {     // continue at payload byte byte_off
        const uint32_t mis = (uint32_t)((uintptr_t)(base + byte_off) & 3u);
        in = base + byte_off; in_len = len - byte_off; skew = 8u * mis;
        wbase = (const uint32_t*)(in - mis);
        buf = (uint64_t)(word(0) >> skew); cnt = 32 - (int32_t)skew; wi = 1;
    }
    NP_HD void refill() { if (cnt <= 32) { buf |= (uint64_t)word(wi++) << cnt; cnt += 32; } }    // afterwards cnt > 32
    NP_HD uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1ull)); }
    NP_HD void drop(int n) { buf >>= n; cnt -= n; }
    NP_HD uint32_t get(int n) { refill(); uint32_t v = peek(n); drop(n); return v; }
    NP_HD uint32_t consumed_bits() const { return wi * 32u - skew - (uint32_t)cnt; }
    NP_HD bool past_end() const { return consumed_bits() > in_len * 8u; }      // consumed bits that were never there
};

// canonical code construction (RFC 1951 3.2.2): count[len], symbols ordered by (len, symbol)
NP_HD int build(uint16_t* count, uint16_t* sym, const uint16_t* lens, int n) {
    uint16_t offs[MAXBITS + 1];
    for (int l = 0; l <= MAXBITS; l++) count[l] = 0;
    for (int s = 0; s < n; s++) count[lens[s]]++;
    if (count[0] == n) return 0;                 // no codes: legal for an unused distance alphabet
    int left = 1;
    for (int l = 1; l <= MAXBITS; l++) { left <<= 1; left -= count[l]; if (left < 0) return left; }
    offs[1] = 0;
    for (int l = 1; l < MAXBITS; l++) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    for (int s = 0; s < n; s++) if (lens[s]) sym[offs[lens[s]]++] = (uint16_t)s;
    return left;                                 // > 0: incomplete code
}
NP_HD uint32_t rev_bits(uint32_t v, int n) {     // reverse the low n bits
#ifdef __CUDA_ARCH__
    return __brev(v) >> (32 - n);
#else
    uint32_t r = 0;
    for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
#endif
}
// RFC 1951 3.2.5 in closed form: lengths 3,4,..,10, 11,13,15,17, 19,23,27,31, ... 227, 258; distances 1,2,3,4, 5,7, 9,13, ... 24577
NP_HD int32_t len_base(int s) {     // s = symbol - 257
    if (s < 8) return 3 + s;
    if (s == 28) return 258;
    return ((4 + (s & 3)) << ((s >> 2) - 1)) + 3;
}
NP_HD int32_t len_extra(int s) { return s < 8 || s == 28 ? 0 : (s - 4) >> 2; }
NP_HD int32_t dist_base(int s) { return s < 4 ? s + 1 : ((2 + (s & 1)) << ((s >> 1) - 1)) + 1; }
NP_HD int32_t dist_extra(int s) { return s < 4 ? 0 : (s >> 1) - 1; }
NP_HD uint32_t lit_entry(int sym, int l) {      // literal/length symbol with an l-bit code
    if (sym < 256) return (uint32_t)sym << 16 | (uint32_t)E_LIT | entry_bits(l, 0);
    if (sym == 256) return (uint32_t)(E_SPECIAL | E_END) | entry_bits(l, 0);
    if (sym >= 286) return (uint32_t)E_SPECIAL | entry_bits(l, 0);
    return (uint32_t)len_base(sym - 257) << 16 | entry_bits(l, len_extra(sym - 257));
}
NP_HD uint32_t dist_entry(int sym, int l) {
    if (sym >= MAXD) return (uint32_t)E_SPECIAL | entry_bits(l, 0);
    return (uint32_t)dist_base(sym) << 16 | entry_bits(l, dist_extra(sym));
}
// lookup table of the codes of up to `bits` bits: index = next `bits` stream bits (LSB first)
NP_HD void build_fast(uint32_t* fast, int bits, const uint16_t* count, const uint16_t* sym, bool dist) {
    for (int i = 0; i < (1 << bits); i++) fast[i] = (uint32_t)(E_SPECIAL | E_LONG);
    uint32_t code = 0; int idx = 0;
    for (int l = 1; l <= bits; l++) {
        for (int k = 0; k < count[l]; k++, idx++, code++) {
            uint32_t r = rev_bits(code, l);
            const uint32_t e = dist ? dist_entry(sym[idx], l) : lit_entry(sym[idx], l);
            for (uint32_t j = r; j < (1u << bits); j += 1u << l) fast[j] = e;
        }
        code <<= 1;
    }
}
NP_HD void walk_state(const uint16_t* count, int steps, int32_t& first, int32_t& index) {
    first = 0; index = 0;
    for (int l = 1; l <= steps; l++) { const int c = count[l]; index += c; first += c; first <<= 1; }
}
NP_HD void build_fast(Tables& t) {
    build_fast(t.fast, FASTBITS, t.lcount, t.lsym, false);
    build_fast(t.dfast, DFASTBITS, t.dcount, t.dsym, true);
    walk_state(t.lcount, FASTBITS, t.lfirst, t.lindex);
    walk_state(t.dcount, DFASTBITS, t.dfirst, t.dindex);
}
// canonical walk, one bit at a time (puff-style): returns the symbol or -1.
// The walk's `first` / `index` after l steps depend only on the code-length histogram, and its `code` is the bit-reversed
// l-bit prefix of the stream: a code known to be longer than `from` bits (it missed the lookup table) resumes there.
NP_HD int decode_slow(Bits& b, const uint16_t* count, const uint16_t* sym, int from = 0, int first0 = 0, int index0 = 0) {
    b.refill();
    int code = 0, first = first0, index = index0;
    uint32_t bits = b.peek(MAXBITS);
    if (from > 0) { code = (int)(rev_bits(bits, from) << 1); bits >>= from; }
    for (int l = from + 1; l <= MAXBITS; l++) {
        code |= (int)(bits & 1u); bits >>= 1;
        int c = count[l];
        if (code - c < first) { b.drop(l); return sym[index + (code - first)]; }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
}
NP_HD int32_t first_bit(uint32_t v) {     // index of the lowest set bit (v != 0)
#ifdef __CUDA_ARCH__
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
// The same walk for a code that missed the lookup table, without consuming it: symbol and code length (0: no such code).
// `bits` = the next MAXBITS stream bits.
NP_HD int walk_long(uint32_t bits, const uint16_t* count, const uint16_t* sym, int from, int first, int index, int* len) {
    int code = (int)(rev_bits(bits, from) << 1);
    bits >>= from;
    for (int l = from + 1; l <= MAXBITS; l++) {
        code |= (int)(bits & 1u); bits >>= 1;
        const int c = count[l];
        if (code - c < first) { *len = l; return sym[index + (code - first)]; }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    *len = 0;
    return -1;
}

// reads a dynamic block header into t (lane 0)
NP_HD int read_dynamic(Bits& b, Tables& t) {
    uint16_t* lens = t.lens();
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int nlen = (int)b.get(5) + 257, ndist = (int)b.get(5) + 1, ncode = (int)b.get(4) + 4;
    if (nlen > 286 || ndist > MAXD) return ERR_LENGTHS;
    for (int i = 0; i < 19; i++) lens[i] = 0;
    for (int i = 0; i < ncode; i++) lens[order[i]] = (uint16_t)b.get(3);
    // the code-length code borrows the distance table's storage while the header is being read
    uint16_t ccount[MAXBITS + 1], csym[19];
    if (build(ccount, csym, lens, 19) != 0) return ERR_LENGTHS;
    int idx = 0;
    while (idx < nlen + ndist) {
        int s = decode_slow(b, ccount, csym);
        if (s < 0) return ERR_CODE;
        if (s < 16) lens[idx++] = (uint16_t)s;
        else {
            int prev = 0, rep;
            if (s == 16) { if (idx == 0) return ERR_LENGTHS; prev = lens[idx - 1]; rep = 3 + (int)b.get(2); }
            else if (s == 17) rep = 3 + (int)b.get(3);
            else rep = 11 + (int)b.get(7);
            if (idx + rep > nlen + ndist) return ERR_LENGTHS;
            while (rep--) lens[idx++] = (uint16_t)prev;
        }
    }
    if (lens[256] == 0) return ERR_LENGTHS;
    int e = build(t.lcount, t.lsym, lens, nlen);
    if (e < 0 || (e > 0 && nlen - t.lcount[0] != 1)) return ERR_LENGTHS;
    e = build(t.dcount, t.dsym, lens + nlen, ndist);
    if (e < 0 || (e > 0 && ndist - t.dcount[0] != 1)) return ERR_LENGTHS;
    build_fast(t);
    return OK;
}
NP_HD void set_fixed(Tables& t) {
    uint16_t* lens = t.lens();
    int s = 0;
    for (; s < 144; s++) lens[s] = 8;
    for (; s < 256; s++) lens[s] = 9;
    for (; s < 280; s++) lens[s] = 7;
    for (; s < 288; s++) lens[s] = 8;
    build(t.lcount, t.lsym, lens, 288);
    for (s = 0; s < MAXD; s++) lens[s] = 5;
    build(t.dcount, t.dsym, lens, MAXD);
    build_fast(t);
}

// Resolves a batch of decoded symbols (t.batch[0..nsym)) at output position pos.  Returns the new position, or a
// negative error code.  W: warp backend (see inflate_block).
template <class W>
NP_HD int32_t resolve_batch(uint8_t* out, uint32_t out_len, int32_t pos, int32_t nsym, const Tables& t, W& w) {
    for (int32_t base = 0; base < nsym; base += w.width()) {
        const int32_t i = base + w.lane();
        const bool valid = i < nsym;
        const uint32_t e = valid ? t.batch[i] : 0u;
        const bool lit = valid && (e & 0x8000u) != 0u;
        const int32_t len = !valid ? 0 : lit ? 1 : (int32_t)(e & 0x1ffu);
        const int32_t dist = (int32_t)(e >> 16);
        int32_t total = 0;
        const int32_t dst = pos + w.exscan(len, &total);
        int32_t err = 0;
        if (valid && !lit && dist > dst) err = ERR_DIST;
        if (pos + total > (int32_t)out_len) err = ERR_OUTPUT;
        if (w.any(err != 0)) return -(w.any(err == ERR_OUTPUT) ? (int32_t)ERR_OUTPUT : (int32_t)ERR_DIST);
        bool done = !valid;
        if (lit) { out[dst] = (uint8_t)(e & 0xffu); done = true; }
        w.sync();
        for (;;) {
            const uint32_t open = w.ballot(!done);
            if (!open) break;
            const int32_t upto = w.shfl(dst, first_bit(open));       // everything before the first open symbol is final
            if (!done) {
                const int32_t src = dst - dist;
                if (dist >= len) {
                    if (src + len <= upto) {
                        int32_t j = 0;
                        for (; j + 8 <= len; j += 8) {               // loads first, then stores: eight bytes in flight
                            uint8_t v[8];
                            for (int u = 0; u < 8; u++) v[u] = out[src + j + u];
                            for (int u = 0; u < 8; u++) out[dst + j + u] = v[u];
                        }
                        for (; j < len; j++) out[dst + j] = out[src + j];
                        done = true;
                    }
                } else if (dst <= upto) {
                    // the match overlaps its own output: only the `dist` bytes before dst are read
                    for (int32_t j = 0, k = 0; j < len; j++) { out[dst + j] = out[src + k]; k = k + 1 == dist ? 0 : k + 1; }
                    done = true;
                }
            }
            w.sync();
        }
        pos += total;
    }
    return pos;
}

// ---- one decoder = one lane walking one BGZF block -------------------------------------------------------------------
// Several lanes of a warp decode DIFFERENT blocks in lockstep (the decode chain is serial per block and issue-bound:
// with one decoder per warp 31 lanes idle through ~65 warp instructions per symbol); the batches they produce are then
// resolved one block after the other by the whole warp (resolve_batch).  A decoder's state lives in its lane's
// registers, its code tables in its own Tables slot.
struct Decoder {
    Bits b; const uint8_t* in; uint8_t* out;
    uint32_t in_len, out_len;
    int32_t pos;               // output bytes resolved so far
    int32_t last;              // the current deflate block is the stream's last one
    int32_t phase;             // PH_*
    int32_t err;               // first error (OK)
};
enum { PH_IDLE = 0, PH_HEADER = 1, PH_SYMBOLS = 2, PH_DONE = 3 };
NP_HD void dec_start(Decoder& d, const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len) {
    d.b.seek(in, in_len, 0);
    d.in = in; d.out = out; d.in_len = in_len; d.out_len = out_len;
    d.pos = 0; d.last = 0; d.err = OK; d.phase = PH_HEADER;
}
// deflate block header (RFC 1951 3.2.3); a stored block is copied by the decoder's own lane (incompressible data: rare)
NP_HD void dec_header(Decoder& d, Tables& t) {
    Bits& b = d.b;
    d.last = (int32_t)b.get(1);
    const int32_t type = (int32_t)b.get(2);
    if (type == 0) {
        b.drop(b.cnt & 7);                                   // to the next byte boundary
        const uint32_t l = b.get(16), nl = b.get(16);
        if ((l ^ 0xffffu) != nl) { d.err = ERR_STORED; d.phase = PH_DONE; return; }
        // the stored bytes start at the reader's byte position (whole bytes left in the bit buffer belong to them)
        const int32_t src = (int32_t)(b.in - d.in) + (int32_t)(b.consumed_bits() >> 3), len = (int32_t)l;
        if (src + len > (int32_t)d.in_len) { d.err = ERR_INPUT; d.phase = PH_DONE; return; }
        if (d.pos + len > (int32_t)d.out_len) { d.err = ERR_OUTPUT; d.phase = PH_DONE; return; }
        for (int32_t i = 0; i < len; i++) d.out[d.pos + i] = d.in[src + i];
        d.pos += len;
        b.seek(d.in, d.in_len, (uint32_t)(src + len));
        d.phase = d.last ? PH_DONE : PH_HEADER;
        return;
    }
    if (type == 3) { d.err = ERR_BTYPE; d.phase = PH_DONE; return; }
    int e = OK;
    if (type == 1) set_fixed(t); else e = read_dynamic(b, t);
    if (e) { d.err = e; d.phase = PH_DONE; return; }
    d.phase = PH_SYMBOLS;
}
// entries of codes longer than the lookup index: the canonical walk behind the table
NP_HD uint32_t long_lit_entry(uint32_t lo, const Tables& t) {
    int l; const int sym = walk_long(lo & 0x7fffu, t.lcount, t.lsym, FASTBITS, t.lfirst, t.lindex, &l);
    return l ? lit_entry(sym, l) : (uint32_t)E_SPECIAL;
}
NP_HD uint32_t long_dist_entry(uint32_t lo, const Tables& t) {
    int l; const int sym = walk_long(lo & 0x7fffu, t.dcount, t.dsym, DFASTBITS, t.dfirst, t.dindex, &l);
    return l ? dist_entry(sym, l) : (uint32_t)E_SPECIAL;
}
// Decodes up to BATCH symbols into t.batch; returns their number.  Nothing is written to the output.
// Per symbol: one refill test (the bit buffer then holds > 32 valid bits: a literal/length code + extra bits needs <= 20),
// one table load, one drop of code and extra bits together, one flag test for everything rare (long codes, the end of the
// block, invalid symbols); a match repeats that for its distance (<= 28 bits).  The kernel is bound by instruction issue
// on this one lane, so the common path is kept a short straight line.
NP_HD int32_t dec_batch(Decoder& d, Tables& t) {
    Bits b = d.b;                                               // the reader's state in registers for the loop
    int32_t nsym = 0;
    uint32_t special = 0;
    do {
        NP_NO_UNROLL
        while (b.cnt <= 32) b.refill();                         // (loops, not ifs: the compiler keeps them as branches)
        uint32_t lo = (uint32_t)b.buf;
        uint32_t e = t.fast[lo & ((1u << FASTBITS) - 1u)];
        b.drop((int)entry_take(e));
        if (NP_UNLIKELY(e & E_SPECIAL)) {
            if (e & E_LONG) { e = long_lit_entry(lo, t); b.drop((int)entry_take(e)); }
            if (e & E_SPECIAL) { special = e; break; }
        }
        if (e & E_LIT) { t.batch[nsym++] = 0x8000u | e >> 16; continue; }
        const uint32_t len = entry_value(e, lo);
        NP_NO_UNROLL
        while (NP_UNLIKELY(b.cnt < 28)) b.refill();             // a distance code and its extra bits: 28 bits at most
        lo = (uint32_t)b.buf;
        e = t.dfast[lo & ((1u << DFASTBITS) - 1u)];
        if (NP_UNLIKELY(e & E_SPECIAL)) {
            if (e & E_LONG) e = long_dist_entry(lo, t);
            if (e & E_SPECIAL) { special = E_SPECIAL; break; }
        }
        b.drop((int)entry_take(e));
        t.batch[nsym++] = len | entry_value(e, lo) << 16;
    } while (nsym < BATCH);
    d.b = b;
    bool end = false;
    if (special) { if (special & E_END) end = true; else d.err = ERR_CODE; }
    if (!d.err && b.past_end()) d.err = ERR_INPUT;
    if (d.err) { d.phase = PH_DONE; return 0; }              // an inconsistent batch is not resolved
    if (end) d.phase = d.last ? PH_DONE : PH_HEADER;
    return nsym;
}
NP_HD int dec_status(const Decoder& d) { return d.err ? d.err : d.pos == (int32_t)d.out_len ? OK : ERR_SIZE; }

// Inflates one BGZF block with ONE decoder (lane 0) — the test build's driver (tests/emu) and the reference for the
// kernel's multi-decoder loop (bgzf_inflate.cu), which runs the same dec_* / resolve_batch steps.
// W: warp backend with lane(), width(), bcast(int32_t v) (value of lane 0), sync(), exscan(v, &total) (exclusive prefix
// sum over the lanes), any(pred), ballot(pred), shfl(v, lane).  Returns OK or an ERR_* (same value on every lane).
template <class W>
NP_HD int inflate_block(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len, Tables& t, W& w) {
    const bool lead = w.lane() == 0;
    Decoder d;
    dec_start(d, in, in_len, out, out_len);
    for (;;) {
        int32_t nsym = 0;
        if (lead) {
            while (d.phase == PH_HEADER) dec_header(d, t);
            if (d.phase == PH_SYMBOLS) nsym = dec_batch(d, t);
        }
        nsym = w.bcast(nsym);
        const int32_t pos = w.bcast(d.pos);
        w.sync();                                                // the batch (and a stored block's bytes) are visible to every lane
        if (nsym > 0) {
            const int32_t np = resolve_batch(out, out_len, pos, nsym, t, w);
            if (lead) { if (np < 0) { d.err = -np; d.phase = PH_DONE; } else d.pos = np; }
        }
        if (w.bcast(lead ? d.phase : 0) == PH_DONE) break;
    }
    return w.bcast(lead ? dec_status(d) : 0);
}

}  // namespace npz
