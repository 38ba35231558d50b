// bgzf_inflate.h — DEFLATE (RFC 1951) decoder for BGZF blocks, one WARP per block.
//
// the step immediately before the polishing path.  The reference inflates the BAM on
// the host through htslib (bgzf.c -> zlib inflate), ~67 MB/s of compressed input per core, twice per
// contig (contig.c:170-180 and :688-704); BGZF blocks are independent <= 64 KiB deflate members, so a
// whole BAM region inflates as thousands of independent streams.
//
// Division of labour inside a warp: lane 0 owns the bit reader and decodes Huffman symbols (a strictly
// sequential dependency chain); literals are stored by lane 0, LZ77 matches are copied by all lanes
// (the copy is the byte-heavy part).  Code tables live in shared memory, one small region per warp:
//   literal/length and distance symbols ordered by code (canonical Huffman, RFC 1951 3.2.2), code-length
//   histograms, and a 9-bit lookup table for literal/length codes (one shared-memory load decodes the
//   common symbols; longer codes fall back to the canonical walk).
//
// The decoder is NP_HD code over a warp backend (lane id, broadcast, barrier) so that tests/emu can run
// it on the CPU with a 1-lane backend and compare with zlib; the product compiles it into
// k_bgzf_inflate (bgzf_inflate.cu).  Nothing here is a CPU fallback of the product.
#pragma once
#include <stdint.h>
#include "device_logic.h"

namespace npz {

enum { OK = 0, ERR_BTYPE = 1, ERR_STORED = 2, ERR_LENGTHS = 3, ERR_CODE = 4, ERR_DIST = 5, ERR_OUTPUT = 6,
       ERR_INPUT = 7, ERR_SIZE = 8 };

struct Block {             // one BGZF block: where its raw deflate payload lives, where its output goes
    uint64_t in_off;       // offset of the deflate payload (after the gzip header) in the compressed buffer
    uint32_t in_len;       // payload bytes (BSIZE + 1 - header - 8 trailer bytes)
    uint32_t out_len;      // ISIZE
    uint64_t out_off;      // offset of the block's bytes in the output buffer
};

enum { MAXBITS = 15, MAXL = 288, MAXD = 30, FASTBITS = 9, DFASTBITS = 8 };

struct Tables {            // per-warp shared memory (2.9 KB)
    uint16_t lcount[MAXBITS + 1], lsym[MAXL];
    uint16_t dcount[MAXBITS + 1], dsym[MAXD + 2];
    uint16_t fast[1 << FASTBITS];          // literal/length: symbol << 4 | code length (0: not in the table)
    uint16_t dfast[1 << DFASTBITS];        // distance: likewise
    uint16_t lens[MAXL + MAXD + 2];        // scratch: code lengths while a dynamic header is read
};

struct Bits {              // LSB-first bit reader over the compressed bytes (lane 0 only)
    const uint8_t* p; const uint8_t* end;
    uint64_t buf; int32_t cnt; int32_t overrun;
    NP_HD static uint32_t load32(const uint8_t* q) {     // 4 bytes at any alignment, little endian
#ifdef __CUDA_ARCH__
        const uint32_t* a = (const uint32_t*)((uintptr_t)q & ~(uintptr_t)3);   // two aligned words + funnel shift; the
        return __funnelshift_r(a[0], a[1], 8u * (uint32_t)((uintptr_t)q & 3u)); // second word may lie past q+4 (buffer slack)
#else
        return (uint32_t)q[0] | (uint32_t)q[1] << 8 | (uint32_t)q[2] << 16 | (uint32_t)q[3] << 24;
#endif
    }
    NP_HD void refill() {
        if (cnt <= 32 && p + 4 <= end) { buf |= (uint64_t)load32(p) << cnt; p += 4; cnt += 32; return; }
        while (cnt <= 56) {
            if (p < end) buf |= (uint64_t)(*p++) << cnt;
            else overrun += 8;                   // zeros past the end: an error only if they get consumed
            cnt += 8;
        }
    }
    NP_HD uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1ull)); }
    NP_HD void drop(int n) { buf >>= n; cnt -= n; }
    NP_HD uint32_t get(int n) { if (cnt < n) refill(); uint32_t v = peek(n); drop(n); return v; }
    NP_HD bool past_end() const { return overrun > cnt; }      // consumed bits that were never there
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
// lookup table of the codes of up to `bits` bits: index = next `bits` stream bits (LSB first)
NP_HD void build_fast(uint16_t* fast, int bits, const uint16_t* count, const uint16_t* sym) {
    for (int i = 0; i < (1 << bits); i++) fast[i] = 0;
    uint32_t code = 0; int idx = 0;
    for (int l = 1; l <= bits; l++) {
        for (int k = 0; k < count[l]; k++, idx++, code++) {
            uint32_t r = rev_bits(code, l);
            uint16_t e = (uint16_t)((uint32_t)sym[idx] << 4 | (uint32_t)l);
            for (uint32_t j = r; j < (1u << bits); j += 1u << l) fast[j] = e;
        }
        code <<= 1;
    }
}
NP_HD void build_fast(Tables& t) {
    build_fast(t.fast, FASTBITS, t.lcount, t.lsym);
    build_fast(t.dfast, DFASTBITS, t.dcount, t.dsym);
}
// canonical walk, one bit at a time (puff-style): returns the symbol or -1
NP_HD int decode_slow(Bits& b, const uint16_t* count, const uint16_t* sym) {
    if (b.cnt < MAXBITS) b.refill();
    int code = 0, first = 0, index = 0;
    uint32_t bits = b.peek(MAXBITS);
    for (int l = 1; l <= MAXBITS; l++) {
        code |= (int)(bits & 1u); bits >>= 1;
        int c = count[l];
        if (code - c < first) { b.drop(l); return sym[index + (code - first)]; }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
}
NP_HD int decode_dist(Bits& b, const Tables& t) {
    if (b.cnt < MAXBITS) b.refill();
    uint32_t e = t.dfast[b.peek(DFASTBITS)];
    if (e & 15u) { b.drop((int)(e & 15u)); return (int)(e >> 4); }
    return decode_slow(b, t.dcount, t.dsym);
}
NP_HD int decode_lit(Bits& b, const Tables& t) {
    if (b.cnt < MAXBITS) b.refill();
    uint32_t e = t.fast[b.peek(FASTBITS)];
    if (e & 15u) { b.drop((int)(e & 15u)); return (int)(e >> 4); }
    return decode_slow(b, t.lcount, t.lsym);
}

// RFC 1951 3.2.5 tables in closed form (a local array would be rebuilt on the stack at every call):
// lengths 3,4,..,10, 11,13,15,17, 19,23,27,31, ... 227, 258; distances 1,2,3,4, 5,7, 9,13, 17,25, ... 24577
NP_HD int32_t len_base(int s) {     // s = symbol - 257
    if (s < 8) return 3 + s;
    if (s == 28) return 258;
    return ((4 + (s & 3)) << ((s >> 2) - 1)) + 3;
}
NP_HD int32_t len_extra(int s) { return s < 8 || s == 28 ? 0 : (s - 4) >> 2; }
NP_HD int32_t dist_base(int s) { return s < 4 ? s + 1 : ((2 + (s & 1)) << ((s >> 1) - 1)) + 1; }
NP_HD int32_t dist_extra(int s) { return s < 4 ? 0 : (s >> 1) - 1; }

// reads a dynamic block header into t (lane 0)
NP_HD int read_dynamic(Bits& b, Tables& t) {
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int nlen = (int)b.get(5) + 257, ndist = (int)b.get(5) + 1, ncode = (int)b.get(4) + 4;
    if (nlen > 286 || ndist > MAXD) return ERR_LENGTHS;
    for (int i = 0; i < 19; i++) t.lens[i] = 0;
    for (int i = 0; i < ncode; i++) t.lens[order[i]] = (uint16_t)b.get(3);
    // the code-length code borrows the distance table's storage while the header is being read
    uint16_t ccount[MAXBITS + 1], csym[19];
    if (build(ccount, csym, t.lens, 19) != 0) return ERR_LENGTHS;
    int idx = 0;
    while (idx < nlen + ndist) {
        int s = decode_slow(b, ccount, csym);
        if (s < 0) return ERR_CODE;
        if (s < 16) t.lens[idx++] = (uint16_t)s;
        else {
            int prev = 0, rep;
            if (s == 16) { if (idx == 0) return ERR_LENGTHS; prev = t.lens[idx - 1]; rep = 3 + (int)b.get(2); }
            else if (s == 17) rep = 3 + (int)b.get(3);
            else rep = 11 + (int)b.get(7);
            if (idx + rep > nlen + ndist) return ERR_LENGTHS;
            while (rep--) t.lens[idx++] = (uint16_t)prev;
        }
    }
    if (t.lens[256] == 0) return ERR_LENGTHS;
    int e = build(t.lcount, t.lsym, t.lens, nlen);
    if (e < 0 || (e > 0 && nlen - t.lcount[0] != 1)) return ERR_LENGTHS;
    e = build(t.dcount, t.dsym, t.lens + nlen, ndist);
    if (e < 0 || (e > 0 && ndist - t.dcount[0] != 1)) return ERR_LENGTHS;
    build_fast(t);
    return OK;
}
NP_HD void set_fixed(Tables& t) {
    int s = 0;
    for (; s < 144; s++) t.lens[s] = 8;
    for (; s < 256; s++) t.lens[s] = 9;
    for (; s < 280; s++) t.lens[s] = 7;
    for (; s < 288; s++) t.lens[s] = 8;
    build(t.lcount, t.lsym, t.lens, 288);
    for (s = 0; s < MAXD; s++) t.lens[s] = 5;
    build(t.dcount, t.dsym, t.lens, MAXD);
    build_fast(t);
}

// Inflates one BGZF block.  W: warp backend with lane(), width(), bcast(int32_t v) (value of lane 0) and sync().
// (A shared-memory mirror of the recent output was tried for near matches and measured slower: the kernel is bound by
// instruction issue — 65 warp instructions per symbol with ~5 active lanes — not by the L2 round trip of the copies.)
// Returns OK or an ERR_* (same value on every lane).
template <class W>
NP_HD int inflate_block(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len, Tables& t, W& w) {
    const bool lead = w.lane() == 0;
    Bits b{in, in + in_len, 0ull, 0, 0};
    int32_t pos = 0;                  // output bytes written so far (kept identical on all lanes)
    int32_t last = 0;
    while (!last) {
        int32_t type = 0, err = OK;
        if (lead) { last = (int32_t)b.get(1); type = (int32_t)b.get(2); }
        last = w.bcast(last); type = w.bcast(type);
        if (type == 0) {
            // stored block: LEN, ~LEN, bytes — copied by all lanes
            int32_t len = 0, src = 0;
            if (lead) {
                b.drop(b.cnt & 7);                                   // to the next byte boundary
                uint32_t l = b.get(16), nl = b.get(16);
                if ((l ^ 0xffffu) != nl) err = ERR_STORED;
                // bytes still sitting in the bit buffer belong to the stored data: give them back
                int32_t back = (b.cnt - b.overrun) / 8; if (back < 0) back = 0;
                src = (int32_t)(b.p - in) - back;
                len = (int32_t)l;
                if (src + len > (int32_t)in_len) err = ERR_INPUT;
                if (pos + len > (int32_t)out_len) err = ERR_OUTPUT;
                b.p = in + src + (err ? 0 : len); b.buf = 0; b.cnt = 0; b.overrun = 0;
            }
            err = w.bcast(err); if (err) return err;
            len = w.bcast(len); src = w.bcast(src);
            for (int32_t i = w.lane(); i < len; i += w.width()) out[pos + i] = in[src + i];
            pos += len;
            w.sync();
            continue;
        }
        if (type == 3) return ERR_BTYPE;
        if (lead) { if (type == 1) set_fixed(t); else err = read_dynamic(b, t); }
        err = w.bcast(err); if (err) return err;
        for (;;) {
            // lane 0 runs through literals until it meets a match or the end of the block
            int32_t len = 0, dist = 0, p = pos, state = 0;          // state: 0 match, 1 end of block, >1 error
            if (lead) {
                for (;;) {
                    int s = decode_lit(b, t);
                    if (s < 0) { state = 1 + ERR_CODE; break; }
                    if (s < 256) {
                        if (p >= (int32_t)out_len) { state = 1 + ERR_OUTPUT; break; }
                        out[p++] = (uint8_t)s;
                        continue;
                    }
                    if (s == 256) { state = 1; break; }
                    s -= 257;
                    if (s >= 29) { state = 1 + ERR_CODE; break; }
                    len = len_base(s) + (int32_t)b.get(len_extra(s));
                    int ds = decode_dist(b, t);
                    if (ds < 0 || ds >= MAXD) { state = 1 + ERR_CODE; break; }
                    dist = dist_base(ds) + (int32_t)b.get(dist_extra(ds));
                    if (dist > p) state = 1 + ERR_DIST;
                    else if (p + len > (int32_t)out_len) state = 1 + ERR_OUTPUT;
                    break;
                }
                if (b.past_end()) state = 1 + ERR_INPUT;
            }
            state = w.bcast(state);
            p = w.bcast(p);
            if (state > 1) return state - 1;
            pos = p;
            if (state == 1) break;
            len = w.bcast(len); dist = w.bcast(dist);
            w.sync();                                                // lane 0's literals are visible to the copiers
            // source index of byte i: i for disjoint ranges, i mod dist when the match overlaps its own output (only
            // the `dist` bytes before pos are read then) — no byte is both read and written within one match
            if (dist >= len) { for (int32_t i = w.lane(); i < len; i += w.width()) out[pos + i] = out[pos - dist + i]; }
            else { for (int32_t i = w.lane(); i < len; i += w.width()) out[pos + i] = out[pos - dist + i % dist]; }
            pos += len;
            w.sync();
        }
    }
    return pos == (int32_t)out_len ? OK : ERR_SIZE;
}

}  // namespace npz
