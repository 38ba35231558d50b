// synth.cpp — seeded synthetic polishing inputs: a truth genome, a draft
// derived from it by SNV / short-indel errors, and paired short reads drawn from the truth whose
// alignments to the DRAFT are emitted analytically from the known edit script (no aligner).
// Output: draft FASTA + coordinate-sorted BAM (own BGZF writer), or the packed shard directly.
// Used by bench.py, the tests and the CLI's `simulate` command; not part of the polishing path.
#include "hostio.h"
#include "errors.h"

#include <zlib.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) { next(); next(); }
    inline uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    inline double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    inline uint32_t below(uint32_t n) { return (uint32_t)((next() >> 32) * (uint64_t)n >> 32); }
    inline double normal() { double u = uni(), v = uni(); if (u < 1e-300) u = 1e-300; return std::sqrt(-2.0 * std::log(u)) * std::cos(6.283185307179586 * v); }
};

static const char kBases[4] = {'A', 'C', 'G', 'T'};
static inline uint8_t nt16(char c) { switch (c) { case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8; default: return 15; } }

struct ContigSim {
    std::string name, draft;              // draft with lowercase marks
    std::vector<uint8_t> recs;            // BAM records (block_size prefixed), unsorted
    std::vector<std::pair<int32_t, uint32_t>> order;   // (pos, offset into recs), sorted by pos
};

static inline int reg2bin(int64_t beg, int64_t end) {   // SAM spec 5.3
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

template <class T> static inline void put(std::vector<uint8_t>& v, T x) { size_t o = v.size(); v.resize(o + sizeof(T)); memcpy(v.data() + o, &x, sizeof(T)); }

static void simulate_contig(const np_synth_params& P, int32_t idx, int64_t L, ContigSim& out) {
    Rng rng(P.seed * 1000003ull + (uint64_t)idx * 7919ull + 17);
    char nm[64]; snprintf(nm, sizeof nm, "ctg%06d", idx);
    out.name = nm;
    // ---- truth: i.i.d. ACGT (GC 0.41) with ~2 % of bases inside homopolymer / dinucleotide runs
    std::string truth((size_t)L, 'A');
    for (int64_t i = 0; i < L;) {
        double u = rng.uni();
        if (u < 0.0025 && i + 16 < L) {           // low-complexity run, mean length ~8
            int len = 4 + (int)rng.below(10);
            char a = kBases[rng.below(4)], b = rng.uni() < 0.7 ? a : kBases[rng.below(4)];
            for (int k = 0; k < len && i < L; k++, i++) truth[(size_t)i] = (k & 1) ? b : a;
        } else {
            double g = rng.uni();
            truth[(size_t)i++] = g < 0.295 ? 'A' : g < 0.59 ? 'T' : g < 0.795 ? 'G' : 'C';
        }
    }
    // ---- draft = truth + errors; dpos[i] = draft coordinate of truth base i (or -1), dextra[i] =
    //      draft-only bases inserted before truth base i
    std::vector<int32_t> dpos((size_t)L, -1);
    std::vector<uint8_t> dextra((size_t)L, 0);
    std::string draft; draft.reserve((size_t)L + (size_t)L / 100 + 16);
    std::vector<int32_t> errsites;
    for (int64_t i = 0; i < L; i++) {
        double u = rng.uni();
        if (i > 2 && i + 4 < L && u < P.draft_indel * 0.5) {            // draft lost 1-3 truth bases
            int len = 1 + (int)rng.below(3);
            errsites.push_back((int32_t)draft.size());
            i += len - 1;                                               // those truth bases keep dpos = -1
            continue;
        }
        if (i > 2 && i + 4 < L && u < P.draft_indel) {                  // draft gained 1-3 bases
            int len = 1 + (int)rng.below(3);
            dextra[(size_t)i] = (uint8_t)len;
            errsites.push_back((int32_t)draft.size());
            for (int k = 0; k < len; k++) draft.push_back(kBases[rng.below(4)]);
        }
        char b = truth[(size_t)i];
        if (u >= P.draft_indel && u < P.draft_indel + P.draft_snv) {
            char c; do { c = kBases[rng.below(4)]; } while (c == b);
            b = c;
            errsites.push_back((int32_t)draft.size());
        }
        dpos[(size_t)i] = (int32_t)draft.size();
        draft.push_back(b);
    }
    const int64_t DL = (int64_t)draft.size();
    // ---- lowercase marks (task-2 style input): around some error sites and at random
    if (P.lowercase_frac > 0) {
        for (int32_t s : errsites) if (rng.uni() < 0.5) {
            int a = (int)rng.below(3), b = (int)rng.below(4);
            for (int64_t q = std::max<int64_t>(0, s - a); q <= std::min<int64_t>(DL - 1, s + b); q++) draft[(size_t)q] = (char)tolower(draft[(size_t)q]);
        }
        for (int64_t q = 0; q < DL; q++) if (rng.uni() < P.lowercase_frac / 3.0) {
            int len = 1 + (int)rng.below(6);
            for (int k = 0; k < len && q < DL; k++, q++) draft[(size_t)q] = (char)tolower(draft[(size_t)q]);
        }
    }
    out.draft = draft;
    // ---- paired reads from the truth
    const int RL = P.read_len;
    int64_t npairs = (int64_t)(P.depth * (double)L / (2.0 * RL) + 0.5);
    out.recs.reserve((size_t)npairs * 2 * (size_t)(60 + RL + RL / 2));
    std::vector<uint8_t> qop; std::vector<int32_t> qdp; std::string seq, qual;
    std::vector<uint32_t> cigar;
    struct Mate { int32_t pos, end; size_t rec_off; bool ok; };
    for (int64_t pr = 0; pr < npairs; pr++) {
        int frag = (int)(350 + 35 * rng.normal());
        if (frag < RL + 10) frag = RL + 10;
        if (frag >= L) frag = (int)L - 1;
        if (frag < RL) continue;
        int64_t fs = (int64_t)rng.below((uint32_t)(L - frag));
        uint8_t mapq = rng.uni() < 0.03 ? (uint8_t)rng.below(41) : 60;
        Mate m[2];
        for (int mate = 0; mate < 2; mate++) {
            int64_t a = mate == 0 ? fs : fs + frag - RL;
            // per read base: op (0 M, 1 I, 4 S) and draft position; D ops are derived from gaps
            seq.clear(); qual.clear(); qop.clear(); qdp.clear();
            for (int64_t i = a; i < a + RL && i < L; i++) {
                char b = truth[(size_t)i];
                double u = rng.uni();
                if (u < P.read_indel * 0.5) continue;                       // read lost this base
                if (u < P.read_indel) {                                     // read gained a base before it
                    seq.push_back(kBases[rng.below(4)]); qop.push_back(1); qdp.push_back(-1);
                } else if (u < P.read_indel + P.read_sub) {
                    char c; do { c = kBases[rng.below(4)]; } while (c == b); b = c;
                }
                seq.push_back(b);
                if (dpos[(size_t)i] >= 0) { qop.push_back(0); qdp.push_back(dpos[(size_t)i]); }
                else { qop.push_back(1); qdp.push_back(-1); }
            }
            int n = (int)seq.size();
            // optional soft clip at one end (2 % of reads)
            int lead_clip = 0, trail_clip = 0;
            if (rng.uni() < 0.02) { int k = 5 + (int)rng.below(36); if (rng.uni() < 0.5) lead_clip = k; else trail_clip = k; }
            // alignment must start and end on an M base
            int s0 = lead_clip, s1 = n - 1 - trail_clip;
            while (s0 < n && qop[(size_t)s0] != 0) s0++;
            while (s1 >= 0 && qop[(size_t)s1] != 0) s1--;
            m[mate].ok = false;
            if (s0 > s1 || n < 30) continue;
            cigar.clear();
            auto push = [&](uint32_t op, uint32_t len) {
                if (!len) return;
                if (!cigar.empty() && (cigar.back() & 0xf) == op) cigar.back() += len << 4; else cigar.push_back(len << 4 | op);
            };
            push(4, (uint32_t)s0);
            int32_t prev = -1;
            for (int q = s0; q <= s1; q++) {
                if (qop[(size_t)q] == 0) {
                    if (prev >= 0 && qdp[(size_t)q] > prev + 1) push(2, (uint32_t)(qdp[(size_t)q] - prev - 1));
                    push(0, 1); prev = qdp[(size_t)q];
                } else push(1, 1);
            }
            push(4, (uint32_t)(n - 1 - s1));
            int32_t pos = qdp[(size_t)s0], end = prev + 1;
            qual.resize((size_t)n);
            for (int q = 0; q < n; q++) { double u = rng.uni(); qual[(size_t)q] = (char)(u < 0.85 ? 37 : u < 0.95 ? 22 : 12); }
            // ---- BAM record
            char qn[32]; int ql = snprintf(qn, sizeof qn, "p%lld", (long long)pr) + 1;
            std::vector<uint8_t>& v = out.recs;
            size_t ro = v.size();
            int32_t bs = 32 + ql + 4 * (int)cigar.size() + (n + 1) / 2 + n;
            put<int32_t>(v, bs);
            put<int32_t>(v, idx - 0);            // refID patched by the writer (contig index base)
            put<int32_t>(v, pos);
            v.push_back((uint8_t)ql); v.push_back(mapq);
            put<uint16_t>(v, (uint16_t)reg2bin(pos, end));
            put<uint16_t>(v, (uint16_t)cigar.size());
            put<uint16_t>(v, 0);                 // flag, patched below
            put<int32_t>(v, n);
            put<int32_t>(v, idx);                // next refID
            put<int32_t>(v, 0);                  // next pos, patched below
            put<int32_t>(v, 0);                  // tlen, patched below
            v.insert(v.end(), qn, qn + ql);
            for (uint32_t c : cigar) put<uint32_t>(v, c);
            for (int q = 0; q < n; q += 2) v.push_back((uint8_t)(nt16(seq[(size_t)q]) << 4 | (q + 1 < n ? nt16(seq[(size_t)q + 1]) : 0)));
            v.insert(v.end(), qual.begin(), qual.end());
            m[mate] = {pos, end, ro, true};
        }
        for (int mate = 0; mate < 2; mate++) {
            if (!m[mate].ok) continue;
            const Mate& o = m[1 - mate];
            uint16_t flag = 0x1 | (mate == 0 ? 0x40 : 0x80) | (mate == 0 ? 0x20 : 0x10);
            int32_t mpos = 0, tlen = 0;
            if (o.ok) { flag |= 0x2; mpos = o.pos; int32_t lo = std::min(m[0].pos, m[1].pos), hi = std::max(m[0].end, m[1].end); tlen = mate == 0 ? hi - lo : -(hi - lo); }
            else { flag |= 0x8; flag &= (uint16_t)~0x20; }
            uint8_t* p = out.recs.data() + m[mate].rec_off + 4;
            memcpy(p + 14, &flag, 2); memcpy(p + 24, &mpos, 4); memcpy(p + 28, &tlen, 4);
            out.order.emplace_back(m[mate].pos, (uint32_t)m[mate].rec_off);
        }
    }
    std::stable_sort(out.order.begin(), out.order.end(), [](const std::pair<int32_t, uint32_t>& a, const std::pair<int32_t, uint32_t>& b) { return a.first < b.first; });
}

static std::vector<int64_t> contig_lengths(const np_synth_params& P) {
    std::vector<int64_t> len((size_t)P.n_contigs);
    Rng rng(P.seed ^ 0xC0FFEEull);
    for (int i = 0; i < P.n_contigs; i++) {
        if (P.min_len > 0 && P.max_len >= P.min_len) {
            double lo = std::log((double)P.min_len), hi = std::log((double)P.max_len);
            len[(size_t)i] = (int64_t)std::exp(lo + (hi - lo) * rng.uni());
        } else len[(size_t)i] = P.contig_len;
    }
    return len;
}

// ---- BGZF writer ----------------------------------------------------------------------------
struct BgzfWriter {
    FILE* f; int level; std::vector<uint8_t> buf; bool ok = true;
    void flush_block(const uint8_t* p, size_t n) {
        uint8_t out[70000];
        z_stream zs; memset(&zs, 0, sizeof zs);
        deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        zs.next_in = const_cast<Bytef*>(p); zs.avail_in = (uInt)n;
        zs.next_out = out + 18; zs.avail_out = sizeof(out) - 18 - 8;
        int rc = deflate(&zs, Z_FINISH);
        size_t clen = zs.total_out;
        deflateEnd(&zs);
        if (rc != Z_STREAM_END) { ok = false; return; }
        static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
        memcpy(out, hdr, 16);
        uint16_t bsize = (uint16_t)(clen + 25);
        memcpy(out + 16, &bsize, 2);
        uint32_t crc = (uint32_t)crc32(crc32(0, nullptr, 0), p, (uInt)n), isz = (uint32_t)n;
        memcpy(out + 18 + clen, &crc, 4); memcpy(out + 22 + clen, &isz, 4);
        if (fwrite(out, 1, clen + 26, f) != clen + 26) ok = false;
    }
    void write(const void* p, size_t n) {
        const uint8_t* q = (const uint8_t*)p;
        while (n) {
            size_t take = std::min(n, (size_t)0xff00 - buf.size());
            buf.insert(buf.end(), q, q + take); q += take; n -= take;
            if (buf.size() == 0xff00) { flush_block(buf.data(), buf.size()); buf.clear(); }
        }
    }
    void finish() {
        if (!buf.empty()) { flush_block(buf.data(), buf.size()); buf.clear(); }
        flush_block(nullptr, 0);   // EOF marker block
    }
};

static void run_parallel(int n, int threads, const std::function<void(int)>& fn) {
    std::atomic<int> next(0);
    auto work = [&]() { for (;;) { int i = next.fetch_add(1); if (i >= n) break; fn(i); } };
    int nt = std::max(1, std::min(threads, n));
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}

}  // namespace

namespace np {

// Packed shard straight from the generator (contigs [lo,hi) of the synthetic genome).
bool synth_shard(const np_synth_params& P, int32_t lo, int32_t hi, int with_qual, int threads, Shard& out, std::string& err) {
    std::vector<int64_t> len = contig_lengths(P);
    if (lo < 0 || hi > P.n_contigs || lo > hi) { err = "synth_shard: bad contig range"; return false; }
    int n = hi - lo;
    std::vector<Shard> parts((size_t)n);
    std::vector<std::string> errs((size_t)n);
    run_parallel(n, threads, [&](int k) {
        ContigSim cs;
        simulate_contig(P, lo + k, len[(size_t)(lo + k)], cs);
        Shard& s = parts[(size_t)k];
        s.with_qual = with_qual != 0;
        s.qual_mode = with_qual;
        s.names.push_back(cs.name);
        s.ctg_seq.assign(cs.draft.begin(), cs.draft.end());
        s.begin_contig(s.ctg_seq.data(), s.ctg_seq.size());
        s.rec_off.push_back(0); if (with_qual) s.qual_off.push_back(0);
        for (auto& o : cs.order) {
            const uint8_t* p = cs.recs.data() + o.second + 4;
            BamRec r;
            memcpy(&r.tid, p, 4); memcpy(&r.pos, p + 4, 4);
            uint8_t l_name = p[8]; r.mapq = p[9];
            uint16_t nc, fl; memcpy(&nc, p + 12, 2); memcpy(&fl, p + 14, 2);
            r.n_cigar = nc; r.flag = fl;
            memcpy(&r.l_qseq, p + 16, 4); memcpy(&r.isize, p + 28, 4);
            const uint8_t* q = p + 32 + l_name;
            r.cigar = (const uint32_t*)q; r.seq = q + 4 * (size_t)nc; r.qual = r.seq + ((size_t)r.l_qseq + 1) / 2;
            if (!shard_pack_record(r, s, errs[(size_t)k])) return;
        }
    });
    for (auto& e : errs) if (!e.empty()) { err = e; return false; }
    out = Shard();
    out.with_qual = with_qual != 0;
    out.qual_mode = with_qual;
    out.ctg_off.push_back(0); out.ctg_read_off.push_back(0); out.rec_off.push_back(0);
    if (with_qual) out.qual_off.push_back(0);
    for (auto& s : parts) {
        out.names.push_back(s.names[0]);
        out.fasta_rank.push_back((int32_t)out.fasta_rank.size());
        out.alg_bytes += s.alg_bytes; out.qual_bytes += s.qual_bytes;
        out.ctg_seq.insert(out.ctg_seq.end(), s.ctg_seq.begin(), s.ctg_seq.end());
        out.ctg_off.push_back((int64_t)out.ctg_seq.size());
        uint32_t rb = (uint32_t)(out.rec.size() / 16), qb = (uint32_t)(out.qual.size() / 16);
        out.rec.insert(out.rec.end(), s.rec.begin(), s.rec.end());
        for (size_t i = 1; i < s.rec_off.size(); i++) out.rec_off.push_back(rb + s.rec_off[i]);
        if (with_qual) {
            out.qual.insert(out.qual.end(), s.qual.begin(), s.qual.end());
            for (size_t i = 1; i < s.qual_off.size(); i++) out.qual_off.push_back(qb + s.qual_off[i]);
        }
        out.ctg_read_off.push_back((int64_t)out.rec_off.size() - 1);
        s = Shard();
    }
    return true;
}

}  // namespace np

extern "C" int32_t np_synth_write(const np_synth_params* P, const char* fasta_path, const char* bam_path) {
    if (!P || !fasta_path || !bam_path || P->n_contigs <= 0) { np::set_error("np_synth_write: bad arguments"); return NP_ERR_ARG; }
    std::vector<int64_t> len = contig_lengths(*P);
    FILE* fa = fopen(fasta_path, "wb");
    FILE* fb = fopen(bam_path, "wb");
    if (!fa || !fb) { np::set_error("np_synth_write: cannot open outputs"); if (fa) fclose(fa); if (fb) fclose(fb); return NP_ERR_IO; }
    BgzfWriter w{fb, P->compress_level};
    // header
    std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    std::vector<std::string> names;
    for (int i = 0; i < P->n_contigs; i++) { char nm[64]; snprintf(nm, sizeof nm, "ctg%06d", i); names.push_back(nm); }
    // draft lengths are only known after simulation: simulate in batches, header needs lengths first
    // -> two passes over the (deterministic) generator would double the cost; instead keep every
    //    contig's records in memory batch by batch and write the header once all drafts are known.
    std::vector<ContigSim> sims((size_t)P->n_contigs);
    int threads = (int)std::max(1u, std::thread::hardware_concurrency());
    run_parallel(P->n_contigs, threads, [&](int k) { simulate_contig(*P, k, len[(size_t)k], sims[(size_t)k]); });
    for (int i = 0; i < P->n_contigs; i++) text += "@SQ\tSN:" + names[(size_t)i] + "\tLN:" + std::to_string(sims[(size_t)i].draft.size()) + "\n";
    w.write("BAM\1", 4);
    int32_t lt = (int32_t)text.size(); w.write(&lt, 4); w.write(text.data(), text.size());
    int32_t nref = P->n_contigs; w.write(&nref, 4);
    for (int i = 0; i < P->n_contigs; i++) {
        int32_t ln = (int32_t)names[(size_t)i].size() + 1; w.write(&ln, 4); w.write(names[(size_t)i].c_str(), (size_t)ln);
        int32_t L = (int32_t)sims[(size_t)i].draft.size(); w.write(&L, 4);
    }
    for (int i = 0; i < P->n_contigs; i++) {
        ContigSim& cs = sims[(size_t)i];
        fprintf(fa, ">%s\n", cs.name.c_str());
        fwrite(cs.draft.data(), 1, cs.draft.size(), fa);
        fputc('\n', fa);
        for (auto& o : cs.order) {
            const uint8_t* p = cs.recs.data() + o.second;
            int32_t bs; memcpy(&bs, p, 4);
            w.write(p, (size_t)bs + 4);
        }
        cs = ContigSim();
    }
    w.finish();
    bool ok = w.ok && !ferror(fa) && !ferror(fb);
    fclose(fa); fclose(fb);
    if (!ok) { np::set_error("np_synth_write: write failed"); return NP_ERR_IO; }
    return NP_OK;
}
