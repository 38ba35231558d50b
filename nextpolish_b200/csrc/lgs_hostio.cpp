// lgs_hostio.cpp — the stage in front of the long-read first pass, on the host: BAM records of one contig -> the clipped,
// anchored alignment strings of its consensus windows (np2_window_batch, include/nextpolish2_b200.h).  Restates the record
// loop of ctg_cns_core (source/lib/ctg_cns.c:3444-3566) for the part that does not need the large-indel machinery:
//   window geometry   cal_win_len :2800, loop :3455-3457,3594      record filters   :3478-3515 (flags 0xD04, aligned fraction)
//   split-read gap    set_satags :2158, check_indel :2463         strings          bam2aln :2403
//   clipping          clip_aln :2809                               anchoring        get_align_shift :139 (exact 8-mers)
//   coverage caps     :3544-3545 (needs the running coverage of get_align_tags :1232-1234)
//   the draft as the reference sees it: read_ref's 2-bit packing and unpacking (bseq.c:87-123, nt_table :7-16) — a base that
//   is not A/C/G/T/U becomes 'A' and sets the low bit of the base before it inside the same 16-base word.
// A contig longer than INS_MIN_CHECK_LEN (100 kb) on which a supplementary / secondary record carries a split-read gap would
// switch the reference's large-indel path on (:3503-3504,3567-3583): not built — the loader reports it (code -10).
// The clipping and anchoring steps keep the reference's exact index arithmetic on purpose (its "too short" branch, the order in
// which start / end / length are adjusted, unsigned comparisons): every quirk changes which reads enter a window, so
// those two functions follow the reference statement by statement; everything around them (record reading, SA parsing, window
// bookkeeping, coverage, merge of several BAMs) is this engine's own.
// Pinned against the reference's own functions run by oracle/ref2_shim.c (np2_ref_contig_windows): window ranges, alignment
// counts and a hash over every alignment string (tests/test_lgs_from_bam.py).  Host plumbing: no compute of the path here.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "hostio.h"
#include "../../include/nextpolish2_b200.h"

namespace np2x { void set_error(const std::string& m); }

namespace {
enum { INS_MIN_CHECK_LEN = 100000, INS_RADOM_LEN = 15000000, MAX_GAP_LEN = 30000 };
const char NT16[] = "=ACMGRSVTWYHKDBN";

struct Aln {                        // alignment (ctg_cns.h:150-162), the fields this stage touches
    int64_t shift = 0, aln_len = 0, t_s = 0, t_e = 0, q_s = 0, q_e = 0;
    std::string t, q;
};
struct Pos { uint32_t s, e; };
struct Gap { Pos gap; uint32_t fs, ds, score; };

uint64_t fnv(uint64_t h, const void* p, size_t n) {
    const unsigned char* c = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ULL; }
    return h;
}

// read_ref + bit2seq1: what ctg_cns_core's rfseq holds
std::string two_bit_roundtrip(const std::string& seq) {
    static const uint8_t nt[128] = {
        4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4, 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
        4,0,4,1,4,4,4,2,4,4,4,4,4,4,4,4, 4,4,4,4,3,3,4,4,4,4,4,4,4,4,4,4, 4,0,4,1,4,4,4,2,4,4,4,4,4,4,4,4, 4,4,4,4,3,3,4,4,4,4,4,4,4,4,4,4};
    std::string out(seq.size(), 'A');
    for (size_t w0 = 0; w0 < seq.size(); w0 += 16) {
        uint32_t buffer = 0;
        const size_t n = std::min<size_t>(16, seq.size() - w0);
        for (size_t i = 0; i < n; i++) buffer = buffer << 2 | nt[(uint8_t)seq[w0 + i] & 127];
        for (size_t i = 0; i < n; i++) out[w0 + i] = "ACGT"[(buffer >> (2 * (n - 1 - i))) & 3];
    }
    return out;
}

int32_t cal_win_len(int w, int s, uint64_t l) {                                      // :2800-2807
    int b = (int)l;
    if (l > (uint64_t)w) {
        const int n = (int)((float)(l - s) / (w - s) + 0.999);
        b = (int)((float)(l + (uint64_t)(n - 1) * s) / n + 0.999);
    }
    return b;
}
uint32_t cig(const uint32_t* c, uint32_t i) { uint32_t v; memcpy(&v, c + i, 4); return v; }
int32_t l_qseq_from_cigar(const uint32_t* c, uint32_t n) {                           // cal_l_qseq_from_cigar :2337-2352: S H M = X I
    int32_t r = 0;
    for (uint32_t i = 0; i < n; i++) { const uint32_t op = cig(c, i) & 15, len = cig(c, i) >> 4; if (op == 4 || op == 5 || op == 0 || op == 7 || op == 8 || op == 1) r += (int32_t)len; }
    return r;
}
int32_t cal_l_qseq(const np::BamRec& r) {                                            // :2354-2366
    if (!r.l_qseq) return l_qseq_from_cigar(r.cigar, r.n_cigar);
    const uint32_t op0 = cig(r.cigar, 0) & 15;
    if (op0 == 4) return r.l_qseq;
    int32_t rlen = r.l_qseq;
    if (op0 == 5) rlen += (int32_t)(cig(r.cigar, 0) >> 4);
    const uint32_t last = cig(r.cigar, r.n_cigar - 1);
    if ((last & 15) == 5) rlen += (int32_t)(last >> 4);
    return rlen;
}
uint32_t cigar_clip(const np::BamRec& r, int end) {                                  // cigarint2ul :2314-2320
    const uint32_t c = cig(r.cigar, end ? r.n_cigar - 1 : 0);
    return ((c & 15) == 4 || (c & 15) == 5) ? c >> 4 : 0;
}
int64_t bam_endpos(const np::BamRec& r) {                                            // htslib: M D N = X consume the reference
    int64_t l = 0;
    for (uint32_t i = 0; i < r.n_cigar; i++) { const uint32_t op = cig(r.cigar, i) & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) l += cig(r.cigar, i) >> 4; }
    return r.pos + (l ? l : 1);
}
uint32_t cigarstr2ul(const char* s, int end) {                                       // :2368-2385
    if (end) {
        int index = 0;
        while (*(s + 1) != '\0') { if (*s >= '0' && *s <= '9') index++; else index = 0; s++; }
        s -= index;
    }
    uint32_t result = 0;
    while (*s >= '0' && *s <= '9') { result = result * 10 + (uint32_t)(*s - '0'); s++; }
    if (*s != 'H' && *s != 'S') result = 0;
    return result;
}
int32_t cigarstr2rlen(const char* s) {                                               // :2387-2400
    uint32_t rlen = 0, clen = 0;
    for (; *s; s++) { if (*s >= '0' && *s <= '9') clen = clen * 10 + (uint32_t)(*s - '0'); else { if (*s == 'M' || *s == 'D') rlen += clen; clen = 0; } }
    return (int32_t)rlen;
}
uint32_t mabs(uint32_t x, uint32_t y) { return x > y ? x - y : y - x; }
void check_indel(Gap* g, int32_t rlen, const Pos* rfp1, const Pos* rdp1, const Pos* rfp2, const Pos* rdp2) {   // :2463-2492
    int l = 0;
    const int32_t mclen = (int32_t)(rlen * 0.1);
    if (rfp1->s > rfp2->s) { l = 1; std::swap(rfp1, rfp2); std::swap(rdp1, rdp2); }
    if (rfp2->e > rfp1->e && rdp2->e > rdp1->e && (int64_t)rdp1->s < mclen && (int64_t)rdp2->e > (int64_t)rlen - mclen &&
        mabs(rfp2->s, rfp1->e) < MAX_GAP_LEN && mabs(rdp2->s, rdp1->e) < MAX_GAP_LEN && rfp1->s != rfp2->s) {
        const uint32_t score = rdp1->s + (uint32_t)rlen - rdp2->e + mabs(rfp2->s, rfp1->e) + mabs(rdp2->s, rdp1->e);
        if (score < g->score || !g->score) {
            g->score = score; g->ds = l ? rdp1->s : rdp2->s; g->fs = l ? rfp1->s : rfp2->s;
            if (rfp1->e < rfp2->s) { g->gap.s = rfp1->e; g->gap.e = rfp2->s; } else { g->gap.s = rfp2->s; g->gap.e = rfp1->e; }
        }
    }
}
// the Z value of aux tag SA, or null
const char* aux_sa(const np::BamRec& r, size_t* len) {
    const uint8_t* p = r.aux; const uint8_t* end = r.aux + r.l_aux;
    while (p + 3 <= end) {
        const char t0 = (char)p[0], t1 = (char)p[1], ty = (char)p[2];
        p += 3;
        size_t sz = 0;
        switch (ty) {
        case 'A': case 'c': case 'C': sz = 1; break;
        case 's': case 'S': sz = 2; break;
        case 'i': case 'I': case 'f': sz = 4; break;
        case 'd': sz = 8; break;
        case 'Z': case 'H': { const uint8_t* q = p; while (q < end && *q) q++; if (t0 == 'S' && t1 == 'A' && ty == 'Z') { *len = (size_t)(q - p); return (const char*)p; } p = q + 1; continue; }
        case 'B': { if (p + 5 > end) return nullptr; const char st = (char)p[0]; uint32_t n; memcpy(&n, p + 1, 4); const size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4; p += 5 + es * (size_t)n; continue; }
        default: return nullptr;
        }
        p += sz;
    }
    return nullptr;
}
// bam2aln :2403-2457: returns the reference position after the alignment, -1 on an operation the reference rejects
int64_t bam2aln(Aln& a, const std::string& rf, const np::BamRec& r) {
    uint32_t rdi = 0; int64_t rfi = a.t_s;
    for (uint32_t i = 0; i < r.n_cigar; i++) {
        uint32_t n = cig(r.cigar, i) >> 4; const uint32_t c = cig(r.cigar, i) & 15;
        switch (c) {
        case 4: case 5: rdi += n; break;
        case 3: a.t_s += n; rfi += n; break;
        case 0: while (n--) { a.t.push_back(rfi < (int64_t)rf.size() ? rf[(size_t)rfi] : 'N'); rfi++; a.q.push_back(NT16[(r.seq[rdi >> 1] >> ((~rdi & 1) << 2)) & 15]); rdi++; } break;
        case 1: while (n--) { a.t.push_back('-'); a.q.push_back(NT16[(r.seq[rdi >> 1] >> ((~rdi & 1) << 2)) & 15]); rdi++; } break;
        case 2: while (n--) { a.t.push_back(rfi < (int64_t)rf.size() ? rf[(size_t)rfi] : 'N'); rfi++; a.q.push_back('-'); } break;
        default: return -1;
        }
    }
    a.aln_len = (int64_t)a.t.size();
    return rfi;
}
void clip_aln(Aln& a, int32_t s, int32_t e, int l) {                                 // :2809-2827 (works on columns 0 .. aln_len of t / q)
    int64_t s_ = 0, e_ = a.aln_len - 1;
    while (a.t_s < s) { if (a.t[(size_t)s_++] != '-') a.t_s++; if (l && a.q[(size_t)(s_ - 1)] != '-') a.q_s++; }
    while (a.t[(size_t)s_] == '-') s_++;
    while (a.t_e > e) { if (a.t[(size_t)e_--] != '-') a.t_e--; if (l && a.q[(size_t)(e_ + 1)] != '-') a.q_e--; }
    if (e_ > s_ + 500) {
        a.aln_len = e_ - s_ + 1;
        a.t.erase(0, (size_t)s_); a.q.erase(0, (size_t)s_);                          // memmove to the front
    } else a.aln_len = 10;
}
void get_align_shift(Aln& a, int k, int l) {                                         // :139-197
    int64_t i = 0, j = 0;
    while (i < a.aln_len) {
        if (a.t[(size_t)i] == a.q[(size_t)i]) j++; else j = 0;
        if (a.t[(size_t)i] != '-') a.t_s++;
        if (l && a.q[(size_t)i] != '-') a.q_s++;
        if (j == k) { a.t_s -= k; a.shift = i - k + 1; a.aln_len = a.aln_len - i + k - 1; if (l) a.q_s -= k; break; }
        i++;
    }
    if (j == k) {
        i = a.aln_len + i - k; j = 0;
        int64_t t = 0;
        while (i >= 0) {
            if (a.t[(size_t)i] == a.q[(size_t)i]) j++; else j = 0;
            if (a.t[(size_t)i] != '-') a.t_e--;
            if (l && a.q[(size_t)i] != '-') a.q_e--;
            if (j == k) { a.t_e += k; a.aln_len = a.aln_len - t + k - 1; if (l) a.q_e += k; break; }
            i--; t++;
        }
    } else a.aln_len = 0;
}

struct Window {
    int32_t s, e; int64_t beg, rege; bool closed = false; int32_t n_empty = 0;
    std::vector<uint16_t> cov;                                                       // msa[].coverage of get_align_tags
    std::vector<uint32_t> aln_t_s, aln_len; std::vector<uint64_t> str_off;
    std::string t, q;
    uint64_t hash = 14695981039346656037ULL;
    void add(const Aln& a) {
        const uint32_t ts = (uint32_t)a.t_s, n = (uint32_t)a.aln_len;
        aln_t_s.push_back(ts); aln_len.push_back(n); str_off.push_back(t.size());
        t.append(a.t, (size_t)a.shift, (size_t)n); q.append(a.q, (size_t)a.shift, (size_t)n);
        hash = fnv(hash, &ts, 4); hash = fnv(hash, &n, 4);
        hash = fnv(hash, a.t.data() + a.shift, n); hash = fnv(hash, a.q.data() + a.shift, n);
        int64_t te = (int64_t)a.t_s - 1;                                             // coverage as get_align_tags counts it (:1232-1234)
        for (uint32_t i = 0; i < n; i++) {
            const char tc = a.t[(size_t)a.shift + i];
            bool first = false;
            if (tc != '-') { te++; first = true; }
            if (first && a.q[(size_t)a.shift + i] != 'M' && te >= 0 && te < (int64_t)cov.size()) cov[(size_t)te]++;
        }
    }
};
}  // namespace

struct np2_windows {
    std::vector<Window> win;
    int32_t read_type = 1, min_cov = 4;
    std::vector<int32_t> win_len, win_aln0; std::vector<uint32_t> aln_t_s, aln_len; std::vector<uint64_t> str_off; std::string t, q;
};

extern "C" {

void np2_windows_free(np2_windows* w) { delete w; }
int32_t np2_windows_count(const np2_windows* w) { return w ? (int32_t)w->win.size() : 0; }
void np2_windows_info(const np2_windows* w, int32_t i, int32_t* start, int32_t* end, int32_t* n_alignments, uint64_t* hash) {
    if (!w || i < 0 || (size_t)i >= w->win.size()) return;
    const Window& x = w->win[(size_t)i];
    if (start) *start = x.s;
    if (end) *end = x.e;
    if (n_alignments) *n_alignments = (int32_t)x.aln_len.size() + x.n_empty;
    if (hash) *hash = x.hash;
}
void np2_windows_starts(const np2_windows* w, int32_t* out) {
    if (!w || !out) return;
    for (size_t i = 0; i < w->win.size(); i++) out[i] = w->win[i].s;
}
void np2_windows_batch(const np2_windows* w, np2_window_batch* out) {
    if (!w || !out) return;
    out->n_windows = (int32_t)w->win.size(); out->win_len = w->win_len.data(); out->win_aln0 = w->win_aln0.data();
    out->read_type = w->read_type; out->min_cov = w->min_cov;
    out->aln_t_s = w->aln_t_s.data(); out->aln_len = w->aln_len.data(); out->str_off = w->str_off.data();
    out->t_str = w->t.data(); out->q_str = w->q.data(); out->str_bytes = (int64_t)w->t.size();
}

np2_windows* np2_windows_from_bam(const char* fasta, const char* bam, const char* contig, int32_t read_type, int32_t window, int32_t overlap) {
    const char* one[1] = {bam};
    return np2_windows_from_bams(fasta, bam ? one : nullptr, 1, contig, read_type, window, overlap);
}

// Several sorted, indexed BAMs (the driver maps the long reads in parts and lists the part BAMs, source/nextPolish:209-211,
// nextpolish2.py -l): records are taken in the order of the reference's merge iterator (bam_merge_iter, bsort.c:174-199,
// 1428-1461): by position, forward strand before reverse, then by the file's place in the list; records of one file keep
// their order.  One BAM is streamed; with several, the contig's records are held in memory for the merge.
np2_windows* np2_windows_from_bams(const char* fasta, const char* const* bams, int32_t n_bams, const char* contig, int32_t read_type,
                                   int32_t window, int32_t overlap) {
    std::string err;
    if (!fasta || !bams || n_bams < 1 || !contig || window <= overlap || overlap < 0) { np2x::set_error("np2_windows_from_bam: bad arguments"); return nullptr; }
    for (int32_t i = 0; i < n_bams; i++) if (!bams[i]) { np2x::set_error("np2_windows_from_bam: bad arguments"); return nullptr; }
    const char* bam = bams[0];
    std::vector<np::FastaRecord> fr;
    if (!np::fasta_load(fasta, {contig}, fr, err) || fr.size() != 1) { np2x::set_error("np2_windows_from_bam: " + (err.empty() ? std::string("contig not in the FASTA") : err)); return nullptr; }
    const std::string rf = two_bit_roundtrip(fr[0].seq);
    const int64_t L = (int64_t)rf.size();
    np::BamFile bf;
    if (!bf.open(bam, err)) { np2x::set_error("np2_windows_from_bam: " + err); return nullptr; }
    int tid = -1;
    for (size_t i = 0; i < bf.header().names.size(); i++) if (bf.header().names[i] == contig) tid = (int)i;
    np2_windows* W = new np2_windows();
    W->read_type = read_type; W->min_cov = 4;                                        // ctg_cns_core passes min_cov 4 (:3587)
    const double max_clip_ratio = read_type == 3 ? 0.1 : 0.7;                        // :3443
    // windows (:3448, :3455-3457, :3594); every window starts with itself as the first alignment (:3458-3468)
    const int32_t b = cal_win_len(window, overlap, (uint64_t)L);
    for (int32_t s = 0, e = 0; e < L;) {
        e = s + b > L ? (int32_t)L : s + b;
        Window x;
        x.s = s; x.e = e; x.beg = s > 0 ? s - 1 : 0;                                  // the region string "ctg:s-rege" is 1-based (:3473)
        x.rege = s == 0 ? (e > INS_RADOM_LEN ? e : INS_RADOM_LEN) : e;
        x.cov.assign((size_t)(e - s) + 1, 0);
        Aln self; self.t = rf.substr((size_t)s, (size_t)(e - s)); self.q = self.t; self.aln_len = e - s; self.t_s = 0;
        x.add(self);
        W->win.push_back(std::move(x));
        s = e - overlap;
    }
    int32_t rc = 0;
    uint64_t voff = 0; bool has = false;
    if (n_bams == 1 && tid >= 0 && L > 0) {
        if (!bf.bai_first_offset(tid, &voff, &has, err)) { np2x::set_error("np2_windows_from_bam: " + err + " (the BAM needs its .bai index)"); delete W; return nullptr; }
    }
    {
        auto process = [&](const np::BamRec& r) -> bool {
            if (r.n_cigar == 0) return true;
            const int64_t endpos = bam_endpos(r);
            const int32_t l_qseq = cal_l_qseq(r);
            Pos rfp1{(uint32_t)r.pos, (uint32_t)endpos}, rdp1{cigar_clip(r, 0), (uint32_t)l_qseq - cigar_clip(r, 1)}, rfp2, rdp2;
            Gap g; memset(&g, 0, sizeof g);
            size_t salen = 0;
            if (const char* sa = aux_sa(r, &salen)) {                                // set_satags :2158-2179 + the loop :3491-3502
                std::string z(sa, salen);
                const uint8_t strand = r.flag & 16 ? 1 : 0;
                size_t p = 0;
                while (p < z.size()) {
                    size_t semi = z.find(';', p); if (semi == std::string::npos) semi = z.size();
                    std::vector<std::string> f; size_t a0 = p;
                    while (a0 <= semi) { size_t c = z.find(',', a0); if (c == std::string::npos || c > semi) c = semi; f.emplace_back(z, a0, c - a0); a0 = c + 1; }
                    if (f.size() >= 4 && f[0] == contig && (uint8_t)(f[2][0] == '+' ? 0 : 1) == strand) {
                        rfp2.s = (uint32_t)(atoll(f[1].c_str()) - 1); rfp2.e = rfp2.s + (uint32_t)cigarstr2rlen(f[3].c_str());
                        rdp2.s = cigarstr2ul(f[3].c_str(), 0); rdp2.e = (uint32_t)l_qseq - cigarstr2ul(f[3].c_str(), 1);
                        check_indel(&g, l_qseq, &rfp1, &rdp1, &rfp2, &rdp2);
                    }
                    p = semi + 1;
                }
            }
            Aln full; bool have_full = false; int64_t full_end = 0;
            for (Window& x : W->win) {
                if (!(r.pos < x.rege && endpos > x.beg)) continue;                    // htslib's region iterator
                if (x.closed) continue;
                if (r.pos >= x.e) { x.closed = true; }                                // `if (p >= e) rege = 0` (:3478): nothing after it is aligned
                if (!x.closed && (r.flag & 0xD04) && L > INS_MIN_CHECK_LEN && g.score) { rc = -10; return false; }   // sup_aln (:3503-3504)
                if (r.flag & 0xD04) continue;
                if (!g.score && (double)(rdp1.e - rdp1.s) / (double)l_qseq <= max_clip_ratio) continue;   // :3510
                if (x.closed) continue;
                if (!have_full) {
                    full.t_s = r.pos; full.t_e = endpos; full.q_s = rdp1.s; full.q_e = rdp1.e;
                    full_end = bam2aln(full, rf, r);
                    have_full = true;
                }
                if (full_end != endpos) { rc = -12; return false; }                   // "bamaln error" (:3527-3530): an operation bam2aln rejects
                Aln a = full;
                if (a.t_s < x.s || a.t_e > x.e) clip_aln(a, x.s, x.e, (int)g.score);
                get_align_shift(a, 8, (int)g.score);
                if ((uint32_t)a.t_s > (uint32_t)a.t_e - 500u) continue;               // unsigned in the reference (:3540): an alignment that ends before position 500 of the contig always passes
                a.t_s -= x.s; a.t_e -= x.s;
                const uint16_t c0 = x.cov[(size_t)a.t_s], c1 = x.cov[(size_t)a.t_e];
                if ((c0 > 3000 && c1 > 3000) || (c0 > 500 && c1 > 500 && (double)(rdp1.e - rdp1.s) < l_qseq * 0.9)) continue;   // :3544-3545
                if (a.aln_len <= 0) {                                                 // no exact 8-mer: the reference appends an EMPTY tag list (it changes
                    const uint32_t ts = (uint32_t)a.t_s, n0 = 0;                      // nothing downstream); counted and hashed like the reference does,
                    x.hash = fnv(fnv(x.hash, &ts, 4), &n0, 4); x.n_empty++;           // not handed to the first pass
                    continue;
                }
                x.add(a);
            }
            return true;
        };
        if (n_bams == 1) {
            auto visit = [&](const np::BamRec& r) -> bool {
                if (r.tid != tid) return r.tid < tid && r.tid >= 0;                  // records before the contig in the first chunk's block: skip; after: stop
                return process(r);
            };
            if (has && !bf.scan(voff, 4, visit, err) && rc == 0) { np2x::set_error("np2_windows_from_bam: " + err); delete W; return nullptr; }
        } else {
            struct Owned { int32_t pos; uint16_t flag; uint32_t n_cigar; int32_t l_qseq, l_aux; size_t off; };
            std::vector<std::vector<Owned>> recs((size_t)n_bams);
            std::vector<std::vector<uint8_t>> bytes((size_t)n_bams);
            for (int32_t k = 0; k < n_bams; k++) {
                np::BamFile f;
                if (!f.open(bams[k], err)) { np2x::set_error("np2_windows_from_bam: " + err); delete W; return nullptr; }
                int t = -1;
                for (size_t i = 0; i < f.header().names.size(); i++) if (f.header().names[i] == contig) t = (int)i;
                uint64_t vo = 0; bool hs = false;
                if (t < 0 || L == 0) continue;
                if (!f.bai_first_offset(t, &vo, &hs, err)) { np2x::set_error("np2_windows_from_bam: " + err + " (every BAM needs its .bai index)"); delete W; return nullptr; }
                if (!hs) continue;
                auto collect = [&](const np::BamRec& r) -> bool {
                    if (r.tid != t) return r.tid < t && r.tid >= 0;
                    Owned o{r.pos, r.flag, r.n_cigar, r.l_qseq, r.l_aux, bytes[(size_t)k].size()};
                    const size_t nc = 4 * (size_t)r.n_cigar, ns = ((size_t)r.l_qseq + 1) / 2;
                    bytes[(size_t)k].insert(bytes[(size_t)k].end(), (const uint8_t*)r.cigar, (const uint8_t*)r.cigar + nc);
                    bytes[(size_t)k].insert(bytes[(size_t)k].end(), r.seq, r.seq + ns);
                    bytes[(size_t)k].insert(bytes[(size_t)k].end(), r.aux, r.aux + r.l_aux);
                    recs[(size_t)k].push_back(o);
                    return true;
                };
                if (!f.scan(vo, 4, collect, err)) { np2x::set_error("np2_windows_from_bam: " + err); delete W; return nullptr; }
            }
            std::vector<size_t> head((size_t)n_bams, 0);
            for (;;) {
                int best = -1;
                for (int32_t k = 0; k < n_bams; k++) {
                    if (head[(size_t)k] >= recs[(size_t)k].size()) continue;
                    if (best < 0) { best = k; continue; }
                    const Owned& a = recs[(size_t)k][head[(size_t)k]]; const Owned& b = recs[(size_t)best][head[(size_t)best]];
                    const int ra = a.flag & 16 ? 1 : 0, rb = b.flag & 16 ? 1 : 0;
                    if (a.pos < b.pos || (a.pos == b.pos && ra < rb)) best = k;      // equal position and strand: the earlier file wins
                }
                if (best < 0) break;
                const Owned& o = recs[(size_t)best][head[(size_t)best]++];
                const uint8_t* p = bytes[(size_t)best].data() + o.off;
                np::BamRec r;
                r.tid = tid; r.pos = o.pos; r.mapq = 0; r.flag = o.flag; r.n_cigar = o.n_cigar; r.l_qseq = o.l_qseq; r.isize = 0;
                r.cigar = (const uint32_t*)p; r.seq = p + 4 * (size_t)o.n_cigar; r.qual = nullptr;
                r.aux = r.seq + ((size_t)o.l_qseq + 1) / 2; r.l_aux = o.l_aux;
                if (!process(r)) break;
            }
        }
    }
    if (rc == -10) { np2x::set_error("np2_windows_from_bam: a split-read gap on a contig longer than 100 kb needs the reference's large-indel path, which is not built (code -10)"); delete W; return nullptr; }
    if (rc == -12) { np2x::set_error("np2_windows_from_bam: CIGAR operation outside M/I/D/N/S/H (the reference stops with \"bamaln error\") (code -12)"); delete W; return nullptr; }
    // flatten
    W->win_aln0.push_back(0);
    for (const Window& x : W->win) {
        W->win_len.push_back(x.e - x.s);
        const uint64_t base = W->t.size();
        for (size_t i = 0; i < x.aln_len.size(); i++) { W->aln_t_s.push_back(x.aln_t_s[i]); W->aln_len.push_back(x.aln_len[i]); W->str_off.push_back(base + x.str_off[i]); }
        W->t += x.t; W->q += x.q;
        W->win_aln0.push_back((int32_t)W->aln_len.size());
    }
    for (Window& x : W->win) { std::string().swap(x.t); std::string().swap(x.q); std::vector<uint16_t>().swap(x.cov); }
    return W;
}

// link_consensus_fast (ctg_cns.c:3053-3119): the windows' first-pass consensus joined into one sequence.  Neighbouring windows
// overlap by `overlap` positions; around the middle of the overlap the two consensus lists are walked against each other
// until k = 50 consecutive bases agree in contig position and letter, and the windows are cut there.  The reference walks
// its lists in backtrack order (index 0 = the window's last base): B(i, j) below is that view of the forward arrays.
// Returns the length written, -1 when cap is too small, -7 when two windows cannot be linked (the reference would run
// off its arrays or stop on its assert).
int64_t np2_link_windows_fast(int32_t n_windows, const int32_t* win_start, const int64_t* win_off, const uint32_t* pos, const char* base,
                              int32_t overlap, char* out_seq, int64_t cap) {
    if (n_windows < 0 || (n_windows > 0 && (!win_start || !win_off || !pos || !base || !out_seq))) { np2x::set_error("np2_link_windows_fast: bad arguments"); return -6; }
    const int k = 50;
    const int64_t s = overlap / 2;
    std::vector<int64_t> len((size_t)n_windows), lstrip((size_t)n_windows, 0), rstrip((size_t)n_windows, 0);
    for (int32_t i = 0; i < n_windows; i++) len[(size_t)i] = win_off[i + 1] - win_off[i];
    auto P = [&](int32_t i, int64_t j) -> int64_t { return (int64_t)pos[win_off[i] + (len[(size_t)i] - 1 - j)]; };
    auto Bc = [&](int32_t i, int64_t j) -> char { return base[win_off[i] + (len[(size_t)i] - 1 - j)]; };
    auto in = [&](int32_t i, int64_t j) { return j >= 0 && j < len[(size_t)i]; };
    bool bad = false;
    int l = 0; int32_t last_c = -1, last_n = -1;
    for (int32_t i = n_windows - 1; i > 0 && !bad; i--) {
        const int32_t c = i, nx = i - 1;                                             // consensus, consensusnext
        int64_t& rs = rstrip[(size_t)c]; int64_t& ls = lstrip[(size_t)nx];
        rs = ls = s;
        if (len[(size_t)c] <= s || len[(size_t)nx] <= s) { bad = true; break; }
        #define NP2_CK(ix, jx) if (!in(ix, jx)) { bad = true; break; }
        while (true) { NP2_CK(c, len[(size_t)c] - rs) if (!(P(c, len[(size_t)c] - rs) < P(c, len[(size_t)c] - 1) + s)) break; rs++; }
        if (bad) break;
        while (true) { NP2_CK(c, len[(size_t)c] - rs) if (!(P(c, len[(size_t)c] - rs) > P(c, len[(size_t)c] - 1) + s)) break; rs--; }
        if (bad) break;
        while (true) { NP2_CK(nx, ls) if (!(P(nx, ls) < P(nx, 0) - s)) break; ls--; }
        if (bad) break;
        while (true) { NP2_CK(nx, ls) if (!(P(nx, ls) > P(nx, 0) - s)) break; ls++; }
        if (bad) break;
        l = 0;
        const int64_t p = (int64_t)win_start[c] - (int64_t)win_start[nx];            // uncorrected_len difference
        int64_t guard = 0;
        while (l < k) {
            NP2_CK(nx, ls) NP2_CK(c, len[(size_t)c] - rs)
            if (++guard > 4 * (len[(size_t)c] + len[(size_t)nx]) + 1000) { bad = true; break; }
            const int64_t j = P(nx, ls) - P(c, len[(size_t)c] - rs);
            if (j == p && Bc(c, len[(size_t)c] - rs) == Bc(nx, ls)) { l++; ls--; rs++; }
            else { l = 0; if (j >= p) ls++; else ls--; }
        }
        #undef NP2_CK
        last_c = c; last_n = nx;
    }
    if (bad) { np2x::set_error("np2_link_windows_fast: neighbouring windows do not link (no 50 agreeing bases around the middle of their overlap) (code -7)"); return -7; }
    if (n_windows > 1) { rstrip[(size_t)last_c] -= k; lstrip[(size_t)last_n] += k; }    // only the pair handled last, as in the reference (:3094-3098)
    int64_t n = 0;
    for (int32_t i = 0; i < n_windows; i++)
        for (int64_t j = len[(size_t)i] - rstrip[(size_t)i] - 1; j >= lstrip[(size_t)i]; j--) {
            if (n >= cap) { np2x::set_error("np2_link_windows_fast: output capacity too small"); return -1; }
            out_seq[n++] = Bc(i, j);
        }
    return n;
}

}  // extern "C"
