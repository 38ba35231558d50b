// hostio.cpp — FASTA / BGZF / BAM / BAI readers and the packed-shard builder (host side).
// See hostio.h for the role of this file relative to the reference's htslib calls.
#include "hostio.h"
#include <cstdlib>

#include <zlib.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <thread>
#include <unordered_map>

namespace np {

// ------------------------------------------------------------------------------------------
// FASTA
// ------------------------------------------------------------------------------------------
static bool read_file(const std::string& path, std::string& out, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    fseeko(f, 0, SEEK_END);
    off_t n = ftello(f);
    fseeko(f, 0, SEEK_SET);
    out.resize((size_t)n);
    size_t got = n ? fread(&out[0], 1, (size_t)n, f) : 0;
    fclose(f);
    if (got != (size_t)n) { err = "short read on " + path; return false; }
    return true;
}

struct FaiEntry { std::string name; int64_t len, off, linebases, linewidth; };

static bool fai_read(const std::string& fasta, std::vector<FaiEntry>& out) {
    std::ifstream in(fasta + ".fai");
    if (!in) return false;
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        FaiEntry e;
        char name[4096];
        long long a, b, c, d;
        if (sscanf(line.c_str(), "%4095s %lld %lld %lld %lld", name, &a, &b, &c, &d) != 5) return false;
        e.name = name; e.len = a; e.off = b; e.linebases = c; e.linewidth = d;
        out.push_back(e);
    }
    return !out.empty();
}

static void parse_fasta_text(const std::string& text, std::vector<FastaRecord>& out) {
    size_t i = 0, n = text.size();
    while (i < n) {
        if (text[i] != '>') {  // skip junk before the first header
            while (i < n && text[i] != '\n') i++;
            i++;
            continue;
        }
        size_t e = text.find('\n', i);
        if (e == std::string::npos) e = n;
        size_t ns = i + 1, ne = ns;
        while (ne < e && !isspace((unsigned char)text[ne])) ne++;
        FastaRecord r;
        r.name = text.substr(ns, ne - ns);
        i = e + 1;
        size_t j = i;
        {   // the record's text ends at the next '>' that starts a line: reserve once
            const void* nx = memchr(text.data() + i, '>', n - i);
            r.seq.reserve(nx ? (size_t)((const char*)nx - (text.data() + i)) : n - i);
        }
        while (j < n && text[j] != '>') {
            const void* nl = memchr(text.data() + j, '\n', n - j);
            size_t le = nl ? (size_t)((const char*)nl - text.data()) : n;
            const char* p = text.data() + j;
            const size_t len = le - j;
            unsigned bad = 0;                                    // any byte outside isgraph (33..126)?
            for (size_t k = 0; k < len; k++) bad |= (unsigned)((unsigned char)(p[k] - 33) >= 94u);
            if (!bad) r.seq.append(p, len);                      // whole line at once (the common case)
            else for (size_t k = 0; k < len; k++) if (isgraph((unsigned char)p[k])) r.seq.push_back(p[k]);
            j = le + 1;
        }
        i = j;
        out.push_back(std::move(r));
    }
}

bool fasta_names(const std::string& path, std::vector<std::string>& names,
                 std::vector<int64_t>& lengths, std::string& err) {
    std::vector<FaiEntry> fai;
    if (fai_read(path, fai)) {
        for (auto& e : fai) { names.push_back(e.name); lengths.push_back(e.len); }
        return true;
    }
    std::string text;
    if (!read_file(path, text, err)) return false;
    std::vector<FastaRecord> recs;
    parse_fasta_text(text, recs);
    for (auto& r : recs) { names.push_back(r.name); lengths.push_back((int64_t)r.seq.size()); }
    return true;
}

bool fasta_load_flat(const std::string& path, std::vector<std::string>& names, std::vector<int64_t>& off,
                     uint8_t* (*grow)(void* ctx, size_t bytes), void* ctx, std::string& err) {
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) { err = "cannot open " + path; return false; }
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); err = "cannot stat " + path; return false; }
    const size_t n = (size_t)st.st_size;
    names.clear(); off.assign(1, 0);
    if (n == 0) { close(fd); return true; }
    // read() into a buffer the calling thread keeps (grow-only): a mapping made and torn down per call costs page-table
    // work and TLB shootdowns across all of the process's threads — the pipelined front ends call this once per job
    static thread_local std::vector<char> tbuf;
    if (tbuf.size() < n) tbuf.resize(n + n / 8);
    for (size_t got = 0; got < n;) {
        const ssize_t k = pread(fd, tbuf.data() + got, n - got, (off_t)got);
        if (k <= 0) { close(fd); err = "cannot read " + path; return false; }
        got += (size_t)k;
    }
    close(fd);
    const char* text = tbuf.data();
    uint8_t* dst = grow(ctx, n + 16);                 // the sequences are never longer than the file
    if (!dst) { err = "out of memory"; return false; }
    size_t i = 0, w = 0;
    bool in_rec = false;
    while (i < n) {
        const void* nl = memchr(text + i, '\n', n - i);
        const size_t le = nl ? (size_t)((const char*)nl - text) : n;
        if (text[i] == '>') {
            if (in_rec) off.push_back((int64_t)w);
            size_t ns = i + 1, ne = ns;
            while (ne < le && !isspace((unsigned char)text[ne])) ne++;
            names.emplace_back(text + ns, ne - ns);
            in_rec = true;
        } else if (in_rec) {
            const char* p = text + i;
            const size_t len = le - i;
            unsigned bad = 0;                                    // any byte outside isgraph (33..126)?
            for (size_t k = 0; k < len; k++) bad |= (unsigned)((unsigned char)(p[k] - 33) >= 94u);
            if (!bad) { memcpy(dst + w, p, len); w += len; }
            else for (size_t k = 0; k < len; k++) if (isgraph((unsigned char)p[k])) dst[w++] = (uint8_t)p[k];
        }
        i = le + 1;
    }
    if (in_rec) off.push_back((int64_t)w);
    return true;
}

bool fasta_load(const std::string& path, const std::vector<std::string>& names,
                std::vector<FastaRecord>& out, std::string& err) {
    std::vector<FaiEntry> fai;
    if (!names.empty() && fai_read(path, fai)) {
        std::unordered_map<std::string, size_t> idx;
        for (size_t i = 0; i < fai.size(); i++) idx.emplace(fai[i].name, i);
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) { err = "cannot open " + path; return false; }
        bool stale = false;
        for (auto& nm : names) {
            auto it = idx.find(nm);
            if (it == idx.end()) { fclose(f); err = "contig " + nm + " not in " + path + ".fai"; return false; }
            const FaiEntry& e = fai[it->second];
            int64_t nlines = e.linebases > 0 ? (e.len + e.linebases - 1) / e.linebases : 0;
            int64_t bytes = e.len + nlines * (e.linewidth - e.linebases) + 8;
            std::string raw((size_t)bytes, '\0');
            fseeko(f, (off_t)e.off, SEEK_SET);
            size_t got = fread(&raw[0], 1, raw.size(), f);
            FastaRecord r;
            r.name = nm;
            r.seq.reserve((size_t)e.len);
            for (size_t k = 0; k < got && (int64_t)r.seq.size() < e.len; k++) {
                if (raw[k] == '>') break;
                if (isgraph((unsigned char)raw[k])) r.seq.push_back(raw[k]);
            }
            // a stale or foreign .fai (or a short read) yields a truncated / shifted record: htslib fails the fetch; here the
            // index is distrusted and the whole file is parsed instead
            if ((int64_t)r.seq.size() != e.len) { stale = true; break; }
            out.push_back(std::move(r));
        }
        fclose(f);
        if (!stale) return true;
        out.clear();
    }
    std::string text;
    if (!read_file(path, text, err)) return false;
    std::vector<FastaRecord> all;
    parse_fasta_text(text, all);
    if (names.empty()) { out = std::move(all); return true; }
    std::unordered_map<std::string, size_t> idx;
    for (size_t i = 0; i < all.size(); i++) idx.emplace(all[i].name, i);
    for (auto& nm : names) {
        auto it = idx.find(nm);
        if (it == idx.end()) { err = "contig " + nm + " not in " + path; return false; }
        out.push_back(all[it->second]);
    }
    return true;
}

// ------------------------------------------------------------------------------------------
// BGZF stream: batches of blocks inflated in parallel into one contiguous buffer
// ------------------------------------------------------------------------------------------
namespace {

struct BlockDesc { size_t coff, csize, isize, uoff; };

static bool bgzf_block_at(const uint8_t* d, size_t size, size_t coff, BlockDesc& b) {
    if (coff + 18 > size) return false;
    const uint8_t* p = d + coff;
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return false;
    uint32_t xlen = p[10] | (p[11] << 8);
    size_t x = 12, xe = 12 + xlen;
    uint32_t bsize = 0;
    bool found = false;
    while (x + 4 <= xe && coff + x + 4 <= size) {
        uint32_t slen = p[x + 2] | (p[x + 3] << 8);
        if (p[x] == 'B' && p[x + 1] == 'C' && slen == 2) { bsize = p[x + 4] | (p[x + 5] << 8); found = true; break; }
        x += 4 + slen;
    }
    if (!found) return false;
    b.coff = coff;
    b.csize = (size_t)bsize + 1;
    if (coff + b.csize > size) return false;
    const uint8_t* t = p + b.csize - 4;
    b.isize = (size_t)t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
    return true;
}

static bool inflate_block(const uint8_t* d, const BlockDesc& b, uint8_t* out, z_stream& zs) {
    const uint8_t* p = d + b.coff;
    uint32_t xlen = p[10] | (p[11] << 8);
    size_t hdr = 12 + xlen;
    if (b.isize == 0) return true;
    inflateReset(&zs);
    zs.next_in = const_cast<Bytef*>(p + hdr);
    zs.avail_in = (uInt)(b.csize - hdr - 8);
    zs.next_out = out;
    zs.avail_out = (uInt)b.isize;
    int rc = inflate(&zs, Z_FINISH);
    if (!(rc == Z_STREAM_END && zs.avail_out == 0)) return false;
    // the block trailer's CRC32 of the inflated bytes, as htslib checks it (bgzf.c: bgzf_uncompress / inflate_block)
    const uint8_t* t = p + b.csize - 8;
    const uint32_t want = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
    return (uint32_t)crc32(crc32(0L, Z_NULL, 0), out, (uInt)b.isize) == want;
}

struct Stream {
    const uint8_t* data; size_t size; size_t coff; int threads;
    std::vector<uint8_t> buf;
    size_t pos = 0;
    bool eof = false;
    std::vector<BlockDesc> last;     // blocks of the most recent batch
    struct Seg { int64_t boff; size_t coff, isize; };
    std::vector<Seg> segs;           // blocks that still have bytes in buf (for voffset())

    bool fill(std::string& err, size_t batch_blocks = 512) {
        if (eof) return false;
        // compact
        if (pos > 0) {
            size_t rem = buf.size() - pos;
            if (rem) memmove(buf.data(), buf.data() + pos, rem);
            buf.resize(rem);
            for (auto& g : segs) g.boff -= (int64_t)pos;
            size_t k = 0;
            while (k < segs.size() && segs[k].boff + (int64_t)segs[k].isize <= 0) k++;
            segs.erase(segs.begin(), segs.begin() + (long)k);
            pos = 0;
        }
        last.clear();
        size_t total = 0;
        while (last.size() < batch_blocks && coff < size) {
            BlockDesc b;
            if (!bgzf_block_at(data, size, coff, b)) { err = "corrupt BGZF block header"; eof = true; return false; }
            b.uoff = total;
            total += b.isize;
            coff += b.csize;
            last.push_back(b);
        }
        if (last.empty()) { eof = true; return false; }
        size_t last_base = buf.size();
        for (auto& b : last) { Seg g; g.boff = (int64_t)(last_base + b.uoff); g.coff = b.coff; g.isize = b.isize; segs.push_back(g); }
        buf.resize(last_base + total);
        uint8_t* base = buf.data() + last_base;
        int nt = std::max(1, std::min<int>(threads, (int)last.size()));
        std::atomic<size_t> next(0);
        std::atomic<bool> ok(true);
        auto work = [&]() {
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -15) != Z_OK) { ok = false; return; }
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= last.size()) break;
                if (!inflate_block(data, last[i], base + last[i].uoff, zs)) ok = false;
            }
            inflateEnd(&zs);
        };
        if (nt == 1) work();
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; t++) th.emplace_back(work);
            for (auto& t : th) t.join();
        }
        if (!ok) { err = "BGZF inflate failed"; eof = true; return false; }
        return true;
    }
    // make at least n unread bytes available; false at EOF
    bool need(size_t n, std::string& err) {
        while (buf.size() - pos < n) {
            if (!fill(err)) return false;
        }
        return true;
    }
    // virtual offset of the current read position
    uint64_t voffset() const {
        for (const auto& g : segs)
            if ((int64_t)pos >= g.boff && (int64_t)pos < g.boff + (int64_t)g.isize)
                return ((uint64_t)g.coff << 16) | (uint64_t)((int64_t)pos - g.boff);
        return (uint64_t)coff << 16;   // at the end of everything inflated so far
    }
};

static inline int32_t rd_i32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
static inline uint16_t rd_u16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

}  // namespace

bool bgzf_scan(const uint8_t* data, size_t size, std::vector<npz::Block>& blocks, int64_t& total, std::string& err,
               std::vector<uint64_t>* coffs, size_t begin, size_t end) {
    blocks.clear();
    if (coffs) coffs->clear();
    total = 0;
    size_t coff = begin;
    if (end > size) end = size;
    while (coff < end) {
        BlockDesc b;
        if (!bgzf_block_at(data, size, coff, b)) { err = "corrupt BGZF block header at offset " + std::to_string(coff); return false; }
        const uint8_t* p = data + coff;
        uint32_t xlen = p[10] | (p[11] << 8);
        size_t hdr = 12 + xlen;
        if (b.csize < hdr + 8) { err = "BGZF block shorter than its header"; return false; }
        npz::Block o;
        o.in_off = coff + hdr - begin;           // relative to `begin`: the caller ships data + begin
        o.in_len = (uint32_t)(b.csize - hdr - 8);
        o.out_len = (uint32_t)b.isize;
        o.out_off = (uint64_t)total;
        blocks.push_back(o);
        if (coffs) coffs->push_back((uint64_t)coff);
        total += (int64_t)b.isize;
        coff += b.csize;
    }
    return true;
}

bool BamFile::bai_record_starts(std::vector<std::vector<uint64_t>>& starts, std::string& err) const {
    std::string raw;
    if (!read_file(path_ + ".bai", raw, err)) return false;
    const uint8_t* p = (const uint8_t*)raw.data();
    size_t n = raw.size(), x = 8;
    if (n < 8 || memcmp(p, "BAI\1", 4) != 0) { err = "bad BAI magic"; return false; }
    int32_t n_ref = rd_i32(p + 4);
    starts.assign((size_t)std::max(0, n_ref), {});
    for (int32_t r = 0; r < n_ref; r++) {
        std::vector<uint64_t>& v = starts[(size_t)r];
        if (x + 4 > n) { err = "truncated BAI"; return false; }
        int32_t n_bin = rd_i32(p + x); x += 4;
        for (int32_t b = 0; b < n_bin; b++) {
            if (x + 8 > n) { err = "truncated BAI"; return false; }
            uint32_t bin; memcpy(&bin, p + x, 4);
            int32_t n_chunk = rd_i32(p + x + 4); x += 8;
            if (x + 16 * (size_t)n_chunk > n) { err = "truncated BAI"; return false; }
            if (bin != 37450)
                for (int32_t c = 0; c < n_chunk; c++) { uint64_t beg; memcpy(&beg, p + x + 16 * (size_t)c, 8); v.push_back(beg); }
            x += 16 * (size_t)n_chunk;
        }
        if (x + 4 > n) { err = "truncated BAI"; return false; }
        int32_t n_intv = rd_i32(p + x); x += 4;
        if (x + 8 * (size_t)n_intv > n) { err = "truncated BAI"; return false; }
        for (int32_t i = 0; i < n_intv; i++) { uint64_t o; memcpy(&o, p + x + 8 * (size_t)i, 8); if (o) v.push_back(o); }
        x += 8 * (size_t)n_intv;
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
    }
    return true;
}

BamFile::~BamFile() {
    if (data_) munmap(const_cast<uint8_t*>(data_), size_);
    if (fd_ >= 0) close(fd_);
}

bool BamFile::open(const std::string& path, std::string& err) {
    path_ = path;
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) { err = "cannot open " + path; return false; }
    struct stat st;
    if (fstat(fd_, &st) != 0 || st.st_size == 0) { err = "cannot stat " + path; return false; }
    size_ = (size_t)st.st_size;
    void* m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
    if (m == MAP_FAILED) { err = "mmap failed on " + path; return false; }
    data_ = (const uint8_t*)m;
    return read_header(err);
}

bool BamFile::read_header(std::string& err) {
    Stream s{data_, size_, 0, 1};
    if (!s.fill(err, 4)) { if (err.empty()) err = "empty BAM"; return false; }
    if (!s.need(12, err)) { err = "truncated BAM header"; return false; }
    if (memcmp(s.buf.data() + s.pos, "BAM\1", 4) != 0) { err = path_ + " is not a BAM file"; return false; }
    int32_t l_text = rd_i32(s.buf.data() + s.pos + 4);
    s.pos += 8;
    // the text may span many blocks: consume it piecewise
    int64_t remaining = l_text;
    while (remaining > 0) {
        if (s.buf.size() == s.pos && !s.fill(err, 64)) { err = "truncated BAM header text"; return false; }
        size_t take = (size_t)std::min<int64_t>(remaining, (int64_t)(s.buf.size() - s.pos));
        s.pos += take;
        remaining -= (int64_t)take;
    }
    if (!s.need(4, err)) { err = "truncated BAM header"; return false; }
    int32_t n_ref = rd_i32(s.buf.data() + s.pos);
    s.pos += 4;
    for (int32_t i = 0; i < n_ref; i++) {
        if (!s.need(4, err)) { err = "truncated BAM header"; return false; }
        int32_t l_name = rd_i32(s.buf.data() + s.pos);
        s.pos += 4;
        if (!s.need((size_t)l_name + 4, err)) { err = "truncated BAM header"; return false; }
        hdr_.names.emplace_back((const char*)s.buf.data() + s.pos, (size_t)std::max(0, l_name - 1));
        s.pos += (size_t)l_name;
        hdr_.lengths.push_back(rd_i32(s.buf.data() + s.pos));
        s.pos += 4;
    }
    // position of the first record
    hdr_.first_record_voffset = s.voffset();
    return true;
}

bool BamFile::scan(uint64_t voff, int threads, const std::function<bool(const BamRec&)>& visit,
                   std::string& err) {
    if (voff == 0) voff = hdr_.first_record_voffset;
    Stream s{data_, size_, (size_t)(voff >> 16), std::max(1, threads)};
    if ((voff >> 16) >= size_) return true;
    if (!s.fill(err)) return err.empty();
    s.pos = (size_t)(voff & 0xffff);
    for (;;) {
        if (!s.need(4, err)) return err.empty();
        int32_t bs = rd_i32(s.buf.data() + s.pos);
        if (bs < 32) { err = "corrupt BAM record"; return false; }
        if (!s.need((size_t)bs + 4, err)) { if (err.empty()) err = "truncated BAM record"; return false; }
        const uint8_t* p = s.buf.data() + s.pos + 4;
        BamRec r;
        r.tid = rd_i32(p);
        r.pos = rd_i32(p + 4);
        uint8_t l_name = p[8];
        r.mapq = p[9];
        r.n_cigar = rd_u16(p + 12);
        r.flag = rd_u16(p + 14);
        r.l_qseq = rd_i32(p + 16);
        r.isize = rd_i32(p + 28);
        const uint8_t* q = p + 32 + l_name;
        r.cigar = (const uint32_t*)q;
        r.seq = q + 4 * (size_t)r.n_cigar;
        r.qual = r.seq + ((size_t)r.l_qseq + 1) / 2;
        if ((size_t)(r.qual + r.l_qseq - p) > (size_t)bs) { err = "corrupt BAM record layout"; return false; }
        r.aux = r.qual + r.l_qseq;
        r.l_aux = (int32_t)((size_t)bs - (size_t)(r.aux - p));
        s.pos += (size_t)bs + 4;
        if (!visit(r)) return true;
    }
}

bool BamFile::bai_first_offset(int tid, uint64_t* voff, bool* has_reads, std::string& err) const {
    std::string raw;
    if (!read_file(path_ + ".bai", raw, err)) return false;
    const uint8_t* p = (const uint8_t*)raw.data();
    size_t n = raw.size(), x = 8;
    if (n < 8 || memcmp(p, "BAI\1", 4) != 0) { err = "bad BAI magic"; return false; }
    int32_t n_ref = rd_i32(p + 4);
    if (tid < 0 || tid >= n_ref) { *has_reads = false; return true; }
    uint64_t best = ~0ull;
    for (int32_t r = 0; r <= tid; r++) {
        if (x + 4 > n) { err = "truncated BAI"; return false; }
        int32_t n_bin = rd_i32(p + x); x += 4;
        for (int32_t b = 0; b < n_bin; b++) {
            if (x + 8 > n) { err = "truncated BAI"; return false; }
            uint32_t bin; memcpy(&bin, p + x, 4);
            int32_t n_chunk = rd_i32(p + x + 4); x += 8;
            if (x + 16 * (size_t)n_chunk > n) { err = "truncated BAI"; return false; }
            if (r == tid && bin != 37450) {
                for (int32_t c = 0; c < n_chunk; c++) {
                    uint64_t beg; memcpy(&beg, p + x + 16 * (size_t)c, 8);
                    best = std::min(best, beg);
                }
            }
            x += 16 * (size_t)n_chunk;
        }
        if (x + 4 > n) { err = "truncated BAI"; return false; }
        int32_t n_intv = rd_i32(p + x); x += 4 + 8 * (size_t)n_intv;
    }
    *has_reads = best != ~0ull;
    *voff = best;
    return true;
}

// ------------------------------------------------------------------------------------------
// Packed shard
// ------------------------------------------------------------------------------------------
void Shard::view(np_shard_view* v) const {
    v->n_contigs = (int32_t)names.size();
    v->n_reads = (int64_t)rec_off.size() - 1;
    v->ctg_off = ctg_off.data();
    v->ctg_seq = ctg_seq.data();
    v->ctg_read_off = ctg_read_off.data();
    v->rec_off = rec_off.data();
    v->rec = rec.data();
    v->qual_off = with_qual ? qual_off.data() : nullptr;
    static const uint8_t kNoQual[16] = {0};
    v->qual = with_qual ? (qual.empty() ? kNoQual : qual.data()) : nullptr;   // non-null even when the sparse stream is empty
}

void Shard::begin_contig(const uint8_t* seq, size_t len) {
    if (qual_mode != 2) return;
    cur_lc.assign(len + 1, 0);
    for (size_t i = 0; i < len; i++) cur_lc[i + 1] = cur_lc[i] + (seq[i] >= 97 && seq[i] <= 122 ? 1 : 0);
}

bool shard_pack_record(const BamRec& r, Shard& s, std::string& err) {
    if (r.l_qseq > 65535 || r.n_cigar > 65535) {
        err = "read longer than 65535 bases / CIGAR ops: not a short-read record";
        return false;
    }
    // 2 bits per base when the read is made of A/C/G/T only (nt16 codes 1,2,4,8)
    bool two = s.two_bit && r.l_qseq > 0;
    for (int32_t i = 0; two && i < (int32_t)r.l_qseq; i++) {
        uint32_t c = (r.seq[i >> 1] >> ((~i & 1) << 2)) & 0xfu;
        two = c == 1 || c == 2 || c == 4 || c == 8;
    }
    const size_t seq_bytes = two ? ((size_t)r.l_qseq + 3) / 4 : ((size_t)r.l_qseq + 1) / 2;
    size_t body = 16 + 4 * (size_t)r.n_cigar + seq_bytes;
    size_t padded = (body + 15) & ~(size_t)15;
    size_t o = s.rec.size();
    if ((o + padded) / 16 > 0xffffffffull) { err = "shard record stream exceeds 64 GiB"; return false; }
    s.rec.resize(o + padded, 0);
    uint8_t* d = s.rec.data() + o;
    int32_t pos = r.pos; memcpy(d, &pos, 4);
    uint16_t flag = r.flag; memcpy(d + 4, &flag, 2);
    d[6] = r.mapq; d[7] = two ? 1 : 0;
    int32_t isz = r.isize; memcpy(d + 8, &isz, 4);
    uint16_t lq = (uint16_t)r.l_qseq, nc = (uint16_t)r.n_cigar;
    memcpy(d + 12, &lq, 2); memcpy(d + 14, &nc, 2);
    memcpy(d + 16, r.cigar, 4 * (size_t)r.n_cigar);
    uint8_t* ds = d + 16 + 4 * (size_t)r.n_cigar;
    if (!two) memcpy(ds, r.seq, seq_bytes);
    else for (int32_t i = 0; i < (int32_t)r.l_qseq; i++) {
        uint32_t c = (r.seq[i >> 1] >> ((~i & 1) << 2)) & 0xfu;
        uint32_t v = c == 1 ? 0u : c == 2 ? 1u : c == 4 ? 2u : 3u;
        ds[i >> 2] |= (uint8_t)(v << (6 - 2 * (i & 3)));
    }
    s.rec_off.push_back((uint32_t)((o + padded) / 16));
    // algorithmic bytes are defined on the 4-bit form, whatever encoding is shipped
    s.alg_bytes += (int64_t)(16 + 4 * (size_t)r.n_cigar + ((size_t)r.l_qseq + 1) / 2);
    if (s.with_qual) {
        bool keep = true;
        if (s.qual_mode == 2) {
            int64_t end = r.pos;
            for (uint32_t i = 0; i < r.n_cigar; i++) {
                uint32_t c; memcpy(&c, (const uint8_t*)r.cigar + 4 * (size_t)i, 4);
                uint32_t op = c & 0xf;
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) end += c >> 4;
            }
            int64_t L = (int64_t)s.cur_lc.size() - 1, a = r.pos < 0 ? 0 : r.pos, b = end > L ? L : end;
            keep = a < b && s.cur_lc[(size_t)b] - s.cur_lc[(size_t)a] > 0;
        }
        if (!keep) { s.qual_off.push_back(s.qual_off.back()); return true; }
        s.qual_bytes += r.l_qseq;
        size_t qo = s.qual.size(), qp = ((size_t)r.l_qseq + 15) & ~(size_t)15;
        s.qual.resize(qo + qp, 0);
        memcpy(s.qual.data() + qo, r.qual, (size_t)r.l_qseq);
        s.qual_off.push_back((uint32_t)((qo + qp) / 16));
    }
    return true;
}

bool shard_load(const std::string& fasta, const std::string& bam,
                const std::vector<std::string>& names, int with_qual, int threads,
                Shard& out, std::string& err) {
    std::vector<FastaRecord> recs;
    if (!fasta_load(fasta, names, recs, err)) return false;
    BamFile bf;
    bool have_bam = !bam.empty() && access(bam.c_str(), F_OK) == 0;
    if (have_bam && !bf.open(bam, err)) return false;
    std::unordered_map<std::string, int> tid_of;
    if (have_bam)
        for (size_t i = 0; i < bf.header().names.size(); i++) tid_of.emplace(bf.header().names[i], (int)i);

    // contigs in BAM tid order (those absent from the BAM header go last, with no reads)
    struct Slot { int tid; size_t rec_idx; };
    std::vector<Slot> slots;
    for (size_t i = 0; i < recs.size(); i++) {
        auto it = tid_of.find(recs[i].name);
        slots.push_back({it == tid_of.end() ? 0x7fffffff : it->second, i});
    }
    std::stable_sort(slots.begin(), slots.end(), [](const Slot& a, const Slot& b) { return a.tid < b.tid; });

    out = Shard();
    out.with_qual = with_qual != 0;
    out.qual_mode = with_qual;
    out.ctg_off.push_back(0);
    out.rec_off.push_back(0);
    if (with_qual) out.qual_off.push_back(0);
    std::unordered_map<int, int> slot_of_tid;
    for (size_t k = 0; k < slots.size(); k++) {
        const FastaRecord& r = recs[slots[k].rec_idx];
        out.names.push_back(r.name);
        out.fasta_rank.push_back((int32_t)slots[k].rec_idx);
        out.ctg_seq.insert(out.ctg_seq.end(), r.seq.begin(), r.seq.end());
        out.ctg_off.push_back((int64_t)out.ctg_seq.size());
        if (slots[k].tid != 0x7fffffff) slot_of_tid.emplace(slots[k].tid, (int)k);
    }
    std::vector<int64_t> counts(slots.size(), 0);

    if (have_bam) {
        bool all = names.empty();
        bool ok = true;
        int cur_slot = -1;     // records must arrive grouped by contig, in slot order
        auto visit_all = [&](const BamRec& r) -> bool {
            if (r.tid < 0) return false;                 // unplaced reads: end of sorted part
            auto it = slot_of_tid.find(r.tid);
            if (it == slot_of_tid.end() || r.n_cigar == 0) return true;
            if (it->second < cur_slot) { err = "BAM is not coordinate sorted"; ok = false; return false; }
            if (it->second != cur_slot)
                out.begin_contig(out.ctg_seq.data() + out.ctg_off[(size_t)it->second], (size_t)(out.ctg_off[(size_t)it->second + 1] - out.ctg_off[(size_t)it->second]));
            cur_slot = it->second;
            if (!shard_pack_record(r, out, err)) { ok = false; return false; }
            counts[(size_t)cur_slot]++;
            return true;
        };
        if (all) {
            if (!bf.scan(0, threads, visit_all, err) || !ok) return false;
        } else {
            for (size_t k = 0; k < slots.size() && ok; k++) {
                int tid = slots[k].tid;
                if (tid == 0x7fffffff) continue;
                uint64_t voff = 0; bool has = true; std::string e2;
                if (!bf.bai_first_offset(tid, &voff, &has, e2)) { voff = 0; has = true; }  // no index: scan
                if (!has) continue;
                out.begin_contig(out.ctg_seq.data() + out.ctg_off[k], (size_t)(out.ctg_off[k + 1] - out.ctg_off[k]));
                auto visit_one = [&](const BamRec& r) -> bool {
                    if (r.tid < 0 || r.tid > tid) return false;
                    if (r.tid < tid || r.n_cigar == 0) return true;
                    if (!shard_pack_record(r, out, err)) { ok = false; return false; }
                    counts[k]++;
                    return true;
                };
                if (!bf.scan(voff, threads, visit_one, err) || !ok) return false;
            }
        }
    }
    out.ctg_read_off.push_back(0);
    for (size_t k = 0; k < slots.size(); k++) out.ctg_read_off.push_back(out.ctg_read_off.back() + counts[k]);
    return true;
}

bool bam_insert_estimate(const std::string& bam, uint32_t count_read_ins, uint32_t max_ins_len,
                         uint32_t* mean, int32_t* read_len, std::string& err) {
    BamFile bf;
    if (!bf.open(bam, err)) return false;
    uint32_t sum = 0, count = 1;
    int32_t rl = 0;
    auto visit = [&](const BamRec& r) -> bool {
        if (!(count < count_read_ins)) return false;
        if (r.isize > 0 && (uint32_t)r.isize < max_ins_len) {   // config.c:89 (signed < unsigned)
            sum += (uint32_t)r.isize;
            if (rl == 0) rl = r.l_qseq;
            count++;
        }
        return true;
    };
    if (!bf.scan(0, 1, visit, err)) return false;
    *mean = sum / count;
    *read_len = rl;
    return true;
}

}  // namespace np
