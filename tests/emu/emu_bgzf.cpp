// emu_bgzf.cpp — TEST BUILD ONLY.  Runs the warp inflate of bgzf_inflate.h with a one-lane backend on the CPU so
// that the decoder (the same source nvcc compiles into k_bgzf_inflate) can be compared with zlib without a GPU.
#include <cstring>
#include <string>
#include <vector>
#include "../../nextpolish_b200/csrc/bgzf_inflate.h"
#include "../../nextpolish_b200/csrc/hostio.h"

namespace {
struct OneLane {
    int32_t lane() const { return 0; }
    int32_t width() const { return 1; }
    int32_t bcast(int32_t v) const { return v; }
    void sync() const {}
    int32_t exscan(int32_t v, int32_t* total) const { *total = v; return 0; }
    bool any(bool p) const { return p; }
    uint32_t ballot(bool p) const { return p ? 1u : 0u; }
    int32_t shfl(int32_t v, int32_t) const { return v; }
};
}  // namespace

extern "C" int np_emu_bgzf_inflate(const uint8_t* comp, int64_t comp_bytes, uint8_t* out, int64_t out_cap, int64_t* out_bytes,
                                   int32_t* n_blocks) {
    std::vector<npz::Block> blocks;
    std::string err;
    int64_t total = 0;
    if (!np::bgzf_scan(comp, (size_t)comp_bytes, blocks, total, err)) return -100;
    *out_bytes = total;
    if (n_blocks) *n_blocks = (int32_t)blocks.size();
    if (!out) return 0;
    if (out_cap < total) return -101;
    npz::Tables t;
    OneLane w;
    for (size_t i = 0; i < blocks.size(); i++) {
        const npz::Block& b = blocks[i];
        if (!b.out_len) continue;
        memset(&t, 0xA5, sizeof t);                       // poison: the decoder must initialise what it reads
        int rc = npz::inflate_block(comp + b.in_off, b.in_len, out + b.out_off, b.out_len, t, w);
        if (rc != npz::OK) return rc;
    }
    return 0;
}

// Record-start virtual offsets the .bai index knows for reference `tid` (BamFile::bai_record_starts) — the anchors of
// the device-side record walk (devload.cu); exported for the CPU test that checks them against an independent parse.
extern "C" int64_t np_emu_bai_record_starts(const char* bam, int32_t tid, uint64_t* out, int64_t cap) {
    np::BamFile bf;
    std::string err;
    if (!bf.open(bam, err)) return -1;
    std::vector<std::vector<uint64_t>> starts;
    if (!bf.bai_record_starts(starts, err)) return -2;
    if (tid < 0 || tid >= (int32_t)starts.size()) return 0;
    const auto& v = starts[(size_t)tid];
    for (size_t i = 0; i < v.size() && (int64_t)i < cap; i++) out[i] = v[i];
    return (int64_t)v.size();
}

// fasta_load_flat (hostio.cpp; the draft reader of np_shard_load_gpu / np_multi) for the CPU test that compares it with an
// independent parse: sequences concatenated into `seq`, their offsets in `off`, names joined by '\n' in `names`.
extern "C" int64_t np_emu_fasta_flat(const char* path, uint8_t* seq, int64_t cap, int64_t* off, int32_t max_ctg, char* names,
                                     int64_t names_cap, int32_t* n_ctg) {
    struct Ctx { std::vector<uint8_t> buf; } ctx;
    std::vector<std::string> nm;
    std::vector<int64_t> o;
    std::string err;
    auto grow = [](void* c, size_t bytes) -> uint8_t* { auto* x = (Ctx*)c; if (x->buf.size() < bytes) x->buf.resize(bytes); return x->buf.data(); };
    if (!np::fasta_load_flat(path, nm, o, grow, &ctx, err)) return -1;
    *n_ctg = (int32_t)nm.size();
    if ((int32_t)nm.size() > max_ctg || (nm.empty() ? 0 : o.back()) > cap) return -2;
    for (size_t i = 0; i < o.size(); i++) off[i] = o[i];
    if (!nm.empty() && o.back() > 0) memcpy(seq, ctx.buf.data(), (size_t)o.back());
    std::string joined;
    for (auto& s : nm) { joined += s; joined += '\n'; }
    if ((int64_t)joined.size() + 1 > names_cap) return -3;
    memcpy(names, joined.c_str(), joined.size() + 1);
    return nm.empty() ? 0 : o.back();
}
