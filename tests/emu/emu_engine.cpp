// emu_engine.cpp — TEST BUILD ONLY. Compiles the engine's kernel bodies (device_logic.h,
// engine_impl.h) with g++ and drives every "launch" with a plain loop, so the exact code that
// nvcc turns into sm_100a kernels can be unit-tested on machines without a GPU.
// It is built into tests/_emu/libnp_emu.so by tests/conftest.py and is never linked into,
// loaded by, or shipped with the product library.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "../../nextpolish_b200/csrc/engine_task2.h"
#include "../../nextpolish_b200/csrc/engine_v2.h"
#include "../../include/nextpolish_b200.h"

namespace {
struct EmuOps {
    void atomic_max(int32_t* p, int32_t v) { if (*p < v) *p = v; }
    void atomic_min(int32_t* p, int32_t v) { if (*p > v) *p = v; }
    void atomic_or(uint32_t* p, uint32_t v) { *p |= v; }
    void atomic_add(int32_t* p, int32_t v) { *p += v; }
    int32_t atomic_add_ret(int32_t* p, int32_t v) { int32_t o = *p; *p += v; return o; }
    uint32_t atomic_cas_u32(uint32_t* p, uint32_t cmp, uint32_t v) { uint32_t o = *p; if (o == cmp) *p = v; return o; }
    void atomic_add_u32(uint32_t* p, uint32_t v) { *p += v; }
    void atomic_min_u32(uint32_t* p, uint32_t v) { if (*p > v) *p = v; }
    void atomic_max_u32(uint32_t* p, uint32_t v) { if (*p < v) *p = v; }
    int32_t block_exscan(int32_t, int32_t*) { return 0; }      // one "thread" per phase: nothing before it
    int32_t reserve(int32_t* ctr, int32_t n) { int32_t o = *ctr; *ctr += n; return o; }
};
struct EmuBackend {
    std::map<std::string, std::vector<uint8_t>> pool;
    template <class T> T* buf(const char* name, size_t count) {
        auto& v = pool[name];
        size_t bytes = count * sizeof(T) + 64;
        if (v.size() < bytes) v.assign(bytes, 0xCD);   // poison: catch reads of unwritten data
        return (T*)v.data();
    }
    void zero(void* p, size_t bytes) { memset(p, 0, bytes); }
    void fill_ff(void* p, size_t bytes) { memset(p, 0xff, bytes); }
    template <class F> void launch(const char*, int64_t n, const F& f) {
        EmuOps ops;
        for (int64_t i = 0; i < n; i++) f(i, ops);
    }
    template <class F> void launch_full(const char* nm, int64_t n, const F& f) { launch(nm, n, f); }
    void exscan_i32(const int32_t* in, int32_t* out, int64_t n) {
        int64_t s = 0;
        for (int64_t i = 0; i < n; i++) { int32_t v = in[i]; out[i] = (int32_t)s; s += v; }
    }
    void exscan2_i32(const int32_t* in_a, int32_t* out_a, const int32_t* in_b, int32_t* out_b, int64_t n) {
        exscan_i32(in_a, out_a, n); exscan_i32(in_b, out_b, n);
    }
    void exscan_ncol(const int32_t* ins, int32_t* out, int64_t G) {
        int64_t s = 0;
        for (int64_t i = 0; i <= G; i++) { out[i] = (int32_t)s; s += 1 + ins[i]; }
    }
    void exscan_keep(const uint8_t* obase, int32_t* out, int64_t C) {
        int64_t s = 0;
        for (int64_t i = 0; i <= C; i++) { out[i] = (int32_t)s; s += obase[i] != 3 ? 1 : 0; }
    }
    void inclsum_i32(const int32_t* in, int32_t* out, int64_t n) {
        int64_t s = 0;
        for (int64_t i = 0; i < n; i++) { s += in[i]; out[i] = (int32_t)s; }
    }
    void inclmax_i32(const int32_t* in, int32_t* out, int64_t n) {
        int32_t m = INT32_MIN;
        for (int64_t i = 0; i < n; i++) { m = std::max(m, in[i]); out[i] = m; }
    }
    int32_t read_i32(const int32_t* p) { return *p; }
    void read_many(const int32_t* const* ptrs, int n, int32_t* out) { for (int i = 0; i < n; i++) out[i] = *ptrs[i]; }
    const int32_t* upload_i32(const char* name, const int32_t* h, size_t n) {
        int32_t* p = buf<int32_t>(name, n + 1);
        if (n) memcpy(p, h, n * sizeof(int32_t));
        return p;
    }
    // the diff kernel's glue, one "warp" (group of 32 reads) at a time
    void diff_pass(const npw::DiffPass& f) {
        EmuOps ops;
        for (int32_t grp = 0; grp < f.g.n_groups; grp++) {
            npw::DiffSink s[32]; npd::Rec rc[32]; int32_t cs[32], n[32], cnt[32];
            int32_t inc = 0, used = 0, nfit = 0;
            for (int l = 0; l < 32; l++) cnt[l] = f.walk((int64_t)grp * 32 + l, f.d.rec, f.g.dd, s[l], rc[l], cs[l], n[l], ops);
            for (int l = 0; l < 32; l++) { inc += cnt[l]; if (inc <= npw::DIFF_GROUP_SLOTS) { used = inc; nfit = l + 1; } }
            f.g.gcnt[grp] = used;
            int32_t obase = 0;
            if (nfit < 32) obase = ops.atomic_add_ret(f.g.pool_n, inc - used);
            inc = 0;
            for (int l = 0; l < 32; l++) {
                inc += cnt[l];
                int32_t base = npw::diff_group_slot(grp, inc, cnt[l]);
                if (base < 0) base = f.g.n_groups * npw::DIFF_GROUP_SLOTS + obase + (inc - cnt[l] - used);
                const int64_t r = (int64_t)grp * 32 + l;
                if (r < f.d.n_reads) f.commit(r, base, s[l], rc[l], cs[l], n[l], ops);
            }
        }
    }
    // the tile kernels with one "thread" per tile (column_pass.h): aggregates + scan, then the column walk
    void tile_aggregates(const npe::Dev& d, const npc::ColGlobals& g) {
        npc::Sums run{0, 0, 0};
        for (int32_t w = 0; w < g.n_tiles; w++) {
            const npc::Tile t = npc::tile_of(d, g, w);
            npc::Sums a{0, 0, 0};
            for (int32_t s = 0; s < npc::TT; s++) {
                const npc::Slice sl = npc::slice_of(d, t, s);
                const npc::Sums x = npc::slice_sums(d, g, t, sl, g.cov + sl.ca);
                a.cov += x.cov; a.tbl += x.tbl; a.str += x.str;
            }
            g.tile_cov[w] = run.cov; g.tile_tbl[w] = run.tbl; g.tile_str[w] = run.str;
            run.cov += a.cov; run.tbl += a.tbl; run.str += a.str;
        }
        g.tile_cov[g.n_tiles] = run.cov; g.tile_tbl[g.n_tiles] = run.tbl; g.tile_str[g.n_tiles] = run.str;
    }
    void column_pass(const npe::Dev& d, const npc::ColGlobals& g) {
        EmuOps ops;
        for (int32_t w = 0; w < g.n_tiles; w++) {
            const npc::Tile t = npc::tile_of(d, g, w);
            npc::Sums run{g.tile_cov[w], g.tile_tbl[w], g.tile_str[w]};
            for (int32_t s = 0; s < npc::TT; s++) {
                const npc::Slice sl = npc::slice_of(d, t, s);
                const npc::Sums x = npc::slice_sums(d, g, t, sl, g.cov + sl.ca);
                npc::slice_walk(d, g, t, sl, g.cov + sl.ca, run.cov, run.tbl, run.str, ops);
                run.cov += x.cov; run.tbl += x.tbl; run.str += x.str;
            }
        }
    }
};
}  // namespace

extern "C" int np_emu_run_impl(const np_shard_view* v, int task, const Configure* cfg,
                          uint8_t* out_seq, int64_t out_cap, int64_t* out_off, int32_t* stats, int variant);
extern "C" int np_emu_run(const np_shard_view* v, int task, const Configure* cfg,
                          uint8_t* out_seq, int64_t out_cap, int64_t* out_off, int32_t* stats) {
    return np_emu_run_impl(v, task, cfg, out_seq, out_cap, out_off, stats, 1);
}
// variant 1: general kernels (engine_impl.h); variant 2: diff pass + window kernel + fallback (engine_v2.h)
static int emu_run(const np_shard_view* v, int task, const Configure* cfg, uint8_t* out_seq, int64_t out_cap, int64_t* out_off,
                   int32_t* stats, int variant, PolishPoint* pts, int64_t pts_cap, int64_t* pts_off);
extern "C" int np_emu_run_impl(const np_shard_view* v, int task, const Configure* cfg,
                          uint8_t* out_seq, int64_t out_cap, int64_t* out_off, int32_t* stats, int variant) {
    return emu_run(v, task, cfg, out_seq, out_cap, out_off, stats, variant, nullptr, 0, nullptr);
}
// the same run with the PolishPoint trace (Configure.trace_polish_open is forced on)
extern "C" int np_emu_run_points(const np_shard_view* v, int task, const Configure* cfg, uint8_t* out_seq, int64_t out_cap,
                                 int64_t* out_off, int variant, PolishPoint* pts, int64_t pts_cap, int64_t* pts_off) {
    return emu_run(v, task, cfg, out_seq, out_cap, out_off, nullptr, variant, pts, pts_cap, pts_off);
}
static int emu_run(const np_shard_view* v, int task, const Configure* cfg, uint8_t* out_seq, int64_t out_cap, int64_t* out_off,
                   int32_t* stats, int variant, PolishPoint* pts, int64_t pts_cap, int64_t* pts_off) {
    npe::Dev d;
    memset(&d, 0, sizeof(d));
    std::vector<int32_t> goff((size_t)v->n_contigs + 1);
    for (int i = 0; i <= v->n_contigs; i++) goff[(size_t)i] = (int32_t)v->ctg_off[i];
    d.n_ctg = v->n_contigs; d.n_reads = v->n_reads; d.G = goff[(size_t)v->n_contigs];
    d.ctg_seq = v->ctg_seq; d.ctg_goff = goff.data(); d.ctg_read_off = v->ctg_read_off;
    d.rec_off = v->rec_off; d.rec = v->rec; d.qual_off = v->qual_off; d.qual = v->qual;
    d.P.trim_len_edge = cfg->trim_len_edge; d.P.ext_len_edge = cfg->ext_len_edge;
    d.P.min_map_quality = cfg->min_map_quality; d.P.rate = cfg->indel_balance_factor_sgs;
    d.P.min_count_ratio_skip = cfg->min_count_ratio_skip; d.P.min_len_ldr = cfg->min_len_ldr;
    d.P.min_len_inter_kmer = cfg->min_len_inter_kmer; d.P.max_len_kmer = cfg->max_len_kmer;
    d.P.max_count_kmer = cfg->max_count_kmer; d.P.max_clip_ratio_sgs = cfg->max_clip_ratio_sgs;
    d.P.read_tlen = cfg->read_tlen;
    d.P.trace = pts ? 1 : 0;
    EmuBackend be;
    npe::RunStats st;
    npe::V2Stats vs; memset(&vs, 0, sizeof(vs));
    int err = task == 1 ? (variant == 2 ? npe::run_score_chain_v2(be, d, v->ctg_off, &st, &vs) : npe::run_score_chain(be, d, &st))
                        : npe::run_kmer_count(be, d, &st, task);
    if (err) return err > 0 ? -err : err;
    if (st.out_bytes > out_cap) return -1000;
    memcpy(out_seq, d.out, (size_t)st.out_bytes);
    for (int i = 0; i <= v->n_contigs; i++) out_off[i] = d.out_off[i];
    if (pts) {
        if (d.n_pts > pts_cap) return -1001;
        static_assert(sizeof(PolishPoint) == sizeof(npe::TracePoint), "PolishPoint layout");
        if (d.n_pts > 0) memcpy(pts, d.pts, (size_t)d.n_pts * sizeof(PolishPoint));
        for (int i = 0; i <= v->n_contigs; i++) pts_off[i] = d.pts_off[i];
    }
    if (stats && variant == 2) { stats[4] = vs.W; stats[5] = vs.n_win; stats[6] = vs.smem; stats[7] = vs.unresolved_windows; }
    if (stats) { stats[0] = st.C; stats[1] = st.T; stats[2] = (int32_t)st.table_entries; stats[3] = (int32_t)st.sym_words; }
    return 0;
}
