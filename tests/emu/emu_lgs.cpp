// emu_lgs.cpp — TEST BUILD ONLY.  Compiles the kernel bodies of the long-read first pass (nextpolish_b200/csrc/
// lgs_first_pass.h) with g++ and drives every "launch" with a plain loop — optionally in a shuffled order, so that any
// dependence on the order in which threads run (atomic bucket fills, segments) shows up — so the code that nvcc turns
// into sm_100a kernels can be checked on machines without a GPU.  Built into tests/_emu/ by tests/test_lgs_first_pass.py;
// never linked into, loaded by, or shipped with the product libraries.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <string>
#include <vector>
#include "../../nextpolish_b200/csrc/lgs_first_pass.h"
#include "../../include/nextpolish2_b200.h"

namespace {
struct EmuOps {
    void atomic_or(uint32_t* p, uint32_t v) { *p |= v; }
    void atomic_add_u32(uint32_t* p, uint32_t v) { *p += v; }
    void atomic_max_u32(uint32_t* p, uint32_t v) { if (*p < v) *p = v; }
    void atomic_add(int32_t* p, int32_t v) { *p += v; }
    int32_t atomic_add_ret(int32_t* p, int32_t v) { const int32_t o = *p; *p += v; return o; }
};
struct EmuBackend {
    std::map<std::string, std::vector<uint8_t>> pool;
    uint64_t shuffle_seed = 0;
    int64_t launches = 0;
    bool good() const { return true; }
    template <class T> T* host(const char* name, size_t count) { return buf<T>((std::string("host:") + name).c_str(), count); }
    template <class T> T* buf(const char* name, size_t count) {
        auto& v = pool[name];
        const size_t bytes = count * sizeof(T) + 64;
        v.assign(bytes, 0xCD);                         // poison on every call: reads of unwritten data show up
        return (T*)v.data();
    }
    template <class T> const T* upload(const char* name, const T* h, size_t count) {
        T* p = buf<T>((std::string("in:") + name).c_str(), count);
        if (count) memcpy(p, h, count * sizeof(T));
        return p;
    }
    void zero(void* p, size_t bytes) { memset(p, 0, bytes); }
    void fill_ff(void* p, size_t bytes) { memset(p, 0xff, bytes); }
    template <class F> void launch(const char*, int64_t n, const F& f) {
        EmuOps ops;
        launches++;
        if (!shuffle_seed) { for (int64_t i = 0; i < n; i++) f(i, ops); return; }
        std::vector<int64_t> order((size_t)n);
        for (int64_t i = 0; i < n; i++) order[(size_t)i] = i;
        std::mt19937_64 rng(shuffle_seed + (uint64_t)launches);
        std::shuffle(order.begin(), order.end(), rng);
        for (int64_t i : order) f(i, ops);
    }
    // ChainWarp with one lane and a local scratch (the parallel gather of the product build becomes a loop)
    struct OneLane {
        np2::Gath g[np2::GMAX];
        int32_t lane() const { return 0; }
        int32_t lanes() const { return 1; }
        void sync() const {}
        np2::Gath* scratch() { return g; }
    };
    int chain_mode = 1;
    bool warp_chain() const { return chain_mode == 1; }
    template <class F> void launch_warps(const char*, int64_t n, const F& f) {
        launches++;
        std::vector<int64_t> order((size_t)n);
        for (int64_t i = 0; i < n; i++) order[(size_t)i] = i;
        if (shuffle_seed) { std::mt19937_64 rng(shuffle_seed + (uint64_t)launches); std::shuffle(order.begin(), order.end(), rng); }
        for (int64_t i : order) { OneLane w; memset(w.g, 0xCD, sizeof w.g); f(i, w); }
    }
    void exscan_i32(const int32_t* in, int32_t* out, int64_t n) {
        int64_t s = 0;
        for (int64_t i = 0; i < n; i++) { const int32_t v = in[i]; out[i] = (int32_t)s; s += v; }
    }
    void download(void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); }
    int32_t read_i32(const int32_t* p) { return *p; }
};
}  // namespace

// the product ABI's call (np2_first_pass) on the emulated backend; stats as np2_engine_last_stats
extern "C" int64_t np2_emu_first_pass_batch(const np2_window_batch* b, uint32_t* out_pos, char* out_base, uint8_t* out_qv, int64_t cap,
                                            int64_t* out_off, uint64_t shuffle_seed, int64_t* stats) {
    EmuBackend be;
    be.shuffle_seed = shuffle_seed & 0xffffffffu;
    be.chain_mode = (shuffle_seed >> 32) & 1 ? 0 : 1;          // bit 32 of the seed: the thread-per-segment chain
    np2::Batch hb;
    hb.n_win = b->n_windows; hb.win_len = b->win_len; hb.win_aln0 = b->win_aln0; hb.read_type = b->read_type; hb.min_cov = b->min_cov;
    hb.aln_t_s = b->aln_t_s; hb.aln_len = b->aln_len; hb.str_off = b->str_off; hb.t_str = b->t_str; hb.q_str = b->q_str; hb.str_bytes = b->str_bytes;
    np2::Stats st{0, 0, 0, 0};
    const int64_t rc = np2::run_first_pass(be, hb, out_pos, (uint8_t*)out_base, out_qv, cap, out_off, &st);
    if (stats) { stats[0] = st.n_seg; stats[1] = st.reruns; stats[2] = st.iterations; stats[3] = st.n_rec; }
    return rc;
}
