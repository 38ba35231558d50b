// lgs_consensus.cu — nextpolish2.so: the long-read consensus path's first slice on the GPU (include/nextpolish2_b200.h).
// Kernel bodies: lgs_first_pass.h (one functor per kernel, launched through k_np2 below).  sm_100a only; no CPU path.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "lgs_first_pass.h"
#include "../../include/nextpolish2_b200.h"

namespace { thread_local std::string g_err; }
namespace np2x { void set_error(const std::string& m) { g_err = m; } }      // shared with lgs_hostio.cpp
namespace {

struct CudaOps {
    __device__ __forceinline__ void atomic_or(uint32_t* p, uint32_t v) { atomicOr(p, v); }
    __device__ __forceinline__ void atomic_add_u32(uint32_t* p, uint32_t v) { atomicAdd(p, v); }
    __device__ __forceinline__ void atomic_max_u32(uint32_t* p, uint32_t v) { atomicMax(p, v); }
    __device__ __forceinline__ void atomic_add(int32_t* p, int32_t v) { atomicAdd(p, v); }
    __device__ __forceinline__ int32_t atomic_add_ret(int32_t* p, int32_t v) { return atomicAdd(p, v); }
};

template <class F>
__global__ void __launch_bounds__(256) k_np2(int64_t n, F f) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { CudaOps ops; f(i, ops); }
}

struct WarpCtx {                     // ChainWarp's view of its warp (lgs_first_pass.h)
    np2::Gath* g;
    __device__ __forceinline__ int32_t lane() const { return (int32_t)(threadIdx.x & 31); }
    __device__ __forceinline__ int32_t lanes() const { return 32; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ np2::Gath* scratch() const { return g; }
};
template <class F>
__global__ void __launch_bounds__(256) k_np2_warp(int64_t n_warps, F f) {
    __shared__ np2::Gath scratch[8][np2::GMAX];
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp < n_warps) { WarpCtx w{scratch[threadIdx.x >> 5]}; f(warp, w); }      // uniform per warp
}

struct Backend {
    int device = 0;
    int chain_mode = 1;                  // 1: warp per segment (ChainWarp), 0: thread per segment (Chain); NEXTPOLISH_B200_LGS_CHAIN=thread|warp
    bool timing = false;                 // NEXTPOLISH_B200_LGS_TIMING=1: CUDA events around every launch
    struct Timed { std::string name; cudaEvent_t a, b; };
    std::vector<Timed> timed;
    cudaStream_t stream = nullptr;
    struct Buf { void* p = nullptr; size_t bytes = 0; };
    std::map<std::string, Buf> pool;
    std::map<std::string, std::vector<uint8_t>> hpool;
    void* cub_tmp = nullptr; size_t cub_bytes = 0;
    bool ok = true; std::string msg;
    int64_t launches = 0;

    void fail(const char* what, cudaError_t e) { if (ok) { ok = false; msg = std::string(what) + ": " + cudaGetErrorString(e); } }
#define NP2_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) fail(#x, e_); } while (0)
    bool good() const { return ok; }
    template <class T> T* host(const char* name, size_t count) { auto& v = hpool[name]; if (v.size() < count * sizeof(T) + 8) v.resize(count * sizeof(T) + 8); return (T*)v.data(); }
    template <class T> T* buf(const char* name, size_t count) {
        Buf& b = pool[name];
        const size_t bytes = count * sizeof(T) + 256;
        if (b.bytes < bytes) {
            if (b.p) { NP2_TRY(cudaStreamSynchronize(stream)); NP2_TRY(cudaFree(b.p)); b.p = nullptr; b.bytes = 0; }
            const size_t want = bytes + bytes / 8;
            NP2_TRY(cudaMalloc(&b.p, want));
            b.bytes = b.p ? want : 0;
        }
        return (T*)b.p;
    }
    template <class T> const T* upload(const char* name, const T* h, size_t count) {
        T* p = buf<T>(name, count);
        if (ok && p && count) NP2_TRY(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, stream));
        return p;
    }
    void zero(void* p, size_t bytes) { if (ok && p && bytes) NP2_TRY(cudaMemsetAsync(p, 0, bytes, stream)); }
    void fill_ff(void* p, size_t bytes) { if (ok && p && bytes) NP2_TRY(cudaMemsetAsync(p, 0xff, bytes, stream)); }
    bool warp_chain() const { return chain_mode == 1; }
    void t_begin(const char* name) {
        if (!timing) return;
        Timed t; t.name = name;
        NP2_TRY(cudaEventCreate(&t.a)); NP2_TRY(cudaEventCreate(&t.b));
        NP2_TRY(cudaEventRecord(t.a, stream));
        timed.push_back(t);
    }
    void t_end() { if (timing && !timed.empty()) NP2_TRY(cudaEventRecord(timed.back().b, stream)); }
    template <class F> void launch(const char* name, int64_t n, const F& f) {
        if (!ok || n <= 0) return;
        t_begin(name);
        k_np2<F><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, f);
        NP2_TRY(cudaGetLastError());
        t_end();
        launches++;
    }
    template <class F> void launch_warps(const char* name, int64_t n_warps, const F& f) {
        if (!ok || n_warps <= 0) return;
        t_begin(name);
        k_np2_warp<F><<<(unsigned)((n_warps + 7) / 8), 256, 0, stream>>>(n_warps, f);
        NP2_TRY(cudaGetLastError());
        t_end();
        launches++;
    }
    void exscan_i32(const int32_t* in, int32_t* out, int64_t n) {
        if (!ok || n <= 0) return;
        size_t need = 0;
        NP2_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, (int)n, stream));
        if (need > cub_bytes) {
            if (cub_tmp) { NP2_TRY(cudaStreamSynchronize(stream)); NP2_TRY(cudaFree(cub_tmp)); cub_tmp = nullptr; }
            NP2_TRY(cudaMalloc(&cub_tmp, need + 1024));
            cub_bytes = cub_tmp ? need + 1024 : 0;
        }
        t_begin("lgs_scan");
        if (ok) NP2_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, need, in, out, (int)n, stream));
        t_end();
        launches += 2;
    }
    void download(void* dst, const void* src, size_t bytes) {
        if (!ok || !bytes) return;
        NP2_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
        NP2_TRY(cudaStreamSynchronize(stream));
    }
    int32_t read_i32(const int32_t* p) { int32_t v = 0; download(&v, p, 4); return v; }
    void clear_timed() { for (auto& t : timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); } timed.clear(); }
    void release() {
        clear_timed();
        for (auto& kv : pool) if (kv.second.p) cudaFree(kv.second.p);
        pool.clear();
        if (cub_tmp) cudaFree(cub_tmp);
        cub_tmp = nullptr;
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
    }
};
}  // namespace

struct np2_engine { Backend be; np2::Stats last{0, 0, 0, 0}; };

extern "C" {

const char* np2_last_error(void) { return g_err.c_str(); }

np2_engine* np2_engine_create(int32_t device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_err = "np2_engine_create: no usable CUDA device; this engine has no CPU path"; return nullptr; }
    if (device < 0 || device >= ndev) { g_err = "np2_engine_create: bad device index"; return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { g_err = "np2_engine_create: cudaSetDevice failed"; return nullptr; }
    np2_engine* e = new np2_engine();
    e->be.device = device;
    if (cudaStreamCreateWithFlags(&e->be.stream, cudaStreamNonBlocking) != cudaSuccess) { g_err = "np2_engine_create: cudaStreamCreate failed"; delete e; return nullptr; }
    return e;
}

void np2_engine_destroy(np2_engine* e) {
    if (!e) return;
    cudaSetDevice(e->be.device);
    e->be.release();
    delete e;
}

int64_t np2_first_pass(np2_engine* e, const np2_window_batch* b, uint32_t* out_pos, char* out_base, uint8_t* out_qv, int64_t cap, int64_t* out_off) {
    if (!e || !b || !out_off || b->n_windows < 0) { g_err = "np2_first_pass: bad arguments"; return -6; }
    if (b->n_windows == 0) { out_off[0] = 0; return 0; }
    if (!out_pos || !out_base || !b->win_len || !b->win_aln0 || !b->aln_t_s || !b->aln_len || !b->str_off || !b->t_str || !b->q_str) { g_err = "np2_first_pass: bad arguments"; return -6; }
    if (cudaSetDevice(e->be.device) != cudaSuccess) { g_err = "np2_first_pass: cudaSetDevice failed"; return -6; }
    e->be.ok = true; e->be.msg.clear();
    e->be.clear_timed();
    const char* cm = getenv("NEXTPOLISH_B200_LGS_CHAIN");
    e->be.chain_mode = (cm && strcmp(cm, "thread") == 0) ? 0 : 1;
    const char* tm = getenv("NEXTPOLISH_B200_LGS_TIMING");
    e->be.timing = tm && tm[0] == '1';
    np2::Batch hb;
    hb.n_win = b->n_windows; hb.win_len = b->win_len; hb.win_aln0 = b->win_aln0; hb.read_type = b->read_type; hb.min_cov = b->min_cov;
    hb.aln_t_s = b->aln_t_s; hb.aln_len = b->aln_len; hb.str_off = b->str_off; hb.t_str = b->t_str; hb.q_str = b->q_str; hb.str_bytes = b->str_bytes;
    const int64_t rc = np2::run_first_pass(e->be, hb, out_pos, (uint8_t*)out_base, out_qv, cap, out_off, &e->last);
    if (rc == -6) g_err = "np2_first_pass: " + (e->be.msg.empty() ? std::string("internal error") : e->be.msg);
    else if (rc < 0) g_err = "np2_first_pass: input rejected (code " + std::to_string(rc) + ", see include/nextpolish2_b200.h)";
    return rc;
}

int64_t np2_engine_launch_count(np2_engine* e) { return e ? e->be.launches : 0; }

int32_t np2_engine_kernel_times(np2_engine* e, const char** names, float* ms, int32_t cap) {
    if (!e) return 0;
    cudaSetDevice(e->be.device);
    cudaStreamSynchronize(e->be.stream);
    int32_t n = 0;
    for (auto& t : e->be.timed) {
        if (n >= cap) break;
        float v = 0;
        if (cudaEventElapsedTime(&v, t.a, t.b) != cudaSuccess) v = -1;
        names[n] = t.name.c_str(); ms[n] = v; n++;
    }
    return n;
}
void np2_engine_last_stats(np2_engine* e, int64_t out[4]) {
    if (!e || !out) return;
    out[0] = e->last.n_seg; out[1] = e->last.reruns; out[2] = e->last.iterations; out[3] = e->last.n_rec;
}

}  // extern "C"
