// engine.cu — CUDA backend of the polishing engine and the device half of the C ABI.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo (see __graft_entry__.build()).
//
// Floating point: the score chain is evaluated with separate multiply / subtract / add in
// double (contig.c:448); the file is compiled with -fmad=false so no FMA contraction can
// change a rounding relative to the reference's x86-64 SSE2 code.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "engine_task2.h"
#include "engine_v2.h"
#include "errors.h"
#include "stream_wait.h"
#include "hostio.h"
#include "../../include/nextpolish_b200.h"

namespace {

struct CudaOps {
    __host__ __device__ __forceinline__ void atomic_max(int32_t* p, int32_t v) {
#ifdef __CUDA_ARCH__
        atomicMax(p, v);
#else
        (void)p; (void)v;
#endif
    }
    __host__ __device__ __forceinline__ void atomic_min(int32_t* p, int32_t v) {
#ifdef __CUDA_ARCH__
        atomicMin(p, v);
#else
        (void)p; (void)v;
#endif
    }
    __host__ __device__ __forceinline__ void atomic_add(int32_t* p, int32_t v) {
#ifdef __CUDA_ARCH__
        atomicAdd(p, v);
#else
        (void)p; (void)v;
#endif
    }
    __host__ __device__ __forceinline__ int32_t atomic_add_ret(int32_t* p, int32_t v) {
#ifdef __CUDA_ARCH__
        return atomicAdd(p, v);
#else
        (void)p; (void)v; return 0;
#endif
    }
    __host__ __device__ __forceinline__ uint32_t atomic_cas_u32(uint32_t* p, uint32_t cmp, uint32_t v) {
#ifdef __CUDA_ARCH__
        return atomicCAS(p, cmp, v);
#else
        (void)p; (void)cmp; (void)v; return 0;
#endif
    }
    __host__ __device__ __forceinline__ void atomic_add_u32(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
        atomicAdd(p, v);
#else
        (void)p; (void)v;
#endif
    }
    __host__ __device__ __forceinline__ void atomic_min_u32(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
        atomicMin(p, v);
#else
        (void)p; (void)v;
#endif
    }
    __host__ __device__ __forceinline__ void atomic_max_u32(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
        atomicMax(p, v);
#else
        (void)p; (void)v;
#endif
    }
    // exclusive prefix sum of one value per thread over the CTA (every thread calls it; contains barriers)
    __host__ __device__ __forceinline__ int32_t block_exscan(int32_t v, int32_t* scratch) {
#ifdef __CUDA_ARCH__
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        int32_t inc = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) scratch[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int32_t w = lane < nw ? scratch[lane] : 0, wi = w;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            if (lane < nw) scratch[lane] = wi - w;
        }
        __syncthreads();
        return inc - v + scratch[wid];
#else
        (void)scratch; (void)v; return 0;
#endif
    }
    // warp-aggregated reservation of n items from a global counter: EVERY lane of the warp must call it
    __host__ __device__ __forceinline__ int32_t reserve(int32_t* ctr, int32_t n) {
#ifdef __CUDA_ARCH__
        const int lane = threadIdx.x & 31;
        int32_t inc = n;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        int32_t base = 0;
        const int32_t total = __shfl_sync(0xffffffffu, inc, 31);
        if (lane == 31 && total > 0) base = atomicAdd(ctr, total);
        base = __shfl_sync(0xffffffffu, base, 31);
        return base + inc - n;
#else
        (void)ctr; (void)n; return 0;
#endif
    }
    __host__ __device__ __forceinline__ void atomic_or(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
        atomicOr(p, v);
#else
        (void)p; (void)v;
#endif
    }
};

template <class F>
__global__ void __launch_bounds__(256) k_items(int64_t n, F f) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { CudaOps ops; f(i, ops); }
}

// every thread of the grid calls f (i may be >= n): for functors with warp-collective operations
template <class F>
__global__ void __launch_bounds__(256) k_items_full(int64_t n, F f) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    (void)n;
    CudaOps ops; f(i, ops);
}

// Exclusive prefix sums of short arrays (region / window / contig counts: a few thousand entries): one CTA per array, one
// launch for up to two arrays, instead of the two launches and the tile-state traffic of a device-wide scan each.
struct SmallScan2 { const int32_t* in[2]; int32_t* out[2]; };
__global__ void __launch_bounds__(1024) k_small_scan(SmallScan2 a, int32_t n) {
    __shared__ int32_t scratch[32];
    __shared__ int32_t carry_s;
    const int32_t* in = a.in[blockIdx.x]; int32_t* out = a.out[blockIdx.x];
    CudaOps ops;
    int32_t carry = 0;
    for (int32_t base = 0; base < n; base += 1024) {
        const int32_t i = base + (int32_t)threadIdx.x;
        const int32_t v = i < n ? in[i] : 0;
        const int32_t ex = ops.block_exscan(v, scratch);
        if (i < n) out[i] = carry + ex;
        if (threadIdx.x == 1023) carry_s = carry + ex + v;
        __syncthreads();
        carry = carry_s;
    }
}

// ---- tile kernels of the column pass (column_pass.h): one CTA per tile of TW draft positions ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
    } while (!done);
}

// The tile's cov slice (and nothing else: colbase and the draft are read through L1) is staged into shared memory by
// one bulk copy while the threads compute their slices.  Slices longer than the staged area read global memory.
struct CovStage {
    const int32_t* base;     // cov of column c at base[c - c0]
    int32_t c0, c1;          // staged columns [c0, c1)
    __device__ __forceinline__ const int32_t* at(const int32_t* gcov, int32_t ca, int32_t cb) const {
        return (ca >= c0 && cb <= c1) ? base + (ca - c0) : gcov + ca;
    }
};
__device__ __forceinline__ CovStage stage_cov(const npc::ColGlobals& g, const npc::Tile& t, int32_t* s_cov, uint32_t bar) {
    CovStage cs;
    const int32_t c0 = t.cbeg & ~3;                                   // 16-byte aligned source
    int32_t n = ((t.cend - c0) + 3) & ~3;
    if (n > npc::COV_CAP) n = npc::COV_CAP & ~3;
    cs.base = s_cov; cs.c0 = c0; cs.c1 = c0 + n;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, (uint32_t)n * 4u);
        bulk_g2s(smem_u32(s_cov), g.cov + c0, (uint32_t)n * 4u, bar);
    }
    return cs;
}

// ---- the streaming diff pass: one WARP per group of 32 consecutive reads, no block-wide barrier.  The group's entries
// go to the group's own pool region (exclusive prefix of the per-read counts by warp shuffles, no atomics); only a group
// with more than 128 entries takes space for its last reads from the overflow area (one atomic per such warp).
// Measured and dropped (round 2, B200, 1 M reads): staging the group's records and / or its 2-bit draft slice into shared
// memory by per-warp bulk copies made the kernel slower (0.113 ms direct, 0.128-0.156 ms staged): every warp visits
// its bytes exactly once, so the copy only adds its own latency in front of the walk and takes L1 capacity away.
constexpr int kDiffThreads = 256, kDiffWarps = kDiffThreads / 32;
__global__ void __launch_bounds__(kDiffThreads) k_diff(npw::DiffPass f) {
    const npe::Dev& d = f.d;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t group = (int64_t)blockIdx.x * kDiffWarps + wid;
    const int64_t r = group * 32 + lane;
    if (group * 32 >= d.n_reads) return;
    CudaOps ops;
    npw::DiffSink s; npd::Rec rc; int32_t cs, n;
    const int32_t cnt = f.walk(r, d.rec, f.g.dd, s, rc, cs, n, ops);
    int32_t inc = cnt;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    int32_t base = npw::diff_group_slot((int32_t)group, inc, cnt);
    const unsigned fit = __ballot_sync(0xffffffffu, base >= 0);
    const int nfit = __popc(fit);                                          // the fitting reads are the first nfit lanes
    const int32_t used = __shfl_sync(0xffffffffu, inc, nfit > 0 ? nfit - 1 : 0);
    if (lane == 0) f.g.gcnt[group] = nfit > 0 ? used : 0;
    if (nfit < 32) {                                                       // rare: the tail of the group goes to the overflow area
        const int32_t total = __shfl_sync(0xffffffffu, inc, 31), before = nfit > 0 ? used : 0;
        int32_t obase = 0;
        if (lane == 0) obase = atomicAdd(f.g.pool_n, total - before);
        obase = __shfl_sync(0xffffffffu, obase, 0);
        if (base < 0) base = f.g.n_groups * npw::DIFF_GROUP_SLOTS + obase + (inc - cnt - before);
    }
    if (r < d.n_reads) f.commit(r, base, s, rc, cs, n, ops);
}

__global__ void __launch_bounds__(npc::TT) k_tile_agg(npe::Dev d, npc::ColGlobals g) {
    __shared__ __align__(128) int32_t s_cov[npc::COV_CAP];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int32_t s_red[3 * (npc::TT / 32)];
    const int32_t w = (int32_t)blockIdx.x, tid = (int32_t)threadIdx.x;
    const npc::Tile t = npc::tile_of(d, g, w);
    const uint32_t bar = smem_u32(&s_bar);
    const CovStage cs = stage_cov(g, t, s_cov, bar);
    const npc::Slice sl = npc::slice_of(d, t, tid);
    __syncthreads();                                   // the barrier is initialised
    mbar_wait(bar, 0);
    npc::Sums x = npc::slice_sums(d, g, t, sl, cs.at(g.cov, sl.ca, sl.cb));
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x.cov += __shfl_xor_sync(0xffffffffu, x.cov, o); x.tbl += __shfl_xor_sync(0xffffffffu, x.tbl, o); x.str += __shfl_xor_sync(0xffffffffu, x.str, o);
    }
    constexpr int NW = npc::TT / 32;
    if ((tid & 31) == 0) { s_red[tid >> 5] = x.cov; s_red[NW + (tid >> 5)] = x.tbl; s_red[2 * NW + (tid >> 5)] = x.str; }
    __syncthreads();
    if (tid == 0) {
        int32_t sa = 0, sb = 0, sc = 0;
        for (int i = 0; i < NW; i++) { sa += s_red[i]; sb += s_red[NW + i]; sc += s_red[2 * NW + i]; }
        g.tile_cov[w] = sa; g.tile_tbl[w] = sb; g.tile_str[w] = sc;
    }
}
// exclusive prefix of the three tile aggregates, in place; [n_tiles] = totals.  One CTA: every thread owns a contiguous
// segment of tiles (serial sum), the 1024 segment sums are scanned with warp shuffles, then every thread rewrites its segment.
__global__ void __launch_bounds__(1024) k_tile_scan(npc::ColGlobals g) {
    __shared__ int32_t s_w[3][32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int32_t n = g.n_tiles + 1, per = (n + 1023) / 1024;
    const int32_t lo = tid * per < n ? tid * per : n, hi = lo + per < n ? lo + per : n;
    int32_t* arr[3] = {g.tile_cov, g.tile_tbl, g.tile_str};
    int32_t sum[3] = {0, 0, 0};
    for (int32_t i = lo; i < hi; i++) if (i < g.n_tiles) { sum[0] += arr[0][i]; sum[1] += arr[1][i]; sum[2] += arr[2][i]; }
    int32_t inc[3] = {sum[0], sum[1], sum[2]};
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1)
        #pragma unroll
        for (int k = 0; k < 3; k++) { const int32_t t = __shfl_up_sync(0xffffffffu, inc[k], o); if (lane >= o) inc[k] += t; }
    if (lane == 31) { s_w[0][wid] = inc[0]; s_w[1][wid] = inc[1]; s_w[2][wid] = inc[2]; }
    __syncthreads();
    int32_t run[3];
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        int32_t wpre = 0;
        for (int i = 0; i < 32; i++) if (i < wid) wpre += s_w[k][i];
        run[k] = wpre + inc[k] - sum[k];
    }
    for (int32_t i = lo; i < hi; i++) {
        #pragma unroll
        for (int k = 0; k < 3; k++) { const int32_t v = i < g.n_tiles ? arr[k][i] : 0; arr[k][i] = run[k]; run[k] += v; }
    }
}
__global__ void __launch_bounds__(npc::TT) k_col_pass(npe::Dev d, npc::ColGlobals g) {
    __shared__ __align__(128) int32_t s_cov[npc::COV_CAP];
    __shared__ __align__(8) unsigned long long s_bar;
    constexpr int NW = npc::TT / 32;
    __shared__ int32_t s_w[3 * NW];
    const int32_t w = (int32_t)blockIdx.x, tid = (int32_t)threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const npc::Tile t = npc::tile_of(d, g, w);
    const uint32_t bar = smem_u32(&s_bar);
    const CovStage cs = stage_cov(g, t, s_cov, bar);
    const npc::Slice sl = npc::slice_of(d, t, tid);
    const int32_t carry_cov = g.tile_cov[w], carry_tbl = g.tile_tbl[w], carry_str = g.tile_str[w];
    __syncthreads();
    mbar_wait(bar, 0);
    const int32_t* cov = cs.at(g.cov, sl.ca, sl.cb);
    const npc::Sums x = npc::slice_sums(d, g, t, sl, cov);
    // block-wide exclusive scan of the three sums (warp shuffles + one shared-memory hop)
    int32_t ia = x.cov, ib = x.tbl, ic = x.str;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o), tc = __shfl_up_sync(0xffffffffu, ic, o);
        if (lane >= o) { ia += ta; ib += tb; ic += tc; }
    }
    if (lane == 31) { s_w[wid] = ia; s_w[NW + wid] = ib; s_w[2 * NW + wid] = ic; }
    __syncthreads();
    int32_t wa = 0, wb = 0, wc = 0;
    #pragma unroll
    for (int i = 0; i < NW; i++) if (i < wid) { wa += s_w[i]; wb += s_w[NW + i]; wc += s_w[2 * NW + i]; }
    CudaOps ops;
    npc::slice_walk(d, g, t, sl, cov, carry_cov + wa + ia - x.cov, carry_tbl + wb + ib - x.tbl, carry_str + wc + ic - x.str, ops);
}

struct NcolOp {   // 1 + insertion length (0 past the end): column count of a position
    int32_t G;
    __host__ __device__ __forceinline__ int32_t operator()(int32_t v) const { return 1 + v; }
};
struct KeepOp {   // a column is emitted unless its chosen base is the gap symbol (contig.c:751)
    __host__ __device__ __forceinline__ int32_t operator()(uint8_t b) const { return b != npd::SYM_GAP ? 1 : 0; }
};
struct MaxOp { __host__ __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; } };

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fail(#x, e_); } } while (0)

struct CudaBackend {
    cudaStream_t stream = nullptr;
    struct Buf { void* p = nullptr; size_t bytes = 0; };
    std::map<std::string, Buf> pool;
    void* cub_tmp = nullptr; size_t cub_bytes = 0;
    int32_t* h_scalar = nullptr;   // pinned
    bool ok = true; std::string msg;
    // per-launch timing
    struct Timed { const char* name; cudaEvent_t a, b; };
    std::vector<Timed> timed; size_t n_timed = 0; bool timing = false;   // off unless np_engine_set_timing(e, 1)
    int32_t launches = 0;

    void fail(const char* what, cudaError_t e) {
        if (ok) { ok = false; msg = std::string(what) + ": " + cudaGetErrorString(e); }
    }
    template <class T> T* buf(const char* name, size_t count) {
        Buf& b = pool[name];
        size_t bytes = count * sizeof(T) + 256;
        if (b.bytes < bytes) {
            if (b.p) { CUDA_TRY(np_wait::stream_wait(stream)); CUDA_TRY(cudaFree(b.p)); b.p = nullptr; }
            size_t want = bytes + bytes / 8;
            CUDA_TRY(cudaMalloc(&b.p, want));
            b.bytes = b.p ? want : 0;
        }
        return (T*)b.p;
    }
    void zero(void* p, size_t bytes) { if (ok && p) CUDA_TRY(cudaMemsetAsync(p, 0, bytes, stream)); }
    void fill_ff(void* p, size_t bytes) { if (ok && p) CUDA_TRY(cudaMemsetAsync(p, 0xff, bytes, stream)); }
    // NEXTPOLISH_B200_NVTX=1: one NVTX range per launch (named like the kernels_ms keys), for timeline profilers
    bool nvtx = getenv("NEXTPOLISH_B200_NVTX") && getenv("NEXTPOLISH_B200_NVTX")[0] == '1';
    void begin_timed(const char* name) {
        if (nvtx) nvtxRangePushA(name);
        if (!timing) return;
        if (n_timed == timed.size()) {
            Timed t; t.name = name;
            CUDA_TRY(cudaEventCreate(&t.a)); CUDA_TRY(cudaEventCreate(&t.b));
            timed.push_back(t);
        }
        timed[n_timed].name = name;
        CUDA_TRY(cudaEventRecord(timed[n_timed].a, stream));
    }
    void end_timed() {
        if (nvtx) nvtxRangePop();
        if (!timing) return;
        CUDA_TRY(cudaEventRecord(timed[n_timed].b, stream));
        n_timed++;
    }
    template <class F> void launch(const char* name, int64_t n, const F& f) {
        if (!ok || n <= 0) return;
        begin_timed(name);
        int64_t blocks = (n + 255) / 256;
        k_items<F><<<(unsigned)blocks, 256, 0, stream>>>(n, f);
        CUDA_TRY(cudaGetLastError());
        launches++;
        end_timed();
    }
    template <class F> void launch_full(const char* name, int64_t n, const F& f) {
        if (!ok || n <= 0) return;
        begin_timed(name);
        int64_t blocks = (n + 255) / 256;
        k_items_full<F><<<(unsigned)blocks, 256, 0, stream>>>(n, f);
        CUDA_TRY(cudaGetLastError());
        launches++;
        end_timed();
    }
    void cub_reserve(size_t bytes) {
        if (bytes > cub_bytes) {
            if (cub_tmp) { CUDA_TRY(np_wait::stream_wait(stream)); CUDA_TRY(cudaFree(cub_tmp)); }
            CUDA_TRY(cudaMalloc(&cub_tmp, bytes + 1024));
            cub_bytes = bytes + 1024;
        }
    }
    enum { kSmallScan = 1 << 16 };
    // two exclusive sums over arrays of the same length
    void exscan2_i32(const int32_t* in_a, int32_t* out_a, const int32_t* in_b, int32_t* out_b, int64_t n) {
        if (!ok || n <= 0) return;
        if (n > kSmallScan) { exscan_i32(in_a, out_a, n); exscan_i32(in_b, out_b, n); return; }
        begin_timed("scan_small");
        SmallScan2 a{{in_a, in_b}, {out_a, out_b}};
        k_small_scan<<<2, 1024, 0, stream>>>(a, (int32_t)n);
        CUDA_TRY(cudaGetLastError());
        launches++;
        end_timed();
    }
    void exscan_i32(const int32_t* in, int32_t* out, int64_t n) {
        if (!ok || n <= 0) return;
        if (n <= kSmallScan) {
            begin_timed("scan_small");
            SmallScan2 a{{in, in}, {out, out}};
            k_small_scan<<<1, 1024, 0, stream>>>(a, (int32_t)n);
            CUDA_TRY(cudaGetLastError());
            launches++;
            end_timed();
            return;
        }
        size_t need = 0;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, (int)n, stream));
        cub_reserve(need);
        begin_timed("scan_sum");
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, need, in, out, (int)n, stream));
        launches += 2;
        end_timed();
    }
    // colbase[p] = sum_{q<p} (1 + ins[q]) for p in [0, G]: the +1 is applied on the fly (no ncol array)
    void exscan_ncol(const int32_t* ins, int32_t* out, int64_t G) {
        if (!ok) return;
        cub::TransformInputIterator<int32_t, NcolOp, const int32_t*> it(ins, NcolOp{(int32_t)G});
        size_t need = 0;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, it, out, (int)(G + 1), stream));
        cub_reserve(need);
        begin_timed("scan_sum");
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, need, it, out, (int)(G + 1), stream));
        launches += 2;
        end_timed();
    }
    // keepidx[c] = number of emitted columns before c, c in [0, C]
    void exscan_keep(const uint8_t* obase, int32_t* out, int64_t C) {
        if (!ok) return;
        cub::TransformInputIterator<int32_t, KeepOp, const uint8_t*> it(obase, KeepOp());
        size_t need = 0;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, it, out, (int)(C + 1), stream));
        cub_reserve(need);
        begin_timed("scan_sum");
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, need, it, out, (int)(C + 1), stream));
        launches += 2;
        end_timed();
    }
    void inclsum_i32(const int32_t* in, int32_t* out, int64_t n) {
        if (!ok || n <= 0) return;
        size_t need = 0;
        CUDA_TRY(cub::DeviceScan::InclusiveSum(nullptr, need, in, out, (int)n, stream));
        cub_reserve(need);
        begin_timed("scan_sum");
        CUDA_TRY(cub::DeviceScan::InclusiveSum(cub_tmp, need, in, out, (int)n, stream));
        launches += 2;
        end_timed();
    }
    void inclmax_i32(const int32_t* in, int32_t* out, int64_t n) {
        if (!ok || n <= 0) return;
        size_t need = 0;
        CUDA_TRY(cub::DeviceScan::InclusiveScan(nullptr, need, in, out, MaxOp(), (int)n, stream));
        cub_reserve(need);
        begin_timed("scan_max");
        CUDA_TRY(cub::DeviceScan::InclusiveScan(cub_tmp, need, in, out, MaxOp(), (int)n, stream));
        launches += 2;
        end_timed();
    }
    const int32_t* upload_i32(const char* name, const int32_t* h, size_t n) {
        int32_t* p = buf<int32_t>(name, n + 1);
        if (ok && n) {
            // pageable source: the copy is followed by a synchronisation so that the caller's vector may go out of scope
            CUDA_TRY(cudaMemcpyAsync(p, h, n * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
            CUDA_TRY(np_wait::stream_wait(stream));
        }
        return p;
    }
    void diff_pass(const npw::DiffPass& f) {
        if (!ok || f.d.n_reads <= 0) return;
        begin_timed("pileup_diff");
        k_diff<<<(unsigned)((f.d.n_reads + kDiffThreads - 1) / kDiffThreads), kDiffThreads, 0, stream>>>(f);
        CUDA_TRY(cudaGetLastError());
        launches++;
        end_timed();
    }
    void tile_aggregates(const npe::Dev& d, const npc::ColGlobals& g) {
        if (!ok || g.n_tiles <= 0) return;
        begin_timed("tile_agg");
        k_tile_agg<<<(unsigned)g.n_tiles, npc::TT, 0, stream>>>(d, g);
        k_tile_scan<<<1, 1024, 0, stream>>>(g);
        CUDA_TRY(cudaGetLastError());
        launches += 2;
        end_timed();
    }
    void column_pass(const npe::Dev& d, const npc::ColGlobals& g) {
        if (!ok || g.n_tiles <= 0) return;
        begin_timed("pileup_scan");
        k_col_pass<<<(unsigned)g.n_tiles, npc::TT, 0, stream>>>(d, g);
        CUDA_TRY(cudaGetLastError());
        launches++;
        end_timed();
    }
    // reads up to 8 scalars with ONE stream synchronisation
    void read_many(const int32_t* const* ptrs, int n, int32_t* out) {
        for (int i = 0; i < n; i++) out[i] = 0;
        if (!ok) return;
        for (int i = 0; i < n; i++) CUDA_TRY(cudaMemcpyAsync(h_scalar + i, ptrs[i], sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(np_wait::stream_wait(stream));
        if (ok) for (int i = 0; i < n; i++) out[i] = h_scalar[i];
    }
    int32_t read_i32(const int32_t* p) {
        if (!ok) return 0;
        CUDA_TRY(cudaMemcpyAsync(h_scalar, p, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(np_wait::stream_wait(stream));
        return ok ? *h_scalar : 0;
    }
    void release() {
        for (auto& kv : pool) if (kv.second.p) cudaFree(kv.second.p);
        pool.clear();
        if (cub_tmp) cudaFree(cub_tmp);
        for (auto& t : timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
        if (h_scalar) cudaFreeHost(h_scalar);
        if (stream) cudaStreamDestroy(stream);
    }
};

static void params_from_cfg(const Configure* cfg, npd::Params* P) {
    P->trim_len_edge = cfg->trim_len_edge;
    P->ext_len_edge = cfg->ext_len_edge;
    P->min_map_quality = cfg->min_map_quality;
    P->rate = cfg->indel_balance_factor_sgs;
    P->min_count_ratio_skip = cfg->min_count_ratio_skip;
    P->min_len_ldr = cfg->min_len_ldr;
    P->min_len_inter_kmer = cfg->min_len_inter_kmer;
    P->max_len_kmer = cfg->max_len_kmer;
    P->max_count_kmer = cfg->max_count_kmer;
    P->max_clip_ratio_sgs = cfg->max_clip_ratio_sgs;
    P->read_tlen = cfg->read_tlen;
    P->trace = cfg->trace_polish_open ? 1 : 0;
}

}  // namespace

struct np_engine {
    int device = 0;
    CudaBackend be;
    npe::Dev d;
    npe::RunStats st;
    npe::V2Stats vs{};
    bool resident = false, owns_shard = false;
    std::vector<int64_t> h_ctg_off, h_read_off;
    // device copies of the shard (when uploaded)
    void *s_seq = nullptr, *s_goff = nullptr, *s_roff = nullptr, *s_recoff = nullptr, *s_rec = nullptr,
         *s_qoff = nullptr, *s_qual = nullptr;
    size_t cap_seq = 0, cap_goff = 0, cap_roff = 0, cap_recoff = 0, cap_rec = 0, cap_qoff = 0, cap_qual = 0;
    std::vector<int64_t> h_out_off;
    bool ran = false;
    np_engine* sibling = nullptr;      // second engine (stream + scratch) of the pipelined np_polish_host
    cudaEvent_t copy_done = nullptr;
    bool pipelined_last = false;
    int64_t launches_total = 0;        // kernel launches of every run so far
};

static bool dev_reserve(np_engine* e, void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes && *p) return true;
    if (*p) { np_wait::stream_wait(e->be.stream); cudaFree(*p); *p = nullptr; *cap = 0; }
    size_t want = bytes + bytes / 16 + 256;
    cudaError_t er = cudaMalloc(p, want);
    if (er != cudaSuccess) { np::set_error(std::string("cudaMalloc: ") + cudaGetErrorString(er)); return false; }
    *cap = want;
    return true;
}

extern "C" {

np_engine* np_engine_create(int32_t device) {
    int n = 0;
    cudaError_t er = cudaGetDeviceCount(&n);
    if (er != cudaSuccess || n == 0) {
        np::set_error(std::string("np_engine_create: no usable CUDA device (") +
                      (er != cudaSuccess ? cudaGetErrorString(er) : "device count 0") +
                      "); this engine has no CPU path");
        return nullptr;
    }
    if (device < 0 || device >= n) { np::set_error("np_engine_create: bad device index"); return nullptr; }
    if ((er = cudaSetDevice(device)) != cudaSuccess) { np::set_error(cudaGetErrorString(er)); return nullptr; }
    np_engine* e = new np_engine();
    e->device = device;
    memset(&e->d, 0, sizeof(e->d));
    if ((er = cudaStreamCreateWithFlags(&e->be.stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (er = cudaMallocHost((void**)&e->be.h_scalar, 64)) != cudaSuccess) {
        np::set_error(std::string("np_engine_create: ") + cudaGetErrorString(er));
        delete e;
        return nullptr;
    }
    return e;
}

void np_engine_destroy(np_engine* e) {
    if (!e) return;
    if (e->sibling) { np_engine_destroy(e->sibling); e->sibling = nullptr; }
    if (e->copy_done) cudaEventDestroy(e->copy_done);
    cudaSetDevice(e->device);
    np_wait::stream_wait(e->be.stream);
    void* ps[] = {e->s_seq, e->s_goff, e->s_roff, e->s_recoff, e->s_rec, e->s_qoff, e->s_qual};
    for (void* p : ps) if (p) cudaFree(p);
    e->be.release();
    delete e;
}

// Makes contigs [k0, k1) of `v` the engine's resident shard.  Host shards are copied to HBM
// (asynchronously on the engine stream; pinned buffers overlap with other streams' work); device
// shards are adopted in place.  Offsets inside rec_off / qual_off stay absolute: the device base
// pointers are shifted instead of rewriting the offset arrays.
static int32_t set_shard_slice(np_engine* e, const np_shard_view* v, int32_t k0, int32_t k1, bool device_resident,
                               cudaEvent_t bulk_after = nullptr) {
    if (!e || !v || v->n_contigs < 0 || v->n_reads < 0 || k0 < 0 || k1 < k0 || k1 > v->n_contigs) { np::set_error("bad shard"); return NP_ERR_ARG; }
    const int64_t c0 = v->ctg_off[k0], G = v->ctg_off[k1] - c0;
    const int64_t r0 = v->ctg_read_off[k0], R = v->ctg_read_off[k1] - r0;
    const int32_t nc = k1 - k0;
    if (G >= 0x7fffff00ll || R >= 0x7fffff00ll) {
        np::set_error("shard exceeds 2^31 positions/reads: split it into several shards");
        return NP_ERR_LIMIT;
    }
    cudaSetDevice(e->device);
    e->h_ctg_off.resize((size_t)nc + 1); e->h_read_off.resize((size_t)nc + 1);
    std::vector<int32_t> goff((size_t)nc + 1);
    for (int i = 0; i <= nc; i++) {
        e->h_ctg_off[(size_t)i] = v->ctg_off[k0 + i] - c0;
        e->h_read_off[(size_t)i] = v->ctg_read_off[k0 + i] - r0;
        goff[(size_t)i] = (int32_t)e->h_ctg_off[(size_t)i];
    }
    cudaStream_t s = e->be.stream;
    if (!dev_reserve(e, &e->s_goff, &e->cap_goff, goff.size() * 4) ||
        !dev_reserve(e, &e->s_roff, &e->cap_roff, e->h_read_off.size() * 8)) return NP_ERR_CUDA;
    // small metadata goes through the pinned bounce buffer of the engine (the vectors above are pageable)
    cudaMemcpyAsync(e->s_goff, goff.data(), goff.size() * 4, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(e->s_roff, e->h_read_off.data(), e->h_read_off.size() * 8, cudaMemcpyHostToDevice, s);
    np_wait::stream_wait(s);      // goff is a local: it must not go out of scope before the copy ran
    // record / quality byte ranges of the slice
    uint32_t rec_lo = 0, rec_hi = 0, q_lo = 0, q_hi = 0;
    if (device_resident) {
        cudaMemcpyAsync(&rec_lo, v->rec_off + r0, 4, cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(&rec_hi, v->rec_off + r0 + R, 4, cudaMemcpyDeviceToHost, s);
        if (v->qual_off) { cudaMemcpyAsync(&q_lo, v->qual_off + r0, 4, cudaMemcpyDeviceToHost, s); cudaMemcpyAsync(&q_hi, v->qual_off + r0 + R, 4, cudaMemcpyDeviceToHost, s); }
        np_wait::stream_wait(s);
    } else {
        rec_lo = v->rec_off[r0]; rec_hi = v->rec_off[r0 + R];
        if (v->qual_off) { q_lo = v->qual_off[r0]; q_hi = v->qual_off[r0 + R]; }
    }
    const size_t rec_bytes = (size_t)(rec_hi - rec_lo) * 16, qual_bytes = (size_t)(q_hi - q_lo) * 16;
    // the bulk copies below are ordered behind `bulk_after` (an earlier upload on another stream): uploads
    // share one PCIe link, so they are kept in submission order instead of splitting its bandwidth
    if (bulk_after) cudaStreamWaitEvent(s, bulk_after, 0);
    npe::Dev& d = e->d;
    d.n_ctg = nc; d.n_reads = R; d.G = (int32_t)G;
    d.ctg_goff = (const int32_t*)e->s_goff; d.ctg_read_off = (const int64_t*)e->s_roff;
    if (device_resident) {
        d.ctg_seq = v->ctg_seq + c0; d.rec_off = v->rec_off + r0; d.rec = v->rec;
        d.qual_off = v->qual_off ? v->qual_off + r0 : nullptr; d.qual = v->qual;
    } else {
        if (!dev_reserve(e, &e->s_seq, &e->cap_seq, (size_t)G + 16) ||
            !dev_reserve(e, &e->s_recoff, &e->cap_recoff, ((size_t)R + 1) * 4) ||
            !dev_reserve(e, &e->s_rec, &e->cap_rec, rec_bytes + 16)) return NP_ERR_CUDA;
        cudaMemcpyAsync(e->s_seq, v->ctg_seq + c0, (size_t)G, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(e->s_recoff, v->rec_off + r0, ((size_t)R + 1) * 4, cudaMemcpyHostToDevice, s);
        if (rec_bytes) cudaMemcpyAsync(e->s_rec, v->rec + (size_t)rec_lo * 16, rec_bytes, cudaMemcpyHostToDevice, s);
        d.ctg_seq = (const uint8_t*)e->s_seq; d.rec_off = (const uint32_t*)e->s_recoff;
        d.rec = (const uint8_t*)e->s_rec - (size_t)rec_lo * 16;
        d.qual_off = nullptr; d.qual = nullptr;
        if (v->qual_off) {
            if (!dev_reserve(e, &e->s_qoff, &e->cap_qoff, ((size_t)R + 1) * 4) ||
                !dev_reserve(e, &e->s_qual, &e->cap_qual, qual_bytes + 16)) return NP_ERR_CUDA;
            cudaMemcpyAsync(e->s_qoff, v->qual_off + r0, ((size_t)R + 1) * 4, cudaMemcpyHostToDevice, s);
            if (qual_bytes) cudaMemcpyAsync(e->s_qual, v->qual + (size_t)q_lo * 16, qual_bytes, cudaMemcpyHostToDevice, s);
            d.qual_off = (const uint32_t*)e->s_qoff; d.qual = (const uint8_t*)e->s_qual - (size_t)q_lo * 16;
        }
    }
    cudaError_t er = cudaGetLastError();
    if (er != cudaSuccess) { np::set_error(std::string("upload: ") + cudaGetErrorString(er)); return NP_ERR_CUDA; }
    e->resident = true; e->ran = false;
    return NP_OK;
}
static int32_t set_shard_common(np_engine* e, const np_shard_view* v, bool device_resident) {
    if (!v) { np::set_error("bad shard"); return NP_ERR_ARG; }
    return set_shard_slice(e, v, 0, v->n_contigs, device_resident);
}

int32_t np_engine_upload(np_engine* e, const np_shard_view* host_shard) { return set_shard_common(e, host_shard, false); }
int32_t np_engine_adopt_device(np_engine* e, const np_shard_view* dev_shard) { return set_shard_common(e, dev_shard, true); }

int32_t np_engine_run(np_engine* e, int32_t task, const Configure* cfg) {
    if (!e || !cfg || !e->resident) { np::set_error("np_engine_run: no resident shard"); return NP_ERR_ARG; }
    cudaSetDevice(e->device);
    params_from_cfg(cfg, &e->d.P);
    e->be.n_timed = 0; e->be.launches = 0;
    int err;
    if (task == NP_TASK_SCORE_CHAIN) {
        const char* v1 = getenv("NEXTPOLISH_B200_GENERAL_KERNELS");     // debugging / A-B timing only
        if (v1 && v1[0] == '1') err = npe::run_score_chain(e->be, e->d, &e->st, !npe::rate_is_dyadic(e->d.P.rate));
        else err = npe::run_score_chain_v2(e->be, e->d, e->h_ctg_off.data(), &e->st, &e->vs);
    }
    else if (task == NP_TASK_KMER_COUNT || task == NP_TASK_SNP_VALID) {
        if (!e->d.qual_off) { np::set_error("np_engine_run: tasks 2 and 4 need the quality stream (load the shard with_qual)"); return NP_ERR_ARG; }
        err = npe::run_kmer_count(e->be, e->d, &e->st, task);
    } else { np::set_error("np_engine_run: unknown task"); return NP_ERR_ARG; }
    if (!e->be.ok) { np::set_error("CUDA failure: " + e->be.msg); return NP_ERR_CUDA; }
    if (err) {
        char b[160];
        snprintf(b, sizeof b, "device error word 0x%x (1=insertion overflow 2=depth>=65535 4=missing score 8=column string bound 16=no-depth regions share an endpoint 32=region scratch 64=window candidate without qualities)", err);
        np::set_error(b);
        return NP_ERR_LIMIT;
    }
    e->ran = true;
    e->launches_total += e->be.launches;
    return NP_OK;
}

int32_t np_engine_sync(np_engine* e) {
    cudaSetDevice(e->device);
    cudaError_t er = np_wait::stream_wait(e->be.stream);
    if (er != cudaSuccess) { np::set_error(cudaGetErrorString(er)); return NP_ERR_CUDA; }
    return NP_OK;
}
int64_t np_engine_result_bytes(np_engine* e) { return e && e->ran ? e->st.out_bytes : -1; }
const uint8_t* np_engine_result_device(np_engine* e) { return e && e->ran ? e->d.out : nullptr; }
void* np_engine_stream(np_engine* e) { return e ? (void*)e->be.stream : nullptr; }
int32_t np_engine_launch_count(np_engine* e) { return e ? e->be.launches : 0; }
int64_t np_engine_launch_total(np_engine* e) { return e ? e->launches_total : 0; }

int32_t np_engine_download(np_engine* e, uint8_t* out_seq, int64_t out_cap, int64_t* out_off) {
    if (!e || !e->ran) { np::set_error("np_engine_download: nothing to download"); return NP_ERR_ARG; }
    if (out_cap < e->st.out_bytes) { np::set_error("np_engine_download: buffer too small"); return NP_ERR_ARG; }
    cudaSetDevice(e->device);
    cudaStream_t s = e->be.stream;
    cudaMemcpyAsync(out_seq, e->d.out, (size_t)e->st.out_bytes, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(out_off, e->d.out_off, ((size_t)e->d.n_ctg + 1) * 8, cudaMemcpyDeviceToHost, s);
    cudaError_t er = np_wait::stream_wait(s);
    if (er != cudaSuccess) { np::set_error(cudaGetErrorString(er)); return NP_ERR_CUDA; }
    return NP_OK;
}

// per-contig offsets of the last run's result (n_contigs + 1 entries), without the bytes
int32_t np_engine_result_offsets(np_engine* e, int64_t* out_off) {
    if (!e || !e->ran || !out_off) { np::set_error("np_engine_result_offsets: nothing to read"); return NP_ERR_ARG; }
    cudaSetDevice(e->device);
    cudaMemcpyAsync(out_off, e->d.out_off, ((size_t)e->d.n_ctg + 1) * 8, cudaMemcpyDeviceToHost, e->be.stream);
    cudaError_t er = np_wait::stream_wait(e->be.stream);
    if (er != cudaSuccess) { np::set_error(cudaGetErrorString(er)); return NP_ERR_CUDA; }
    return NP_OK;
}

// PolishPoint trace of the last run (only when cfg->trace_polish_open was set): count, and a download of the points
// (contig-relative positions, contig order) with their per-contig offsets.
int64_t np_engine_point_count(np_engine* e) { return e && e->ran ? e->d.n_pts : -1; }
int32_t np_engine_points(np_engine* e, PolishPoint* out, int64_t cap, int64_t* off) {
    if (!e || !e->ran || !e->d.P.trace) { np::set_error("np_engine_points: the last run did not trace (Configure.trace_polish_open)"); return NP_ERR_ARG; }
    if (cap < e->d.n_pts) { np::set_error("np_engine_points: buffer too small"); return NP_ERR_ARG; }
    static_assert(sizeof(PolishPoint) == sizeof(npe::TracePoint), "PolishPoint layout");
    cudaSetDevice(e->device);
    cudaStream_t s = e->be.stream;
    if (e->d.n_pts > 0) cudaMemcpyAsync(out, e->d.pts, (size_t)e->d.n_pts * sizeof(PolishPoint), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(off, e->d.pts_off, ((size_t)e->d.n_ctg + 1) * 8, cudaMemcpyDeviceToHost, s);
    cudaError_t er = np_wait::stream_wait(s);
    if (er != cudaSuccess) { np::set_error(cudaGetErrorString(er)); return NP_ERR_CUDA; }
    return NP_OK;
}

void np_engine_set_timing(np_engine* e, int32_t on) { if (e) { e->be.timing = on != 0; e->be.n_timed = 0; } }

int32_t np_engine_window_stats(np_engine* e, int32_t* out5) {
    if (!e) return NP_ERR_ARG;
    out5[0] = e->vs.W; out5[1] = e->vs.n_win; out5[2] = e->vs.smem; out5[3] = e->vs.unresolved_windows; out5[4] = e->vs.fallback_cols;
    return NP_OK;
}

int32_t np_engine_copy_result(np_engine* e, void* dst_device, int64_t dst_cap) {
    if (!e || !e->ran || dst_cap < e->st.out_bytes) { np::set_error("np_engine_copy_result: bad arguments"); return NP_ERR_ARG; }
    cudaSetDevice(e->device);
    cudaError_t er = cudaMemcpyAsync(dst_device, e->d.out, (size_t)e->st.out_bytes, cudaMemcpyDeviceToDevice, e->be.stream);
    if (er != cudaSuccess) { np::set_error(cudaGetErrorString(er)); return NP_ERR_CUDA; }
    return NP_OK;
}

// The gather form: a 16-byte header (int64 byte count, 8 bytes of padding) followed by the polished bytes, written on the
// engine stream entirely from device memory (the count comes from out_off[n_contigs]): one buffer per rank is what the
// single collective of the path moves.
int32_t np_engine_pack_result(np_engine* e, void* dst_device, int64_t dst_cap) {
    if (!e || !e->ran || dst_cap < e->st.out_bytes + 16) { np::set_error("np_engine_pack_result: bad arguments"); return NP_ERR_ARG; }
    cudaSetDevice(e->device);
    cudaStream_t s = e->be.stream;
    cudaMemsetAsync(dst_device, 0, 16, s);
    cudaMemcpyAsync(dst_device, e->d.out_off + e->d.n_ctg, 8, cudaMemcpyDeviceToDevice, s);
    cudaError_t er = cudaMemcpyAsync((uint8_t*)dst_device + 16, e->d.out, (size_t)e->st.out_bytes, cudaMemcpyDeviceToDevice, s);
    if (er != cudaSuccess) { np::set_error(cudaGetErrorString(er)); return NP_ERR_CUDA; }
    return NP_OK;
}

int32_t np_engine_kernel_times(np_engine* e, const char** names, float* ms, int32_t cap) {
    if (!e) return 0;
    cudaSetDevice(e->device);
    np_wait::stream_wait(e->be.stream);
    int32_t n = 0;
    for (size_t i = 0; i < e->be.n_timed && n < cap; i++, n++) {
        names[n] = e->be.timed[i].name;
        float t = 0;
        cudaEventElapsedTime(&t, e->be.timed[i].a, e->be.timed[i].b);
        ms[n] = t;
    }
    return n;
}

// Upload + run + download.  Shards with several contigs are cut into two halves handled by two engines
// (two streams): the second half's host-to-device copy runs while the first half is being polished.
int32_t np_polish_host(np_engine* e, int32_t task, const np_shard_view* host_shard,
                       const Configure* cfg, uint8_t* out_seq, int64_t out_cap, int64_t* out_off) {
    if (!e || !host_shard) { np::set_error("np_polish_host: bad arguments"); return NP_ERR_ARG; }
    const np_shard_view* v = host_shard;
    const int32_t n = v->n_contigs;
    const int64_t G = n > 0 ? v->ctg_off[n] - v->ctg_off[0] : 0;
    // The two-engine pipeline is opt-in: on the 5 Mb bench shard the doubled fixed cost of two runs
    // (host syncs, small kernels) outweighs the hidden copy time (measured 7.2 vs 6.9 ms per step).
    const char* pipe = getenv("NEXTPOLISH_B200_PIPELINE");
    if (n < 2 || G < (1 << 20) || !(pipe && pipe[0] == '1')) {
        int32_t rc = np_engine_upload(e, v);
        if (rc != NP_OK) return rc;
        rc = np_engine_run(e, task, cfg);
        if (rc != NP_OK) return rc;
        return np_engine_download(e, out_seq, out_cap, out_off);
    }
    if (!e->sibling) {
        e->sibling = np_engine_create(e->device);
        if (!e->sibling) return NP_ERR_CUDA;
    }
    np_engine* b = e->sibling;
    int32_t kmid = 1;                                   // first contig index of the second half
    while (kmid < n - 1 && v->ctg_off[kmid] - v->ctg_off[0] < G / 2) kmid++;
    int32_t rc = set_shard_slice(e, v, 0, kmid, false);
    if (rc == NP_OK) {
        // the second half's copy must not share PCIe with the first: order it behind the first copy
        if (!e->copy_done) cudaEventCreateWithFlags(&e->copy_done, cudaEventDisableTiming);
        cudaEventRecord(e->copy_done, e->be.stream);
        cudaStreamWaitEvent(b->be.stream, e->copy_done, 0);
        rc = set_shard_slice(b, v, kmid, n, false);
    }
    if (rc == NP_OK) rc = np_engine_run(e, task, cfg);               // overlaps with the second copy
    if (rc == NP_OK) rc = np_engine_run(b, task, cfg);
    if (rc != NP_OK) return rc;
    const int64_t na = np_engine_result_bytes(e), nb = np_engine_result_bytes(b);
    if (na + nb > out_cap) { np::set_error("np_polish_host: output buffer too small"); return NP_ERR_ARG; }
    rc = np_engine_download(e, out_seq, out_cap, out_off);
    if (rc == NP_OK) rc = np_engine_download(b, out_seq + na, out_cap - na, out_off + kmid);
    if (rc != NP_OK) return rc;
    for (int32_t k = kmid; k <= n; k++) out_off[k] += na;
    e->pipelined_last = true;
    return NP_OK;
}

// ---------------------------------------------------------------------------------------------
// Streaming front end: jobs (task, host shard) are submitted in order; the upload of a job is enqueued at
// submission on its own slot (engine + stream + buffers) and overlaps the kernels of the jobs before it.
// This is the double-buffered pinned-ring of SURVEY.md 7.3 H7: contig blocks stream through HBM.
// ---------------------------------------------------------------------------------------------
struct np_stream {
    int device = 0;
    std::vector<np_engine*> slot;
    struct Job { int64_t ticket; int32_t task; const np_shard_view* v; Configure cfg; uint8_t* out; int64_t cap; int64_t* off; int s; int32_t rc; };
    std::vector<Job> q;                 // submitted, unfinished, oldest first
    std::map<int64_t, int32_t> finished;
    cudaEvent_t last_upload = nullptr; bool have_upload = false;
    int64_t next_ticket = 0;
};

static void stream_process_oldest(np_stream* st) {
    np_stream::Job j = st->q.front();
    st->q.erase(st->q.begin());
    np_engine* e = st->slot[(size_t)j.s];
    if (j.rc == NP_OK) j.rc = np_engine_run(e, j.task, &j.cfg);
    if (j.rc == NP_OK) j.rc = np_engine_download(e, j.out, j.cap, j.off);
    st->finished[j.ticket] = j.rc;
}

np_stream* np_stream_create(int32_t device, int32_t depth) {
    if (depth < 1) depth = 1;
    if (depth > 8) depth = 8;
    np_stream* st = new np_stream();
    st->device = device;
    for (int i = 0; i < depth; i++) {
        np_engine* e = np_engine_create(device);
        if (!e) { for (np_engine* x : st->slot) np_engine_destroy(x); delete st; return nullptr; }
        st->slot.push_back(e);
    }
    cudaEventCreateWithFlags(&st->last_upload, cudaEventDisableTiming);
    return st;
}
void np_stream_destroy(np_stream* st) {
    if (!st) return;
    while (!st->q.empty()) stream_process_oldest(st);
    for (np_engine* e : st->slot) np_engine_destroy(e);
    if (st->last_upload) cudaEventDestroy(st->last_upload);
    delete st;
}
// Returns a ticket >= 0 (or a negative NP_ERR_*).  host_shard, out_seq and out_off must stay valid until
// np_stream_wait(ticket) returned; cfg is copied.  When every slot is busy the oldest job is finished first.
int64_t np_stream_submit(np_stream* st, int32_t task, const np_shard_view* host_shard, const Configure* cfg,
                         uint8_t* out_seq, int64_t out_cap, int64_t* out_off) {
    if (!st || !host_shard || !cfg) { np::set_error("np_stream_submit: bad arguments"); return NP_ERR_ARG; }
    if (st->q.size() == st->slot.size()) stream_process_oldest(st);
    std::vector<char> used(st->slot.size(), 0);
    for (const auto& j : st->q) used[(size_t)j.s] = 1;
    int s = 0;
    while (used[(size_t)s]) s++;
    np_stream::Job j{st->next_ticket++, task, host_shard, *cfg, out_seq, out_cap, out_off, s, NP_OK};
    np_engine* e = st->slot[(size_t)s];
    j.rc = set_shard_slice(e, host_shard, 0, host_shard->n_contigs, false, st->have_upload ? st->last_upload : nullptr);
    cudaEventRecord(st->last_upload, e->be.stream);
    st->have_upload = true;
    st->q.push_back(j);
    return j.ticket;
}
// Finishes every job up to and including `ticket` (kernels + download); returns that job's status.
int32_t np_stream_wait(np_stream* st, int64_t ticket) {
    if (!st) return NP_ERR_ARG;
    while (!st->finished.count(ticket)) {
        if (st->q.empty()) { np::set_error("np_stream_wait: unknown ticket"); return NP_ERR_ARG; }
        stream_process_oldest(st);
    }
    int32_t rc = st->finished[ticket];
    st->finished.erase(ticket);
    return rc;
}
int64_t np_stream_launch_count(np_stream* st) {      // kernel launches of every job finished so far
    int64_t n = 0;
    if (st) for (np_engine* e : st->slot) n += e->launches_total;
    return n;
}

// ---------------------------------------------------------------------------------------------
// Reference ABI: one contig per call (nextpolish1.py:181-189, main.c:12-26)
// ---------------------------------------------------------------------------------------------
static np_engine* process_engine() {
    // created lazily in the calling process: nextpolish1.py forks its Pool after config_init
    // (nextpolish1.py:219-223), and a CUDA context must not cross fork().
    static np_engine* eng = nullptr;
    static pid_t owner = 0;
    if (eng && owner == getpid()) return eng;
    int dev = 0;
    if (const char* s = getenv("NEXTPOLISH_B200_DEVICE")) dev = atoi(s);
    eng = np_engine_create(dev);
    owner = getpid();
    if (!eng) {
        fprintf(stderr, "nextpolish_b200: %s\n", np_last_error());
        exit(1);   // the reference's fatal-error convention (contig.c:86-89)
    }
    return eng;
}

static PolishResult* run_one_contig(const char* tigname, Configure* cfg, int task) {
    if (!tigname || !cfg || !cfg->fastafn) { fprintf(stderr, "nextpolish_b200: bad arguments\n"); exit(1); }
    np_engine* e = process_engine();
    const char* names[1] = {tigname};
    const int wq = task == NP_TASK_KMER_COUNT ? 2 : task == NP_TASK_SNP_VALID ? 1 : 0;
    // the contig's shard is built on the GPU from the BAM's compressed bytes when <bam>.bai exists (devload.cu);
    // otherwise (or with NEXTPOLISH_B200_HOST_LOAD=1) by the host packer
    np_dev_shard* ds = nullptr;
    const char* hl = getenv("NEXTPOLISH_B200_HOST_LOAD");
    if (cfg->bamfn && !(hl && hl[0] == '1')) ds = np_shard_load_gpu(e->device, cfg->fastafn, cfg->bamfn, names, 1, wq);
    np::Shard sh; std::string err;
    np_shard_view v;
    int32_t rc;
    if (ds) {
        np_dev_shard_view(ds, &v);
        rc = np_engine_adopt_device(e, &v);
    } else {
        std::vector<std::string> nm{std::string(tigname)};
        if (!np::shard_load(cfg->fastafn, cfg->bamfn ? cfg->bamfn : "", nm, wq, 4, sh, err)) {
            fprintf(stderr, "nextpolish_b200: %s\n", err.c_str());
            exit(1);
        }
        sh.view(&v);
        rc = np_engine_upload(e, &v);
    }
    PolishResult* res = polishresult_init();
    if (rc == NP_OK) rc = np_engine_run(e, task, cfg);
    if (rc == NP_ERR_LIMIT && wq == 2 && strstr(np_last_error(), "0x40")) {
        // a window candidate without qualities: the sparse quality stream (only reads that overlap a lowercase draft base)
        // was not enough for this contig — load every read's qualities and run again
        if (ds) { np_dev_shard_free(ds); ds = nullptr; }
        std::vector<std::string> nm{std::string(tigname)};
        if (!np::shard_load(cfg->fastafn, cfg->bamfn ? cfg->bamfn : "", nm, 1, 4, sh, err)) { fprintf(stderr, "nextpolish_b200: %s\n", err.c_str()); exit(1); }
        sh.view(&v);
        rc = np_engine_upload(e, &v);
        if (rc == NP_OK) rc = np_engine_run(e, task, cfg);
    }
    if (rc != NP_OK) { fprintf(stderr, "nextpolish_b200: %s\n", np_last_error()); exit(1); }
    int64_t cap = np_engine_result_bytes(e) + 1;
    res->contig = (char*)calloc(1, (size_t)cap + 1);
    int64_t off[2] = {0, 0};
    rc = np_engine_download(e, (uint8_t*)res->contig, cap, off);
    if (rc != NP_OK) { fprintf(stderr, "nextpolish_b200: %s\n", np_last_error()); exit(1); }
    res->length = (int32_t)off[1];
    res->contig[off[1]] = '\0';
    if (cfg->trace_polish_open) {                       // contig.c:792-797: the change trace travels with the result
        const int64_t np_ = np_engine_point_count(e);
        res->data = (PolishPoint*)calloc((size_t)(np_ > 0 ? np_ : 1), sizeof(PolishPoint));
        int64_t poff[2] = {0, 0};
        if (np_engine_points(e, res->data, np_ > 0 ? np_ : 0, poff) != NP_OK) { fprintf(stderr, "nextpolish_b200: %s\n", np_last_error()); exit(1); }
        res->datalength = (int32_t)np_;
    }
    if (ds) np_dev_shard_free(ds);
    return res;
}

PolishResult* score_chain(const char* tigname, Configure* configure) { return run_one_contig(tigname, configure, NP_TASK_SCORE_CHAIN); }
PolishResult* kmer_count(const char* tigname, Configure* configure) { return run_one_contig(tigname, configure, NP_TASK_KMER_COUNT); }

static PolishResult* out_of_scope(const char* what) {
    fprintf(stderr, "nextpolish_b200: %s is outside this engine's scope (SURVEY.md section 8f); "
                    "use the reference nextpolish1.so for tasks 3-5\n", what);
    exit(1);
    return nullptr;
}
PolishResult* snp_phase(const char*, Configure*) { return out_of_scope("snp_phase"); }
PolishResult* snp_valid(const char* tigname, Configure* configure) { return run_one_contig(tigname, configure, NP_TASK_SNP_VALID); }
PolishResult* lgspolish(const char*, Configure*) { return out_of_scope("lgspolish"); }

}  // extern "C"
