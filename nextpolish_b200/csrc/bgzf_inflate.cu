// bgzf_inflate.cu — GPU inflate of BGZF blocks (one warp per block) and its C ABI.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo (nextpolish_b200/csrc/Makefile).
#include <cuda_runtime.h>
#include <unistd.h>

#include <cstdio>
#include <algorithm>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "bgzf_inflate.h"
#include "bgzf_inflate_dev.h"
#include "errors.h"
#include "stream_wait.h"
#include "hostio.h"
#include "../../include/nextpolish_b200.h"

namespace {

struct WarpBackend {
    __device__ __forceinline__ int32_t lane() const { return (int32_t)(threadIdx.x & 31u); }
    __device__ __forceinline__ int32_t width() const { return 32; }
    __device__ __forceinline__ int32_t bcast(int32_t v) const { return __shfl_sync(0xffffffffu, v, 0); }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ int32_t exscan(int32_t v, int32_t* total) const {
        const int32_t lane = (int32_t)(threadIdx.x & 31u);
        int32_t inc = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        *total = __shfl_sync(0xffffffffu, inc, 31);
        return inc - v;
    }
    __device__ __forceinline__ bool any(bool p) const { return __any_sync(0xffffffffu, p) != 0; }
    __device__ __forceinline__ uint32_t ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
    __device__ __forceinline__ int32_t shfl(int32_t v, int32_t src) const { return __shfl_sync(0xffffffffu, v, src); }
};

// kDecoders: blocks decoded in lockstep by one warp (lanes 0 .. kDecoders-1).  Measured on B200 (52 MB BAM,
                                   // 4225 blocks): 1 decoder x 8 warps 4.3 ms, 2 x 4 warps 5.3 ms, 4 x 2 warps 7.8 ms — with only
                                   // ~29 blocks per SM the kernel runs at the latency of one block, and the batches of a warp's
                                   // blocks are resolved one after the other.  One decoder per LANE (32 blocks per warp in lockstep,
                                   // each lane resolving its own batch, 2.4 KB of tables per lane) was correct but ran at 27.5 ms:
                                   // per-lane byte copies through L2 put every block's round at ~90 us

// Persistent warps; every decoder lane pulls BGZF blocks from an atomic ticket.  One round of a warp: decoders that need
// one parse their deflate block header, all decoders decode a batch of symbols in lockstep, then the warp resolves the
// batches one block after the other.
template <int kDecoders, int kWarpsPerCta, int kMinCtas>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kMinCtas) k_bgzf_inflate(const uint8_t* comp, const npz::Block* blocks, int32_t n_blocks,
                                                                         uint8_t* out, int32_t* status, int32_t* ticket) {
    __shared__ npz::Tables tabs[kWarpsPerCta][kDecoders];
    const int wid = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31u);
    WarpBackend w;
    const bool dec = lane < kDecoders;
    npz::Tables& mine = tabs[wid][dec ? lane : 0];
    npz::Decoder d;
    d.phase = npz::PH_IDLE; d.err = npz::OK; d.pos = 0; d.out = nullptr; d.out_len = 0;
    int32_t blk = -1;
    bool exhausted = !dec;
    for (;;) {
        // finished blocks report, idle decoders fetch the next block
        if (dec && d.phase == npz::PH_DONE) { status[blk] = npz::dec_status(d); d.phase = npz::PH_IDLE; }
        while (dec && !exhausted && d.phase == npz::PH_IDLE) {
            blk = atomicAdd(ticket, 1);
            if (blk >= n_blocks) { exhausted = true; break; }
            const npz::Block b = blocks[blk];
            if (b.out_len) npz::dec_start(d, comp + b.in_off, b.in_len, out + b.out_off, b.out_len);
            else status[blk] = npz::OK;
        }
        if (!w.any(dec && d.phase != npz::PH_IDLE)) break;
        int32_t nsym = 0;
        if (dec) {
            while (d.phase == npz::PH_HEADER) npz::dec_header(d, mine);
            if (d.phase == npz::PH_SYMBOLS) nsym = npz::dec_batch(d, mine);
        }
        w.sync();                                                    // batches and stored bytes are visible to every lane
        #pragma unroll
        for (int k = 0; k < kDecoders; k++) {
            const int32_t nk = w.shfl(nsym, k);
            if (nk == 0) continue;
            const int32_t pk = w.shfl(d.pos, k);
            const uint32_t olen = (uint32_t)w.shfl((int32_t)d.out_len, k);
            const unsigned long long op = (unsigned long long)(uintptr_t)d.out;
            uint8_t* ok = (uint8_t*)(uintptr_t)((unsigned long long)(uint32_t)w.shfl((int32_t)(op & 0xffffffffull), k) |
                                                (unsigned long long)(uint32_t)w.shfl((int32_t)(op >> 32), k) << 32);
            const int32_t np = npz::resolve_batch(ok, olen, pk, nk, tabs[wid][k], w);
            if (lane == k) { if (np < 0) { d.err = -np; d.phase = npz::PH_DONE; } else d.pos = np; }
        }
    }
}

// decoders per warp: NEXTPOLISH_B200_INFLATE_DECODERS = 1 (default), 2 or 4
static int inflate_decoders() {
    static int k = [] { const char* e = getenv("NEXTPOLISH_B200_INFLATE_DECODERS"); const int v = e ? atoi(e) : 1; return v == 2 || v == 4 ? v : 1; }();
    return k;
}
static void launch_inflate(const uint8_t* comp, const npz::Block* blocks, int32_t nb, uint8_t* out, int32_t* status, int32_t* ticket,
                           int sms, cudaStream_t stream) {
    const int k = inflate_decoders();
    const int per_cta = k == 1 ? 8 : k == 2 ? 8 : 8;                 // blocks in flight per CTA: decoders x warps
    int ctas = (nb + per_cta - 1) / per_cta;
    if (ctas > sms * 4) ctas = sms * 4;                               // persistent warps pull blocks from the ticket
    if (k == 1) k_bgzf_inflate<1, 8, 4><<<ctas, 8 * 32, 0, stream>>>(comp, blocks, nb, out, status, ticket);
    else if (k == 2) k_bgzf_inflate<2, 4, 4><<<ctas, 4 * 32, 0, stream>>>(comp, blocks, nb, out, status, ticket);
    else k_bgzf_inflate<4, 2, 4><<<ctas, 2 * 32, 0, stream>>>(comp, blocks, nb, out, status, ticket);
}

struct InflateCtx {     // buffers kept across calls (per process)
    void *d_comp = nullptr, *d_out = nullptr, *d_blocks = nullptr, *d_status = nullptr;
    size_t cap_comp = 0, cap_out = 0, cap_blocks = 0, cap_status = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};
static InflateCtx g_ctx;

static bool reserve(void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes && *p) return true;
    if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    if (cudaMalloc(p, want) != cudaSuccess) return false;
    *cap = want;
    return true;
}

}  // namespace

namespace {
// Copies a byte range into pinned memory with a few helper threads.  The helpers are created once and live for the
// process (spawning and joining seven threads per load meant stack mappings made and torn down under the process-wide
// memory-map lock, next to six other loading threads); one copy runs at a time: with several loads in flight
// simultaneous copies would put slots x helpers on the host's cores at once.
class StagePool {
public:
    void copy(uint8_t* dst, const uint8_t* src, size_t n) {
        std::lock_guard<std::mutex> one(job_mu_);
        start_threads();
        const size_t parts = th_.size() + 1, step = (n + parts - 1) / parts;
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = dst; src_ = src; n_ = n; step_ = step; pending_ = (int)th_.size(); gen_++;
        }
        cv_.notify_all();
        memcpy(dst, src, std::min(step, n));                          // the caller takes the first part
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }
private:
    void start_threads() {
        if (started_) return;
        started_ = true;
        const unsigned hw = std::thread::hardware_concurrency();
        const int nt = hw >= 16 ? 7 : hw >= 8 ? 3 : (hw >= 4 ? 1 : 0);
        for (int t = 0; t < nt; t++) th_.emplace_back([this, t] { run(t + 1); });
        for (auto& x : th_) x.detach();                               // process lifetime: they sleep on the condition variable
    }
    void run(int part) {
        uint64_t seen = 0;
        for (;;) {
            uint8_t* d; const uint8_t* s; size_t n, step;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_; d = dst_; s = src_; n = n_; step = step_;
            }
            const size_t a = (size_t)part * step, b = std::min(n, a + step);
            if (a < b) memcpy(d + a, s + a, b - a);
            {
                std::lock_guard<std::mutex> lk(mu_);
                pending_--;
            }
            done_.notify_all();
        }
    }
    std::mutex job_mu_, mu_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> th_;
    bool started_ = false;
    uint8_t* dst_ = nullptr; const uint8_t* src_ = nullptr; size_t n_ = 0, step_ = 0;
    int pending_ = 0; uint64_t gen_ = 0;
};
void stage_copy(uint8_t* dst, const uint8_t* src, size_t n) {
    // never destroyed (its threads outlive static destructors); rebuilt in a forked child, whose copy has no threads
    static std::mutex mu; static StagePool* pool = nullptr; static pid_t owner = 0;
    StagePool* p;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!pool || owner != getpid()) { pool = new StagePool(); owner = getpid(); }
        p = pool;
    }
    p->copy(dst, src, n);
}
}  // namespace

namespace npz_dev {
cudaMemPool_t thread_pool(int device) {
    static thread_local std::vector<std::pair<int, cudaMemPool_t>> pools;
    for (auto& p : pools) if (p.first == device) return p.second;
    cudaMemPool_t pool = nullptr;
    cudaMemPoolProps props;
    memset(&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); cudaDeviceGetDefaultMemPool(&pool, device); }
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    pools.emplace_back(device, pool);
    return pool;
}
// Asynchronous inflate of `blocks` (payload offsets relative to comp_host) into d_out (device) on `stream`:
// inflate_launch enqueues the copies and the kernel and returns; inflate_finish synchronises the stream and
// checks every block's status.  Used by callers that keep the bytes in HBM (devload.cu) and overlap host work.
int32_t inflate_launch(InflateJob& j, const uint8_t* comp_host, size_t comp_bytes, const std::vector<npz::Block>& blocks,
                       uint8_t* d_out, cudaStream_t stream, std::string& err, bool src_pinned) {
    j.nb = blocks.size(); j.stream = stream;
    if (j.nb == 0) return NP_OK;
    const size_t nb = j.nb;
    int dev0 = 0;
    cudaGetDevice(&dev0);
    cudaMemPool_t pool = thread_pool(dev0);
    if (cudaMallocFromPoolAsync(&j.d_comp, comp_bytes + 16, pool, stream) != cudaSuccess || cudaMallocFromPoolAsync(&j.d_blocks, nb * sizeof(npz::Block), pool, stream) != cudaSuccess ||
        cudaMallocFromPoolAsync(&j.d_status, (nb + 1) * 4, pool, stream) != cudaSuccess) { err = "cudaMalloc failed"; return NP_ERR_CUDA; }
    // The compressed bytes usually sit in pageable memory (an mmap of the BAM), which the driver would stage with one
    // thread (~11 GB/s measured).  Larger ranges are copied into a grow-only pinned buffer by a few host threads
    // and go up as one asynchronous DMA; the buffer is reused by the next job only after inflate_finish synchronised.
    static thread_local void* pinned = nullptr; static thread_local size_t pinned_bytes = 0;   // per host thread: loads may run concurrently
    const uint8_t* src = comp_host;
    if (!src_pinned && comp_bytes >= (4u << 20)) {
        if (pinned_bytes < comp_bytes) {
            if (pinned) { cudaFreeHost(pinned); pinned = nullptr; pinned_bytes = 0; }
            size_t want = comp_bytes + comp_bytes / 4;
            if (cudaMallocHost(&pinned, want) == cudaSuccess) pinned_bytes = want; else { pinned = nullptr; cudaGetLastError(); }
        }
        if (pinned) {
            // one range at a time, copied by a few persistent helper threads (StagePool above)
            stage_copy((uint8_t*)pinned, comp_host, comp_bytes);
            src = (const uint8_t*)pinned;
        }
    }
    cudaMemcpyAsync(j.d_comp, src, comp_bytes, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(j.d_blocks, blocks.data(), nb * sizeof(npz::Block), cudaMemcpyHostToDevice, stream);
    cudaMemsetAsync(j.d_status, 0xff, nb * 4, stream);
    cudaMemsetAsync((int32_t*)j.d_status + nb, 0, 4, stream);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaEventCreate(&j.e0); cudaEventCreate(&j.e1);
    cudaEventRecord(j.e0, stream);
    launch_inflate((const uint8_t*)j.d_comp, (const npz::Block*)j.d_blocks, (int32_t)nb, d_out, (int32_t*)j.d_status, (int32_t*)j.d_status + nb, sms, stream);
    cudaEventRecord(j.e1, stream);
    return NP_OK;
}
int32_t inflate_finish(InflateJob& j, float* kernel_ms, std::string& err) {
    if (j.nb == 0) return NP_OK;
    std::vector<int32_t> st(j.nb);
    cudaMemcpyAsync(st.data(), j.d_status, j.nb * 4, cudaMemcpyDeviceToHost, j.stream);
    cudaError_t er = np_wait::stream_wait(j.stream);
    if (kernel_ms && er == cudaSuccess) cudaEventElapsedTime(kernel_ms, j.e0, j.e1);
    cudaEventDestroy(j.e0); cudaEventDestroy(j.e1);
    cudaFreeAsync(j.d_comp, j.stream); cudaFreeAsync(j.d_blocks, j.stream); cudaFreeAsync(j.d_status, j.stream);
    j.nb = 0;
    if (er != cudaSuccess) { err = cudaGetErrorString(er); return NP_ERR_CUDA; }
    for (size_t i = 0; i < st.size(); i++)
        if (st[i] != npz::OK) { err = "BGZF block " + std::to_string(i) + " failed with inflate error " + std::to_string(st[i]); return NP_ERR_IO; }
    return NP_OK;
}
}  // namespace npz_dev

extern "C" {

// Inflates a whole BGZF byte range (concatenated blocks, e.g. a BAM file or the chunk of one contig) on the GPU:
// block headers are parsed on the host, the compressed bytes are copied to HBM, one warp inflates each block,
// the result is copied back.  kernel_ms (optional) receives the device time of the inflate kernel alone.
int32_t np_bgzf_inflate(int32_t device, const uint8_t* comp, int64_t comp_bytes, uint8_t* out, int64_t out_cap,
                        int64_t* out_bytes, int32_t* n_blocks_out, float* kernel_ms) {
    if (!comp || comp_bytes < 0 || !out_bytes) { np::set_error("np_bgzf_inflate: bad arguments"); return NP_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        np::set_error("np_bgzf_inflate: no usable CUDA device; this engine has no CPU path");
        return NP_ERR_CUDA;
    }
    cudaSetDevice(device);
    std::vector<npz::Block> blocks;
    std::string err;
    int64_t total = 0;
    if (!np::bgzf_scan(comp, (size_t)comp_bytes, blocks, total, err)) { np::set_error("np_bgzf_inflate: " + err); return NP_ERR_IO; }
    *out_bytes = total;
    if (n_blocks_out) *n_blocks_out = (int32_t)blocks.size();
    if (!out) return NP_OK;                                   // size query
    if (out_cap < total) { np::set_error("np_bgzf_inflate: output buffer too small"); return NP_ERR_ARG; }
    if (blocks.empty() || total == 0) return NP_OK;
    InflateCtx& c = g_ctx;
    if (!c.stream) {
        cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking);
        cudaEventCreate(&c.e0); cudaEventCreate(&c.e1);
    }
    const size_t nb = blocks.size();
    if (!reserve(&c.d_comp, &c.cap_comp, (size_t)comp_bytes + 16) || !reserve(&c.d_out, &c.cap_out, (size_t)total + 16) ||
        !reserve(&c.d_blocks, &c.cap_blocks, nb * sizeof(npz::Block)) || !reserve(&c.d_status, &c.cap_status, (nb + 1) * 4)) {
        np::set_error("np_bgzf_inflate: cudaMalloc failed");
        return NP_ERR_CUDA;
    }
    cudaMemcpyAsync(c.d_comp, comp, (size_t)comp_bytes, cudaMemcpyHostToDevice, c.stream);
    cudaMemcpyAsync(c.d_blocks, blocks.data(), nb * sizeof(npz::Block), cudaMemcpyHostToDevice, c.stream);
    cudaMemsetAsync(c.d_status, 0xff, (nb + 1) * 4, c.stream);
    cudaMemsetAsync((int32_t*)c.d_status + nb, 0, 4, c.stream);          // ticket counter
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaEventRecord(c.e0, c.stream);
    launch_inflate((const uint8_t*)c.d_comp, (const npz::Block*)c.d_blocks, (int32_t)nb, (uint8_t*)c.d_out, (int32_t*)c.d_status, (int32_t*)c.d_status + nb, sms, c.stream);
    cudaEventRecord(c.e1, c.stream);
    std::vector<int32_t> st(nb);
    cudaMemcpyAsync(st.data(), c.d_status, nb * 4, cudaMemcpyDeviceToHost, c.stream);
    cudaMemcpyAsync(out, c.d_out, (size_t)total, cudaMemcpyDeviceToHost, c.stream);
    cudaError_t er = np_wait::stream_wait(c.stream);
    if (er != cudaSuccess) { np::set_error(std::string("np_bgzf_inflate: ") + cudaGetErrorString(er)); return NP_ERR_CUDA; }
    if (kernel_ms) cudaEventElapsedTime(kernel_ms, c.e0, c.e1);
    for (size_t i = 0; i < nb; i++)
        if (st[i] != npz::OK) {
            char m[96];
            snprintf(m, sizeof m, "np_bgzf_inflate: block %zu failed with inflate error %d", i, st[i]);
            np::set_error(m);
            return NP_ERR_IO;
        }
    return NP_OK;
}

}  // extern "C"
