// bgzf_inflate.cu — GPU inflate of BGZF blocks (one warp per block) and its C ABI.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo (nextpolish_b200/csrc/Makefile).
#include <cuda_runtime.h>

#include <cstdio>
#include <algorithm>
#include <cstring>
#include <thread>
#include <string>
#include <vector>

#include "bgzf_inflate.h"
#include "bgzf_inflate_dev.h"
#include "errors.h"
#include "hostio.h"
#include "../../include/nextpolish_b200.h"

namespace {

struct WarpBackend {
    __device__ __forceinline__ int32_t lane() const { return (int32_t)(threadIdx.x & 31u); }
    __device__ __forceinline__ int32_t width() const { return 32; }
    __device__ __forceinline__ int32_t bcast(int32_t v) const { return __shfl_sync(0xffffffffu, v, 0); }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};

constexpr int kWarpsPerCta = 8;

// One warp per BGZF block; blocks are handed out through an atomic ticket so that the long blocks of a batch
// do not wait behind a static assignment.
__global__ void __launch_bounds__(kWarpsPerCta * 32, 4) k_bgzf_inflate(const uint8_t* comp, const npz::Block* blocks, int32_t n_blocks,
                                                                    uint8_t* out, int32_t* status, int32_t* ticket) {
    __shared__ npz::Tables tabs[kWarpsPerCta];
    const int wid = (int)(threadIdx.x >> 5);
    WarpBackend w;
    for (;;) {
        int32_t i = 0;
        if (w.lane() == 0) i = atomicAdd(ticket, 1);
        i = w.bcast(i);
        if (i >= n_blocks) break;
        const npz::Block b = blocks[i];
        int rc = npz::OK;
        if (b.out_len) rc = npz::inflate_block(comp + b.in_off, b.in_len, out + b.out_off, b.out_len, tabs[wid], w);
        if (w.lane() == 0) status[i] = rc;
        w.sync();
    }
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

namespace npz_dev {
// Asynchronous inflate of `blocks` (payload offsets relative to comp_host) into d_out (device) on `stream`:
// inflate_launch enqueues the copies and the kernel and returns; inflate_finish synchronises the stream and
// checks every block's status.  Used by callers that keep the bytes in HBM (devload.cu) and overlap host work.
int32_t inflate_launch(InflateJob& j, const uint8_t* comp_host, size_t comp_bytes, const std::vector<npz::Block>& blocks,
                       uint8_t* d_out, cudaStream_t stream, std::string& err) {
    j.nb = blocks.size(); j.stream = stream;
    if (j.nb == 0) return NP_OK;
    const size_t nb = j.nb;
    if (cudaMallocAsync(&j.d_comp, comp_bytes + 16, stream) != cudaSuccess || cudaMallocAsync(&j.d_blocks, nb * sizeof(npz::Block), stream) != cudaSuccess ||
        cudaMallocAsync(&j.d_status, (nb + 1) * 4, stream) != cudaSuccess) { err = "cudaMalloc failed"; return NP_ERR_CUDA; }
    // The compressed bytes usually sit in pageable memory (an mmap of the BAM), which the driver would stage with one
    // thread (~11 GB/s measured).  Larger ranges are copied into a grow-only pinned buffer by a few host threads
    // and go up as one asynchronous DMA; the buffer is reused by the next job only after inflate_finish synchronised.
    static thread_local void* pinned = nullptr; static thread_local size_t pinned_bytes = 0;   // per host thread: loads may run concurrently
    const uint8_t* src = comp_host;
    if (comp_bytes >= (4u << 20)) {
        if (pinned_bytes < comp_bytes) {
            if (pinned) { cudaFreeHost(pinned); pinned = nullptr; pinned_bytes = 0; }
            size_t want = comp_bytes + comp_bytes / 4;
            if (cudaMallocHost(&pinned, want) == cudaSuccess) pinned_bytes = want; else { pinned = nullptr; cudaGetLastError(); }
        }
        if (pinned) {
            unsigned hw = std::thread::hardware_concurrency();
            const size_t nt = hw >= 8 ? 4 : (hw >= 4 ? 2 : 1);
            std::vector<std::thread> th;
            const size_t step = (comp_bytes + nt - 1) / nt;
            uint8_t* const pin = (uint8_t*)pinned;      // a thread_local is not captured: the helper threads need the value
            for (size_t t = 1; t < nt; t++) {
                const size_t a = t * step, b = std::min(comp_bytes, a + step);
                if (a < b) th.emplace_back([=] { memcpy(pin + a, comp_host + a, b - a); });
            }
            memcpy(pinned, comp_host, std::min(step, comp_bytes));
            for (auto& x : th) x.join();
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
    int ctas = (int)((nb + kWarpsPerCta - 1) / kWarpsPerCta);
    if (ctas > sms * 8) ctas = sms * 8;
    cudaEventCreate(&j.e0); cudaEventCreate(&j.e1);
    cudaEventRecord(j.e0, stream);
    k_bgzf_inflate<<<ctas, kWarpsPerCta * 32, 0, stream>>>((const uint8_t*)j.d_comp, (const npz::Block*)j.d_blocks, (int32_t)nb, d_out,
                                                          (int32_t*)j.d_status, (int32_t*)j.d_status + nb);
    cudaEventRecord(j.e1, stream);
    return NP_OK;
}
int32_t inflate_finish(InflateJob& j, float* kernel_ms, std::string& err) {
    if (j.nb == 0) return NP_OK;
    std::vector<int32_t> st(j.nb);
    cudaMemcpyAsync(st.data(), j.d_status, j.nb * 4, cudaMemcpyDeviceToHost, j.stream);
    cudaError_t er = cudaStreamSynchronize(j.stream);
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
    int ctas = (int)((nb + kWarpsPerCta - 1) / kWarpsPerCta);
    if (ctas > sms * 8) ctas = sms * 8;                                   // persistent warps pull blocks from the ticket
    cudaEventRecord(c.e0, c.stream);
    k_bgzf_inflate<<<ctas, kWarpsPerCta * 32, 0, c.stream>>>((const uint8_t*)c.d_comp, (const npz::Block*)c.d_blocks, (int32_t)nb,
                                                            (uint8_t*)c.d_out, (int32_t*)c.d_status, (int32_t*)c.d_status + nb);
    cudaEventRecord(c.e1, c.stream);
    std::vector<int32_t> st(nb);
    cudaMemcpyAsync(st.data(), c.d_status, nb * 4, cudaMemcpyDeviceToHost, c.stream);
    cudaMemcpyAsync(out, c.d_out, (size_t)total, cudaMemcpyDeviceToHost, c.stream);
    cudaError_t er = cudaStreamSynchronize(c.stream);
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
