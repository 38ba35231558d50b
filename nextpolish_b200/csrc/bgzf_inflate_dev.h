// bgzf_inflate_dev.h — device-resident form of the BGZF inflate (bgzf_inflate.cu) for callers that keep the inflated
// bytes in HBM and overlap host work with the copy and the kernel (devload.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "bgzf_inflate.h"

namespace npz_dev {

// Stream-ordered allocations come from a memory pool PRIVATE to the calling host thread (and device): loads issued from
// different threads (the files pipeline runs one worker per slot) never reuse each other's freed blocks, so the
// allocator never orders one worker's stream behind the other's.  Release threshold = keep everything cached.
cudaMemPool_t thread_pool(int device);

struct InflateJob {
    void *d_comp = nullptr, *d_blocks = nullptr, *d_status = nullptr;
    size_t nb = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t stream = nullptr;
};
// Enqueues the upload of the compressed bytes (block payload offsets are relative to comp_host) and the inflate
// kernel on `stream`, output into d_out (device); returns without waiting.
// src_pinned: comp_host lies in page-locked memory (e.g. a BAM mapping registered with cudaHostRegister): it is
// uploaded by one asynchronous DMA as it is; otherwise larger ranges are first copied into a pinned staging buffer.
int32_t inflate_launch(InflateJob& j, const uint8_t* comp_host, size_t comp_bytes, const std::vector<npz::Block>& blocks,
                       uint8_t* d_out, cudaStream_t stream, std::string& err, bool src_pinned = false);
// Synchronises the stream, releases the job's buffers and checks every block's status.
int32_t inflate_finish(InflateJob& j, float* kernel_ms, std::string& err);

}  // namespace npz_dev
