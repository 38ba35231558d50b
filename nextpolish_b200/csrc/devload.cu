// devload.cu — BAM bytes -> packed shard, built in HBM (second piece).
//
// The reference decodes the BAM on the host through htslib (bgzf.c inflate + bam_read1, contig.c:170-180 and
// :688-704, twice per contig); our host packer (hostio.cpp) does the same work once and is the real-world
// bottleneck (~2-3 Mbp/s against ~3000 Mbp/s for the polishing kernels).  Here the compressed byte range of the
// wanted contigs goes to the GPU as it is on disk and everything else happens there:
//
//   k_bgzf_inflate   one warp per BGZF block                                            (bgzf_inflate.h)
//   k_rec_walk       one thread per ANCHOR interval: record boundaries.  A BAM record only says how long it is,
//                    so boundaries are a sequential chain; the .bai index knows a true record start for every
//                    chunk and every 16 kb window (bai_record_starts), and every chain must land exactly on the
//                    next anchor — a chain that starts on a true record start and is followed faithfully IS the
//                    file's record sequence, so this is a proof, not a heuristic
//   k_rec_meta       one thread per record: fields, contig slot, 2-bit / 4-bit decision, packed size, whether the
//                    qualities are shipped (sparse mode: the span contains a lowercase draft base)
//   scans            record / quality offsets (CUB)
//   k_rec_pack       one warp per kept record: header, CIGAR, bases, qualities (coalesced byte traffic)
//
// The result is byte-identical to the host packer's shard (tests/test_devload.py) and is adopted by the engine in
// place (np_engine_adopt_device).
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <sys/stat.h>

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "bgzf_inflate.h"
#include "bgzf_inflate_dev.h"
#include "errors.h"
#include "stream_wait.h"
#include "hostio.h"
#include "../../include/nextpolish_b200.h"

namespace {

__device__ __forceinline__ uint32_t ld32(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
__device__ __forceinline__ uint32_t ld16(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }

enum { DL_ERR_CHAIN = 1, DL_ERR_RECORD = 2, DL_ERR_LONG = 4, DL_ERR_ORDER = 8 };

// One walk per anchor interval: counts the records and writes their offsets into the interval's scratch range
// (capacity = interval bytes / 36, a record being at least 4 + 32 bytes; range starts computed on the host).
// A record only says how long it is, so the walk is a chain of dependent loads; one WARP per interval streams the
// interval through shared memory in 4 KiB windows (16-byte cp.async copies, double-buffered: the window that follows is
// in flight while lane 0 follows the chain through the current one): no global round trip on the chain at all, where the
// first version paid one per record and the second one per window.
constexpr int kWalkWin = 4096, kWalkWarps = 4;
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__global__ void __launch_bounds__(kWalkWarps * 32) k_rec_walk(const uint8_t* U, int64_t total, const int64_t* anchors, const int64_t* cap_base, int32_t n_int,
                                                             int32_t* cnt, int64_t* scratch, int32_t* err) {
    __shared__ __align__(16) uint8_t win[kWalkWarps][2][kWalkWin + 16];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int32_t i = (int32_t)(blockIdx.x * kWalkWarps + wid);
    if (i >= n_int) return;
    int64_t off = anchors[i];
    const int64_t end = anchors[i + 1];
    int32_t n = 0;
    int64_t* out = scratch + cap_base[i];
    const int64_t cap = cap_base[i + 1] - cap_base[i];
    bool bad = false;
    // a window holds the kWalkWin + 16 bytes from `b` on (U comes from the allocator: 16-byte aligned windows; the inflate
    // buffer has 16 bytes of slack; chunks past it read as zeros)
    auto fetch = [&](int buf, int64_t b) {
        uint8_t* w = win[wid][buf];
        for (int k = lane; k < (kWalkWin + 16) / 16; k += 32) {
            const int64_t a = b + 16 * k;
            if (a + 16 <= total + 16) cp_async16(w + 16 * k, U + a);
            else *(uint4*)(w + 16 * k) = make_uint4(0u, 0u, 0u, 0u);
        }
        cp_async_commit();
    };
    int64_t base = off & ~(int64_t)15;
    int cur = 0;
    if (off < end) fetch(0, base);
    while (off < end && !bad) {
        fetch(cur ^ 1, base + kWalkWin);                                // the window that follows, in flight during the walk
        cp_async_wait<1>();                                             // the current window has landed
        __syncwarp();
        if (lane == 0) {
            const uint8_t* w = win[wid][cur];
            while (off < end && off + 4 <= base + kWalkWin + 16) {
                if (off + 4 > total) { bad = true; break; }
                const uint8_t* p = w + (off - base);
                const uint32_t bs = (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
                if (bs < 32u || off + 4 + (int64_t)bs > total || n >= cap) { bad = true; break; }
                out[n++] = off;
                off += 4 + (int64_t)bs;
            }
        }
        off = __shfl_sync(0xffffffffu, off, 0);
        bad = __shfl_sync(0xffffffffu, (int)bad, 0) != 0;
        __syncwarp();                                                   // lane 0's reads of this window come before the next fetch into it
        const int64_t nb = base + kWalkWin;
        if (off + 4 <= nb + kWalkWin + 16) { base = nb; cur ^= 1; }     // the chain continues inside the prefetched window
        else {                                                          // a record longer than a window: start over at its end
            cp_async_wait<0>();
            __syncwarp();
            base = off & ~(int64_t)15;
            if (off < end && !bad) fetch(cur, base);
        }
    }
    cp_async_wait<0>();
    if (lane == 0) {
        if (bad || off != end) atomicOr(err, DL_ERR_CHAIN);
        cnt[i] = n;
    }
}
// scratch ranges -> one dense array of record offsets (one warp per interval)
__global__ void k_rec_compact(const int64_t* scratch, const int64_t* cap_base, const int32_t* cnt, const int32_t* base, int32_t n_int,
                              int64_t* rec_start) {
    int32_t i = (int32_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = (int32_t)(threadIdx.x & 31u);
    if (i >= n_int) return;
    const int64_t* in = scratch + cap_base[i];
    int64_t* out = rec_start + base[i];
    for (int32_t j = lane; j < cnt[i]; j += 32) out[j] = in[j];
}

struct MetaArgs {
    const uint8_t* U; const int64_t* rec_start; int32_t n_rec;
    const int32_t* slot_of_tid; int32_t n_ref;
    const int64_t* slot_goff;      // [n_slots + 1] offset of every slot's contig in the concatenated draft
    const int32_t* lcpre;          // [G + 1] lowercase prefix counts of the draft (qual mode 2) or null
    int32_t qual_mode, two_bit;
    int32_t *keep, *units, *qunits, *slot; uint8_t* enc;
    int32_t* err;
};
__global__ void k_rec_meta(MetaArgs a) {
    int32_t r = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x);
    if (r >= a.n_rec) return;
    const uint8_t* p = a.U + a.rec_start[r];
    const uint32_t bs = ld32(p);
    p += 4;
    const int32_t tid = (int32_t)ld32(p), pos = (int32_t)ld32(p + 4);
    const uint32_t l_name = p[8], n_cigar = ld16(p + 12);
    const int32_t l_seq = (int32_t)ld32(p + 16);
    int32_t keep = 0, units = 0, qunits = 0, slot = -1; uint32_t enc = 0;
    if (l_seq < 0 || 32ull + l_name + 4ull * n_cigar + ((uint64_t)l_seq + 1) / 2 + (uint64_t)l_seq > (uint64_t)bs) atomicOr(a.err, DL_ERR_RECORD);
    else if (tid >= 0 && tid < a.n_ref && (slot = a.slot_of_tid[tid]) >= 0 && n_cigar > 0) {
        keep = 1;
        if (l_seq > 65535) atomicOr(a.err, DL_ERR_LONG);
        const uint8_t* cig = p + 32 + l_name;
        const uint8_t* seq = cig + 4 * n_cigar;
        bool two = a.two_bit && l_seq > 0;
        for (int32_t i = 0; two && i < l_seq; i++) {
            uint32_t c = (seq[i >> 1] >> ((~i & 1) << 2)) & 0xfu;
            two = c == 1 || c == 2 || c == 4 || c == 8;
        }
        enc = two ? 1u : 0u;
        const uint32_t seq_bytes = two ? ((uint32_t)l_seq + 3) / 4 : ((uint32_t)l_seq + 1) / 2;
        units = (int32_t)((16u + 4u * n_cigar + seq_bytes + 15u) / 16u);
        bool kq = a.qual_mode == 1;
        if (a.qual_mode == 2) {
            int64_t end = pos;
            for (uint32_t i = 0; i < n_cigar; i++) {
                uint32_t c = ld32(cig + 4 * i), op = c & 0xfu;
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) end += c >> 4;
            }
            const int64_t g0 = a.slot_goff[slot], L = a.slot_goff[slot + 1] - g0;
            const int64_t x = pos < 0 ? 0 : pos, y = end > L ? L : end;
            kq = x < y && a.lcpre[g0 + y] - a.lcpre[g0 + x] > 0;
        }
        if (kq) qunits = (l_seq + 15) / 16;
    }
    if (r > 0 && keep) {     // coordinate-sorted input: contig slots never go backwards
        const uint8_t* q = a.U + a.rec_start[r - 1] + 4;
        const int32_t ptid = (int32_t)ld32(q);
        if (ptid >= 0 && ptid < a.n_ref && a.slot_of_tid[ptid] > slot && ld16(q + 12) > 0) atomicOr(a.err, DL_ERR_ORDER);
    }
    a.keep[r] = keep; a.units[r] = units; a.qunits[r] = qunits; a.slot[r] = slot; a.enc[r] = (uint8_t)enc;
}

struct PackArgs {
    const uint8_t* U; const int64_t* rec_start; int32_t n_rec;
    const int32_t *keep, *kidx, *uoff, *quoff, *qunits, *slot; const uint8_t* enc;
    uint32_t* rec_off; uint8_t* rec; uint32_t* qual_off; uint8_t* qual;
    int32_t* slot_count; int32_t n_keep, total_units, total_qunits;
    int64_t u_bytes;            // inflated bytes in U (the buffer has 16 more)
};
// One WARP per record, 32 records per warp and round (round 2; one thread per record with byte loops ran at 1.1 ms per
// 5 Mb shard because neighbouring threads walked different 280-byte records).  ncu on the first warp-per-record builds
// showed the kernel waiting on chains of small dependent loads (long-scoreboard stalls 23 per issue, 390 GB/s): eight
// per-record scalars, then the header byte, CIGAR, bases, qualities, each its own 32-byte request.  So
//   * the per-record scalars of 32 consecutive records are loaded at once, one record per lane (coalesced), and handed
//     round by shuffles; the offset tables and the per-contig counts are written from that form, too;
//   * a record is staged in shared memory with ONE request of 16 bytes per lane (512 bytes; longer records take a second
//     round or, beyond the staging size, are read in place), and the request for the NEXT record is issued before the
//     current one is processed;
//   * from the staged copy the lanes write consecutive output bytes.
constexpr int kPackWarps = 8, kPackStage = 1024;
struct PackRec { int64_t start; int32_t uoff, quoff, qunits; uint32_t enc; };
__device__ __forceinline__ void pack_one(const PackArgs& a, const PackRec& m, const uint8_t* p, int lane) {
    // the fixed 32-byte part of the record: one byte per lane, fields through shuffles
    const uint32_t hb = p[lane];
    const uint32_t l_name = __shfl_sync(0xffffffffu, hb, 8);
    const uint32_t n_cigar = __shfl_sync(0xffffffffu, hb, 12) | __shfl_sync(0xffffffffu, hb, 13) << 8;
    const int32_t l_seq = (int32_t)(__shfl_sync(0xffffffffu, hb, 16) | __shfl_sync(0xffffffffu, hb, 17) << 8 |
                                    __shfl_sync(0xffffffffu, hb, 18) << 16 | __shfl_sync(0xffffffffu, hb, 19) << 24);
    uint8_t* __restrict__ d = a.rec + (size_t)m.uoff * 16;
    // header: pos, flag, mapq, enc, isize, l_qseq, n_cigar (include/nextpolish_b200.h): byte i of the packed header
    // comes from byte src[i] of the BAM record (0xff: not a copy)
    {
        const uint64_t src_lo = 0xff090f0e07060504ull, src_hi = 0x0d0cffff1f1e1d1cull;     // d[0..7], d[8..15]
        const uint32_t sidx = lane < 16 ? (uint32_t)(((lane < 8 ? src_lo : src_hi) >> (8 * (lane & 7))) & 0xffu) : 0u;
        uint32_t v = __shfl_sync(0xffffffffu, hb, (int)(sidx & 31u));
        if (lane == 7) v = m.enc;
        if (lane == 12) v = (uint32_t)l_seq & 0xffu;
        if (lane == 13) v = ((uint32_t)l_seq >> 8) & 0xffu;
        if (lane < 16) d[lane] = (uint8_t)v;
    }
    const uint8_t* cig = p + 32 + l_name;
    for (uint32_t i = (uint32_t)lane; i < 4 * n_cigar; i += 32) d[16 + i] = cig[i];
    const uint8_t* seq = cig + 4 * n_cigar;
    uint8_t* __restrict__ ds = d + 16 + 4 * n_cigar;
    if (!m.enc) { for (int32_t i = lane; i < (l_seq + 1) / 2; i += 32) ds[i] = seq[i]; }
    else {
        // four bases (two 4-bit bytes) -> one 2-bit byte; bases past l_seq contribute zero bits
        for (int32_t bq = lane; bq < (l_seq + 3) / 4; bq += 32) {
            const uint32_t b0 = seq[2 * bq], b1 = 4 * bq + 2 < l_seq ? seq[2 * bq + 1] : 0u;
            uint32_t o = 0;
            #pragma unroll
            for (int32_t j = 0; j < 4; j++) {
                const uint32_t c = ((j < 2 ? b0 : b1) >> ((~j & 1) << 2)) & 0xfu;
                const uint32_t two = c == 1 ? 0u : c == 2 ? 1u : c == 4 ? 2u : 3u;
                if (4 * bq + j < l_seq) o |= two << (6 - 2 * j);
            }
            ds[bq] = (uint8_t)o;
        }
    }
    if (a.qual_off && m.qunits) {
        const uint8_t* q = seq + (l_seq + 1) / 2;
        uint8_t* __restrict__ dq = a.qual + (size_t)m.quoff * 16;
        for (int32_t i = lane; i < l_seq; i += 32) dq[i] = q[i];
    }
}
__global__ void __launch_bounds__(kPackWarps * 32) k_rec_pack(PackArgs a) {
    __shared__ __align__(16) uint8_t stage[kPackWarps][kPackStage];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int32_t w0 = (int32_t)(blockIdx.x * kPackWarps + wid), nw = (int32_t)(gridDim.x * kPackWarps);
    if (w0 == 0 && lane == 0) {     // closing entries of the offset arrays
        a.rec_off[a.n_keep] = (uint32_t)a.total_units;
        if (a.qual_off) a.qual_off[a.n_keep] = (uint32_t)a.total_qunits;
    }
    uint8_t* st = stage[wid];
    // the first 512 bytes from the 16-byte boundary at or before a record (the inflate buffer has 16 bytes of slack)
    auto head = [&](int64_t start) {
        const int64_t at = (start & ~(int64_t)15) + 16 * lane;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (at + 16 <= a.u_bytes + 16) v = *(const uint4*)(a.U + at);
        return v;
    };
    for (int32_t c = w0; (int64_t)c * 32 < a.n_rec; c += nw) {
        // this lane's record of the round: its scalars, its entries of the offset tables, its contig's count
        const int32_t r = c * 32 + lane;
        const bool mine = r < a.n_rec && a.keep[r];
        PackRec m{0, 0, 0, 0, 0u};
        if (mine) {
            m.start = a.rec_start[r]; m.uoff = a.uoff[r]; m.enc = a.enc[r];
            const int32_t k = a.kidx[r];
            a.rec_off[k] = (uint32_t)m.uoff;
            if (a.qual_off) { m.quoff = a.quoff[r]; m.qunits = a.qunits[r]; a.qual_off[k] = (uint32_t)m.quoff; }
            atomicAdd(&a.slot_count[a.slot[r]], 1);
        }
        uint32_t todo = __ballot_sync(0xffffffffu, mine);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (todo) v = head(__shfl_sync(0xffffffffu, m.start, __ffs((int)todo) - 1));
        while (todo) {
            const int j = __ffs((int)todo) - 1;
            todo &= todo - 1;
            PackRec x;
            x.start = __shfl_sync(0xffffffffu, m.start, j); x.uoff = __shfl_sync(0xffffffffu, m.uoff, j);
            x.quoff = __shfl_sync(0xffffffffu, m.quoff, j); x.qunits = __shfl_sync(0xffffffffu, m.qunits, j);
            x.enc = __shfl_sync(0xffffffffu, m.enc, j);
            *(uint4*)(st + 16 * lane) = v;
            __syncwarp();
            if (todo) v = head(__shfl_sync(0xffffffffu, m.start, __ffs((int)todo) - 1));       // the next record's bytes: in flight from here
            const int64_t base = x.start & ~(int64_t)15;
            const uint8_t* q = st + (int32_t)(x.start - base);
            const int32_t need = (int32_t)(x.start - base) + 4 + (int32_t)((uint32_t)q[0] | (uint32_t)q[1] << 8 | (uint32_t)q[2] << 16 | (uint32_t)q[3] << 24);
            if (need <= kPackStage) {
                // a record's bytes never end later than the buffer's: chunks that start at or past `need` are not loaded
                if (need > 512) {
                    if (512 + 16 * lane < need) *(uint4*)(st + 512 + 16 * lane) = *(const uint4*)(a.U + base + 512 + 16 * lane);
                    __syncwarp();
                }
                pack_one(a, x, q + 4, lane);
            } else pack_one(a, x, a.U + x.start + 4, lane);        // a long record: read in place
            __syncwarp();                                            // the staging buffer is reused by the next record
        }
    }
}
__global__ void k_lower_flags(const uint8_t* seq, int64_t n, int32_t* f) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f[i] = seq[i] >= 97 && seq[i] <= 122 ? 1 : 0; else if (i == n) f[i] = 0;
}

// Stream-ordered allocations from the device's default memory pool, whose release threshold is raised once so
// that freed blocks stay cached: a load allocates ~20 buffers (276 MB of inflated bytes among them) and
// cudaMalloc / cudaFree would cost more than the kernels.
// One allocation/work stream per (host thread, device): loads issued from different threads (the files pipeline runs
// one worker thread per slot) or for different devices never share a stream, an allocation or a cached BAM handle.
static std::mutex g_pool_mu;
static void pool_setup(int device) {
    static std::set<int> done;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (done.count(device)) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done.insert(device);
}
static cudaStream_t thread_stream(int device) {
    static thread_local std::map<int, cudaStream_t> streams;
    auto it = streams.find(device);
    if (it != streams.end()) return it->second;
    cudaStream_t s = nullptr;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    streams[device] = s;
    return s;
}
struct Dbuf {
    void* p = nullptr; cudaStream_t st = nullptr;
    ~Dbuf() { if (p) cudaFreeAsync(p, st); }
    bool alloc(size_t bytes, cudaStream_t s) {
        st = s;
        int dev = 0;
        cudaGetDevice(&dev);
        return cudaMallocFromPoolAsync(&p, bytes + 256, npz_dev::thread_pool(dev), s) == cudaSuccess;
    }
    template <class T> T* as() const { return (T*)p; }
};

static bool exscan(const int32_t* in, int32_t* out, int n, cudaStream_t s) {
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, n, s);
    Dbuf tmp;
    if (!tmp.alloc(need, s)) return false;
    cub::DeviceScan::ExclusiveSum(tmp.p, need, in, out, n, s);
    return true;                   // tmp is freed in stream order
}

}  // namespace

struct np_dev_shard {
    int device = 0;
    cudaStream_t stream = nullptr;     // the loading thread's stream: every buffer below was allocated on it
    std::vector<std::string> names;
    std::vector<int64_t> ctg_off, ctg_read_off;
    std::vector<int32_t> fasta_rank;
    Dbuf seq, rec_off, rec, qual_off, qual;
    int64_t n_reads = 0, rec_bytes = 0, qual_bytes = 0, seq_bytes = 0;
    bool with_qual = false;
    float ms_inflate = 0, ms_total = 0;
    int64_t comp_bytes = 0, inflated_bytes = 0;
};

extern "C" {

void np_dev_shard_free(np_dev_shard* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) np_wait::stream_wait(s->stream);
    delete s;
}

static np_dev_shard* load_gpu_impl(int32_t device, const char* fasta, const char* bam, const char* const* names_in,
                                   int32_t n_names, int32_t with_qual, const uint8_t* const* pre_seq, const int64_t* pre_len);

np_dev_shard* np_shard_load_gpu(int32_t device, const char* fasta, const char* bam, const char* const* names_in,
                                int32_t n_names, int32_t with_qual) {
    return load_gpu_impl(device, fasta, bam, names_in, n_names, with_qual, nullptr, nullptr);
}
// The same load for callers that already hold the draft in host memory (np_multi parses the FASTA once for all GPUs):
// names[i] has the bases seq[i][0 .. len[i]).
np_dev_shard* np_shard_load_gpu_seqs(int32_t device, const char* bam, const char* const* names, const uint8_t* const* seq,
                                     const int64_t* len, int32_t n_names, int32_t with_qual) {
    if (!names || !seq || !len || n_names <= 0) { np::set_error("np_shard_load_gpu_seqs: bad arguments"); return nullptr; }
    return load_gpu_impl(device, "", bam, names, n_names, with_qual, seq, len);
}

}  // extern "C"

static np_dev_shard* load_gpu_impl(int32_t device, const char* fasta, const char* bam, const char* const* names_in,
                                   int32_t n_names, int32_t with_qual, const uint8_t* const* pre_seq, const int64_t* pre_len) {
    using namespace np;
    if (!fasta || !bam) { set_error("np_shard_load_gpu: fasta / bam is NULL"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        set_error("np_shard_load_gpu: no usable CUDA device; this loader has no CPU path (use np_shard_load)");
        return nullptr;
    }
    cudaSetDevice(device);
    // NEXTPOLISH_B200_TRACE=1: host-side phase times on stderr
    const bool trace = getenv("NEXTPOLISH_B200_TRACE") && getenv("NEXTPOLISH_B200_TRACE")[0] == '1';
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tprev = now();
    auto lap = [&](const char* what) { if (trace) { double t = now(); fprintf(stderr, "np_shard_load_gpu: %-32s %8.2f ms\n", what, t - tprev); tprev = t; } };
    std::string err;
    std::vector<std::string> names;
    for (int32_t i = 0; i < n_names; i++) names.emplace_back(names_in[i]);
    const bool all = names.empty();

    // ---- BAM side first: everything up to the launch of the inflate needs only the BAM and its index, so the
    //      FASTA is read and parsed on the host while the GPU copies and inflates
    // The reference ABI polishes one contig per call (nextpolish1.py:181-189): the open BAM (mmap + header), its
    // parsed index and the name table are kept for the last path used instead of being rebuilt for every contig.
    // Entries are immutable once built and handed out as shared_ptr: a thread keeps its BAM alive while another thread
    // replaces the cache slot; the table itself is guarded by a mutex.
    struct BamEntry {
        std::string path; struct stat st;
        BamFile bf;
        std::vector<std::vector<uint64_t>> starts;
        std::unordered_map<std::string, int> tid_of;
    };
    static std::mutex cache_mu;
    static std::vector<std::shared_ptr<BamEntry>> cache;          // most recently used first, at most 4 entries
    struct stat stnow;
    memset(&stnow, 0, sizeof stnow);
    if (stat(bam, &stnow) != 0) { set_error(std::string("np_shard_load_gpu: cannot stat ") + bam); return nullptr; }
    std::shared_ptr<BamEntry> ent;
    {
        std::lock_guard<std::mutex> lk(cache_mu);
        for (size_t i = 0; i < cache.size(); i++) {
            BamEntry& c = *cache[i];
            if (c.path == bam && c.st.st_size == stnow.st_size && c.st.st_mtime == stnow.st_mtime && c.st.st_ino == stnow.st_ino) {
                ent = cache[i];
                cache.erase(cache.begin() + (long)i);
                cache.insert(cache.begin(), ent);
                break;
            }
        }
    }
    if (!ent) {
        ent = std::make_shared<BamEntry>();
        if (!ent->bf.open(bam, err)) { set_error("np_shard_load_gpu: " + err); return nullptr; }
        if (!ent->bf.bai_record_starts(ent->starts, err)) { set_error("np_shard_load_gpu: needs " + std::string(bam) + ".bai (" + err + ")"); return nullptr; }
        ent->starts.resize(ent->bf.header().names.size());
        for (size_t i = 0; i < ent->bf.header().names.size(); i++) ent->tid_of.emplace(ent->bf.header().names[i], (int)i);
        ent->path = bam; ent->st = stnow;
        std::lock_guard<std::mutex> lk(cache_mu);
        cache.insert(cache.begin(), ent);
        if (cache.size() > 4) cache.pop_back();
    }
    BamFile& bf = ent->bf;
    const std::vector<std::vector<uint64_t>>& starts = ent->starts;
    const std::unordered_map<std::string, int>& tid_of = ent->tid_of;
    const int32_t n_ref = (int32_t)bf.header().names.size();
    int tid_min = 0x7fffffff, tid_max = -1;
    if (all) { if (n_ref > 0) { tid_min = 0; tid_max = n_ref - 1; } }
    else for (const auto& nm : names) {
        auto it = tid_of.find(nm);
        if (it != tid_of.end()) { tid_min = std::min(tid_min, it->second); tid_max = std::max(tid_max, it->second); }
    }
    lap("bam open + header + index");

    pool_setup(device);
    cudaStream_t st = thread_stream(device);      // one stream for allocations and work: stream-ordered reuse is safe
    np_dev_shard* S = new np_dev_shard();
    S->device = device; S->stream = st;
    S->with_qual = with_qual != 0;
    npz_dev::InflateJob job;
    auto fail = [&](const std::string& m) {
        std::string ignored;
        if (job.nb) npz_dev::inflate_finish(job, nullptr, ignored);        // an inflate in flight: wait and release its buffers
        set_error("np_shard_load_gpu: " + m);
        np_wait::stream_wait(st);
        np_dev_shard_free(S);
        return (np_dev_shard*)nullptr;
    };
    struct EventPair {
        cudaEvent_t a = nullptr, b = nullptr;
        EventPair() { cudaEventCreate(&a); cudaEventCreate(&b); }
        ~EventPair() { cudaEventDestroy(a); cudaEventDestroy(b); }
    } ev;
    cudaEvent_t e0 = ev.a, e1 = ev.b;
    cudaEventRecord(e0, st);

    // byte range of the wanted contigs: [vbeg, vend) as virtual offsets
    uint64_t vbeg = 0, vend = (uint64_t)bf.size() << 16;
    bool any_reads = false;
    if (all) { vbeg = bf.header().first_record_voffset; any_reads = true; }
    else if (tid_max >= 0) {
        for (int t = tid_min; t <= tid_max && !any_reads; t++) if (!starts[(size_t)t].empty()) { vbeg = starts[(size_t)t][0]; any_reads = true; }
        for (int t = tid_max + 1; t < n_ref; t++) if (!starts[(size_t)t].empty()) { vend = starts[(size_t)t][0]; break; }
    }
    if ((vbeg >> 16) >= bf.size()) any_reads = false;
    std::vector<npz::Block> blocks; std::vector<uint64_t> coffs; int64_t total = 0;
    std::vector<int64_t> anchors;
    Dbuf U;
    if (any_reads) {
        const size_t cbeg = (size_t)(vbeg >> 16);
        const size_t cend = (vend >> 16) >= bf.size() ? bf.size() : (size_t)(vend >> 16) + 1;
        if (!bgzf_scan(bf.data(), bf.size(), blocks, total, err, &coffs, cbeg, cend)) return fail(err);
        // the scan includes the block that starts before `cend`: ship it whole (payload + 8 trailer bytes)
        const size_t ship = blocks.empty() ? 0 : (size_t)blocks.back().in_off + blocks.back().in_len + 8;
        auto u_of = [&](uint64_t v, int64_t* out) -> bool {
            if ((v >> 16) >= bf.size()) { *out = total; return true; }
            auto it = std::lower_bound(coffs.begin(), coffs.end(), v >> 16);
            if (it == coffs.end() || *it != (v >> 16)) return false;
            const npz::Block& b = blocks[(size_t)(it - coffs.begin())];
            if ((v & 0xffff) > b.out_len) return false;
            *out = (int64_t)b.out_off + (int64_t)(v & 0xffff);
            return true;
        };
        int64_t ub = 0, ue = 0;
        if (!u_of(vbeg, &ub) || !u_of(vend, &ue)) return fail("index offsets do not match the BGZF blocks");
        anchors.push_back(ub);
        for (int t = tid_min; t <= tid_max; t++)
            for (uint64_t v : starts[(size_t)t]) {
                if (v <= vbeg || v >= vend) continue;
                int64_t u;
                if (!u_of(v, &u)) return fail("index offsets do not match the BGZF blocks");
                anchors.push_back(u);
            }
        anchors.push_back(ue);
        std::sort(anchors.begin(), anchors.end());
        anchors.erase(std::unique(anchors.begin(), anchors.end()), anchors.end());
        S->comp_bytes = (int64_t)ship; S->inflated_bytes = total;
        lap("bgzf_scan + anchors");
        if (!U.alloc((size_t)total + 16, st)) return fail("cudaMalloc failed");
        if (npz_dev::inflate_launch(job, bf.data() + cbeg, ship, blocks, U.as<uint8_t>(), st, err) != NP_OK) return fail(err);
        lap("H2D (pageable) + inflate launch");
    }

    // ---- FASTA side (the GPU is busy with the copy and the inflate meanwhile)
    // Whole-file loads parse straight into a grow-only pinned buffer of the calling thread (one copy, asynchronous upload);
    // named subsets go through fasta_load (seeks with the .fai).
    struct PinBuf { uint8_t* p = nullptr; size_t cap = 0; };
    static thread_local PinBuf fa_pin, fa_perm;      // parsed whole file; bases gathered in shard order
    std::vector<std::string> fa_names; std::vector<int64_t> fa_off;
    std::vector<FastaRecord> recs;
    const uint8_t* flat = nullptr;
    if (all) {
        auto grow = [](void* ctx, size_t bytes) -> uint8_t* {
            PinBuf& b = *(PinBuf*)ctx;
            if (b.cap >= bytes) return b.p;
            if (b.p) { cudaFreeHost(b.p); b.p = nullptr; b.cap = 0; }
            const size_t want = bytes + bytes / 4 + 4096;
            if (cudaMallocHost((void**)&b.p, want) != cudaSuccess) { cudaGetLastError(); b.p = nullptr; return nullptr; }
            b.cap = want;
            return b.p;
        };
        // the previous upload out of this buffer has completed: every load ends with a stream synchronisation
        if (!fasta_load_flat(fasta, fa_names, fa_off, grow, &fa_pin, err)) return fail(err);
        flat = fa_pin.p;
    } else if (pre_seq) {
        fa_off.push_back(0);
        for (int32_t i = 0; i < n_names; i++) { fa_names.push_back(names[(size_t)i]); fa_off.push_back(fa_off.back() + pre_len[i]); }
    } else {
        if (!fasta_load(fasta, names, recs, err)) return fail(err);
        fa_off.push_back(0);
        for (auto& r : recs) { fa_names.push_back(r.name); fa_off.push_back(fa_off.back() + (int64_t)r.seq.size()); }
    }
    const size_t n_fa = fa_names.size();
    // contigs in BAM tid order (those absent from the BAM header go last, with no reads) — as shard_load
    struct Slot { int tid; size_t rec_idx; };
    std::vector<Slot> slots;
    for (size_t i = 0; i < n_fa; i++) {
        auto it = tid_of.find(fa_names[i]);
        slots.push_back({it == tid_of.end() ? 0x7fffffff : it->second, i});
    }
    std::stable_sort(slots.begin(), slots.end(), [](const Slot& a, const Slot& b) { return a.tid < b.tid; });
    bool identity = flat != nullptr;
    for (size_t k = 0; k < slots.size() && identity; k++) identity = slots[k].rec_idx == k;
    // (when the order changes, or the bases came from fasta_load / the caller: gathered into the thread's pinned buffer)
    uint8_t* ctg_perm = nullptr; size_t perm_at = 0;
    if (!identity) {
        const size_t need = (size_t)fa_off.back() + 16;
        if (fa_perm.cap < need) {
            if (fa_perm.p) { cudaFreeHost(fa_perm.p); fa_perm.p = nullptr; fa_perm.cap = 0; }
            if (cudaMallocHost((void**)&fa_perm.p, need + need / 4 + 4096) != cudaSuccess) { cudaGetLastError(); return fail("cudaMallocHost failed"); }
            fa_perm.cap = need + need / 4 + 4096;
        }
        ctg_perm = fa_perm.p;
    }
    std::vector<int32_t> slot_of_tid((size_t)std::max(1, n_ref), -1);
    S->ctg_off.push_back(0);
    for (size_t k = 0; k < slots.size(); k++) {
        const size_t ri = slots[k].rec_idx;
        const size_t len = (size_t)(fa_off[ri + 1] - fa_off[ri]);
        S->names.push_back(fa_names[ri]);
        S->fasta_rank.push_back((int32_t)ri);
        if (!identity) {
            const uint8_t* src = pre_seq ? pre_seq[ri] : flat ? flat + fa_off[ri] : (const uint8_t*)recs[ri].seq.data();
            memcpy(ctg_perm + perm_at, src, len);
            perm_at += len;
        }
        S->ctg_off.push_back(S->ctg_off.back() + (int64_t)len);
        if (slots[k].tid != 0x7fffffff && slot_of_tid[(size_t)slots[k].tid] < 0) slot_of_tid[(size_t)slots[k].tid] = (int32_t)k;
    }
    const uint8_t* ctg_ptr = identity ? flat : ctg_perm;
    const size_t ctg_bytes = (size_t)S->ctg_off.back();
    const int32_t n_slots = (int32_t)slots.size();
    if (ctg_bytes >= 0x7fffff00ull) return fail("shard exceeds 2^31 positions: load it in several contig groups");
    S->ctg_read_off.assign((size_t)n_slots + 1, 0);
    S->seq_bytes = (int64_t)ctg_bytes;
    lap("fasta_load + contig table");
    if (!S->seq.alloc(ctg_bytes + 16, st)) return fail("cudaMalloc failed");
    if (ctg_bytes) cudaMemcpyAsync(S->seq.p, ctg_ptr, ctg_bytes, cudaMemcpyHostToDevice, st);
    auto finish_empty = [&]() {
        if (!S->rec_off.alloc(16, st) || !S->rec.alloc(16, st) || (with_qual && (!S->qual_off.alloc(16, st) || !S->qual.alloc(16, st)))) return false;
        cudaMemsetAsync(S->rec_off.p, 0, 16, st);
        if (with_qual) cudaMemsetAsync(S->qual_off.p, 0, 16, st);
        return np_wait::stream_wait(st) == cudaSuccess;
    };
    if (!any_reads) { if (!finish_empty()) return fail("cudaMalloc failed"); return S; }
    const int32_t n_int = (int32_t)anchors.size() - 1;

    // ---- record boundaries
    Dbuf d_anch, d_cnt, d_base, d_err, d_capb, d_scratch;
    std::vector<int64_t> cap_base((size_t)n_int + 1, 0);
    for (int32_t i = 0; i < n_int; i++) cap_base[(size_t)i + 1] = cap_base[(size_t)i] + (anchors[(size_t)i + 1] - anchors[(size_t)i]) / 36 + 1;
    if (!d_anch.alloc(anchors.size() * 8, st) || !d_cnt.alloc(((size_t)n_int + 2) * 4, st) || !d_base.alloc(((size_t)n_int + 2) * 4, st) || !d_err.alloc(16, st) ||
        !d_capb.alloc(cap_base.size() * 8, st) || !d_scratch.alloc(((size_t)cap_base.back() + 1) * 8, st)) return fail("cudaMalloc failed");
    cudaMemcpyAsync(d_anch.p, anchors.data(), anchors.size() * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_capb.p, cap_base.data(), cap_base.size() * 8, cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(d_err.p, 0, 16, st);
    cudaMemsetAsync(d_cnt.p, 0, ((size_t)n_int + 2) * 4, st);
    int32_t n_rec = 0, h_err = 0;
    if (n_int > 0) {
        k_rec_walk<<<(n_int + kWalkWarps - 1) / kWalkWarps, kWalkWarps * 32, 0, st>>>(U.as<uint8_t>(), total, d_anch.as<int64_t>(), d_capb.as<int64_t>(), n_int, d_cnt.as<int32_t>(),
                                                   d_scratch.as<int64_t>(), d_err.as<int32_t>());
        if (!exscan(d_cnt.as<int32_t>(), d_base.as<int32_t>(), n_int + 1, st)) return fail("cudaMalloc failed");
        cudaMemcpyAsync(&n_rec, d_base.as<int32_t>() + n_int, 4, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(&h_err, d_err.p, 4, cudaMemcpyDeviceToHost, st);
    }
    if (npz_dev::inflate_finish(job, &S->ms_inflate, err) != NP_OK) return fail(err);      // synchronises the stream
    lap("inflate + record count walk (sync)");
    if (h_err) return fail("record chain does not meet the index anchors (corrupt BAM or stale .bai)");
    if (n_rec == 0) { if (!finish_empty()) return fail("cudaMalloc failed"); return S; }
    Dbuf d_start, d_keep, d_units, d_qunits, d_slot, d_enc, d_kidx, d_uoff, d_quoff, d_sot, d_goff, d_lc, d_lcf, d_scount;
    const size_t nr1 = (size_t)n_rec + 1;
    if (!d_start.alloc(nr1 * 8, st) || !d_keep.alloc(nr1 * 4, st) || !d_units.alloc(nr1 * 4, st) || !d_qunits.alloc(nr1 * 4, st) || !d_slot.alloc(nr1 * 4, st) ||
        !d_enc.alloc(nr1, st) || !d_kidx.alloc(nr1 * 4, st) || !d_uoff.alloc(nr1 * 4, st) || !d_quoff.alloc(nr1 * 4, st) || !d_sot.alloc(slot_of_tid.size() * 4, st) ||
        !d_goff.alloc(((size_t)n_slots + 1) * 8, st) || !d_scount.alloc(((size_t)n_slots + 1) * 4, st)) return fail("cudaMalloc failed");
    k_rec_compact<<<(n_int * 32 + 127) / 128, 128, 0, st>>>(d_scratch.as<int64_t>(), d_capb.as<int64_t>(), d_cnt.as<int32_t>(), d_base.as<int32_t>(), n_int, d_start.as<int64_t>());
    cudaMemcpyAsync(d_sot.p, slot_of_tid.data(), slot_of_tid.size() * 4, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_goff.p, S->ctg_off.data(), ((size_t)n_slots + 1) * 8, cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(d_scount.p, 0, ((size_t)n_slots + 1) * 4, st);
    const int64_t G = (int64_t)ctg_bytes;
    if (with_qual == 2) {
        if (!d_lc.alloc(((size_t)G + 2) * 4, st) || !d_lcf.alloc(((size_t)G + 2) * 4, st)) return fail("cudaMalloc failed");
        k_lower_flags<<<(unsigned)((G + 1 + 255) / 256), 256, 0, st>>>(S->seq.as<uint8_t>(), G, d_lcf.as<int32_t>());
        if (!exscan(d_lcf.as<int32_t>(), d_lc.as<int32_t>(), (int)(G + 1), st)) return fail("cudaMalloc failed");
    }
    const char* e4 = getenv("NEXTPOLISH_B200_4BIT");
    MetaArgs ma{U.as<uint8_t>(), d_start.as<int64_t>(), n_rec, d_sot.as<int32_t>(), n_ref, d_goff.as<int64_t>(),
                with_qual == 2 ? d_lc.as<int32_t>() : nullptr, with_qual, (e4 && e4[0] == '1') ? 0 : 1,
                d_keep.as<int32_t>(), d_units.as<int32_t>(), d_qunits.as<int32_t>(), d_slot.as<int32_t>(), d_enc.as<uint8_t>(), d_err.as<int32_t>()};
    cudaMemsetAsync(d_keep.as<int32_t>() + n_rec, 0, 4, st); cudaMemsetAsync(d_units.as<int32_t>() + n_rec, 0, 4, st);
    cudaMemsetAsync(d_qunits.as<int32_t>() + n_rec, 0, 4, st);
    k_rec_meta<<<(n_rec + 127) / 128, 128, 0, st>>>(ma);
    if (!exscan(d_keep.as<int32_t>(), d_kidx.as<int32_t>(), n_rec + 1, st) || !exscan(d_units.as<int32_t>(), d_uoff.as<int32_t>(), n_rec + 1, st) ||
        !exscan(d_qunits.as<int32_t>(), d_quoff.as<int32_t>(), n_rec + 1, st)) return fail("cudaMalloc failed");
    int32_t tot[3] = {0, 0, 0};
    cudaMemcpyAsync(&tot[0], d_kidx.as<int32_t>() + n_rec, 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&tot[1], d_uoff.as<int32_t>() + n_rec, 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&tot[2], d_quoff.as<int32_t>() + n_rec, 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&h_err, d_err.p, 4, cudaMemcpyDeviceToHost, st);
    np_wait::stream_wait(st);
    lap("index walk + meta + scans (sync)");
    if (h_err) {
        return fail(h_err & DL_ERR_LONG ? "read longer than 65535 bases / CIGAR ops: not a short-read record"
                    : h_err & DL_ERR_ORDER ? "BAM is not coordinate sorted" : "corrupt BAM record layout");
    }
    const int32_t n_keep = tot[0];
    S->n_reads = n_keep; S->rec_bytes = (int64_t)tot[1] * 16; S->qual_bytes = (int64_t)tot[2] * 16;
    if (!S->rec_off.alloc(((size_t)n_keep + 1) * 4, st) || !S->rec.alloc((size_t)S->rec_bytes + 16, st) ||
        (with_qual && (!S->qual_off.alloc(((size_t)n_keep + 1) * 4, st) || !S->qual.alloc((size_t)S->qual_bytes + 16, st)))) return fail("cudaMalloc failed");
    cudaMemsetAsync(S->rec.p, 0, (size_t)S->rec_bytes + 16, st);
    if (with_qual) cudaMemsetAsync(S->qual.p, 0, (size_t)S->qual_bytes + 16, st);
    PackArgs pa{U.as<uint8_t>(), d_start.as<int64_t>(), n_rec, d_keep.as<int32_t>(), d_kidx.as<int32_t>(), d_uoff.as<int32_t>(), d_quoff.as<int32_t>(),
                d_qunits.as<int32_t>(), d_slot.as<int32_t>(), d_enc.as<uint8_t>(), S->rec_off.as<uint32_t>(), S->rec.as<uint8_t>(),
                with_qual ? S->qual_off.as<uint32_t>() : nullptr, with_qual ? S->qual.as<uint8_t>() : nullptr, d_scount.as<int32_t>(), n_keep, tot[1], tot[2], total};
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        const int want = (n_rec + 32 * kPackWarps - 1) / (32 * kPackWarps), cap = sms * 16;   // 32 records per warp and round
        k_rec_pack<<<std::max(1, std::min(want, cap)), kPackWarps * 32, 0, st>>>(pa);
    }
    std::vector<int32_t> counts((size_t)n_slots + 1, 0);
    cudaMemcpyAsync(counts.data(), d_scount.p, ((size_t)n_slots + 1) * 4, cudaMemcpyDeviceToHost, st);
    cudaEventRecord(e1, st);
    cudaError_t er = np_wait::stream_wait(st);
    if (er != cudaSuccess) return fail(cudaGetErrorString(er));
    lap("pack (sync)");
    if (trace) {
        cudaMemPool_t pool = npz_dev::thread_pool(device); unsigned long long used = 0, resv = 0; size_t fr = 0, tt = 0;
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &resv);
        cudaMemGetInfo(&fr, &tt);
        fprintf(stderr, "np_shard_load_gpu: pool used %.1f MB reserved %.1f MB, device free %.1f MB\n", used / 1e6, resv / 1e6, fr / 1e6);
    }
    cudaEventElapsedTime(&S->ms_total, e0, e1);
    for (int32_t k = 0; k < n_slots; k++) S->ctg_read_off[(size_t)k + 1] = S->ctg_read_off[(size_t)k] + counts[(size_t)k];
    return S;
}

extern "C" {

void np_dev_shard_view(const np_dev_shard* s, np_shard_view* v) {
    memset(v, 0, sizeof(*v));
    v->n_contigs = (int32_t)s->names.size();
    v->n_reads = s->n_reads;
    v->ctg_off = s->ctg_off.data();
    v->ctg_read_off = s->ctg_read_off.data();
    v->ctg_seq = s->seq.as<uint8_t>();
    v->rec_off = s->rec_off.as<uint32_t>();
    v->rec = s->rec.as<uint8_t>();
    v->qual_off = s->with_qual ? s->qual_off.as<uint32_t>() : nullptr;
    v->qual = s->with_qual ? s->qual.as<uint8_t>() : nullptr;
}
const char* np_dev_shard_contig_name(const np_dev_shard* s, int32_t i) { return s && i >= 0 && i < (int32_t)s->names.size() ? s->names[(size_t)i].c_str() : nullptr; }
int32_t np_dev_shard_contig_rank(const np_dev_shard* s, int32_t i) { return s && i >= 0 && i < (int32_t)s->fasta_rank.size() ? s->fasta_rank[(size_t)i] : -1; }
// sizes: {rec bytes, qual bytes, draft bytes, compressed bytes shipped, inflated bytes}; times: {inflate kernel ms, whole load ms (device)}
void np_dev_shard_stats(const np_dev_shard* s, int64_t* sizes5, float* ms2) {
    if (sizes5) { sizes5[0] = s->rec_bytes; sizes5[1] = s->qual_bytes; sizes5[2] = s->seq_bytes; sizes5[3] = s->comp_bytes; sizes5[4] = s->inflated_bytes; }
    if (ms2) { ms2[0] = s->ms_inflate; ms2[1] = s->ms_total; }
}
// Copies the device arrays to host buffers of the sizes np_dev_shard_stats / the view report (tests, debugging).
int32_t np_dev_shard_download(const np_dev_shard* s, uint8_t* ctg_seq, uint32_t* rec_off, uint8_t* rec, uint32_t* qual_off, uint8_t* qual) {
    cudaSetDevice(s->device);
    if (ctg_seq && s->seq_bytes) cudaMemcpy(ctg_seq, s->seq.p, (size_t)s->seq_bytes, cudaMemcpyDeviceToHost);
    if (rec_off) cudaMemcpy(rec_off, s->rec_off.p, ((size_t)s->n_reads + 1) * 4, cudaMemcpyDeviceToHost);
    if (rec && s->rec_bytes) cudaMemcpy(rec, s->rec.p, (size_t)s->rec_bytes, cudaMemcpyDeviceToHost);
    if (s->with_qual && qual_off) cudaMemcpy(qual_off, s->qual_off.p, ((size_t)s->n_reads + 1) * 4, cudaMemcpyDeviceToHost);
    if (s->with_qual && qual && s->qual_bytes) cudaMemcpy(qual, s->qual.p, (size_t)s->qual_bytes, cudaMemcpyDeviceToHost);
    cudaError_t er = cudaDeviceSynchronize();
    if (er != cudaSuccess) { np::set_error(cudaGetErrorString(er)); return NP_ERR_CUDA; }
    return NP_OK;
}

}  // extern "C"
