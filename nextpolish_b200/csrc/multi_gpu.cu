// multi_gpu.cu — one input polished on several GPUs of one box: contigs are independent units, so the
// contig list of ONE draft is cut into contiguous blocks of balanced cumulative length — what the reference's driver does
// for its worker jobs (blc_genome, source/nextPolish:93-117, consumed by nextpolish1.py -b/-i:148-161).  Every GPU owns a
// contiguous range of blocks and works through it with a few pipelined slots (engine + host thread each: the file reads,
// upload and inflate of one block overlap the kernels of the others; a block's reads are ONE contiguous compressed byte
// range of the BAM: devload.cu) and appends each block's polished bytes to its result buffer in HBM.  At the end the
// path's single collective gathers the result buffers on the first GPU: grouped ncclSend / ncclRecv over NVLink (exact
// byte counts: this is one process, the sizes are known on the host), then one download.
//
// Blocks are also what bounds memory: a block never exceeds the block budget (NEXTPOLISH_B200_BLOCK_MBP million draft
// bases, default 8), so genome size is limited by the result buffers (1 B per base), not by one shard's 2^31 limits.
//
// Communicators from ncclCommInitAll.  NCCL is bound at run time (dlopen of libnccl.so.2: the library that a host
// application such as PyTorch already loaded is reused, and nextpolish1.so keeps loading on boxes without NCCL, where only
// the multi-GPU form of this entry point reports an error).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "errors.h"
#include "stream_wait.h"
#include "hostio.h"
#include "../../include/nextpolish_b200.h"

namespace {
struct Nccl {
    void* h = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool load(std::string& err) {
        if (h) return true;
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define NP_SYM(name) name = (decltype(name))dlsym(h, "nccl" #name); if (!name) { err = "libnccl: missing nccl" #name; return false; }
        NP_SYM(CommInitAll) NP_SYM(CommDestroy) NP_SYM(Send) NP_SYM(Recv) NP_SYM(GroupStart) NP_SYM(GroupEnd) NP_SYM(GetErrorString)
#undef NP_SYM
        return true;
    }
};
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

// One persistent host thread per GPU: the loader keeps per-thread state (stream, private memory pool, pinned staging
// buffers), which must survive from one run to the next.
struct GpuWorker {
    std::thread th;
    std::mutex mu; std::condition_variable cv;
    std::function<void()> job; bool has_job = false, done = true, quit = false;
    void loop() {
        np_wait::use_blocking_waits(true);
        for (;;) {
            std::function<void()> j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return quit || has_job; });
                if (quit) return;
                j = std::move(job); has_job = false;
            }
            j();
            { std::lock_guard<std::mutex> lk(mu); done = true; }
            cv.notify_all();
        }
    }
    void start() { th = std::thread([this] { loop(); }); }
    void submit(std::function<void()> j) {
        { std::lock_guard<std::mutex> lk(mu); job = std::move(j); has_job = true; done = false; }
        cv.notify_all();
    }
    void wait() { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return done; }); }
    void stop() {
        { std::lock_guard<std::mutex> lk(mu); quit = true; }
        cv.notify_all();
        if (th.joinable()) th.join();
    }
};

struct np_multi {
    std::vector<int> dev;
    int slots = 3;                                     // pipelined slots per GPU
    std::vector<GpuWorker*> workers;                   // [gpu * slots + s]
    std::vector<np_engine*> eng;                       // [gpu * slots + s]
    std::vector<ncclComm_t> comm;
    Nccl nccl;
    std::vector<void*> d_res; std::vector<size_t> d_res_cap;   // per GPU: polished bytes of its blocks
    // result of the last run (FASTA order), kept until the next run
    uint8_t* h_out = nullptr; size_t h_cap = 0;       // pinned
    void* d_gather = nullptr; size_t d_cap = 0;        // on dev[0]
    std::vector<std::string> names; std::vector<const char*> name_ptrs;
    std::vector<int64_t> start, len;
};

extern "C" {

void np_multi_destroy(np_multi* m) {
    if (!m) return;
    for (GpuWorker* w : m->workers) { w->stop(); delete w; }
    for (size_t i = 0; i < m->comm.size(); i++) if (m->comm[i]) m->nccl.CommDestroy(m->comm[i]);
    for (size_t i = 0; i < m->eng.size(); i++) if (m->eng[i]) np_engine_destroy(m->eng[i]);
    for (size_t g = 0; g < m->d_res.size(); g++) if (m->d_res[g]) { cudaSetDevice(m->dev[g]); cudaFree(m->d_res[g]); }
    if (!m->dev.empty()) cudaSetDevice(m->dev[0]);
    if (m->d_gather) cudaFree(m->d_gather);
    if (m->h_out) cudaFreeHost(m->h_out);
    delete m;
}

np_multi* np_multi_create(const int32_t* devices, int32_t n_devices) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { np::set_error("np_multi_create: no usable CUDA device; this engine has no CPU path"); return nullptr; }
    if (n_devices < 1 || n_devices > ndev) { np::set_error("np_multi_create: bad device count"); return nullptr; }
    np_multi* m = new np_multi();
    for (int i = 0; i < n_devices; i++) m->dev.push_back(devices ? devices[i] : i);
    if (const char* ev = getenv("NEXTPOLISH_B200_SLOTS")) { const int v = atoi(ev); if (v >= 1 && v <= 4) m->slots = v; }
    std::string err;
    if (n_devices > 1) {
        setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);          // NCCL logs to stdout by default: the CLI's stdout is the FASTA
        if (!m->nccl.load(err)) { np::set_error("np_multi_create: " + err); delete m; return nullptr; }
        m->comm.assign((size_t)n_devices, nullptr);
        ncclResult_t rc = m->nccl.CommInitAll(m->comm.data(), n_devices, m->dev.data());
        if (rc != ncclSuccess) { np::set_error(std::string("ncclCommInitAll: ") + m->nccl.GetErrorString(rc)); m->comm.clear(); np_multi_destroy(m); return nullptr; }
    }
    m->d_res.assign((size_t)n_devices, nullptr); m->d_res_cap.assign((size_t)n_devices, 0);
    for (int i = 0; i < n_devices * m->slots; i++) {
        np_engine* e = np_engine_create(m->dev[(size_t)(i / m->slots)]);
        if (!e) { np_multi_destroy(m); return nullptr; }
        m->eng.push_back(e);
    }
    for (int i = 0; i < n_devices * m->slots; i++) { GpuWorker* w = new GpuWorker(); m->workers.push_back(w); w->start(); }
    return m;
}

// Contiguous blocks of the contig list (BAM reference order) with balanced cumulative length: block b ends at the first
// contig whose cumulative length reaches (b + 1) / n of the total.  part[i] = block of contig i (order of `lengths`).
void np_partition_contiguous(const int64_t* lengths, int32_t n_contigs, int32_t n_parts, int32_t* part) {
    int64_t total = 0;
    for (int32_t i = 0; i < n_contigs; i++) total += lengths[i];
    int64_t acc = 0; int32_t b = 0;
    for (int32_t i = 0; i < n_contigs; i++) {
        part[i] = b;
        acc += lengths[i];
        while (b < n_parts - 1 && acc * n_parts >= (int64_t)(b + 1) * total) b++;
    }
}

int32_t np_multi_run(np_multi* m, int32_t task, const char* fasta, const char* bam, const Configure* cfg, np_files_result* out) {
    return np_multi_run_names(m, task, fasta, bam, cfg, nullptr, -1, out);
}

// The same run restricted to the contigs `names` (a worker's block: nextpolish1.py -b/-i, :148-161); n_names < 0: every
// contig of the draft.  Names the draft does not hold are an error (the reference dereferences NULL there).  The result
// lists the selected contigs in FASTA order.
int32_t np_multi_run_names(np_multi* m, int32_t task, const char* fasta, const char* bam, const Configure* cfg,
                           const char* const* names, int32_t n_names, np_files_result* out) {
    if (!m || !fasta || !bam || !cfg || !out || (n_names > 0 && !names)) { np::set_error("np_multi_run: bad arguments"); return NP_ERR_ARG; }
    const int n = (int)m->dev.size(), K = m->slots;
    std::string err;
    const double t0 = now_ms();
    // the draft is read and parsed ONCE (one pass over an mmap of the file) and shared by every loader
    std::vector<std::string> fa_names; std::vector<int64_t> fa_off, fa_len;
    std::vector<uint8_t> draft;
    auto grow = [](void* ctx, size_t bytes) -> uint8_t* { auto& v = *(std::vector<uint8_t>*)ctx; v.resize(bytes); return v.data(); };
    if (!np::fasta_load_flat(fasta, fa_names, fa_off, grow, &draft, err)) { np::set_error("np_multi_run: " + err); return NP_ERR_IO; }
    for (size_t i = 0; i + 1 < fa_off.size(); i++) fa_len.push_back(fa_off[i + 1] - fa_off[i]);
    // contigs in BAM reference order (those the BAM does not know last), as the loaders order them
    np::BamFile bf;
    if (!bf.open(bam, err)) { np::set_error("np_multi_run: " + err); return NP_ERR_IO; }
    std::unordered_map<std::string, int> tid_of;
    for (size_t i = 0; i < bf.header().names.size(); i++) tid_of.emplace(bf.header().names[i], (int)i);
    // sel: FASTA indices of the contigs to polish, in FASTA order; sel_pos: their position in the result
    std::vector<int32_t> sel, sel_pos(fa_names.size(), -1);
    if (n_names < 0) { for (size_t i = 0; i < fa_names.size(); i++) sel.push_back((int32_t)i); }
    else {
        std::unordered_map<std::string, int32_t> fa_idx;
        for (size_t i = 0; i < fa_names.size(); i++) fa_idx.emplace(fa_names[i], (int32_t)i);
        std::vector<uint8_t> want(fa_names.size(), 0);
        for (int32_t k = 0; k < n_names; k++) {
            auto it = fa_idx.find(names[k] ? names[k] : "");
            if (it == fa_idx.end()) { np::set_error(std::string("np_multi_run: contig not in the draft: ") + (names[k] ? names[k] : "(null)")); return NP_ERR_ARG; }
            want[(size_t)it->second] = 1;
        }
        for (size_t i = 0; i < fa_names.size(); i++) if (want[i]) sel.push_back((int32_t)i);
    }
    for (size_t k = 0; k < sel.size(); k++) sel_pos[(size_t)sel[k]] = (int32_t)k;
    const int32_t nc = (int32_t)sel.size();
    std::vector<int32_t> order(sel);
    auto tid = [&](int32_t i) { auto it = tid_of.find(fa_names[(size_t)i]); return it == tid_of.end() ? 0x7fffffff : it->second; };
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return tid(a) < tid(b); });
    std::vector<int64_t> lens((size_t)nc);
    int64_t total_len = 0;
    for (int32_t k = 0; k < nc; k++) { lens[(size_t)k] = fa_len[(size_t)order[(size_t)k]]; total_len += lens[(size_t)k]; }
    // blocks: at least one per slot when the contig list allows it, none above the block budget (a single contig larger
    // than the budget is a block of its own)
    double block_mbp = 8.0;
    if (const char* ev = getenv("NEXTPOLISH_B200_BLOCK_MBP")) { const double v = atof(ev); if (v > 0) block_mbp = v; }
    if (const char* ev = getenv("NEXTPOLISH_B200_SHARD_MBP")) { const double v = atof(ev); if (v > 0 && v < block_mbp) block_mbp = v; }
    int64_t nb64 = (int64_t)((double)total_len / (block_mbp * 1e6)) + 1;
    if (nb64 < (int64_t)n * K) nb64 = (int64_t)n * K;
    if (nb64 > nc) nb64 = nc > 0 ? nc : 1;
    const int32_t NB = (int32_t)nb64;
    std::vector<int32_t> part((size_t)nc, 0);
    np_partition_contiguous(lens.data(), nc, NB, part.data());
    struct Block { std::vector<const char*> names; std::vector<int32_t> rank; int64_t bases = 0; int gpu = 0;
                   int64_t off = 0, bytes = 0; std::vector<int64_t> ctg_off; std::vector<int32_t> slot_rank; std::vector<std::string> slot_name; };
    std::vector<Block> blocks((size_t)NB);
    for (int32_t k = 0; k < nc; k++) {
        Block& b = blocks[(size_t)part[(size_t)k]];
        b.names.push_back(fa_names[(size_t)order[(size_t)k]].c_str()); b.rank.push_back(order[(size_t)k]); b.bases += lens[(size_t)k];
    }
    // GPU g owns the blocks [first[g], first[g+1]): contiguous ranges with balanced bases
    std::vector<int64_t> bl((size_t)NB);
    for (int32_t b = 0; b < NB; b++) bl[(size_t)b] = blocks[(size_t)b].bases;
    std::vector<int32_t> gpu_of((size_t)NB, 0);
    np_partition_contiguous(bl.data(), NB, n, gpu_of.data());
    std::vector<int32_t> first((size_t)n + 1, NB);
    for (int32_t b = NB - 1; b >= 0; b--) { blocks[(size_t)b].gpu = gpu_of[(size_t)b]; first[(size_t)gpu_of[(size_t)b]] = b; }
    for (int g = n - 1; g >= 0; g--) if (first[(size_t)g] == NB && g + 1 <= n) first[(size_t)g] = first[(size_t)g + 1];
    first[(size_t)n] = NB;
    // per-GPU result buffers (polished length ~ draft length; 1.25x + slack)
    for (int g = 0; g < n; g++) {
        int64_t bases = 0;
        for (int32_t b = first[(size_t)g]; b < first[(size_t)g + 1]; b++) bases += blocks[(size_t)b].bases;
        const size_t want = (size_t)bases + (size_t)bases / 4 + (size_t)4096 * (size_t)(first[(size_t)g + 1] - first[(size_t)g] + 1);
        if (want > m->d_res_cap[(size_t)g]) {
            cudaSetDevice(m->dev[(size_t)g]);
            if (m->d_res[(size_t)g]) cudaFree(m->d_res[(size_t)g]);
            m->d_res[(size_t)g] = nullptr; m->d_res_cap[(size_t)g] = 0;
            if (cudaMalloc(&m->d_res[(size_t)g], want + want / 8) != cudaSuccess) { np::set_error("np_multi_run: cudaMalloc failed"); return NP_ERR_CUDA; }
            m->d_res_cap[(size_t)g] = want + want / 8;
        }
    }

    // ---- every slot pulls blocks of its GPU: load, polish, append the polished bytes to the GPU's result buffer
    const int wq = task == NP_TASK_KMER_COUNT ? 2 : task == NP_TASK_SNP_VALID ? 1 : 0;
    std::vector<std::mutex> gmu((size_t)n);
    std::vector<int32_t> next((size_t)n); std::vector<int64_t> used((size_t)n, 0), h2d_of((size_t)n, 0);
    for (int g = 0; g < n; g++) next[(size_t)g] = first[(size_t)g];
    std::mutex emu; int32_t first_rc = NP_OK; std::string first_msg;
    auto slot_work = [&](int si) {
        const int g = si / K;
        np_engine* e = m->eng[(size_t)si];
        cudaSetDevice(m->dev[(size_t)g]);
        for (;;) {
            int32_t b;
            { std::lock_guard<std::mutex> lk(gmu[(size_t)g]); if (first_rc != NP_OK || next[(size_t)g] >= first[(size_t)g + 1]) return; b = next[(size_t)g]++; }
            Block& B = blocks[(size_t)b];
            if (B.names.empty()) continue;
            std::vector<const uint8_t*> seqs; std::vector<int64_t> ls;
            for (int32_t fr : B.rank) { seqs.push_back(draft.data() + fa_off[(size_t)fr]); ls.push_back(fa_off[(size_t)fr + 1] - fa_off[(size_t)fr]); }
            np_dev_shard* ds = np_shard_load_gpu_seqs(m->dev[(size_t)g], bam, B.names.data(), seqs.data(), ls.data(), (int32_t)seqs.size(), wq);
            int32_t r = ds ? NP_OK : NP_ERR_IO;
            np_shard_view v;
            if (ds) { np_dev_shard_view(ds, &v); r = np_engine_adopt_device(e, &v); }
            if (r == NP_OK) r = np_engine_run(e, task, cfg);
            if (r == NP_OK) {
                B.bytes = np_engine_result_bytes(e);
                B.ctg_off.assign((size_t)v.n_contigs + 1, 0);
                r = np_engine_result_offsets(e, B.ctg_off.data());
            }
            if (r == NP_OK) {
                int64_t sizes[5];
                np_dev_shard_stats(ds, sizes, nullptr);
                { std::lock_guard<std::mutex> lk(gmu[(size_t)g]); B.off = used[(size_t)g]; used[(size_t)g] += B.bytes; h2d_of[(size_t)g] += sizes[3] + sizes[2]; }
                if ((size_t)(B.off + B.bytes) > m->d_res_cap[(size_t)g]) { r = NP_ERR_LIMIT; np::set_error("np_multi_run: result buffer too small"); }
                else r = np_engine_copy_result(e, (uint8_t*)m->d_res[(size_t)g] + B.off, B.bytes);
                if (r == NP_OK) r = np_engine_sync(e);
                for (int32_t i = 0; i < v.n_contigs; i++) { B.slot_rank.push_back(np_dev_shard_contig_rank(ds, i)); B.slot_name.push_back(np_dev_shard_contig_name(ds, i)); }
            }
            if (r != NP_OK) { std::lock_guard<std::mutex> lk(emu); if (first_rc == NP_OK) { first_rc = r; first_msg = "GPU " + std::to_string(m->dev[(size_t)g]) + ": " + np_last_error(); } }
            if (ds) np_dev_shard_free(ds);
            if (r != NP_OK) return;
        }
    };
    for (int si = 0; si < n * K; si++) m->workers[(size_t)si]->submit([&slot_work, si] { slot_work(si); });
    for (int si = 0; si < n * K; si++) m->workers[(size_t)si]->wait();
    if (first_rc != NP_OK) { np::set_error("np_multi_run (" + first_msg + ")"); return first_rc; }
    const double t1 = now_ms();

    // ---- the single collective: result buffers of every GPU -> the first GPU (exact sizes), then one download
    std::vector<int64_t> goff((size_t)n + 1, 0);
    int64_t h2d = 0;
    for (int g = 0; g < n; g++) { goff[(size_t)g + 1] = goff[(size_t)g] + used[(size_t)g]; h2d += h2d_of[(size_t)g]; }
    const int64_t total = goff[(size_t)n];
    cudaSetDevice(m->dev[0]);
    if (n > 1 && (size_t)total + 16 > m->d_cap) {
        if (m->d_gather) cudaFree(m->d_gather);
        m->d_gather = nullptr; m->d_cap = 0;
        const size_t want = (size_t)total + (size_t)total / 8 + 4096;
        if (cudaMalloc(&m->d_gather, want) != cudaSuccess) { np::set_error("np_multi_run: cudaMalloc failed"); return NP_ERR_CUDA; }
        m->d_cap = want;
    }
    if ((size_t)total + 16 > m->h_cap) {
        if (m->h_out) cudaFreeHost(m->h_out);
        m->h_out = nullptr; m->h_cap = 0;
        const size_t want = (size_t)total + (size_t)total / 8 + 4096;
        if (cudaMallocHost((void**)&m->h_out, want) != cudaSuccess) { np::set_error("np_multi_run: cudaMallocHost failed"); return NP_ERR_CUDA; }
        m->h_cap = want;
    }
    std::vector<ncclResult_t> nrc((size_t)n, ncclSuccess);
    auto gather = [&](int g) {
        cudaSetDevice(m->dev[(size_t)g]);
        cudaStream_t s = (cudaStream_t)np_engine_stream(m->eng[(size_t)g * (size_t)K]);
        if (g == 0) {
            const uint8_t* src = (const uint8_t*)m->d_res[0];
            if (n > 1) {
                if (used[0]) cudaMemcpyAsync(m->d_gather, m->d_res[0], (size_t)used[0], cudaMemcpyDeviceToDevice, s);
                m->nccl.GroupStart();
                for (int p = 1; p < n; p++)
                    if (used[(size_t)p]) { ncclResult_t r = m->nccl.Recv((uint8_t*)m->d_gather + goff[(size_t)p], (size_t)used[(size_t)p], ncclUint8, p, m->comm[0], s); if (r != ncclSuccess) nrc[0] = r; }
                ncclResult_t r = m->nccl.GroupEnd(); if (r != ncclSuccess) nrc[0] = r;
                src = (const uint8_t*)m->d_gather;
            }
            if (total) cudaMemcpyAsync(m->h_out, src, (size_t)total, cudaMemcpyDeviceToHost, s);
        } else if (used[(size_t)g]) {
            ncclResult_t r = m->nccl.Send(m->d_res[(size_t)g], (size_t)used[(size_t)g], ncclUint8, 0, m->comm[(size_t)g], s);
            if (r != ncclSuccess) nrc[(size_t)g] = r;
        }
        if (np_wait::stream_wait(s) != cudaSuccess) nrc[(size_t)g] = ncclUnhandledCudaError;
    };
    for (int g = 0; g < n; g++) m->workers[(size_t)g * (size_t)K]->submit([&gather, g] { gather(g); });
    for (int g = 0; g < n; g++) m->workers[(size_t)g * (size_t)K]->wait();
    for (int g = 0; g < n; g++) if (nrc[(size_t)g] != ncclSuccess) { np::set_error(std::string("np_multi_run: gather failed: ") + (n > 1 ? m->nccl.GetErrorString(nrc[(size_t)g]) : "CUDA error")); return NP_ERR_CUDA; }
    const double t2 = now_ms();

    // ---- results in FASTA order
    m->names.assign((size_t)nc, std::string()); m->start.assign((size_t)nc, 0); m->len.assign((size_t)nc, 0);
    for (const Block& B : blocks)
        for (size_t i = 0; i < B.slot_rank.size(); i++) {
            // the loader reports, for every slot of the shard, its rank inside the name list it was given
            const int32_t fr = sel_pos[(size_t)B.rank[(size_t)B.slot_rank[i]]];
            m->names[(size_t)fr] = B.slot_name[i];
            m->start[(size_t)fr] = goff[(size_t)B.gpu] + B.off + B.ctg_off[i];
            m->len[(size_t)fr] = B.ctg_off[i + 1] - B.ctg_off[i];
        }
    m->name_ptrs.clear();
    for (auto& nm : m->names) m->name_ptrs.push_back(nm.c_str());
    out->task = task; out->n_contigs = nc; out->names = m->name_ptrs.data(); out->seq = m->h_out;
    out->start = m->start.data(); out->len = m->len.data();
    out->h2d_bytes = h2d; out->d2h_bytes = total;
    out->load_ms = (float)(t1 - t0); out->polish_ms = (float)(t2 - t1);      // blocks (parse, load, polish) / gather + download
    (void)NB;
    return NP_OK;
}

}  // extern "C"
