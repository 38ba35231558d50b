// multi_gpu.cu — one input polished on several GPUs of one box: contigs are independent units, so the
// contig list of ONE draft is cut into one contiguous block per GPU by cumulative length — what the reference's driver
// does for its worker jobs (blc_genome, source/nextPolish:93-117, consumed by nextpolish1.py -b/-i:148-161) — every GPU
// builds and polishes its own shard (its slice of the BAM is one contiguous compressed byte range: devload.cu), and the
// polished bytes are gathered on the first GPU with the path's single collective: grouped ncclSend / ncclRecv over
// NVLink (exact byte counts: this is one process, the sizes are known on the host), then one download.
//
// One host thread per GPU; communicators from ncclCommInitAll.  NCCL is bound at run time (dlopen of libnccl.so.2:
// the library that a host application such as PyTorch already loaded is reused, and nextpolish1.so keeps loading on
// boxes without NCCL, where only this entry point reports an error).
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
    std::vector<GpuWorker*> workers;
    std::vector<np_engine*> eng;
    std::vector<ncclComm_t> comm;
    Nccl nccl;
    // result of the last run (FASTA order), kept until the next run
    uint8_t* h_out = nullptr; size_t h_cap = 0;       // pinned
    void* d_gather = nullptr; size_t d_cap = 0;        // on dev[0]
    std::vector<std::string> names; std::vector<const char*> name_ptrs;
    std::vector<int64_t> start, len;
    float ms_gather = 0;
};

extern "C" {

// runs f(g) for every GPU on its own persistent thread and waits for all of them
static void on_all_gpus(np_multi* m, const std::function<void(int)>& f) {
    const int n = (int)m->workers.size();
    for (int g = 0; g < n; g++) m->workers[(size_t)g]->submit([&f, g] { f(g); });
    for (int g = 0; g < n; g++) m->workers[(size_t)g]->wait();
}

void np_multi_destroy(np_multi* m) {
    if (!m) return;
    for (GpuWorker* w : m->workers) { w->stop(); delete w; }
    for (size_t i = 0; i < m->comm.size(); i++) if (m->comm[i]) m->nccl.CommDestroy(m->comm[i]);
    for (size_t i = 0; i < m->eng.size(); i++) if (m->eng[i]) np_engine_destroy(m->eng[i]);
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
    std::string err;
    if (n_devices > 1) {
        setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);          // NCCL logs to stdout by default: the CLI's stdout is the FASTA
        if (!m->nccl.load(err)) { np::set_error("np_multi_create: " + err); delete m; return nullptr; }
        m->comm.assign((size_t)n_devices, nullptr);
        ncclResult_t rc = m->nccl.CommInitAll(m->comm.data(), n_devices, m->dev.data());
        if (rc != ncclSuccess) { np::set_error(std::string("ncclCommInitAll: ") + m->nccl.GetErrorString(rc)); m->comm.clear(); np_multi_destroy(m); return nullptr; }
    }
    for (int i = 0; i < n_devices; i++) {
        np_engine* e = np_engine_create(m->dev[(size_t)i]);
        if (!e) { np_multi_destroy(m); return nullptr; }
        m->eng.push_back(e);
    }
    for (int i = 0; i < n_devices; i++) { GpuWorker* w = new GpuWorker(); m->workers.push_back(w); w->start(); }
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

// One round: GPU g loads and polishes block names_of[g] (empty: idle); the polished bytes of all GPUs are gathered on the
// first GPU and downloaded to m->h_out + h_base.  Fills names / start / len (FASTA ranks) of the round's contigs.
static int32_t multi_round(np_multi* m, int32_t task, const char* bam, const Configure* cfg,
                           const std::vector<std::vector<const char*>>& names_of, const std::vector<std::vector<int32_t>>& rank_of,
                           const uint8_t* flat, const std::vector<int64_t>& fa_off,
                           int64_t h_base, int64_t* h_used, int64_t* h2d_total) {
    const int n = (int)m->dev.size();
    const int wq = task == NP_TASK_KMER_COUNT ? 2 : task == NP_TASK_SNP_VALID ? 1 : 0;
    std::vector<int32_t> rc((size_t)n, NP_OK); std::vector<std::string> msg((size_t)n);
    std::vector<np_dev_shard*> ds((size_t)n, nullptr);
    std::vector<int64_t> nbytes((size_t)n, 0), h2d_of((size_t)n, 0);
    std::vector<std::vector<int64_t>> off((size_t)n);
    auto work = [&](int g) {
        cudaSetDevice(m->dev[(size_t)g]);
        if (names_of[(size_t)g].empty()) return;
        std::vector<const uint8_t*> seqs; std::vector<int64_t> lens;      // the block's contigs inside the draft parsed once by np_multi_run
        for (int32_t fr : rank_of[(size_t)g]) { seqs.push_back(flat + fa_off[(size_t)fr]); lens.push_back(fa_off[(size_t)fr + 1] - fa_off[(size_t)fr]); }
        ds[(size_t)g] = np_shard_load_gpu_seqs(m->dev[(size_t)g], bam, names_of[(size_t)g].data(), seqs.data(), lens.data(), (int32_t)seqs.size(), wq);
        if (!ds[(size_t)g]) { rc[(size_t)g] = NP_ERR_IO; msg[(size_t)g] = np_last_error(); return; }
        np_shard_view v;
        np_dev_shard_view(ds[(size_t)g], &v);
        int32_t r = np_engine_adopt_device(m->eng[(size_t)g], &v);
        if (r == NP_OK) r = np_engine_run(m->eng[(size_t)g], task, cfg);
        if (r == NP_OK) {
            nbytes[(size_t)g] = np_engine_result_bytes(m->eng[(size_t)g]);
            off[(size_t)g].assign((size_t)v.n_contigs + 1, 0);
            r = np_engine_result_offsets(m->eng[(size_t)g], off[(size_t)g].data());
            int64_t sizes[5];
            np_dev_shard_stats(ds[(size_t)g], sizes, nullptr);
            h2d_of[(size_t)g] = sizes[3] + sizes[2];
        }
        if (r != NP_OK) { rc[(size_t)g] = r; msg[(size_t)g] = np_last_error(); }
    };
    on_all_gpus(m, work);
    auto cleanup = [&]() { for (int g = 0; g < n; g++) if (ds[(size_t)g]) { cudaSetDevice(m->dev[(size_t)g]); np_dev_shard_free(ds[(size_t)g]); } };
    for (int g = 0; g < n; g++) if (rc[(size_t)g] != NP_OK) { np::set_error("np_multi_run (GPU " + std::to_string(m->dev[(size_t)g]) + "): " + msg[(size_t)g]); cleanup(); return rc[(size_t)g]; }

    // ---- the single collective: polished bytes of every GPU -> the first GPU (exact sizes), then one download
    std::vector<int64_t> goff((size_t)n + 1, 0);
    for (int g = 0; g < n; g++) { goff[(size_t)g + 1] = goff[(size_t)g] + nbytes[(size_t)g]; *h2d_total += h2d_of[(size_t)g]; }
    const int64_t total = goff[(size_t)n];
    cudaSetDevice(m->dev[0]);
    if ((size_t)total + 16 > m->d_cap) {
        if (m->d_gather) cudaFree(m->d_gather);
        m->d_gather = nullptr; m->d_cap = 0;
        const size_t want = (size_t)total + (size_t)total / 8 + 4096;
        if (cudaMalloc(&m->d_gather, want) != cudaSuccess) { np::set_error("np_multi_run: cudaMalloc failed"); cleanup(); return NP_ERR_CUDA; }
        m->d_cap = want;
    }
    if ((size_t)(h_base + total) + 16 > m->h_cap) {                    // grow the pinned result buffer, keeping earlier rounds
        const size_t want = (size_t)(h_base + total) + (size_t)(h_base + total) / 4 + 4096;
        uint8_t* nb = nullptr;
        if (cudaMallocHost((void**)&nb, want) != cudaSuccess) { np::set_error("np_multi_run: cudaMallocHost failed"); cleanup(); return NP_ERR_CUDA; }
        if (m->h_out) { if (h_base) memcpy(nb, m->h_out, (size_t)h_base); cudaFreeHost(m->h_out); }
        m->h_out = nb; m->h_cap = want;
    }
    std::vector<ncclResult_t> nrc((size_t)n, ncclSuccess);
    auto gather = [&](int g) {
        cudaSetDevice(m->dev[(size_t)g]);
        cudaStream_t s = (cudaStream_t)np_engine_stream(m->eng[(size_t)g]);
        if (g == 0) {
            if (nbytes[0]) cudaMemcpyAsync(m->d_gather, np_engine_result_device(m->eng[0]), (size_t)nbytes[0], cudaMemcpyDeviceToDevice, s);
            if (n > 1) {
                m->nccl.GroupStart();
                for (int p = 1; p < n; p++)
                    if (nbytes[(size_t)p]) { ncclResult_t r = m->nccl.Recv((uint8_t*)m->d_gather + goff[(size_t)p], (size_t)nbytes[(size_t)p], ncclUint8, p, m->comm[0], s); if (r != ncclSuccess) nrc[0] = r; }
                ncclResult_t r = m->nccl.GroupEnd(); if (r != ncclSuccess) nrc[0] = r;
            }
            if (total) cudaMemcpyAsync(m->h_out + h_base, m->d_gather, (size_t)total, cudaMemcpyDeviceToHost, s);
        } else if (nbytes[(size_t)g]) {
            ncclResult_t r = m->nccl.Send(np_engine_result_device(m->eng[(size_t)g]), (size_t)nbytes[(size_t)g], ncclUint8, 0, m->comm[(size_t)g], s);
            if (r != ncclSuccess) nrc[(size_t)g] = r;
        }
        if (cudaStreamSynchronize(s) != cudaSuccess) nrc[(size_t)g] = ncclUnhandledCudaError;
    };
    on_all_gpus(m, gather);
    for (int g = 0; g < n; g++) if (nrc[(size_t)g] != ncclSuccess) { np::set_error(std::string("np_multi_run: gather failed: ") + (n > 1 ? m->nccl.GetErrorString(nrc[(size_t)g]) : "CUDA error")); cleanup(); return NP_ERR_CUDA; }
    for (int g = 0; g < n; g++) {
        if (!ds[(size_t)g]) continue;
        const int32_t ncg = (int32_t)names_of[(size_t)g].size();
        for (int32_t i = 0; i < ncg; i++) {
            // the loader reports, for every slot of the shard, its rank inside the name list it was given
            const int32_t fr = rank_of[(size_t)g][(size_t)np_dev_shard_contig_rank(ds[(size_t)g], i)];
            m->names[(size_t)fr] = np_dev_shard_contig_name(ds[(size_t)g], i);
            m->start[(size_t)fr] = h_base + goff[(size_t)g] + off[(size_t)g][(size_t)i];
            m->len[(size_t)fr] = off[(size_t)g][(size_t)i + 1] - off[(size_t)g][(size_t)i];
        }
    }
    cleanup();
    *h_used = total;
    return NP_OK;
}

// Blocks: n_gpus x rounds contiguous blocks of the contig list (BAM reference order) with balanced cumulative length;
// rounds = what keeps a block under the shard budget (NEXTPOLISH_B200_SHARD_MBP million draft bases, default 256: the
// inflated BAM slice, the packed records and the column arrays of a block stay far below the 2^31 limits of a shard and
// within HBM at any depth a short-read run uses).  Block b runs in round b / n_gpus on GPU b % n_gpus.
int32_t np_multi_run(np_multi* m, int32_t task, const char* fasta, const char* bam, const Configure* cfg, np_files_result* out) {
    if (!m || !fasta || !bam || !cfg || !out) { np::set_error("np_multi_run: bad arguments"); return NP_ERR_ARG; }
    const int n = (int)m->dev.size();
    std::string err;
    // contigs in BAM reference order (those the BAM does not know last), as the loaders order them
    // the draft is read and parsed ONCE (one pass over an mmap of the file) and shared by every GPU's loader
    std::vector<std::string> fa_names; std::vector<int64_t> fa_off, fa_len;
    std::vector<uint8_t> draft;
    auto grow = [](void* ctx, size_t bytes) -> uint8_t* { auto& v = *(std::vector<uint8_t>*)ctx; v.resize(bytes); return v.data(); };
    if (!np::fasta_load_flat(fasta, fa_names, fa_off, grow, &draft, err)) { np::set_error("np_multi_run: " + err); return NP_ERR_IO; }
    for (size_t i = 0; i + 1 < fa_off.size(); i++) fa_len.push_back(fa_off[i + 1] - fa_off[i]);
    np::BamFile bf;
    if (!bf.open(bam, err)) { np::set_error("np_multi_run: " + err); return NP_ERR_IO; }
    std::unordered_map<std::string, int> tid_of;
    for (size_t i = 0; i < bf.header().names.size(); i++) tid_of.emplace(bf.header().names[i], (int)i);
    const int32_t nc = (int32_t)fa_names.size();
    std::vector<int32_t> order((size_t)nc);
    for (int32_t i = 0; i < nc; i++) order[(size_t)i] = i;
    auto tid = [&](int32_t i) { auto it = tid_of.find(fa_names[(size_t)i]); return it == tid_of.end() ? 0x7fffffff : it->second; };
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return tid(a) < tid(b); });
    std::vector<int64_t> lens((size_t)nc);
    int64_t total_len = 0;
    for (int32_t k = 0; k < nc; k++) { lens[(size_t)k] = fa_len[(size_t)order[(size_t)k]]; total_len += lens[(size_t)k]; }
    double budget_mbp = 256.0;
    if (const char* ev = getenv("NEXTPOLISH_B200_SHARD_MBP")) { const double v = atof(ev); if (v > 0) budget_mbp = v; }
    int64_t rounds = (int64_t)((double)total_len / (budget_mbp * 1e6 * n)) + 1;
    if (rounds > nc) rounds = nc > 0 ? nc : 1;
    const int32_t n_blocks = (int32_t)(rounds * n);
    std::vector<int32_t> part((size_t)nc, 0);
    np_partition_contiguous(lens.data(), nc, n_blocks, part.data());

    m->names.assign((size_t)nc, std::string()); m->start.assign((size_t)nc, 0); m->len.assign((size_t)nc, 0);
    int64_t h_base = 0, h2d = 0;
    const double t0 = now_ms();
    for (int64_t r = 0; r < rounds; r++) {
        std::vector<std::vector<const char*>> names_of((size_t)n);
        std::vector<std::vector<int32_t>> rank_of((size_t)n);            // FASTA rank of every contig of the block
        for (int32_t k = 0; k < nc; k++) {
            const int32_t b = part[(size_t)k];
            if (b / n != r) continue;
            names_of[(size_t)(b % n)].push_back(fa_names[(size_t)order[(size_t)k]].c_str());
            rank_of[(size_t)(b % n)].push_back(order[(size_t)k]);
        }
        int64_t used = 0;
        const int32_t rc = multi_round(m, task, bam, cfg, names_of, rank_of, draft.data(), fa_off, h_base, &used, &h2d);
        if (rc != NP_OK) return rc;
        h_base += used;
    }
    const double t1 = now_ms();
    m->name_ptrs.clear();
    for (auto& nm : m->names) m->name_ptrs.push_back(nm.c_str());
    out->task = task; out->n_contigs = nc; out->names = m->name_ptrs.data(); out->seq = m->h_out;
    out->start = m->start.data(); out->len = m->len.data();
    out->h2d_bytes = h2d; out->d2h_bytes = h_base;
    out->load_ms = (float)(t1 - t0); out->polish_ms = (float)rounds;      // wall clock of all rounds; number of rounds
    return NP_OK;
}

}  // extern "C"
