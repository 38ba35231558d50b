// files_pipeline.cu — the from-files front end: draft FASTA + coordinate-sorted BGZF BAM (+ .bai) -> polished bytes
// on the host, pipelined.  This is the batch form of what the reference ABI does one contig at a time
// (score_chain / kmer_count: contig_init + two BAM passes + contig_get_contig, scorechain.c:3-15, kmercount.c:93-126)
// and what the reference's main.c does for a whole FASTA (main.c:12-26).
//
// A pipeline owns `depth` workers (np_engine + host thread each) and 2 x depth job records.  np_files_submit queues a job
// and returns; the next free worker takes the oldest queued job, loads the shard on the GPU (np_shard_load_gpu: compressed
// bytes host -> device, BGZF inflate, record unpack, packing), runs the task and copies the polished bytes to the record's
// pinned host buffer.  Host work and upload of one job overlap the kernels of the others; because workers pull from a
// queue (rather than owning every depth-th ticket) no worker idles while the caller is still waiting for an older job.
#include <cuda_runtime.h>

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "errors.h"
#include "stream_wait.h"
#include "../../include/nextpolish_b200.h"

namespace {
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Slot {            // one job record: request, result, the pinned buffer the result lands in
    enum { IDLE, QUEUED, RUNNING, DONE } state = IDLE;
    // job
    int64_t ticket = -1; int32_t task = 0; std::string fasta, bam; Configure cfg; int32_t with_host_load = 0;
    // result
    int32_t rc = NP_OK; std::string err;
    uint8_t* out = nullptr; size_t out_cap = 0;         // pinned
    std::vector<int64_t> off, start, len;
    std::vector<std::string> names; std::vector<const char*> name_ptrs;
    int64_t h2d = 0, d2h = 0; float load_ms = 0, polish_ms = 0;
};
struct Worker { np_engine* eng = nullptr; std::thread th; };
}  // namespace

struct np_files {
    int device = 0;
    std::vector<Slot*> slots;          // job records, ticket % slots.size()
    std::vector<Worker*> workers;
    std::mutex mu; std::condition_variable cv;
    int64_t next_ticket = 0, next_run = 0;
    bool quit = false;
};

static void run_job(np_files* P, Slot& s, np_engine* eng) {
    cudaSetDevice(P->device);
    s.rc = NP_OK; s.err.clear();
    const double t0 = now_ms();
    const int wq = s.task == NP_TASK_KMER_COUNT ? 2 : s.task == NP_TASK_SNP_VALID ? 1 : 0;
    np_dev_shard* ds = np_shard_load_gpu(P->device, s.fasta.c_str(), s.bam.c_str(), nullptr, 0, wq);
    if (!ds) { s.rc = NP_ERR_IO; s.err = np_last_error(); return; }
    np_shard_view v;
    np_dev_shard_view(ds, &v);
    const double t1 = now_ms();
    int32_t rc = np_engine_adopt_device(eng, &v);
    if (rc == NP_OK) rc = np_engine_run(eng, s.task, &s.cfg);
    if (rc == NP_OK) {
        const int64_t n = np_engine_result_bytes(eng);
        if ((size_t)n + 1 > s.out_cap) {
            if (s.out) cudaFreeHost(s.out);
            s.out = nullptr; s.out_cap = 0;
            size_t want = (size_t)n + (size_t)n / 8 + 4096;
            if (cudaMallocHost((void**)&s.out, want) != cudaSuccess) { rc = NP_ERR_CUDA; np::set_error("np_files: cudaMallocHost failed"); }
            else s.out_cap = want;
        }
        s.off.assign((size_t)v.n_contigs + 1, 0);
        if (rc == NP_OK) rc = np_engine_download(eng, s.out, (int64_t)s.out_cap, s.off.data());
        if (rc == NP_OK) {
            // results in FASTA order, like contig_write_to_file's loop over the .fai (main.c:12-26)
            const int32_t nc = v.n_contigs;
            s.names.assign((size_t)nc, std::string()); s.start.assign((size_t)nc, 0); s.len.assign((size_t)nc, 0);
            for (int32_t i = 0; i < nc; i++) {
                const int32_t r = np_dev_shard_contig_rank(ds, i);
                s.names[(size_t)r] = np_dev_shard_contig_name(ds, i);
                s.start[(size_t)r] = s.off[(size_t)i]; s.len[(size_t)r] = s.off[(size_t)i + 1] - s.off[(size_t)i];
            }
            s.name_ptrs.clear();
            for (auto& nm : s.names) s.name_ptrs.push_back(nm.c_str());
            int64_t sizes[5]; float ms2[2];
            np_dev_shard_stats(ds, sizes, ms2);
            s.h2d = sizes[3] + sizes[2];                       // compressed BAM range + draft bases
            s.d2h = n + (int64_t)(s.off.size() * 8);
        }
    }
    if (rc != NP_OK) { s.rc = rc; s.err = np_last_error(); }
    np_dev_shard_free(ds);
    const double t2 = now_ms();
    s.load_ms = (float)(t1 - t0); s.polish_ms = (float)(t2 - t1);
}

static void worker(np_files* P, Worker* w) {
    np_wait::use_blocking_waits(true);
    for (;;) {
        Slot* s = nullptr;
        {
            std::unique_lock<std::mutex> lk(P->mu);
            P->cv.wait(lk, [&] { return P->quit || P->next_run < P->next_ticket; });
            if (P->next_run >= P->next_ticket) return;            // quit, nothing queued
            s = P->slots[(size_t)(P->next_run % (int64_t)P->slots.size())];
            P->next_run++;
            s->state = Slot::RUNNING;
        }
        run_job(P, *s, w->eng);
        {
            std::lock_guard<std::mutex> lk(P->mu);
            s->state = Slot::DONE;
        }
        P->cv.notify_all();
    }
}

extern "C" {

np_files* np_files_create(int32_t device, int32_t depth) {
    if (depth < 1) depth = 1;
    if (depth > 8) depth = 8;
    np_files* P = new np_files();
    P->device = device;
    for (int i = 0; i < 2 * depth; i++) P->slots.push_back(new Slot());
    for (int i = 0; i < depth; i++) {
        Worker* w = new Worker();
        w->eng = np_engine_create(device);
        if (!w->eng) { delete w; np_files_destroy(P); return nullptr; }
        P->workers.push_back(w);
        w->th = std::thread(worker, P, w);
    }
    return P;
}

void np_files_destroy(np_files* P) {
    if (!P) return;
    {
        std::lock_guard<std::mutex> lk(P->mu);
        P->quit = true;                                  // queued jobs are still run: workers leave when the queue is empty
    }
    P->cv.notify_all();
    for (Worker* w : P->workers) {
        if (w->th.joinable()) w->th.join();
        np_engine_destroy(w->eng);
        delete w;
    }
    cudaSetDevice(P->device);
    for (Slot* s : P->slots) {
        if (s->out) cudaFreeHost(s->out);
        delete s;
    }
    delete P;
}

int64_t np_files_submit(np_files* P, int32_t task, const char* fasta, const char* bam, const Configure* cfg) {
    if (!P || !fasta || !bam || !cfg || (task != NP_TASK_SCORE_CHAIN && task != NP_TASK_KMER_COUNT && task != NP_TASK_SNP_VALID)) { np::set_error("np_files_submit: bad arguments"); return NP_ERR_ARG; }
    int64_t ticket;
    {
        std::lock_guard<std::mutex> lk(P->mu);
        ticket = P->next_ticket;
        Slot& s = *P->slots[(size_t)(ticket % (int64_t)P->slots.size())];
        if (s.state != Slot::IDLE) { np::set_error("np_files_submit: every job record holds an unfinished or unread job: call np_files_wait first"); return NP_ERR_ARG; }
        s.ticket = ticket; s.task = task; s.fasta = fasta; s.bam = bam; s.cfg = *cfg;
        s.cfg.fastafn = s.cfg.bamfn = s.cfg.thirdbamfn = nullptr;
        s.state = Slot::QUEUED;
        P->next_ticket++;
    }
    P->cv.notify_all();
    return ticket;
}

int32_t np_files_wait(np_files* P, int64_t ticket, np_files_result* out) {
    if (!P || ticket < 0) { np::set_error("np_files_wait: unknown ticket"); return NP_ERR_ARG; }
    std::unique_lock<std::mutex> lk(P->mu);
    if (ticket >= P->next_ticket) { np::set_error("np_files_wait: unknown ticket"); return NP_ERR_ARG; }
    Slot& s = *P->slots[(size_t)(ticket % (int64_t)P->slots.size())];
    if (s.ticket != ticket || s.state == Slot::IDLE) { np::set_error("np_files_wait: ticket already consumed"); return NP_ERR_ARG; }
    P->cv.wait(lk, [&] { return s.state == Slot::DONE; });
    s.state = Slot::IDLE;                       // the record's buffers stay untouched until its next submit
    if (s.rc != NP_OK) { np::set_error(s.err); return s.rc; }
    if (out) {
        out->task = s.task; out->n_contigs = (int32_t)s.names.size();
        out->names = s.name_ptrs.data(); out->seq = s.out; out->start = s.start.data(); out->len = s.len.data();
        out->h2d_bytes = s.h2d; out->d2h_bytes = s.d2h; out->load_ms = s.load_ms; out->polish_ms = s.polish_ms;
    }
    return NP_OK;
}

}  // extern "C"
