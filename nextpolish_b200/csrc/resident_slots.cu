// resident_slots.cu — shards that already sit in HBM, polished by several engines at once.
//
// One run of task 1 or task 2 over a 5 Mb shard is a chain of ~60 short kernels (10-120 us each, most of them latency-bound
// at partial occupancy) with a handful of host synchronisations: alone it leaves most of the GPU idle.  Contigs — and
// therefore shards — are independent units (the reference runs one contig per worker process, nextpolish1.py:219-224), so
// the way to fill the machine is the one the from-files pipeline and np_multi already use: `slots` engines, each with its
// own stream, scratch buffers and host thread, working on different shards concurrently.  np_resident_submit hands a job
// (task, device shard view, destination buffer in HBM) to the next slot in round-robin order and returns; the slot adopts
// the shard in place, runs the task and writes the result in the gather form (np_engine_pack_result: 16-byte header with the
// byte count, then the polished bytes) into the caller's device buffer.
#include <cuda_runtime.h>

#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "errors.h"
#include "stream_wait.h"
#include "../../include/nextpolish_b200.h"

namespace {
struct RSlot {
    np_engine* eng = nullptr;
    std::thread th;
    std::mutex mu; std::condition_variable cv;
    enum { IDLE, QUEUED, RUNNING, DONE } state = IDLE;
    bool quit = false;
    int64_t ticket = -1; int32_t task = 0; np_shard_view view; Configure cfg; void* dst = nullptr; int64_t dst_cap = 0;
    int32_t rc = NP_OK; std::string err; int64_t out_bytes = 0;
};
}  // namespace

struct np_resident {
    int device = 0;
    std::vector<RSlot*> slots;
    int64_t next_ticket = 0;
};

static void rworker(np_resident* P, RSlot* sp) {
    RSlot& s = *sp;
    cudaSetDevice(P->device);
    np_wait::use_blocking_waits(true);
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(s.mu);
            s.cv.wait(lk, [&] { return s.quit || s.state == RSlot::QUEUED; });
            if (s.quit) return;
            s.state = RSlot::RUNNING;
        }
        int32_t rc = np_engine_adopt_device(s.eng, &s.view);
        if (rc == NP_OK) rc = np_engine_run(s.eng, s.task, &s.cfg);
        if (rc == NP_OK) { s.out_bytes = np_engine_result_bytes(s.eng); rc = np_engine_pack_result(s.eng, s.dst, s.dst_cap); }
        if (rc == NP_OK) rc = np_engine_sync(s.eng);
        s.rc = rc;
        if (rc != NP_OK) s.err = np_last_error();
        {
            std::lock_guard<std::mutex> lk(s.mu);
            s.state = RSlot::DONE;
        }
        s.cv.notify_all();
    }
}

extern "C" {

void np_resident_destroy(np_resident* P) {
    if (!P) return;
    for (RSlot* s : P->slots) {
        {
            std::unique_lock<std::mutex> lk(s->mu);
            s->cv.wait(lk, [&] { return s->state == RSlot::IDLE || s->state == RSlot::DONE; });
            s->quit = true;
        }
        s->cv.notify_all();
        if (s->th.joinable()) s->th.join();
        np_engine_destroy(s->eng);
        delete s;
    }
    delete P;
}

np_resident* np_resident_create(int32_t device, int32_t slots) {
    if (slots < 1) slots = 1;
    if (slots > 16) slots = 16;
    np_resident* P = new np_resident();
    P->device = device;
    for (int i = 0; i < slots; i++) {
        RSlot* s = new RSlot();
        s->eng = np_engine_create(device);
        if (!s->eng) { delete s; np_resident_destroy(P); return nullptr; }
        P->slots.push_back(s);
        s->th = std::thread(rworker, P, s);
    }
    return P;
}

int64_t np_resident_submit(np_resident* P, int32_t task, const np_shard_view* dev_shard, const Configure* cfg,
                           void* dst_device, int64_t dst_cap) {
    if (!P || !dev_shard || !cfg || !dst_device) { np::set_error("np_resident_submit: bad arguments"); return NP_ERR_ARG; }
    const int64_t ticket = P->next_ticket;
    RSlot& s = *P->slots[(size_t)(ticket % (int64_t)P->slots.size())];
    {
        std::unique_lock<std::mutex> lk(s.mu);
        if (s.state != RSlot::IDLE) { np::set_error("np_resident_submit: every slot holds an unfinished or unread job: call np_resident_wait first"); return NP_ERR_ARG; }
        s.ticket = ticket; s.task = task; s.view = *dev_shard; s.cfg = *cfg; s.dst = dst_device; s.dst_cap = dst_cap;
        s.cfg.fastafn = s.cfg.bamfn = s.cfg.thirdbamfn = nullptr;
        s.state = RSlot::QUEUED;
    }
    s.cv.notify_all();
    P->next_ticket++;
    return ticket;
}

// Blocks until the job is complete on the device (its result is in dst_device); returns its status, the polished byte
// count through out_bytes.  Tickets are consumed once, in any order.
int32_t np_resident_wait(np_resident* P, int64_t ticket, int64_t* out_bytes) {
    if (!P || ticket < 0 || ticket >= P->next_ticket) { np::set_error("np_resident_wait: unknown ticket"); return NP_ERR_ARG; }
    RSlot& s = *P->slots[(size_t)(ticket % (int64_t)P->slots.size())];
    std::unique_lock<std::mutex> lk(s.mu);
    if (s.ticket != ticket || s.state == RSlot::IDLE) { np::set_error("np_resident_wait: ticket already consumed"); return NP_ERR_ARG; }
    s.cv.wait(lk, [&] { return s.state == RSlot::DONE; });
    s.state = RSlot::IDLE;
    if (s.rc != NP_OK) { np::set_error(s.err); return s.rc; }
    if (out_bytes) *out_bytes = s.out_bytes;
    return NP_OK;
}

int64_t np_resident_launch_count(np_resident* P) {     // kernel launches of every job so far
    int64_t n = 0;
    if (P) for (RSlot* s : P->slots) n += np_engine_launch_total(s->eng);
    return n;
}

}  // extern "C"
