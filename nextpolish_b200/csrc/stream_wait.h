// stream_wait.h — how a host thread waits for its CUDA stream.
//
// cudaStreamSynchronize spins on the CPU: the lowest latency between the ~60 short kernels of a run, and what every
// front end uses by default.  The pipelined front ends (np_files, np_resident, np_multi) run one host thread per slot;
// with NEXTPOLISH_B200_SPIN=0 their workers sleep on a blocking-sync event instead and leave their core to the threads
// that have host work.  Measured on B200 with 16 host threads (tools/prof_slots.py): sleeping costs ~0.2 ms per wake-up —
// one resident engine 1.6 -> 3.9 ms per step, eight slots 0.72 -> 1.0 ms — so it stays an option for hosts with fewer
// cores than slots, not the default.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

namespace np_wait {
inline bool& blocking_flag() { static thread_local bool b = false; return b; }
inline void use_blocking_waits(bool on) {
    const char* e = getenv("NEXTPOLISH_B200_SPIN");
    blocking_flag() = on && e && e[0] == '0';
}
inline cudaError_t stream_wait(cudaStream_t s) {
    if (!blocking_flag()) return cudaStreamSynchronize(s);
    static thread_local cudaEvent_t ev = nullptr;
    static thread_local int ev_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!ev || ev_dev != dev) {
        if (ev) cudaEventDestroy(ev);
        ev = nullptr;
        cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) { ev = nullptr; return cudaStreamSynchronize(s); }
        ev_dev = dev;
    }
    cudaError_t e = cudaEventRecord(ev, s);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ev);
}
}  // namespace np_wait
