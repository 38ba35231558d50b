// engine_v2.h — task 1 with the fused shared-memory window kernel (window_kernel.h) and the general
// global-memory kernels of engine_impl.h as the fallback for whatever a window leaves unresolved.
#pragma once
#include <stdlib.h>
#include <vector>
#include "window_kernel.h"

namespace npe {

// One window, all phases, for a backend-provided "thread range" (tid, nt) and barrier.
// CUDA: tid = threadIdx.x, nt = blockDim.x, barrier = __syncthreads; emu: tid = 0, nt = 1, no-op.
#define NP_WINDOW_PHASES(x, tid, nt, ops, BARRIER, STAMP)          \
    npw::ph_compare(x, tid, nt, ops);     BARRIER; STAMP(2);        \
    npw::ph_scan(x, tid, nt, ops);        BARRIER;                  \
    npw::ph_mark_tables(x, tid, nt, ops); BARRIER; STAMP(3);        \
    npw::ph_votes(x, tid, nt, ops);       BARRIER;                  \
    npw::ph_tally(x, tid, nt);            BARRIER; STAMP(4);        \
    npw::ph_chain(x, tid, nt);            /* disjoint columns: */   \
    npw::ph_anchors(x, tid, nt);          BARRIER; STAMP(5);        \
    npw::ph_finish(x, tid, nt, ops);

struct V2Stats { int32_t W, n_win, smem, unresolved_windows, fallback_cols; };

// host_ctg_off: contig offsets on the host (n_ctg + 1 entries)
template <class BE>
int run_score_chain_v2(BE& be, Dev& d, const int64_t* host_ctg_off, RunStats* st, V2Stats* vs) {
    if (!rate_is_dyadic(d.P.rate)) return run_score_chain(be, d, st, true);   // see rate_is_dyadic
    const int64_t R = d.n_reads; const int32_t G = d.G;
    d.task = 1;
    d.err = be.template buf<int32_t>("err", 1);
    be.zero(d.err, sizeof(int32_t));
    d.r_ctg = be.template buf<int32_t>("r_ctg", R + 1);
    d.r_gpos = be.template buf<int32_t>("r_gpos", R + 1);
    d.r_qstart = be.template buf<int32_t>("r_qstart", R + 1);
    d.r_qend = be.template buf<int32_t>("r_qend", R + 1);
    d.r_wend = be.template buf<int32_t>("r_wend", R + 1);
    d.r_hend = be.template buf<int32_t>("r_hend", R + 1);
    d.r_pm = be.template buf<int32_t>("r_pm", R + 1);
    d.r_c0 = be.template buf<int32_t>("r_c0", R + 1);
    d.r_n = be.template buf<int32_t>("r_n", R + 1);
    d.r_bound = be.template buf<int32_t>("r_bound", R + 1);
    d.r_symoff = be.template buf<int32_t>("r_symoff", R + 1);
    d.r_level = be.template buf<uint8_t>("r_level", R + 1);
    d.ins = be.template buf<int32_t>("ins", (size_t)G + 1);
    d.colbase = be.template buf<int32_t>("colbase", (size_t)G + 1);
    d.out_off = be.template buf<int64_t>("out_off", (size_t)d.n_ctg + 1);
    be.zero(d.ins, sizeof(int32_t) * ((size_t)G + 1));
    if (R > 0) {
        be.launch("read_prep", R, ReadPrep{d});
        be.inclmax_i32(d.r_wend, d.r_pm, R);
    }
    be.exscan_ncol(d.ins, d.colbase, (int64_t)G);

    // ---- window plan: largest W whose biggest window fits the shared-memory budget
    npw::WinGlobals g; memset(&g, 0, sizeof(g));
    g.maxneed = be.template buf<int32_t>("w_maxneed", 2);
    g.n_unresolved = g.maxneed + 1;
    g.r_need = be.template buf<uint8_t>("r_need", (size_t)R + 1);
    const int32_t budget = 100 * 1024, hard = 200 * 1024;
    int32_t need = 0; bool fits = false;
    std::vector<int32_t> hw_ctg, hw_p0;
    int32_t kW[3] = {512, 256, 128};
    int wi0 = 0;
    if (const char* ev = getenv("NEXTPOLISH_B200_WINDOW")) {   // tuning only: first window size to try
        int v = atoi(ev);
        if (v >= 64 && v <= 2048 && v % 32 == 0) { kW[0] = v; if (v <= 256) kW[1] = v / 2 > 64 ? v / 2 : 64; if (v <= 128) kW[2] = 64; }
    }
    for (int wi = wi0; wi < 3 && !fits; wi++) {
        g.W = kW[wi];
        hw_ctg.clear(); hw_p0.clear();
        for (int32_t k = 0; k < d.n_ctg; k++)
            for (int64_t p = host_ctg_off[k]; p < host_ctg_off[k + 1]; p += g.W) { hw_ctg.push_back(k); hw_p0.push_back((int32_t)p); }
        g.n_win = (int32_t)hw_ctg.size();
        g.win_ctg = be.upload_i32("w_ctg", hw_ctg.data(), hw_ctg.size());
        g.win_p0 = be.upload_i32("w_p0", hw_p0.data(), hw_p0.size());
        g.win_rlo = be.template buf<int32_t>("w_rlo", (size_t)g.n_win + 1);
        g.win_rhi = be.template buf<int32_t>("w_rhi", (size_t)g.n_win + 1);
        g.win_need = be.template buf<int32_t>("w_need", (size_t)g.n_win + 1);
        be.zero(g.maxneed, 2 * sizeof(int32_t));
        if (g.n_win > 0) be.launch("win_plan", g.n_win, npw::WinPlan{d, g});
        const int32_t* ptrs[2] = {g.maxneed, d.colbase + G};
        int32_t vals[2];
        be.read_many(ptrs, 2, vals);
        need = vals[0]; d.C = vals[1];
        fits = need <= (wi == 2 ? hard : budget);
    }
    const int32_t C = d.C;
    if (!fits) return run_score_chain(be, d, st);          // e.g. extreme depth: general kernels only

    d.mism = be.template buf<uint8_t>("mism", (size_t)C + 1);
    d.obase = be.template buf<uint8_t>("obase", (size_t)C + 1);
    d.oflag = be.template buf<uint8_t>("oflag", (size_t)C + 1);
    d.votes = be.template buf<uint32_t>("votes", (size_t)C + 1);
    d.needi = be.template buf<int32_t>("needi", (size_t)C + 1);
    d.tidx = be.template buf<int32_t>("tidx", (size_t)C + 1);
    d.keepidx = be.template buf<int32_t>("keepidx", (size_t)C + 1);
    be.zero(d.needi, sizeof(int32_t) * ((size_t)C + 1));
    be.zero(g.r_need, (size_t)R + 1);
    if (g.n_win > 0) be.run_windows(d, g, need);

    // windows that left something unresolved bump a counter: the compaction scan over all columns and the
    // general kernels run only then
    int32_t n_unres = be.read_i32(g.n_unresolved);
    d.T = 0;
    if (n_unres > 0) {
        be.exscan_i32(d.needi, d.tidx, (int64_t)C + 1);
        d.T = be.read_i32(d.tidx + C);
    }
    int32_t E = 0, Wd = 0;
    if (d.T > 0) {                                           // fallback: general kernels on the marked stretches
        const int32_t T = d.T;
        d.refsym = be.template buf<uint8_t>("refsym", (size_t)C + 1);
        d.cflag = be.template buf<uint8_t>("cflag", (size_t)C + 1);
        d.colpos = be.template buf<int32_t>("colpos", (size_t)C + 1);
        be.launch("col_init", G, ColInit{d});
        be.launch("col_ends", d.n_ctg, ColEnds{d});
        be.launch("sym_bound", R + 1, SymBound{d, g.r_need});
        be.exscan_i32(d.r_bound, d.r_symoff, R + 1);
        Wd = be.read_i32(d.r_symoff + R);
        d.sym = be.template buf<uint32_t>("sym", (size_t)Wd + 1);
        be.launch("expand", R, Expand{d, 1});
        d.tcols = be.template buf<int32_t>("tcols", (size_t)T + 1);
        d.tcap = be.template buf<int32_t>("tcap", (size_t)T + 1);
        d.toff = be.template buf<int32_t>("toff", (size_t)T + 1);
        d.tnk = be.template buf<int32_t>("tnk", (size_t)T + 1);
        d.bpk = be.template buf<uint16_t>("bpk", (size_t)T * 16);
        d.amax = be.template buf<uint8_t>("amax", (size_t)T + 1);
        be.launch("table_cols", C, TableCols{d});
        be.exscan_i32(d.tcap, d.toff, (int64_t)T + 1);
        E = be.read_i32(d.toff + T);
        d.ktab = be.template buf<uint32_t>("ktab", (size_t)E + 1);
        be.launch("build_table", T, BuildTable{d});
        be.launch("chain_dp", T, ChainDP{d});
    }
    be.exscan_keep(d.obase, d.keepidx, (int64_t)C);
    int32_t total = 0, err = 0;
    {
        const int32_t* ptrs[2] = {d.keepidx + C, d.err};
        int32_t vals[2];
        be.read_many(ptrs, 2, vals);
        total = vals[0]; err = vals[1];
    }
    d.out = be.template buf<uint8_t>("out", (size_t)total + 1);
    if (C > 0) be.launch("emit", C, Emit{d, (uint8_t)(FLAG_ZERO | FLAG_COVERAGE)});
    be.launch("out_offsets", (int64_t)d.n_ctg + 1, OutOffsets{d});
    run_trace(be, d);
    if (st) { st->C = C; st->T = d.T; st->sym_words = Wd; st->table_entries = E; st->out_bytes = total; }
    if (vs) { vs->W = g.W; vs->n_win = g.n_win; vs->smem = need; vs->unresolved_windows = n_unres; vs->fallback_cols = d.T; }
    return err;
}

}  // namespace npe
