// engine_v2.h — task 1 as a pipeline of flat kernels: the streaming diff pass (diff_pass.h: every read against the
// 2-bit draft, once), the column kernels (column_pass.h: tile prefix sums, anchors, tables, votes, score chain) and the
// general global-memory kernels of engine_impl.h as the fallback for stretches the tables cannot hold.
#pragma once
#include <stdlib.h>
#include <vector>
#include "column_pass.h"

namespace npe {

struct V2Stats { int32_t W, n_win, smem, unresolved_windows, fallback_cols; };   // tile width, tiles, -, unresolved stretches, fallback columns

// host_ctg_off: contig offsets on the host (n_ctg + 1 entries)
template <class BE>
int run_score_chain_v2(BE& be, Dev& d, const int64_t* host_ctg_off, RunStats* st, V2Stats* vs) {
    if (!rate_is_dyadic(d.P.rate)) return run_score_chain(be, d, st, true);   // see rate_is_dyadic
    const int64_t R = d.n_reads; const int32_t G = d.G;
    if (R >= (int64_t)npw::DIFF_MAX_READS) return run_score_chain(be, d, st);  // read index does not fit a diff entry
    d.task = 1;
    d.err = be.template buf<int32_t>("err", 1);
    be.zero(d.err, sizeof(int32_t));
    d.r_ctg = be.template buf<int32_t>("r_ctg", R + 1);
    d.r_gpos = be.template buf<int32_t>("r_gpos", R + 1);
    d.r_qstart = be.template buf<int32_t>("r_qstart", R + 1);
    d.r_qend = be.template buf<int32_t>("r_qend", R + 1);
    d.r_wend = be.template buf<int32_t>("r_wend", R + 1);
    d.r_hend = be.template buf<int32_t>("r_hend", R + 1);
    d.r_level = be.template buf<uint8_t>("r_level", R + 1);
    d.ins = be.template buf<int32_t>("ins", (size_t)G + 1);
    d.colbase = be.template buf<int32_t>("colbase", (size_t)G + 1);
    d.out_off = be.template buf<int64_t>("out_off", (size_t)d.n_ctg + 1);
    be.zero(d.ins, sizeof(int32_t) * ((size_t)G + 1));
    if (R > 0) be.launch("read_prep", R, ReadPrep{d});
    be.exscan_ncol(d.ins, d.colbase, (int64_t)G);

    // ---- tiles of TW draft positions, never across a contig
    npc::ColGlobals g; memset(&g, 0, sizeof(g));
    std::vector<int32_t> h_tile_off((size_t)d.n_ctg + 1);
    {
        int64_t nt = 0;
        for (int32_t k = 0; k < d.n_ctg; k++) { h_tile_off[(size_t)k] = (int32_t)nt; nt += (host_ctg_off[k + 1] - host_ctg_off[k] + npc::TW - 1) / npc::TW; }
        h_tile_off[(size_t)d.n_ctg] = (int32_t)nt;
        g.n_tiles = (int32_t)nt;
    }
    g.tile_off = be.upload_i32("tile_off", h_tile_off.data(), h_tile_off.size());
    d.C = be.read_i32(d.colbase + G);
    const int32_t C = d.C;

    // ---- streaming half of the pileup scan: every read against the 2-bit draft, once (diff_pass.h)
    npw::DiffGlobals dg; memset(&dg, 0, sizeof(dg));
    dg.dd = be.template buf<npw::Draft2>("dd", (size_t)G / 16 + 12) + 4;
    dg.rdesc = be.template buf<npw::ReadDesc>("rdesc", (size_t)R + 1);
    dg.n_groups = (int32_t)((R + npw::DIFF_GROUP - 1) / npw::DIFF_GROUP);
    const int64_t n_region = (int64_t)dg.n_groups * npw::DIFF_GROUP_SLOTS;            // 4 slots per read on average
    if (n_region + R / 2 + 4096 >= 0x7ffffff0ll) return run_score_chain(be, d, st);
    dg.pool_cap = (int32_t)(n_region + R / 2 + 4096);
    dg.pool = be.template buf<npw::DiffEnt>("dpool", (size_t)dg.pool_cap);
    dg.pool_n = be.template buf<int32_t>("dpool_n", 2);
    dg.gcnt = be.template buf<int32_t>("dpool_gcnt", (size_t)dg.n_groups + 1);
    dg.cov = be.template buf<int32_t>("cov", (size_t)C + 4);
    dg.disb = be.template buf<uint32_t>("disb", (size_t)C / 32 + 8);
    g.tile_cov = be.template buf<int32_t>("tile_cov", (size_t)g.n_tiles + 2);
    g.tile_tbl = be.template buf<int32_t>("tile_tbl", (size_t)g.n_tiles + 2);
    g.tile_str = be.template buf<int32_t>("tile_str", (size_t)g.n_tiles + 2);
    d.obase = be.template buf<uint8_t>("obase", (size_t)C + 1);
    d.oflag = be.template buf<uint8_t>("oflag", (size_t)C + 1);
    d.keepidx = be.template buf<int32_t>("keepidx", (size_t)C + 1);
    be.launch("pack_draft", (int64_t)G / 16 + 2, npw::PackDraft2{d, dg});
    int32_t n_ent = 0, n_ovf = 0, T = 0;
    for (int attempt = 0;; attempt++) {
        g.cov = dg.cov; g.disb = dg.disb;
        be.zero(dg.pool_n, 2 * sizeof(int32_t));
        be.zero(dg.cov, sizeof(int32_t) * ((size_t)C + 4));
        be.zero(dg.disb, sizeof(uint32_t) * ((size_t)C / 32 + 8));
        be.diff_pass(npw::DiffPass{d, dg});
        be.tile_aggregates(d, g);                                   // tile_agg + tile_scan
        const int32_t* ptrs[3] = {d.err, dg.pool_n, g.tile_tbl + g.n_tiles};
        int32_t vals[3];
        be.read_many(ptrs, 3, vals);
        n_ovf = vals[1]; n_ent = (int32_t)n_region + vals[1]; T = vals[2];
        if (!(vals[0] & npw::ERR_DIFF_POOL)) break;
        // noisy shard: the overflow area was too small.  The pass counted what it needs: grow the pool and repeat it once.
        if (attempt > 0 || n_ent <= dg.pool_cap || n_ent >= 0x7ffffff0 - 16) return run_score_chain(be, d, st);
        dg.pool_cap = n_ent + 16;
        dg.pool = be.template buf<npw::DiffEnt>("dpool", (size_t)dg.pool_cap);
        be.zero(d.err, sizeof(int32_t));
    }

    // ---- column half: anchors + tables, votes, score chain (column_pass.h)
    g.T = T;
    g.refw = be.template buf<uint32_t>("refw", (size_t)C / 8 + 4);
    g.tblb = be.template buf<uint32_t>("tblb", (size_t)C / 32 + 8);
    g.tblp = be.template buf<uint32_t>("tblp", (size_t)C / 32 + 8);
    g.tcol = be.template buf<int32_t>("t_col", (size_t)T + 1);
    g.tvotes = be.template buf<uint32_t>("t_votes", (size_t)T + 1);
    g.tflag = be.template buf<uint8_t>("t_flag", (size_t)T + 1);
    g.tbad = be.template buf<uint8_t>("t_bad", (size_t)T + 1);
    g.tunres = be.template buf<uint8_t>("t_unres", (size_t)T + 1);
    g.te = be.template buf<uint32_t>("t_e", (size_t)T * npc::WK + 1);
    g.tfs = be.template buf<uint32_t>("t_fs", (size_t)T * npc::WK + 1);
    g.bt_base = be.template buf<uint32_t>("bt_base", (size_t)T + 1);
    g.bt_pv = be.template buf<uint32_t>("bt_pv", (size_t)T + 1);
    g.bt_am = be.template buf<uint16_t>("bt_am", (size_t)T + 1);
    g.sstart = be.template buf<int32_t>("s_start", (size_t)T + 1);
    g.n_unresolved = be.template buf<int32_t>("n_unres", 2);
    g.n_start = g.tile_str + g.n_tiles;
    be.zero(g.n_unresolved, 2 * sizeof(int32_t));
    be.zero(g.tbad, (size_t)T + 1);
    be.zero(g.tunres, (size_t)T + 1);
    be.zero(g.te, sizeof(uint32_t) * ((size_t)T * npc::WK + 1));
    be.zero(g.tfs, sizeof(uint32_t) * ((size_t)T * npc::WK + 1));
    be.zero(g.refw, sizeof(uint32_t) * ((size_t)C / 8 + 4));
    be.zero(g.tblb, sizeof(uint32_t) * ((size_t)C / 32 + 8));
    if (g.n_tiles > 0) be.column_pass(d, g);
    if (T > 0) {
        const npc::Votes votes{npc::EntryVotes{d, dg, g}, npc::StartVotes{d, dg, g}, (int64_t)n_ovf};
        be.launch("votes", votes.items(), votes);
        be.launch("chain", T, npc::Chain{d, g});
    }
    be.exscan_keep(d.obase, d.keepidx, (int64_t)C);
    int32_t total = 0, err = 0, n_unres = 0;
    {
        const int32_t* ptrs[3] = {d.keepidx + C, d.err, g.n_unresolved};
        int32_t vals[3];
        be.read_many(ptrs, 3, vals);
        total = vals[0]; err = vals[1]; n_unres = vals[2];
    }
    d.T = 0;
    int32_t E = 0, Wd = 0;
    if (n_unres > 0) {
        // fallback: the general kernels on the stretches the tables could not hold (more than WK distinct 3-mers in a column)
        d.needi = be.template buf<int32_t>("needi", (size_t)C + 1);
        d.tidx = be.template buf<int32_t>("tidx", (size_t)C + 1);
        d.votes = be.template buf<uint32_t>("votes", (size_t)C + 1);
        uint8_t* r_need = be.template buf<uint8_t>("r_need", (size_t)R + 1);
        be.zero(d.needi, sizeof(int32_t) * ((size_t)C + 1));
        be.launch("mark_unresolved", T, npc::MarkUnresolved{d, g});
        be.exscan_i32(d.needi, d.tidx, (int64_t)C + 1);
        d.T = be.read_i32(d.tidx + C);
        be.launch("read_need", R, npc::ReadNeed{d, dg, r_need});
        const int32_t Tf = d.T;
        d.r_pm = be.template buf<int32_t>("r_pm", R + 1);
        d.r_c0 = be.template buf<int32_t>("r_c0", R + 1);
        d.r_n = be.template buf<int32_t>("r_n", R + 1);
        d.r_bound = be.template buf<int32_t>("r_bound", R + 1);
        d.r_symoff = be.template buf<int32_t>("r_symoff", R + 1);
        d.refsym = be.template buf<uint8_t>("refsym", (size_t)C + 1);
        d.cflag = be.template buf<uint8_t>("cflag", (size_t)C + 1);
        d.colpos = be.template buf<int32_t>("colpos", (size_t)C + 1);
        be.inclmax_i32(d.r_wend, d.r_pm, R);
        be.launch("col_init", G, ColInit{d});
        be.launch("col_ends", d.n_ctg, ColEnds{d});
        be.launch("sym_bound", R + 1, SymBound{d, r_need});
        be.exscan_i32(d.r_bound, d.r_symoff, R + 1);
        Wd = be.read_i32(d.r_symoff + R);
        d.sym = be.template buf<uint32_t>("sym", (size_t)Wd + 1);
        be.launch("expand", R, Expand{d, 1});
        d.tcols = be.template buf<int32_t>("tcols", (size_t)Tf + 1);
        d.tcap = be.template buf<int32_t>("tcap", (size_t)Tf + 1);
        d.toff = be.template buf<int32_t>("toff", (size_t)Tf + 1);
        d.tnk = be.template buf<int32_t>("tnk", (size_t)Tf + 1);
        d.bpk = be.template buf<uint16_t>("bpk", (size_t)Tf * 16);
        d.amax = be.template buf<uint8_t>("amax", (size_t)Tf + 1);
        be.launch("table_cols", C, TableCols{d});
        be.exscan_i32(d.tcap, d.toff, (int64_t)Tf + 1);
        E = be.read_i32(d.toff + Tf);
        d.ktab = be.template buf<uint32_t>("ktab", (size_t)E + 1);
        be.launch("build_table", Tf, BuildTable{d});
        be.launch("chain_dp", Tf, ChainDP{d});
        be.exscan_keep(d.obase, d.keepidx, (int64_t)C);
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
    if (vs) { vs->W = npc::TW; vs->n_win = g.n_tiles; vs->smem = 0; vs->unresolved_windows = n_unres; vs->fallback_cols = d.T; }
    return err;
}

}  // namespace npe
