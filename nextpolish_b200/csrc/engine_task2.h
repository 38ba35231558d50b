// engine_task2.h — task 2 (kmer_count, kmercount.c:93-126) as data-parallel passes.
//
//   flag runs      per position : lowercase runs of the draft (FLAG_ZERO, contig.c:94-97)
//   regions        per contig   : contig_get_region x2 over the run list + contig_merge_region
//                                 (contig.c:498-620) -> no-depth regions, k-mer regions
//   insert layout  per read     : contig_create_insert_region (contig.c:182-200) -> ins[], colbase
//   nodepth score  per region   : contig_score_correct(region, 0x12) (contig.c:706-734): level-2
//                                 votes, chain DP, backtrack, then level-1 votes on the still
//                                 unsupported sub-regions, re-score, re-correct
//   split          per region   : ss_spilt_region (kmercount.c:128-173) -> windows
//   window vote    per window   : ss_kmer_correct (kmercount.c:175-261): spanning level-2 reads,
//                                 identical-string tally, 50 x MAPQ60 cap, winner
//   apply + emit                : contig_update_contig (contig.c:811-821), contig_get_contig(FLAG_ZERO)
//
// Regions and windows are independent units (one thread each; they are few and small); the
// order-dependent parts of the reference (first-seen tie-breaks, the MAPQ-60 cap, last-writer-wins
// on window end columns) are kept by iterating reads in BAM order inside a unit and by applying
// window winners in window order.
#pragma once
#include "engine_impl.h"

#ifndef NP_FORCE_REGION_WALK
#define NP_FORCE_REGION_WALK 0      // test builds set 1: always take the literal per-contig merge (the fallback of the flat merge)
#endif

namespace npe {

struct Dev2 {
    Dev d;                       // shared layout / read arrays (task = 2)
    // lowercase runs
    int32_t *rs_flag, *rs_idx;   // per position: run-start flag, its exclusive scan
    int32_t *re_flag, *re_idx;   // per position: run-end flag, its exclusive scan
    int32_t *run_s, *run_e;      // [n_runs]
    int32_t n_runs;
    // region lists per contig (slices [ctg_run_off[k], ctg_run_off[k+1]) * 2 ints)
    int32_t *gnd, *gkm;          // [2*n_runs+2] regions of every run group, written at the group head's slot
    int32_t *gcnt_nd, *gcnt_km;  // [n_runs] regions emitted by the group headed at this run (0: not a head)
    // flat merge of the region lists (RegionDense / RegionHeads / RegionWrite): arrays indexed by list * (n_runs+1) + slot
    int32_t *gvalid, *gpos;      // [2*(n_runs+1)] slot holds a region; exclusive scan (dense index)
    int32_t *dreg, *dctg;        // dense regions per list: (start, end) pairs at dreg + list * 2 * (n_runs+1); contig of each
    uint8_t* dhead;              // dense: region starts a merged region
    int32_t *ghead, *gout;       // per slot: valid && head; exclusive scan (index of the merged region)
    int32_t* reg_bad;            // [1] a contig's list is not in the shape the flat merge handles (see RegionHeads)
    int32_t *nd_reg, *km_reg;    // [2*n_runs+2] (start,end) pairs, global positions (fallback path)
    int32_t *nd_cnt, *km_cnt;    // per contig: number of regions (pairs)
    int32_t *nd_off, *km_off;    // exclusive scans of the counts
    int32_t *ndl, *kml;          // compacted lists [2*NR]
    int32_t NR_nd, NR_km;
    uint8_t* inreg;              // per position (+1): inside (start, end] of some region
    int32_t* r_hpm;              // prefix max of r_hend
    // nodepth scoring scratch
    int32_t *ndmark, *ndidx;     // per column
    int32_t *vcap, *koff;        // per column vote capacity and its scan
    int32_t NRC;                 // number of nodepth-region columns
    uint32_t* ktab2;             // k-mer lists
    int32_t* nk2;                // [NRC] list lengths
    uint32_t* cnt2;              // [NRC] votes
    uint16_t* refk2;             // [NRC] refkmer
    double* sc2;                 // [NRC*16] scores by base code
    uint16_t* kc2;               // [NRC*16]
    uint8_t* ord2;               // [NRC*16] first-seen order of base codes
    uint8_t* ns2;                // [NRC] argmax base code of the column's score list
    int32_t* subbuf;             // [NRC + 2*NR_nd + 2] inner region lists
    // (region, read) pairs of the no-depth pass: the CIGAR walks run one thread per pair
    int32_t *nd_pcnt, *nd_poff;  // per region: candidate reads (level >= 1), scan
    int32_t *nd_sb, *nd_soff;    // per region: symbol-slot bytes (pairs * columns), scan
    int32_t *ndp_read, *ndp_first, *ndp_n;   // per pair: read, first column (relative to the region), symbols
    uint8_t* ndp_sym;            // symbol slots
    int32_t NP_nd;
    int32_t *nd_reg_of, *nd_col_of;   // per no-depth column (compact index): its region, its global column
    // windows
    int32_t *wcnt, *woff;        // per k-mer region: window count, scan
    int32_t *win;                // [2*NW] (start,end)
    int32_t NW;
    int32_t *wcand, *wsoff;      // per window: candidate count; scratch offset (bytes, in 4-byte units)
    int32_t* wscratch;           // strings + tallies
    int32_t* wbest;              // [NW] offset of the winner string in wscratch (bytes), or -1
    int32_t flagzero;            // ss_kmer_correct's flagzero (kmercount.c:175): 1 = parsed reads leave FLAG_ZERO alone and a window
                                 // that found a string is cleared as a whole (snp_valid's first pass, snpvalid.c:18)
    int32_t *wfail, *wfidx;      // snp_valid: window found no string; exclusive scan
    int32_t *fcnt, *foff;        // snp_valid: sub-windows of every failed window (snpvalid.c:37-66), scan
    const int32_t* win1;         // snp_valid: the first pass's windows while the second list is built
    int32_t NW1; int32_t* warn;  // [1] device flag: a cut-point list was odd / inverted (reference behaviour undefined)
    // (window, read) pairs of the vote: walks run one thread per pair
    int32_t *wp_cnt, *wp_off;    // per window: candidates (+1 slot for the stale record), scan
    int32_t *wp_read;            // per pair: read index (-1: unused stale slot)
    int32_t NP_w;
};

enum { ERR_SHARED_ENDPOINT = 16, ERR_REGION_SCRATCH = 32, ERR_QUAL_MISSING = 64 };

// ---- lowercase runs ---------------------------------------------------------------------------
NP_HD bool pos_flagged(const Dev& d, int32_t p) { uint32_t ch = d.ctg_seq[p]; return ch >= 97 && ch <= 122; }

struct RunFlags {
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t p, B&) const {
        const Dev& d = w.d;
        int32_t s = 0, e = 0;
        if (p < d.G && pos_flagged(d, (int32_t)p)) {
            int32_t k = find_contig_i32(d.ctg_goff, d.n_ctg, (int32_t)p);
            int32_t gs = d.ctg_goff[k], ge = d.ctg_goff[k + 1] - 1;
            s = (p == gs) || !pos_flagged(d, (int32_t)p - 1);
            e = (p == ge) || !pos_flagged(d, (int32_t)p + 1);
        }
        w.rs_flag[p] = s; w.re_flag[p] = e;
    }
};
struct RunFill {
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t p, B&) const {
        if (w.rs_flag[p]) w.run_s[w.rs_idx[p]] = (int32_t)p;
        // the run ending at p is the last one started at or before p: no second scan needed
        if (w.re_flag[p]) w.run_e[w.rs_idx[p] + w.rs_flag[p] - 1] = (int32_t)p;
    }
};

// contig_brim_no_extension / contig_brim_with_extension on positions (contig.c:498-517)
NP_HD void brim_pos(const Dev& d, bool with_ext, int32_t ext, int32_t bstart, int32_t bend, int32_t* start, int32_t* end) {
    *start = *start >= bstart + ext ? *start - ext : bstart;
    *end = *end <= bend - ext ? *end + ext : bend;
    if (!with_ext) return;
    while (*start > bstart) {
        int32_t p = *start + 1;
        uint32_t a = p <= bend ? d.ctg_seq[p] : 0, b = d.ctg_seq[p - 1];
        if (a >= 97 && a <= 122) a -= 32;
        uint32_t bu = (b >= 97 && b <= 122) ? b - 32 : b;
        bool same = p <= bend && base_code(a) == base_code(bu);
        if (same || (b >= 97 && b <= 122)) (*start)--; else break;
    }
    while (*end < bend) {
        int32_t p = *end - 1;
        uint32_t a = p >= bstart ? d.ctg_seq[p] : 0, b = d.ctg_seq[p + 1];
        if (a >= 97 && a <= 122) a -= 32;
        uint32_t bu = (b >= 97 && b <= 122) ? b - 32 : b;
        bool same = p >= bstart && base_code(a) == base_code(bu);
        if (same || (b >= 97 && b <= 122)) (*end)++; else break;
    }
}

// contig_get_region (contig.c:519-563) over the lowercase runs of one contig. Appends (start,end)
// pairs to out; returns the number of pairs.
NP_HD int32_t regions_from_runs(const Dev& d, const int32_t* run_s, const int32_t* run_e, int32_t r0, int32_t r1,
                                int32_t gs, int32_t ge, int32_t gap, int32_t con, bool with_ext, int32_t ext, int32_t* out) {
    int32_t n = 0, cur = gs, r = r0;
    while (r < r1) {
        if (run_e[r] < cur) { r++; continue; }                  // run swallowed by an earlier extension
        int32_t qstart = run_s[r] > cur ? run_s[r] : cur;
        int32_t qend = run_e[r];
        int32_t pcon = qend - qstart + 1;
        r++;
        while (r < r1 && run_s[r] - qend - 1 <= gap) { pcon = run_e[r] - run_s[r] + 1; qend = run_e[r]; r++; }
        int32_t i = qend + gap + 1;                              // position where pgap first exceeds gap
        if (i > ge) {                                            // open at the contig end: emitted unconditionally
            brim_pos(d, with_ext, ext, gs, ge, &qstart, &qend);
            out[2 * n] = qstart; out[2 * n + 1] = qend; n++;
            break;
        }
        if (pcon > con) {
            brim_pos(d, with_ext, ext, gs, ge, &qstart, &qend);
            out[2 * n] = qstart; out[2 * n + 1] = qend; n++;
            cur = (qend > i ? qend : i) + 1;
        } else cur = i + 1;
    }
    return n;
}

// contig_merge_region (contig.c:595-620), literal, in place; returns the new pair count
NP_HD int32_t merge_regions(int32_t* l, int32_t npairs) {
    if (npairs == 0) return 0;
    int32_t ps = 0, qs = 0, length = 1;
    for (int32_t i = 0; i < npairs; i++) {
        if (l[2 * ps] >= l[2 * qs + 1]) {
            qs += 1;
            if (qs != ps) { l[2 * qs] = l[2 * ps]; l[2 * qs + 1] = l[2 * ps + 1]; }
            length += 1;
        } else {
            while (l[2 * ps] < l[2 * qs]) qs -= 1;
            l[2 * qs + 1] = l[2 * ps + 1];
        }
        ps += 1;
    }
    return length;
}

// Right end of the region a cluster whose last run ends at position e would be extended to by
// contig_brim_* (contig.c:498-517); it depends on e only.
NP_HD int32_t reach_right(const Dev& d, bool with_ext, int32_t ext, int32_t gs, int32_t ge, int32_t e) {
    int32_t a = e, b = e;
    brim_pos(d, with_ext, ext, gs, ge, &a, &b);
    return b;
}
// The region scan (contig.c:519-563) is sequential because an emitted region's extension can swallow
// the start of the next cluster (`i = qend`).  Between run r and r+1 nothing carries over when the gap
// closes the cluster AND the extension of a cluster ending at r cannot reach run r+1: such "hard
// boundaries" cut the run list into groups that are scanned independently, one thread per group.
NP_HD bool hard_boundary(const Dev& d, const int32_t* run_s, const int32_t* run_e, int32_t r, int32_t gs, int32_t ge,
                         int32_t gap, bool with_ext, int32_t ext) {
    if (run_s[r + 1] - run_e[r] - 1 <= gap) return false;
    return reach_right(d, with_ext, ext, gs, ge, run_e[r]) < run_s[r + 1];
}
struct RegionGroups {    // one thread per lowercase run and variant (0: no-depth regions, 1: k-mer regions)
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t it, B&) const {
        const Dev& d = w.d;
        int32_t r = (int32_t)(it >> 1), variant = (int32_t)(it & 1);
        int32_t* cnt = variant ? w.gcnt_km : w.gcnt_nd;
        int32_t* out = (variant ? w.gkm : w.gnd) + 2 * (size_t)r;
        int32_t k = find_contig_i32(d.ctg_goff, d.n_ctg, w.run_s[r]);
        int32_t gs = d.ctg_goff[k], ge = d.ctg_goff[k + 1] - 1;
        int32_t r0 = w.rs_idx[gs], r1 = w.rs_idx[ge + 1];
        int32_t gap = variant ? d.P.min_len_inter_kmer : 0, con = variant ? 0 : d.P.min_len_ldr;
        bool with_ext = variant != 0; int32_t ext = d.P.ext_len_edge;
        cnt[r] = 0;
        if (r == 0) { w.gvalid[variant * ((size_t)w.n_runs + 1) + w.n_runs] = 0; *w.reg_bad = 0; }   // the scans' total entry
        if (r > r0 && !hard_boundary(d, w.run_s, w.run_e, r - 1, gs, ge, gap, with_ext, ext)) return;   // not a group head
        int32_t rend = r + 1;
        while (rend < r1 && !hard_boundary(d, w.run_s, w.run_e, rend - 1, gs, ge, gap, with_ext, ext)) rend++;
        const int32_t n = regions_from_runs(d, w.run_s, w.run_e, r, rend, gs, ge, gap, con, with_ext, ext, out);
        cnt[r] = n;
        // the group's regions sit in the slots of its first n runs: the head marks every slot of its group
        for (int32_t q = r; q < rend; q++) w.gvalid[variant * ((size_t)w.n_runs + 1) + q] = q - r < n;
    }
};
// ---- flat merge of the region lists ------------------------------------------------------------
// contig_merge_region (contig.c:595-620) walks one contig's list; in the shape every list has unless a region's left
// extension runs across a whole earlier region (starts and ends non-decreasing inside a contig, first region not empty)
// it reduces to: region i opens a merged region iff it is the contig's first or start_i >= end_(i-1); a merged region ends
// with the end of its last member.  That form is data-parallel over ALL regions of ALL contigs (slot order = position
// order = contig order), so the lists are merged and compacted by three flat kernels and two scans instead of one thread
// per contig.  Any other shape raises reg_bad and the per-contig kernels below (the literal walk) redo the lists.
struct RegionDense {     // one thread per slot and list: copy the slot's region to its dense index
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t it, B&) const {
        const Dev& d = w.d;
        const int64_t s = it >> 1; const int variant = (int)(it & 1);
        const size_t N1 = (size_t)w.n_runs + 1, ix = variant * N1 + (size_t)s;
        if (!w.gvalid[ix]) return;
        const int32_t i = w.gpos[ix];
        const int32_t* src = (variant ? w.gkm : w.gnd) + 2 * (size_t)s;
        int32_t* dst = w.dreg + variant * 2 * N1;
        dst[2 * (size_t)i] = src[0]; dst[2 * (size_t)i + 1] = src[1];
        w.dctg[variant * N1 + i] = find_contig_i32(d.ctg_goff, d.n_ctg, src[0]);
    }
};
struct RegionHeads {     // one thread per slot (+ the scans' total entry) and list
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t it, B&) const {
        const int64_t s = it >> 1; const int variant = (int)(it & 1);
        const size_t N1 = (size_t)w.n_runs + 1, ix = variant * N1 + (size_t)s;
        w.ghead[ix] = 0;
        if (s >= w.n_runs || !w.gvalid[ix]) return;
        const size_t i = (size_t)w.gpos[ix];
        const int32_t* reg = w.dreg + variant * 2 * N1;
        const int32_t* ctg = w.dctg + variant * N1;
        const bool first = i == 0 || ctg[i - 1] != ctg[i];
        bool head = true;
        if (first) { if (reg[2 * i] >= reg[2 * i + 1]) *w.reg_bad = 1; }       // the literal walk compares pair 0 with itself
        else {
            if (reg[2 * i] < reg[2 * i - 2] || reg[2 * i + 1] < reg[2 * i - 1]) *w.reg_bad = 1;
            head = reg[2 * i] >= reg[2 * i - 1];
        }
        w.ghead[ix] = head;
        w.dhead[variant * N1 + i] = head;
    }
};
struct RegionWrite {     // one thread per slot and list: merged regions straight into the compact lists
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t it, B&) const {
        const int64_t s = it >> 1; const int variant = (int)(it & 1);
        const size_t N1 = (size_t)w.n_runs + 1, ix = variant * N1 + (size_t)s;
        if (!w.gvalid[ix]) return;
        const size_t i = (size_t)w.gpos[ix];
        const size_t n = (size_t)w.gpos[variant * N1 + (size_t)w.n_runs];     // dense regions of this list
        const int32_t* reg = w.dreg + variant * 2 * N1;
        const int32_t* ctg = w.dctg + variant * N1;
        const uint8_t* hd = w.dhead + variant * N1;
        int32_t* out = variant ? w.kml : w.ndl;
        const size_t j = (size_t)(w.gout[ix] + w.ghead[ix] - 1);
        if (w.ghead[ix]) out[2 * j] = reg[2 * i];
        if (i + 1 >= n || ctg[i + 1] != ctg[i] || hd[i + 1]) out[2 * j + 1] = reg[2 * i + 1];
    }
};
struct ContigRegions {   // one thread per contig and list (0: no-depth, 1: k-mer): concatenate its groups' regions, contig_merge_region
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t it, B&) const {
        const Dev& d = w.d;
        const int64_t k = it >> 1; const bool km_list = (it & 1) != 0;
        int32_t gs = d.ctg_goff[k], ge = d.ctg_goff[k + 1] - 1;
        int32_t a = 0;
        if (ge >= gs) {
            int32_t r0 = w.rs_idx[gs], r1 = w.rs_idx[ge + 1];
            int32_t* dst = (km_list ? w.km_reg : w.nd_reg) + 2 * (size_t)r0 + 2 * (size_t)k;   // slack of one pair per contig
            const int32_t* cnt = km_list ? w.gcnt_km : w.gcnt_nd;
            const int32_t* src = km_list ? w.gkm : w.gnd;
            for (int32_t r = r0; r < r1; r++)
                for (int32_t q = 0; q < cnt[r]; q++, a++) { dst[2 * a] = src[2 * (size_t)r + 2 * q]; dst[2 * a + 1] = src[2 * (size_t)r + 2 * q + 1]; }
            a = merge_regions(dst, a);
        }
        (km_list ? w.km_cnt : w.nd_cnt)[k] = a;
        if (k == 0) (km_list ? w.km_cnt : w.nd_cnt)[d.n_ctg] = 0;
    }
};
struct CompactRegions {  // COMPACT_LANES threads per contig: copy its slices into the dense lists
    enum { COMPACT_LANES = 256 };
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t it, B&) const {
        const Dev& d = w.d;
        const int64_t k = it / COMPACT_LANES; const int32_t t = (int32_t)(it % COMPACT_LANES);
        int32_t gs = d.ctg_goff[k];
        int32_t r0 = w.rs_idx[gs];
        const int32_t* nd = w.nd_reg + 2 * (size_t)r0 + 2 * (size_t)k;
        const int32_t* km = w.km_reg + 2 * (size_t)r0 + 2 * (size_t)k;
        for (int32_t i = t; i < 2 * w.nd_cnt[k]; i += COMPACT_LANES) w.ndl[2 * w.nd_off[k] + i] = nd[i];
        for (int32_t i = t; i < 2 * w.km_cnt[k]; i += COMPACT_LANES) w.kml[2 * w.km_off[k] + i] = km[i];
    }
};
struct RegionDiff {      // mark the positions (start, end] of every region of both lists
    Dev2 w; int which;
    template <class B> NP_HD void operator()(int64_t i, B&) const {
        const int32_t* l = which == 0 ? w.ndl : w.kml;
        for (int32_t p = l[2 * i] + 1; p <= l[2 * i + 1]; p++) w.inreg[p] = 1;
    }
};
struct InsertLen2 {      // contig_create_insert_region (contig.c:182-245)
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t r, B& be) const {
        const Dev& d = w.d;
        if (d.r_level[r] < 1) return;
        Rec rc = load_rec(d.rec, d.rec_off, r);
        int32_t pos = d.r_gpos[r];
        for (int i = 0; i < rc.n_cigar; i++) {
            int op = cig_op(rc.cigar[i]); int32_t len = cig_len(rc.cigar[i]);
            if (op == OP_M || op == OP_D) pos += len;
            else if (op == OP_I && pos >= 0 && pos <= d.G && w.inreg[pos] > 0) be.atomic_max(&d.ins[pos - 1], len);
        }
    }
};
struct ColInit2 {        // live base / flag arrays of task 2 start as the draft's
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t c, B&) const {
        const Dev& d = w.d;
        d.obase[c] = d.refsym[c];
        d.oflag[c] = d.cflag[c];
        w.ndmark[c] = 0; w.vcap[c] = 0;
        if (c == 0) { w.ndmark[d.C] = 0; w.vcap[d.C] = 0; }
    }
};

// ---- candidate reads of a region / window -----------------------------------------------------
// reads of the contig in BAM order whose [gpos, hend) may overlap [a, b): index range [lo, hi)
NP_HD void overlap_range(const Dev2& w, int32_t k, int32_t a, int32_t b, int64_t* lo, int64_t* hi) {
    const Dev& d = w.d;
    int64_t r0 = d.ctg_read_off[k], r1 = d.ctg_read_off[k + 1];
    *hi = lower_bound_i32(d.r_gpos, r0, r1, b);             // first read with gpos >= b
    int64_t l = upper_bound_i32(w.r_hpm, 0, *hi, a);        // first read whose prefix-max hts end > a
    *lo = l < r0 ? r0 : l;
}

// ---- nodepth regions --------------------------------------------------------------------------
// No-depth regions that share an end column (next.start == this.end survives contig_merge_region)
// form a chain: the reference processes them in order and the shared column keeps the votes of
// both passes, so a chain is handled by ONE thread, sequentially.
NP_HD bool chain_head(const Dev2& w, int64_t i) { return i == 0 || w.ndl[2 * i - 1] < w.ndl[2 * i]; }
NP_HD bool chain_next(const Dev2& w, int64_t j) { return j + 1 < w.NR_nd && w.ndl[2 * (j + 1)] <= w.ndl[2 * j + 1]; }

// candidate reads of a region: level >= 1 and overlapping [s, e+1) (contig_parse_region, contig.c:692-698)
struct NdPairCount {     // per region: mark its columns, count its (region, read) pairs
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t i, B& be) const {
        const Dev& d = w.d;
        if (i >= w.NR_nd) { w.nd_pcnt[i] = 0; w.nd_sb[i] = 0; return; }
        int32_t s = w.ndl[2 * i], e = w.ndl[2 * i + 1];
        int32_t k = find_contig_i32(d.ctg_goff, d.n_ctg, s);
        int32_t c0 = d.colbase[s], c1 = d.colbase[e];
        for (int32_t c = c0; c <= c1; c++) { w.ndmark[c] = 1; be.atomic_add(&w.vcap[c], 1); }   // the draft's own vote
        int64_t lo, hi; overlap_range(w, k, s, e + 1, &lo, &hi);
        int32_t n = 0;
        for (int64_t r = lo; r < hi; r++) if (d.r_level[r] >= 1 && d.r_hend[r] > s) n++;
        w.nd_pcnt[i] = n; w.nd_sb[i] = n * (c1 - c0 + 1);
    }
};
struct NdPairFill {
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t i, B&) const {
        const Dev& d = w.d;
        int32_t s = w.ndl[2 * i], e = w.ndl[2 * i + 1];
        int32_t k = find_contig_i32(d.ctg_goff, d.n_ctg, s);
        int64_t lo, hi; overlap_range(w, k, s, e + 1, &lo, &hi);
        int32_t n = 0;
        for (int64_t r = lo; r < hi; r++) if (d.r_level[r] >= 1 && d.r_hend[r] > s) w.ndp_read[w.nd_poff[i] + n++] = (int32_t)r;
    }
};
struct SymSlotVisitor {  // the read's column string over the region, one byte per column
    uint8_t* slot; int32_t c0, ncols, first, n; int32_t* vcap; int32_t inc; int32_t* err;
    template <class B> NP_HD void put(int32_t col, uint32_t s, B& be) {
        int32_t i = col - c0;
        if (i < 0 || i >= ncols) { *err |= ERR_REGION_SCRATCH; return; }
        if (n == 0) first = i;
        if (i != first + n) { *err |= ERR_REGION_SCRATCH; return; }
        slot[i] = (uint8_t)s; n++;
        be.atomic_add(&vcap[col], inc);
    }
};
template <class B> struct SymSlotAdapter {
    SymSlotVisitor* v; B* be;
    NP_HD void sym(int32_t col, uint32_t s, int32_t, bool) { v->put(col, s, *be); }
    NP_HD void overflow() { *v->err |= ERR_INS_OVERFLOW; }
};
struct NdWalk {          // per (region, read) pair: CIGAR walk over the region
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t p, B& be) const {
        const Dev& d = w.d;
        int32_t i = (int32_t)upper_bound_i32(w.nd_poff, 0, w.NR_nd + 1, (int32_t)p) - 1;
        int32_t s = w.ndl[2 * i], e = w.ndl[2 * i + 1];
        int32_t k = find_contig_i32(d.ctg_goff, d.n_ctg, s);
        int32_t c0 = d.colbase[s], ncols = d.colbase[e] - c0 + 1;
        int32_t j = (int32_t)p - w.nd_poff[i];
        int64_t r = w.ndp_read[p];
        Rec rc = load_rec(d.rec, d.rec_off, r);
        // level-1 reads may vote twice on a column shared by two inner sub-regions
        SymSlotVisitor v{w.ndp_sym + (size_t)w.nd_soff[i] + (size_t)j * ncols, c0, ncols, 0, 0, w.vcap, d.r_level[r] == 1 ? 2 : 1, d.err};
        SymSlotAdapter<B> a{&v, &be};
        walk_read(rc, d.ctg_goff[k], s, e, d.r_qstart[r], d.r_qend[r], d.colbase, a);
        w.ndp_first[p] = v.first; w.ndp_n[p] = v.n;
    }
};

struct NdCtx {           // accessors of the per-region scratch
    const Dev2* w;
    NP_HD uint32_t* tab(int32_t c) const { return w->ktab2 + w->koff[c]; }
    NP_HD int32_t& nk(int32_t c) const { return w->nk2[w->ndidx[c]]; }
    NP_HD uint32_t& cnt(int32_t c) const { return w->cnt2[w->ndidx[c]]; }
    NP_HD void add(int32_t c, uint32_t kmer) const {                       // base.c:60-71
        uint32_t* t = tab(c); int32_t n = nk(c);
        int32_t j = 0;
        for (; j < n; j++) if ((t[j] & 0xffffu) == kmer) { t[j] += 1u << 16; break; }
        if (j == n) {
            if (n >= w->koff[c + 1] - w->koff[c]) { *w->d.err |= ERR_REGION_SCRATCH; return; }
            t[n] = kmer | (1u << 16); nk(c) = n + 1;
        }
        cnt(c)++;
    }
};
struct VoteVisitor {
    NdCtx x; uint32_t kmer;
    NP_HD void sym(int32_t col, uint32_t s, int32_t, bool) { kmer = ((kmer & 0xffu) << 4) | s; x.add(col, kmer); }
    NP_HD void overflow() { *x.w->d.err |= ERR_INS_OVERFLOW; }
};

// contig_region_score + contig_region_correct (contig.c:456-496) over columns [c0, c1]
NP_HD void nd_score_correct(const Dev2& w, int32_t c0, int32_t c1, double rate) {
    const Dev& d = w.d;
    // Forward pass with the previous and the current column's score lists in thread-local arrays (the
    // global copies are only what the backtrack needs: winning k-mer per base code, argmax base code).
    double sp[16], sc[16]; uint16_t kc[16]; uint8_t ord[16]; bool hp[16], hc[16];
    int pn = 0; double spmax = 0;
    bool zero_prev = true;
    for (int32_t c = c0; c <= c1; c++) {
        const int32_t ci = w.ndidx[c];
        const uint32_t* t = w.ktab2 + w.koff[c];
        const int32_t nk = w.nk2[ci];
        const uint32_t total = w.cnt2[ci], refk = w.refk2[ci];
        const uint32_t tot = total > 1 ? total - 1 : total;
        const double dec = (double)tot * rate;
        for (int b = 0; b < 16; b++) hc[b] = false;
        int no = 0;
        for (int32_t j = 0; j < nk; j++) {
            uint32_t k = t[j] & 0xffffu, cnt = t[j] >> 16, pv = (k >> 4) & 0xfu;
            double s = 0;
            if (!zero_prev) {
                if (pv == 0) s = spmax;
                else { if (!hp[pv]) *d.err |= ERR_MISSING_SCORE; s = sp[pv]; }
            }
            if (k == refk && total > 1) cnt--;
            s = s + ((double)cnt - dec);
            uint32_t b = k & 0xfu;
            if (!hc[b]) { hc[b] = true; sc[b] = s; kc[b] = (uint16_t)k; ord[no++] = (uint8_t)b; }
            else if (sc[b] < s) { sc[b] = s; kc[b] = (uint16_t)k; }
        }
        int am = ord[0]; double mx = sc[am];                                   // base_max_score (base.c:185-197)
        for (int q = 1; q < no; q++) if (sc[ord[q]] > mx) { mx = sc[ord[q]]; am = ord[q]; }
        uint16_t* gk = w.kc2 + (size_t)ci * 16;
        for (int q = 0; q < no; q++) gk[ord[q]] = kc[ord[q]];
        w.ns2[ci] = (uint8_t)am;                                               // argmax base code of this column
        for (int b = 0; b < 16; b++) { hp[b] = hc[b]; sp[b] = sc[b]; }
        pn = no; spmax = mx; zero_prev = false;
    }
    (void)pn;
    // backtrack (contig.c:473-496)
    uint32_t chosen = w.ns2[w.ndidx[c1]];
    for (int32_t c = c1;; c--) {
        const int32_t ci = w.ndidx[c];
        const uint32_t* t = w.ktab2 + w.koff[c]; const int32_t nk = w.nk2[ci];
        uint32_t total = w.cnt2[ci], support = 0;
        for (int32_t j = 0; j < nk; j++) if ((t[j] & 0xfu) == chosen) support += t[j] >> 16;
        uint8_t fl = d.oflag[c];
        if (total == 1) fl |= FLAG_ZERO; else fl &= (uint8_t)~FLAG_ZERO;
        if (support / (double)total < d.P.min_count_ratio_skip) fl |= FLAG_COVERAGE; else fl &= (uint8_t)~FLAG_COVERAGE;
        d.obase[c] = (uint8_t)chosen; d.oflag[c] = fl;
        if (c == c0) break;
        uint32_t k = w.kc2[(size_t)ci * 16 + chosen], pv = (k >> 4) & 0xfu;
        chosen = (pv == 0) ? w.ns2[w.ndidx[c - 1]] : pv;
    }
}

// contig_get_region over COLUMNS [c0,c1] of region [bs,be] with gap 0, con 0, brim_no_extension
// (the inner call of contig_score_correct, contig.c:722): returns pair count in out
NP_HD int32_t inner_regions(const Dev& d, int32_t c0, int32_t c1, int32_t bs, int32_t be, int32_t ext, int32_t* out) {
    int32_t n = 0, qstart = -1, qend = -1;
    int32_t c = c0;
    while (c <= c1) {
        int32_t i = d.colpos[c];
        if (d.oflag[c] & FLAG_ZERO) { if (qstart == -1) qstart = i; qend = i; }
        else if (qstart != -1) {
            // pgap = 1 > gap = 0; pcon >= 1 > con = 0: always emitted
            int32_t s = qstart >= bs + ext ? qstart - ext : bs, e = qend <= be - ext ? qend + ext : be;
            out[2 * n] = s; out[2 * n + 1] = e; n++;
            if (e > i) c = d.colbase[e];
            qstart = qend = -1;
        }
        c++;
    }
    if (qstart != -1) {
        int32_t s = qstart >= bs + ext ? qstart - ext : bs, e = qend <= be - ext ? qend + ext : be;
        out[2 * n] = s; out[2 * n + 1] = e; n++;
    }
    return n;
}

// Votes of the pairs of region i with filter level `level` on region columns [lo, hi] (relative to the
// region's first column c0), applied COLUMN by column: a column's list length, vote count and its first
// entries stay in registers while the pairs are visited in BAM order (same first-seen order as the
// reference's read-by-read walk, base.c:60-71).  The rolling 3-mer context starts at the pair's first vote
// inside the range (contig.c:255,360-363).
enum { ND_CACHE = 96 };   // pairs of a region whose metadata is cached in thread-local arrays
NP_HD void nd_apply_pairs(const Dev2& w, int32_t c0, int32_t ncols, int32_t i, int32_t level, int32_t lo, int32_t hi,
                          int32_t ss, int32_t se, int32_t ctx_lo = -1) {
    if (ctx_lo < 0) ctx_lo = lo;                   // column (relative) where the rolling 3-mer context starts
    const Dev& d = w.d;
    const int32_t p0 = w.nd_poff[i], p1 = w.nd_poff[i + 1];
    // pair metadata once per call (the column loop below would otherwise re-read it from HBM per column)
    int16_t pf[ND_CACHE], pe[ND_CACHE]; int32_t pidx[ND_CACHE]; int32_t np = 0;
    bool cached = true;
    for (int32_t p = p0; p < p1; p++) {
        const int64_t r = w.ndp_read[p];
        if (d.r_level[r] != level) continue;
        if (ss >= 0 && (d.r_hend[r] <= ss || d.r_gpos[r] >= se + 1)) continue;        // contig_parse_region's overlap test
        const int32_t first = w.ndp_first[p], n = w.ndp_n[p];
        if (n <= 0 || first > hi || first + n - 1 < lo) continue;
        if (np == ND_CACHE || first + n > 32767) { cached = false; break; }
        pf[np] = (int16_t)first; pe[np] = (int16_t)(first + n); pidx[np] = p - p0; np++;
    }
    for (int32_t t = lo; t <= hi; t++) {
        const int32_t c = c0 + t, ci = w.ndidx[c];
        uint32_t* tab = w.ktab2 + w.koff[c];
        const int32_t cap = w.koff[c + 1] - w.koff[c];
        int32_t nk = w.nk2[ci]; uint32_t cnt = w.cnt2[ci];
        const int32_t nloop = cached ? np : p1 - p0;
        for (int32_t q = 0; q < nloop; q++) {
            int32_t first, end, pj;
            if (cached) { first = pf[q]; end = pe[q]; pj = pidx[q]; }
            else {
                const int32_t p = p0 + q; const int64_t r = w.ndp_read[p];
                if (d.r_level[r] != level) continue;
                if (ss >= 0 && (d.r_hend[r] <= ss || d.r_gpos[r] >= se + 1)) continue;
                first = w.ndp_first[p]; end = first + w.ndp_n[p]; pj = q;
            }
            if (t < first || t >= end) continue;
            const uint8_t* slot = w.ndp_sym + (size_t)w.nd_soff[i] + (size_t)pj * ncols;
            const int32_t a = first > ctx_lo ? first : ctx_lo;                         // first vote inside the range
            uint32_t kmer = slot[t];
            if (t - 1 >= a) kmer |= (uint32_t)slot[t - 1] << 4;
            if (t - 2 >= a) kmer |= (uint32_t)slot[t - 2] << 8;
            int32_t j = 0;
            for (; j < nk; j++) if ((tab[j] & 0xffffu) == kmer) { tab[j] += 1u << 16; break; }
            if (j == nk) { if (nk >= cap) { *d.err |= ERR_REGION_SCRATCH; continue; } tab[nk++] = kmer | (1u << 16); }
            cnt++;
        }
        w.nk2[ci] = nk; w.cnt2[ci] = cnt;
    }
}

NP_HD bool nd_isolated(const Dev2& w, int64_t i) { return chain_head(w, i) && !chain_next(w, i); }

struct NdColFill {       // per region: reverse map of its columns (compact index -> region, column)
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t i, B&) const {
        const Dev& d = w.d;
        int32_t c0 = d.colbase[w.ndl[2 * i]], c1 = d.colbase[w.ndl[2 * i + 1]];
        for (int32_t c = c0; c <= c1; c++) { w.nd_reg_of[w.ndidx[c]] = (int32_t)i; w.nd_col_of[w.ndidx[c]] = c; }
    }
};
// First pass of an ISOLATED region (no shared end column), one thread per column: the draft's own vote
// (contig_as_read) and the level-2 votes.  Chained regions keep the sequential path in NodepthScore.
struct NdTallyCols {
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t ci, B&) const {
        const Dev& d = w.d;
        int32_t i = w.nd_reg_of[ci];
        if (!nd_isolated(w, i)) return;
        int32_t c = w.nd_col_of[ci];
        int32_t c0 = d.colbase[w.ndl[2 * i]], c1 = d.colbase[w.ndl[2 * i + 1]], ncols = c1 - c0 + 1, t = c - c0;
        uint32_t kmer = d.obase[c];
        if (t >= 1) kmer |= (uint32_t)d.obase[c - 1] << 4;
        if (t >= 2) kmer |= (uint32_t)d.obase[c - 2] << 8;
        w.refk2[ci] = (uint16_t)kmer;
        w.ktab2[w.koff[c]] = kmer | (1u << 16);
        w.nk2[ci] = 1; w.cnt2[ci] = 1;
        nd_apply_pairs(w, c0, ncols, i, 2, t, t, -1, -1, 0);
    }
};

struct NodepthScore {    // contig_score_correct(region, 0x12), one thread per chain of no-depth regions
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t i0, B&) const {
        const Dev& d = w.d;
        if (!chain_head(w, i0)) return;
        NdCtx x{&w};
        for (int64_t i = i0;; i++) {
            int32_t s = w.ndl[2 * i], e = w.ndl[2 * i + 1];
            int32_t c0 = d.colbase[s], c1 = d.colbase[e], ncols = c1 - c0 + 1;
            if (!nd_isolated(w, i)) {                                          // isolated regions: done by NdTallyCols
                uint32_t kmer = 0;
                for (int32_t c = c0; c <= c1; c++) {                           // contig_as_read
                    int32_t ci = w.ndidx[c];
                    if (!(i > i0 && c == c0)) { w.nk2[ci] = 0; w.cnt2[ci] = 0; }   // shared column keeps its votes
                    kmer = ((kmer & 0xffu) << 4) | d.obase[c];
                    w.refk2[ci] = (uint16_t)kmer;
                    x.add(c, kmer);
                }
                nd_apply_pairs(w, c0, ncols, (int32_t)i, 2, 0, ncols - 1, -1, -1);   // contig_parse_region, level == 2
            }
            nd_score_correct(w, c0, c1, d.P.rate);
            int32_t* sub = w.subbuf + (size_t)w.ndidx[c0] + 2 * (size_t)i;
            int32_t ns = inner_regions(d, c0, c1, s, e, d.P.ext_len_edge, sub);
            ns = merge_regions(sub, ns);
            for (int32_t q = 0; q < ns; q++) {
                int32_t ss = sub[2 * q], se = sub[2 * q + 1];
                nd_apply_pairs(w, c0, ncols, (int32_t)i, 1, d.colbase[ss] - c0, d.colbase[se] - c0, ss, se);   // level == 1 reads on top
                nd_score_correct(w, d.colbase[ss], d.colbase[se], d.P.rate);
            }
            if (!chain_next(w, i)) break;
        }
    }
};

// ---- windows ----------------------------------------------------------------------------------
// ss_spilt_region (kmercount.c:128-173) for one region; out == nullptr: count only
NP_HD int32_t split_region(const Dev& d, int32_t s, int32_t e, int32_t maxlen, int32_t* out) {
    int32_t n = 0, cur = s;
    if (e - s > maxlen) {
        int32_t qstart = -1, qend = -1;
        int32_t c = d.colbase[s], cend = d.colbase[e];
        while (c <= cend && !(d.oflag[c] & FLAG_ZERO)) c++;
        for (; c <= cend; c++) {
            int32_t j = d.colpos[c];
            if (!(d.oflag[c] & FLAG_ZERO)) { if (qstart == -1) qstart = j; qend = j; }
            else if (qstart != -1) {
                int32_t k = (qstart + qend) >> 1;
                if (out) { out[2 * n] = cur; out[2 * n + 1] = k; }
                n++; cur = k;
                qstart = qend = -1;
            }
        }
    }
    if (out) { out[2 * n] = cur; out[2 * n + 1] = e; }
    return n + 1;
}
struct SplitCount {
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t i, B&) const {
        if (i >= w.NR_km) { w.wcnt[i] = 0; return; }
        w.wcnt[i] = split_region(w.d, w.kml[2 * i], w.kml[2 * i + 1], w.d.P.max_len_kmer, nullptr);
    }
};
struct SplitFill {
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t i, B&) const {
        split_region(w.d, w.kml[2 * i], w.kml[2 * i + 1], w.d.P.max_len_kmer, w.win + 2 * (size_t)w.woff[i]);
    }
};

// candidates of window [s,e]: reads of the contig with gpos < s and hend > e+1 (swapped iterator)
NP_HD void window_range(const Dev2& w, int32_t k, int32_t s, int32_t e, int64_t* lo, int64_t* term) {
    const Dev& d = w.d;
    int64_t r0 = d.ctg_read_off[k], r1 = d.ctg_read_off[k + 1];
    *term = lower_bound_i32(d.r_gpos, r0, r1, s);
    int64_t l = upper_bound_i32(w.r_hpm, 0, *term, e + 1);
    *lo = l < r0 ? r0 : l;
}
// Scratch of one window: (ncand + 1) slots [string bytes | length, qual, first column, mapq] — the extra
// slot is for the stale record of kmercount.c:212-216 — followed by (ncand + 1) tally entries
// [slot, num, sum mapq, sum qual].
NP_HD int32_t win_slot_words(int32_t len) { return (len + 3) / 4 + 4; }

struct WindowCount {
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t i, B&) const {
        const Dev& d = w.d;
        if (i >= w.NW) { w.wcand[i] = 0; w.wp_cnt[i] = 0; return; }
        int32_t s = w.win[2 * i], e = w.win[2 * i + 1];
        int32_t k = find_contig_i32(d.ctg_goff, d.n_ctg, s);
        int64_t lo, term; window_range(w, k, s, e, &lo, &term);
        int32_t n = 0;
        for (int64_t r = lo; r < term; r++) if (d.r_hend[r] > e + 1) n++;
        int32_t len = d.colbase[e] - d.colbase[s] + 1;
        w.wp_cnt[i] = n + 1;
        w.wcand[i] = (n + 1) * (win_slot_words(len) + 4);
    }
};
struct WinPairFill {
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t i, B&) const {
        const Dev& d = w.d;
        int32_t s = w.win[2 * i], e = w.win[2 * i + 1];
        int32_t k = find_contig_i32(d.ctg_goff, d.n_ctg, s);
        int64_t lo, term; window_range(w, k, s, e, &lo, &term);
        int32_t n = 0;
        for (int64_t r = lo; r < term; r++) if (d.r_hend[r] > e + 1) w.wp_read[w.wp_off[i] + n++] = (int32_t)r;
        // the record that ended the swapped iterator (first read of the contig with pos >= start)
        w.wp_read[w.wp_off[i] + n] = (term < d.ctg_read_off[k + 1] && d.r_level[term] == 1) ? (int32_t)term : -1;
    }
};

struct KmerVisitor {     // ss_parse_read_kmer's appends (kmercount.c:389-440); flag clearing is deferred
    uint8_t* region; int32_t cap; int32_t length, del, qual, first; const uint8_t* q;
    NP_HD void sym(int32_t col, uint32_t s, int32_t qpos, bool subgap) {
        if (length == 0) first = col;
        if (length < cap) region[length] = (uint8_t)s;
        length++;
        if (subgap) del++;
        if (qpos >= 0 && q) qual += q[qpos];
    }
    int32_t* err;
    NP_HD void overflow() { *err |= ERR_INS_OVERFLOW; }
};
struct WinWalk {         // per (window, read) pair: the read's column string over the window
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t p, B&) const {
        const Dev& d = w.d;
        int32_t i = (int32_t)upper_bound_i32(w.wp_off, 0, w.NW + 1, (int32_t)p) - 1;
        int32_t j = (int32_t)p - w.wp_off[i], ncand = w.wp_cnt[i] - 1;
        int32_t s = w.win[2 * i], e = w.win[2 * i + 1];
        int32_t len = d.colbase[e] - d.colbase[s] + 1, sw = win_slot_words(len);
        int32_t* slot = w.wscratch + w.wsoff[i] + (size_t)j * sw;
        int32_t* meta = slot + (len + 3) / 4;
        int64_t r = w.wp_read[p];
        meta[0] = -1;
        if (r < 0) return;
        if (j < ncand && d.r_level[r] != 2) return;            // only level-2 candidates are parsed (kmercount.c:197)
        for (int32_t t = 0; t < (len + 3) / 4; t++) slot[t] = 0;   // strings are compared word-wise
        int32_t k = find_contig_i32(d.ctg_goff, d.n_ctg, s);
        Rec rc = load_rec(d.rec, d.rec_off, r);
        const uint8_t* qp = d.qual + (size_t)d.qual_off[r] * 16;
        if (d.qual_off[r + 1] == d.qual_off[r] && rc.l_qseq > 0) {
            // sparse quality stream: a spanning candidate always overlaps the window's lowercase column and
            // therefore has qualities; the stale record (slot ncand) may not — it can then never be full
            // length, so its quality sum is irrelevant
            if (j < ncand) { *d.err |= ERR_QUAL_MISSING; return; }
            qp = nullptr;
        }
        KmerVisitor v{(uint8_t*)slot, len, 0, 0, 0, 0, qp, d.err};
        walk_read(rc, d.ctg_goff[k], s, e, d.r_qstart[r], d.r_qend[r], d.colbase, v);
        meta[0] = v.length;
        meta[1] = (v.length > 0 && v.length != v.del) ? v.qual / (v.length - v.del) : 0;   // kmercount.c:457-462
        meta[2] = v.first;
        meta[3] = (int32_t)rc.mapq;
    }
};

struct WindowVote {      // ss_kmer_correct for one window (kmercount.c:188-253) over the pre-walked pairs
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t i, B&) const {
        const Dev& d = w.d;
        int32_t s = w.win[2 * i], e = w.win[2 * i + 1];
        int32_t len = d.colbase[e] - d.colbase[s] + 1, sw = win_slot_words(len), lw = (len + 3) / 4;
        int32_t ncand = w.wp_cnt[i] - 1;
        int32_t* base = w.wscratch + w.wsoff[i];
        int32_t* tal = base + (size_t)(ncand + 1) * sw;             // [slot, num, mapq, qual] entries
        int32_t nstr = 0, count = 0;
        bool broke = false;
        // every parsed read clears FLAG_ZERO on the columns it covers (kmercount.c:398,411,431,437); the
        // ranges of successive reads nearly coincide, so only the part outside the interval cleared so
        // far is touched (a disjoint range is cleared on its own)
        int32_t clr_lo = 0, clr_hi = 0;
        auto clear_range = [&](int32_t a, int32_t b) { for (int32_t c = a; c < b; c++) d.oflag[c] &= (uint8_t)~FLAG_ZERO; };
        // ss_kmer_get_region (kmercount.c:332-363) on pair slot j; returns ks->mapqual after the call
        auto get_region = [&](int32_t j) -> int32_t {
            const int32_t* slot = base + (size_t)j * sw;
            const int32_t* meta = slot + lw;
            int32_t length = meta[0];
            if (length > 0 && !w.flagzero) {
                int32_t a = meta[2], b = meta[2] + length;
                if (clr_hi == clr_lo) { clear_range(a, b); clr_lo = a; clr_hi = b; }
                else if (b < clr_lo || a > clr_hi) clear_range(a, b);
                else {
                    if (a < clr_lo) { clear_range(a, clr_lo); clr_lo = a; }
                    if (b > clr_hi) { clear_range(clr_hi, b); clr_hi = b; }
                }
            }
            if (length != len) return 0;
            int32_t q = 0;
            for (; q < nstr; q++) {
                const int32_t* a = base + (size_t)tal[4 * q] * sw; bool same = true;
                for (int32_t t = 0; t < lw; t++) if (a[t] != slot[t]) { same = false; break; }   // slots are zero padded
                if (same) break;
            }
            if (q == nstr) { tal[4 * q] = j; tal[4 * q + 1] = 1; tal[4 * q + 2] = meta[3]; tal[4 * q + 3] = meta[1]; nstr++; }
            else { tal[4 * q + 1]++; tal[4 * q + 2] += meta[3]; tal[4 * q + 3] += meta[1]; }
            return meta[3];
        };
        for (int32_t j = 0; j < ncand; j++) {
            if (d.r_level[w.wp_read[w.wp_off[i] + j]] != 2) continue;
            int32_t mq = get_region(j);
            if (mq == 60) { count++; if (count >= d.P.max_count_kmer) { broke = true; break; } }
        }
        if (nstr == 0 && !broke && ncand > 0 && w.wp_read[w.wp_off[i] + ncand] >= 0) {
            // kmercount.c:209-219: the stale record is filtered and parsed once per record the second
            // iterator yields
            for (int32_t t = 0; t < ncand; t++) get_region(ncand);
        }
        int32_t best = -1;
        if (nstr > 0) {
            if (w.flagzero) clear_range(d.colbase[s], d.colbase[e] + 1);      // contig_clean_flag(start, end, FLAG_ZERO_N), kmercount.c:222-224
            if (count == d.P.max_count_kmer) {
                int32_t want = 60 * count;
                for (int32_t q = 0; q < nstr; q++) if (tal[4 * q + 2] == want) { best = q; break; }
            }
            if (best < 0) {
                best = 0;
                for (int32_t q = 0; q < nstr; q++) {
                    const int32_t *a = tal + 4 * best + 1, *b = tal + 4 * q + 1;
                    bool less = a[0] != b[0] ? a[0] < b[0] : a[1] != b[1] ? a[1] < b[1] : a[2] < b[2];   // ks_compare
                    if (q != best && less) best = q;
                }
            }
        }
        w.wbest[i] = best < 0 ? -1 : (int32_t)(w.wsoff[i] + tal[4 * best] * sw);
    }
};
struct WindowApply {     // contig_update_contig in window order: a later window wins shared columns
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t i, B&) const {
        const Dev& d = w.d;
        if (w.wbest[i] < 0) return;
        int32_t s = w.win[2 * i], e = w.win[2 * i + 1];
        const uint8_t* str = (const uint8_t*)(w.wscratch + w.wbest[i]);
        int32_t c0 = d.colbase[s], c1 = d.colbase[e];
        bool next_shares = i + 1 < w.NW && w.win[2 * (i + 1)] == e && w.wbest[i + 1] >= 0;
        for (int32_t c = c0; c <= c1; c++) {
            if (c == c1 && next_shares) break;
            d.obase[c] = str[c - c0];
        }
    }
};

// ---- snp_valid's second pass: a window that found no string is cut again (fts_spilt_region, snpvalid.c:37-66) -------
// cut points: the middle of every unflagged stretch that a flagged column follows (twice, or once for the stretch the
// window starts with), then the window's end; ss_kmer_correct reads them as (start, end) pairs.  out == nullptr: count.
// A list that is odd, or has start > end (a window that starts on a flagged column / has no flagged column), makes the
// reference read past its list: undefined there; here the dangling point and inverted pairs are dropped and *warn is set.
NP_HD int32_t fts_split(const Dev& d, int32_t start, int32_t end, int32_t* out, int32_t* warn) {
    int32_t np = 0, npair = 0, pend = 0;
    int32_t qstart = -1, qend = -1;
    auto point = [&](int32_t v) {
        if (np & 1) {
            if (pend <= v) { if (out) { out[2 * npair] = pend; out[2 * npair + 1] = v; } npair++; }
            else *warn = 1;
        } else pend = v;
        np++;
    };
    for (int32_t c = d.colbase[start]; c <= d.colbase[end]; c++) {
        const int32_t i = d.colpos[c];
        if (!(d.oflag[c] & FLAG_ZERO)) { if (qstart == -1) qstart = i; qend = i; }
        else if (qstart != -1) {
            int count = 2;
            if (qstart == start) { qend = start; count--; }
            int32_t mid = (qstart + qend) / 2;
            for (int k = 0; k < count; k++) { point(mid); if (qstart != qend) mid++; }
            qstart = qend = -1;
        }
    }
    point(end);
    if (np & 1) *warn = 1;
    return npair;
}
struct WinFailFlag {     // per first-pass window
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t i, B&) const { w.wfail[i] = (i < w.NW && w.wbest[i] < 0) ? 1 : 0; }
};
struct FtsCount {        // per first-pass window (failed ones count their sub-windows)
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t i, B&) const {
        int32_t n = 0;
        if (i < w.NW1 && w.wfail[i]) n = fts_split(w.d, w.win1[2 * i], w.win1[2 * i + 1], nullptr, w.warn);
        w.fcnt[i] = n;
    }
};
struct FtsFill {
    Dev2 w;
    template <class B> NP_HD void operator()(int64_t i, B&) const {
        int32_t dummy = 0;
        if (w.wfail[i] && w.fcnt[i] > 0) fts_split(w.d, w.win1[2 * i], w.win1[2 * i + 1], w.win + 2 * (size_t)w.foff[i], &dummy);
    }
};

// the vote over the windows w.win[0 .. NW) (kmercount.c:175-261): candidates, column strings, tally, winner, apply
template <class BE>
void run_window_votes(BE& be, Dev2& w) {
    if (w.NW <= 0) return;
    w.wcand = be.template buf<int32_t>("wcand", (size_t)w.NW + 1);
    w.wsoff = be.template buf<int32_t>("wsoff", (size_t)w.NW + 1);
    w.wbest = be.template buf<int32_t>("wbest", (size_t)w.NW + 1);
    w.wp_cnt = be.template buf<int32_t>("wp_cnt", (size_t)w.NW + 1);
    w.wp_off = be.template buf<int32_t>("wp_off", (size_t)w.NW + 1);
    be.launch("window_count", (int64_t)w.NW + 1, WindowCount{w});
    be.exscan2_i32(w.wcand, w.wsoff, w.wp_cnt, w.wp_off, (int64_t)w.NW + 1);
    int32_t WS = 0;
    { const int32_t* ptrs[2] = {w.wsoff + w.NW, w.wp_off + w.NW}; int32_t v[2]; be.read_many(ptrs, 2, v); WS = v[0]; w.NP_w = v[1]; }
    w.wscratch = be.template buf<int32_t>("wscratch", (size_t)WS + 4);
    w.wp_read = be.template buf<int32_t>("wp_read", (size_t)w.NP_w + 1);
    be.launch("window_fill", w.NW, WinPairFill{w});
    be.launch("window_walk", w.NP_w, WinWalk{w});
    be.launch("window_vote", w.NW, WindowVote{w});
    be.launch("window_apply", w.NW, WindowApply{w});
}

// ---- orchestration ------------------------------------------------------------------------------
// mode 2: kmer_count (kmercount.c:93-126); mode 4: snp_valid (snpvalid.c:3-35: the k-mer regions only, a first vote that
// leaves FLAG_ZERO to whole windows, a second vote over the re-cut windows that found no string, no lowercase on output)
template <class BE>
int run_kmer_count(BE& be, Dev& d0, RunStats* st, int mode = 2) {
    Dev2 w; memset(&w, 0, sizeof(w));
    Dev& d = w.d; d = d0;
    const int64_t R = d.n_reads; const int32_t G = d.G;
    d.task = 2;
    d.err = be.template buf<int32_t>("err", 1);
    be.zero(d.err, sizeof(int32_t));
    d.r_ctg = be.template buf<int32_t>("r_ctg", R + 1);
    d.r_gpos = be.template buf<int32_t>("r_gpos", R + 1);
    d.r_qstart = be.template buf<int32_t>("r_qstart", R + 1);
    d.r_qend = be.template buf<int32_t>("r_qend", R + 1);
    d.r_wend = be.template buf<int32_t>("r_wend", R + 1);
    d.r_hend = be.template buf<int32_t>("r_hend", R + 1);
    d.r_pm = be.template buf<int32_t>("r_pm", R + 1);
    w.r_hpm = be.template buf<int32_t>("r_hpm", R + 1);
    d.r_level = be.template buf<uint8_t>("r_level", R + 1);
    d.ins = be.template buf<int32_t>("ins", (size_t)G + 1);
    d.colbase = be.template buf<int32_t>("colbase", (size_t)G + 1);
    d.out_off = be.template buf<int64_t>("out_off", (size_t)d.n_ctg + 1);
    be.zero(d.ins, sizeof(int32_t) * ((size_t)G + 1));
    if (R > 0) {
        be.launch("read_prep", R, ReadPrep{d});
        be.inclmax_i32(d.r_hend, w.r_hpm, R);
    }
    // lowercase runs and region lists
    w.rs_flag = be.template buf<int32_t>("rs_flag", (size_t)G + 2);
    w.re_flag = be.template buf<int32_t>("re_flag", (size_t)G + 2);
    w.rs_idx = be.template buf<int32_t>("rs_idx", (size_t)G + 2);
    be.launch("run_flags", (int64_t)G + 1, RunFlags{w});
    be.exscan_i32(w.rs_flag, w.rs_idx, (int64_t)G + 1);
    w.n_runs = be.read_i32(w.rs_idx + G);
    w.run_s = be.template buf<int32_t>("run_s", (size_t)w.n_runs + 1);
    w.run_e = be.template buf<int32_t>("run_e", (size_t)w.n_runs + 1);
    if (G > 0) be.launch("run_fill", G, RunFill{w});
    size_t regcap = 2 * ((size_t)w.n_runs + (size_t)d.n_ctg + 2);
    w.nd_reg = be.template buf<int32_t>("nd_reg", regcap);
    w.km_reg = be.template buf<int32_t>("km_reg", regcap);
    w.nd_cnt = be.template buf<int32_t>("nd_cnt", (size_t)d.n_ctg + 1);
    w.km_cnt = be.template buf<int32_t>("km_cnt", (size_t)d.n_ctg + 1);
    w.nd_off = be.template buf<int32_t>("nd_off", (size_t)d.n_ctg + 1);
    w.km_off = be.template buf<int32_t>("km_off", (size_t)d.n_ctg + 1);
    w.gnd = be.template buf<int32_t>("gnd", regcap);
    w.gkm = be.template buf<int32_t>("gkm", regcap);
    w.gcnt_nd = be.template buf<int32_t>("gcnt_nd", (size_t)w.n_runs + 1);
    w.gcnt_km = be.template buf<int32_t>("gcnt_km", (size_t)w.n_runs + 1);
    const size_t N1 = (size_t)w.n_runs + 1, nslot = 2 * N1;         // slots (+ the scans' total entry) of both lists
    w.gvalid = be.template buf<int32_t>("gvalid", nslot);
    w.gpos = be.template buf<int32_t>("gpos", nslot);
    w.ghead = be.template buf<int32_t>("ghead", nslot);
    w.gout = be.template buf<int32_t>("gout", nslot);
    w.dreg = be.template buf<int32_t>("dreg", 2 * nslot);
    w.dctg = be.template buf<int32_t>("dctg", nslot);
    w.dhead = be.template buf<uint8_t>("dhead", nslot);
    w.reg_bad = be.template buf<int32_t>("reg_bad", 1);
    // the region lists never outgrow the run list: one region per run at most
    w.ndl = be.template buf<int32_t>("ndl", 2 * (size_t)w.n_runs + 2 * (size_t)d.n_ctg + 4);
    w.kml = be.template buf<int32_t>("kml", 2 * (size_t)w.n_runs + 2 * (size_t)d.n_ctg + 4);
    int32_t bad = NP_FORCE_REGION_WALK;
    w.NR_nd = 0; w.NR_km = 0;
    if (w.n_runs > 0) {
        be.launch("region_groups", 2 * (int64_t)w.n_runs, RegionGroups{w});
        be.exscan2_i32(w.gvalid, w.gpos, w.gvalid + N1, w.gpos + N1, (int64_t)N1);
        be.launch("region_dense", 2 * (int64_t)w.n_runs, RegionDense{w});
        be.launch("region_heads", 2 * (int64_t)N1, RegionHeads{w});
        be.exscan2_i32(w.ghead, w.gout, w.ghead + N1, w.gout + N1, (int64_t)N1);
        be.launch("region_write", 2 * (int64_t)w.n_runs, RegionWrite{w});
        const int32_t* ptrs[3] = {w.gout + w.n_runs, w.gout + N1 + w.n_runs, w.reg_bad}; int32_t v[3];
        be.read_many(ptrs, 3, v); w.NR_nd = v[0]; w.NR_km = v[1]; bad |= v[2];
    }
    if (bad && w.n_runs > 0) {                                      // the literal per-contig walk
        be.launch("contig_regions", 2 * (int64_t)d.n_ctg, ContigRegions{w});
        be.exscan2_i32(w.nd_cnt, w.nd_off, w.km_cnt, w.km_off, (int64_t)d.n_ctg + 1);
        { const int32_t* ptrs[2] = {w.nd_off + d.n_ctg, w.km_off + d.n_ctg}; int32_t v[2]; be.read_many(ptrs, 2, v); w.NR_nd = v[0]; w.NR_km = v[1]; }
        be.launch("compact_regions", (int64_t)d.n_ctg * CompactRegions::COMPACT_LANES, CompactRegions{w});
    }
    // insertion columns inside regions only
    w.inreg = be.template buf<uint8_t>("inreg", (size_t)G + 2);
    be.zero(w.inreg, (size_t)G + 2);
    if (mode == 4) w.NR_nd = 0;                                  // snp_valid has no low-depth re-scoring
    if (w.NR_nd > 0) be.launch("region_diff_nd", w.NR_nd, RegionDiff{w, 0});
    if (w.NR_km > 0) be.launch("region_diff_km", w.NR_km, RegionDiff{w, 1});
    if (R > 0) be.launch("insert_len2", R, InsertLen2{w});
    be.exscan_ncol(d.ins, d.colbase, (int64_t)G);
    d.C = be.read_i32(d.colbase + G);
    const int32_t C = d.C;
    d.refsym = be.template buf<uint8_t>("refsym", (size_t)C + 1);
    d.cflag = be.template buf<uint8_t>("cflag", (size_t)C + 1);
    d.obase = be.template buf<uint8_t>("obase", (size_t)C + 1);
    d.oflag = be.template buf<uint8_t>("oflag", (size_t)C + 1);
    d.colpos = be.template buf<int32_t>("colpos", (size_t)C + 1);
    d.keepidx = be.template buf<int32_t>("keepidx", (size_t)C + 1);
    w.ndmark = be.template buf<int32_t>("ndmark", (size_t)C + 2);
    w.ndidx = be.template buf<int32_t>("ndidx", (size_t)C + 2);
    w.vcap = be.template buf<int32_t>("vcap", (size_t)C + 2);
    w.koff = be.template buf<int32_t>("koff", (size_t)C + 2);
    if (G > 0) {
        be.launch("col_init", G, ColInit{d});
        be.launch("col_ends", d.n_ctg, ColEnds{d});
    }
    if (C > 0) be.launch("col_init2", C, ColInit2{w});
    // no-depth regions
    if (w.NR_nd > 0) {
        w.nd_pcnt = be.template buf<int32_t>("nd_pcnt", (size_t)w.NR_nd + 1);
        w.nd_poff = be.template buf<int32_t>("nd_poff", (size_t)w.NR_nd + 1);
        w.nd_sb = be.template buf<int32_t>("nd_sb", (size_t)w.NR_nd + 1);
        w.nd_soff = be.template buf<int32_t>("nd_soff", (size_t)w.NR_nd + 1);
        be.launch("nodepth_pairs", (int64_t)w.NR_nd + 1, NdPairCount{w});
        be.exscan2_i32(w.nd_pcnt, w.nd_poff, w.nd_sb, w.nd_soff, (int64_t)w.NR_nd + 1);
        int32_t SB = 0;
        { const int32_t* ptrs[2] = {w.nd_poff + w.NR_nd, w.nd_soff + w.NR_nd}; int32_t v[2]; be.read_many(ptrs, 2, v); w.NP_nd = v[0]; SB = v[1]; }
        w.ndp_read = be.template buf<int32_t>("ndp_read", (size_t)w.NP_nd + 1);
        w.ndp_first = be.template buf<int32_t>("ndp_first", (size_t)w.NP_nd + 1);
        w.ndp_n = be.template buf<int32_t>("ndp_n", (size_t)w.NP_nd + 1);
        w.ndp_sym = be.template buf<uint8_t>("ndp_sym", (size_t)SB + 16);
        be.launch("nodepth_fill", w.NR_nd, NdPairFill{w});
        if (w.NP_nd > 0) be.launch("nodepth_walk", w.NP_nd, NdWalk{w});
        be.exscan_i32(w.ndmark, w.ndidx, (int64_t)C + 1);
        be.exscan_i32(w.vcap, w.koff, (int64_t)C + 1);
        int32_t KE = 0, e1 = 0;
        { const int32_t* ptrs[3] = {w.ndidx + C, w.koff + C, d.err}; int32_t v[3]; be.read_many(ptrs, 3, v); w.NRC = v[0]; KE = v[1]; e1 = v[2]; }
        size_t nrc = (size_t)w.NRC + 1;
        w.ktab2 = be.template buf<uint32_t>("ktab2", (size_t)KE + 1);
        w.nk2 = be.template buf<int32_t>("nk2", nrc);
        w.cnt2 = be.template buf<uint32_t>("cnt2", nrc);
        w.refk2 = be.template buf<uint16_t>("refk2", nrc);
        w.sc2 = be.template buf<double>("sc2", nrc * 16);
        w.kc2 = be.template buf<uint16_t>("kc2", nrc * 16);
        w.ord2 = be.template buf<uint8_t>("ord2", nrc * 16);
        w.ns2 = be.template buf<uint8_t>("ns2", nrc);
        w.subbuf = be.template buf<int32_t>("subbuf", nrc + 2 * (size_t)w.NR_nd + 4);
        if (e1) return e1;
        w.nd_reg_of = be.template buf<int32_t>("nd_reg_of", nrc);
        w.nd_col_of = be.template buf<int32_t>("nd_col_of", nrc);
        be.launch("nodepth_colfill", w.NR_nd, NdColFill{w});
        if (w.NRC > 0) be.launch("nodepth_tally", w.NRC, NdTallyCols{w});
        be.launch("nodepth_score", w.NR_nd, NodepthScore{w});
    }
    // windows
    w.NW = 0;
    if (w.NR_km > 0) {
        w.wcnt = be.template buf<int32_t>("wcnt", (size_t)w.NR_km + 1);
        w.woff = be.template buf<int32_t>("woff", (size_t)w.NR_km + 1);
        be.launch("split_count", (int64_t)w.NR_km + 1, SplitCount{w});
        be.exscan_i32(w.wcnt, w.woff, (int64_t)w.NR_km + 1);
        w.NW = be.read_i32(w.woff + w.NR_km);
        w.win = be.template buf<int32_t>("win", 2 * (size_t)w.NW + 2);
        be.launch("split_fill", w.NR_km, SplitFill{w});
        w.flagzero = mode == 4 ? 1 : 0;
        run_window_votes(be, w);
        if (mode == 4 && w.NW > 0) {
            // second pass over the windows that found no string
            const int32_t NW1 = w.NW;
            w.NW1 = NW1; w.win1 = w.win;
            w.wfail = be.template buf<int32_t>("wfail", (size_t)NW1 + 1);
            w.fcnt = be.template buf<int32_t>("fcnt", (size_t)NW1 + 1);
            w.foff = be.template buf<int32_t>("foff", (size_t)NW1 + 1);
            w.warn = be.template buf<int32_t>("snp_warn", 1);
            be.zero(w.warn, sizeof(int32_t));
            be.launch("win_fail", (int64_t)NW1 + 1, WinFailFlag{w});
            be.launch("fts_count", (int64_t)NW1 + 1, FtsCount{w});
            be.exscan_i32(w.fcnt, w.foff, (int64_t)NW1 + 1);
            const int32_t NW2 = be.read_i32(w.foff + NW1);
            if (NW2 > 0) {
                w.win = be.template buf<int32_t>("win2", 2 * (size_t)NW2 + 2);
                be.launch("fts_fill", NW1, FtsFill{w});
                w.NW = NW2; w.flagzero = 0;
                run_window_votes(be, w);
            }
        }
    }
    be.exscan_keep(d.obase, d.keepidx, (int64_t)C);
    int32_t total = 0, err = 0;
    { const int32_t* ptrs[2] = {d.keepidx + C, d.err}; int32_t v[2]; be.read_many(ptrs, 2, v); total = v[0]; err = v[1]; }
    d.out = be.template buf<uint8_t>("out", (size_t)total + 1);
    if (C > 0) be.launch("emit", C, Emit{d, (uint8_t)(mode == 4 ? 0 : FLAG_ZERO)});
    be.launch("out_offsets", (int64_t)d.n_ctg + 1, OutOffsets{d});
    run_trace(be, d);
    if (st) { st->C = C; st->T = w.NW; st->sym_words = w.NR_nd; st->table_entries = w.NR_km; st->out_bytes = total; }
    d0 = d;
    return err;
}

}  // namespace npe
