/*
 * np_oracle.c — TEST INFRASTRUCTURE ONLY (see np_oracle.h).
 *
 * Plain-C restatement of NextPolish's short-read hot path on packed shards: per-column k-mer
 * vote lists, the full sequential score-chain DP over each contig, backtrack, flags and
 * emission (task 1), plus region finding, low-depth re-scoring and the spanning-read window
 * vote (task 2).  It follows the reference's data flow literally (one record per column with
 * a first-seen-ordered k-mer list and a per-base score list) so that it is structurally
 * independent of the GPU engine's anchor/stretch decomposition.
 *
 * Every function cites the reference lines (source/lib/...) it restates.  BAM iteration is
 * replaced by the predicate the htslib iterators implement (records of the contig in file
 * order with pos < end+1 && endpos > start; swapped iterator: pos < start && endpos > end+1,
 * contig.c:1010-1043,1130-1135 + htslib/hts.c:2613-2653).
 *
 * Parity pinning: the reference has no golden vectors for this path; this file is pinned
 * against the reference compiled from source (oracle/_ref) by tests/test_oracle_vs_ref.py
 * and the committed fixtures under tests/golden/.
 *
 * Documented deviations (all in undefined-behaviour territory of the reference):
 *  - records with n_cigar == 0 are dropped by the packer (the reference reads cigar[0] of such
 *    records uninitialised in contig_read_cliprate);
 *  - contig_cut_read's homopolymer loops are bounded to the read (the reference runs into
 *    neighbouring record bytes; both give "no usable interval");
 *  - contig_brim_with_extension reads data[L] / data[-1] at the contig ends; treated as
 *    "different base, not flagged";
 *  - the stale-record fallback of kmercount.c:212-216 is reproduced only when the record that
 *    terminated the first iterator belongs to the same contig.
 */
#include "np_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define BASE_DEL 3
#define FLAG_ZERO 1
#define FLAG_COVERAGE 2
#define MAX_MAPQ 60
#define BAM_CMATCH 0
#define BAM_CINS 1
#define BAM_CDEL 2
#define BAM_CSOFT_CLIP 4
#define BAM_CHARD_CLIP 5

/* base.c:5-15 */
static const char basetostr_[] = "=ACMGRSVTWYHKDBN";
static uint8_t strtobase_(uint8_t c)
{
    switch (c) {
    case '=': return 0;  case 'A': return 1;  case 'C': return 2;  case 'M': return 3;
    case 'G': return 4;  case 'R': return 5;  case 'S': return 6;  case 'V': return 7;
    case 'T': return 8;  case 'W': return 9;  case 'Y': return 10; case 'H': return 11;
    case 'K': return 12; case 'D': return 13; case 'B': return 14; default: return 15;
    }
}

typedef struct { uint16_t kmer, count; } KC;                 /* base.h:28-31 */
typedef struct { uint8_t base; uint16_t kmer; double score; } SC;   /* base.h:33-37 */
typedef struct {                                             /* base.h:40-48 */
    uint8_t base, flag;
    uint16_t refkmer, count;
    KC* k; int nk, capk;
    SC s[16]; int ns;
} Col;

typedef struct {
    int32_t pos; uint16_t flag; uint8_t mapq, enc; int32_t isize; int32_t l_qseq; int32_t n_cigar;
    const uint32_t* cigar; const uint8_t* seq; const uint8_t* qual;
} Rd;

typedef struct {
    const np_shard_view* v; const Configure* cfg;
    int32_t L; const uint8_t* seq; int64_t r0, r1;
    uint8_t *base0, *flag0;       /* per reference position, before column layout */
    int32_t *ins, *colbase, *colpos;
    int32_t C; Col* col;
    int filter_kind;              /* 1: contig_read_fliter1, 0: contig_read_fliter */
    int32_t max_rlen;
    PolishPoint* pts; int64_t pts_cap, npts;   /* optional change trace (contig.c:743-799) */
} Ctg;

typedef struct { int32_t* d; int n, cap; } IList;
static void il_push(IList* l, int32_t x)
{
    if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 64; l->d = realloc(l->d, sizeof(int32_t) * (size_t)l->cap); }
    l->d[l->n++] = x;
}

static inline uint8_t seqi(const uint8_t* s, int32_t i) { return (s[i >> 1] >> ((~i & 1) << 2)) & 0xf; }
/* packed-shard base i as an nt16 code: enc 0 = BAM 4-bit, enc 1 = 2 bits per base (include/nextpolish_b200.h) */
#define rdseq(rd, i) ((rd)->enc ? (uint8_t)(1u << (((rd)->seq[(i) >> 2] >> (6 - 2 * ((i) & 3))) & 3u)) : seqi((rd)->seq, (i)))
static inline int cig_op(uint32_t c) { return (int)(c & 0xf); }
static inline int32_t cig_len(uint32_t c) { return (int32_t)(c >> 4); }

static void get_read(const np_shard_view* v, int64_t r, Rd* rd)
{
    const uint8_t* p = v->rec + (size_t)v->rec_off[r] * 16;
    uint16_t u16;
    memcpy(&rd->pos, p, 4);
    memcpy(&rd->flag, p + 4, 2);
    rd->mapq = p[6];
    rd->enc = p[7];
    memcpy(&rd->isize, p + 8, 4);
    memcpy(&u16, p + 12, 2); rd->l_qseq = u16;
    memcpy(&u16, p + 14, 2); rd->n_cigar = u16;
    rd->cigar = (const uint32_t*)(p + 16);
    rd->seq = p + 16 + 4 * (size_t)rd->n_cigar;
    rd->qual = v->qual ? v->qual + (size_t)v->qual_off[r] * 16 : NULL;
}

/* htslib/sam.c:391-397 (bam_endpos), reference-consuming ops M,D,N,=,X */
static int32_t hts_endpos(const Rd* rd)
{
    int32_t rl = 0;
    if ((rd->flag & 4) || rd->n_cigar == 0) return rd->pos + 1;
    for (int i = 0; i < rd->n_cigar; i++) {
        int op = cig_op(rd->cigar[i]);
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += cig_len(rd->cigar[i]);
    }
    return rd->pos + rl;
}

/* contig.c:632-646 */
static double read_cliprate(const Rd* rd)
{
    int32_t addlen = 0;
    if (cig_op(rd->cigar[0]) == BAM_CSOFT_CLIP) addlen += cig_len(rd->cigar[0]);
    if (cig_op(rd->cigar[rd->n_cigar - 1]) == BAM_CSOFT_CLIP) addlen += cig_len(rd->cigar[rd->n_cigar - 1]);
    return rd->l_qseq > 0 ? addlen / (double)rd->l_qseq : 0;
}

/* contig.c:648-665 (contig_read_fliter) and contig.c:667-677 (contig_read_fliter1) */
static int read_filter(const Ctg* g, const Rd* rd)
{
    int result = 0;
    if (g->filter_kind == 1) return (rd->flag & 0xC04) == 0 ? 1 : 0;
    if ((rd->flag & 0xC04) == 0) {
        int32_t length = rd->isize >= 0 ? rd->isize : -rd->isize;
        double cliprate = read_cliprate(rd);
        if ((length > 0 && length < g->cfg->read_tlen) || cliprate < g->cfg->max_clip_ratio_sgs) {
            result = 1;
            if (rd->mapq >= g->cfg->min_map_quality && (cliprate < g->cfg->max_clip_ratio_sgs + 0.05)) result = 2;
        }
    }
    return result;
}

/* contig.c:333-358 */
static void cut_read(const Ctg* g, const Rd* rd, int32_t* qstart, int32_t* qend)
{
    int32_t addlen = 0, trim = g->cfg->trim_len_edge;
    if (cig_op(rd->cigar[0]) == BAM_CSOFT_CLIP) addlen = cig_len(rd->cigar[0]);
    *qstart = trim + addlen;
    addlen = 0;
    if (cig_op(rd->cigar[rd->n_cigar - 1]) == BAM_CSOFT_CLIP) addlen = cig_len(rd->cigar[rd->n_cigar - 1]);
    *qend = rd->l_qseq - trim - addlen - 1;
    if (trim > 0) {
        while (*qstart < rd->l_qseq && *qstart >= 1 && rdseq(rd, *qstart) == rdseq(rd, *qstart - 1)) (*qstart)++;
        while (*qend >= 0 && *qend + 1 < rd->l_qseq && rdseq(rd, *qend) == rdseq(rd, *qend + 1)) (*qend)--;
    }
}

/* ---- columns ---------------------------------------------------------------------------- */
static void col_add(Col* c, uint16_t kmer)                   /* base.c:60-71 */
{
    for (int i = 0; i < c->nk; i++)
        if (c->k[i].kmer == kmer) { c->k[i].count++; c->count++; return; }
    if (c->nk == c->capk) { c->capk = c->capk ? c->capk * 2 : 4; c->k = realloc(c->k, sizeof(KC) * (size_t)c->capk); }
    c->k[c->nk].kmer = kmer; c->k[c->nk].count = 1; c->nk++;
    c->count++;
}
static SC* col_find_score(Col* c, uint8_t base)              /* seqlist_find + comparescore */
{
    for (int i = 0; i < c->ns; i++) if (c->s[i].base == base) return &c->s[i];
    return NULL;
}
static SC* col_max_score(Col* c)                             /* base.c:185-197 */
{
    SC* q = NULL;
    if (c->ns) { q = &c->s[0]; for (int i = 0; i < c->ns; i++) if (c->s[i].score > q->score) q = &c->s[i]; }
    return q;
}
static SC* col_get_score(Col* c, uint16_t kmer)              /* base.c:171-178 */
{
    if (kmer) return col_find_score(c, kmer & 0xf);
    return col_max_score(c);
}
static void col_add_score(Col* c, uint16_t kmer, double score)   /* base.c:159-169 */
{
    SC* r = col_find_score(c, kmer & 0xf);
    if (!r) { if (c->ns >= 16) { fprintf(stderr, "oracle: score list overflow\n"); exit(2); } r = &c->s[c->ns++]; }
    r->base = kmer & 0xf; r->kmer = kmer; r->score = score;
}
static double col_coverage(const Col* c, uint16_t base)      /* base.c:79-89 */
{
    uint32_t count = 0;
    for (int i = 0; i < c->nk; i++) if ((c->k[i].kmer & 0xf) == base) count += c->k[i].count;
    return count / (double)c->count;
}

/* contig.c:81-102: per-position base code and FLAG_ZERO for lowercase input */
static void ctg_init(Ctg* g, const np_shard_view* v, int ctg, const Configure* cfg)
{
    memset(g, 0, sizeof(*g));
    g->v = v; g->cfg = cfg;
    g->L = (int32_t)(v->ctg_off[ctg + 1] - v->ctg_off[ctg]);
    g->seq = v->ctg_seq + v->ctg_off[ctg];
    g->r0 = v->ctg_read_off[ctg]; g->r1 = v->ctg_read_off[ctg + 1];
    g->base0 = malloc((size_t)g->L + 1); g->flag0 = calloc((size_t)g->L + 1, 1);
    g->ins = calloc((size_t)g->L + 1, sizeof(int32_t));
    g->colbase = malloc(((size_t)g->L + 1) * sizeof(int32_t));
    for (int32_t i = 0; i < g->L; i++) {
        uint8_t q = g->seq[i];
        if (q >= 97 && q <= 122) { q -= 32; g->flag0[i] |= FLAG_ZERO; }
        g->base0[i] = strtobase_(q);
    }
    g->max_rlen = 1;
    for (int64_t r = g->r0; r < g->r1; r++) {
        Rd rd; get_read(v, r, &rd);
        int32_t e = hts_endpos(&rd) - rd.pos;
        if (e > g->max_rlen) g->max_rlen = e;
    }
}
static void ctg_free_cols(Ctg* g)
{
    if (g->col) { for (int32_t c = 0; c < g->C; c++) free(g->col[c].k); free(g->col); g->col = NULL; }
    free(g->colpos); g->colpos = NULL;
}
static void ctg_free(Ctg* g)
{
    ctg_free_cols(g);
    free(g->base0); free(g->flag0); free(g->ins); free(g->colbase);
}
/* Flatten (position, sub-column) into one column array in contig_data_next order
 * (contig.c:385-399); sub-columns are born with base 3 and their anchor's flag
 * (base.c:24, contig.c:232-235). Sub-columns after the last position are unreachable. */
static void ctg_layout(Ctg* g)
{
    ctg_free_cols(g);
    int64_t C = 0;
    for (int32_t i = 0; i < g->L; i++) { g->colbase[i] = (int32_t)C; C += 1 + (i + 1 < g->L ? g->ins[i] : 0); }
    g->colbase[g->L] = (int32_t)C;
    g->C = (int32_t)C;
    g->col = calloc((size_t)C + 1, sizeof(Col));
    g->colpos = malloc(((size_t)C + 1) * sizeof(int32_t));
    for (int32_t i = 0; i < g->L; i++) {
        int32_t c = g->colbase[i], n = g->colbase[i + 1] - c;
        g->col[c].base = g->base0[i]; g->col[c].flag = g->flag0[i]; g->colpos[c] = i;
        for (int32_t j = 1; j < n; j++) { g->col[c + j].base = 3; g->col[c + j].flag = g->flag0[i]; g->colpos[c + j] = i; }
    }
}
static inline int32_t nsub(const Ctg* g, int32_t i) { return g->colbase[i + 1] - g->colbase[i] - 1; }

/* lower bound: first read index in [r0,r1) with pos >= p */
static int64_t first_read_at_or_after(const Ctg* g, int32_t p)
{
    int64_t lo = g->r0, hi = g->r1;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1; Rd rd; get_read(g->v, mid, &rd);
        if (rd.pos < p) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* contig.c:202-245 (flag == 0 form): grow ins[pos-1] to the longest insertion seen */
static void parse_read_insert(Ctg* g, const Rd* rd, int32_t start, int32_t end)
{
    int32_t pos = rd->pos;
    for (int i = 0; i < rd->n_cigar; i++) {
        int op = cig_op(rd->cigar[i]); int32_t len = cig_len(rd->cigar[i]);
        if (op == BAM_CMATCH || op == BAM_CDEL) pos += len;
        else if (op == BAM_CINS) { if (pos > start && pos <= end) if (g->ins[pos - 1] < len) g->ins[pos - 1] = len; }
    }
}
/* contig.c:170-180 / 182-200: reads overlapping [start, end+1) with filter level >= 1 */
static void create_insert(Ctg* g, int32_t start, int32_t end)
{
    int64_t r = first_read_at_or_after(g, start - g->max_rlen);
    for (; r < g->r1; r++) {
        Rd rd; get_read(g->v, r, &rd);
        if (rd.pos >= end + 1) break;
        if (!(hts_endpos(&rd) > start)) continue;
        if (read_filter(g, &rd) >= 1) parse_read_insert(g, &rd, start, end);
    }
}

static inline uint16_t left_kmer(uint16_t kmer, uint8_t base) { return (uint16_t)(((kmer & 0xff) << 4) | base); }  /* contig.c:360-363 */

/* contig.c:373-383 */
static void as_read(Ctg* g, int32_t start, int32_t end)
{
    uint16_t kmer = 0;
    for (int32_t c = g->colbase[start]; c <= g->colbase[end]; c++) {
        g->col[c].refkmer = kmer = left_kmer(kmer, g->col[c].base);
        col_add(&g->col[c], kmer);
    }
}

/* contig.c:247-331 */
static void parse_read(Ctg* g, const Rd* rd, int32_t start, int32_t end)
{
    if (!rd->n_cigar) return;
    uint16_t kmer = 0;
    int32_t pos = rd->pos, qpos = 0, qstart, qend, j, k, len;
    int cur, last = BAM_CINS;
    cut_read(g, rd, &qstart, &qend);
    for (int i = 0; i < rd->n_cigar; i++) {
        len = cig_len(rd->cigar[i]); cur = cig_op(rd->cigar[i]);
        switch (cur) {
        case BAM_CMATCH: case BAM_CDEL:
            for (j = 0; j < len; j++, pos++) {
                if (pos >= start && pos <= end && qpos >= qstart && qpos <= qend) {
                    if (last != BAM_CINS && pos > start && (qpos > qstart || (qpos == qstart && last == BAM_CDEL))) {
                        int32_t n = nsub(g, pos - 1);
                        for (k = 0; k < n; k++) { kmer = left_kmer(kmer, BASE_DEL); col_add(&g->col[g->colbase[pos - 1] + 1 + k], kmer); }
                    }
                    if (cur == BAM_CDEL) kmer = left_kmer(kmer, BASE_DEL);
                    else kmer = left_kmer(kmer, rdseq(rd, qpos));
                    col_add(&g->col[g->colbase[pos]], kmer);
                }
                if (cur != BAM_CDEL) qpos++;
                last = cur;
            }
            break;
        case BAM_CINS:
            if (pos) {
                int32_t n = pos <= g->L ? nsub(g, pos - 1) : 0;
                for (j = 0; j < len; j++, qpos++) {
                    if (pos > start && pos <= end && qpos >= qstart && qpos <= qend) {
                        if (j >= n) { fprintf(stderr, "oracle: insertion longer than its sub-columns\n"); exit(2); }
                        kmer = left_kmer(kmer, rdseq(rd, qpos));
                        col_add(&g->col[g->colbase[pos - 1] + 1 + j], kmer);
                    }
                }
                if (pos > start && pos <= end && qpos > qstart && qpos <= qend + 1) {
                    for (; j < n; j++) { kmer = left_kmer(kmer, BASE_DEL); col_add(&g->col[g->colbase[pos - 1] + 1 + j], kmer); }
                }
                last = cur;
            } else { qpos += len; qstart += len; last = cur; }
            break;
        case BAM_CHARD_CLIP: case BAM_CSOFT_CLIP:
            qpos += len;
            break;
        }
        if (pos > end) break;
    }
}

/* contig.c:688-704 */
static void parse_region(Ctg* g, int32_t start, int32_t end, int filterlevel)
{
    int64_t r = first_read_at_or_after(g, start - g->max_rlen);
    for (; r < g->r1; r++) {
        Rd rd; get_read(g->v, r, &rd);
        if (rd.pos >= end + 1) break;
        if (!(hts_endpos(&rd) > start)) continue;
        if (read_filter(g, &rd) == filterlevel) parse_read(g, &rd, start, end);
    }
}

/* contig.c:424-454 */
static void calculate_score(Col* cur, Col* prev, double rate)
{
    cur->ns = 0;
    double score = 0;
    uint16_t temp, count, total = cur->count;
    if (total > 1) total--;
    for (int i = 0; i < cur->nk; i++) {
        KC* p = &cur->k[i];
        temp = p->kmer >> 4;
        SC* ps = (temp & 0xf) == 0 ? col_max_score(prev) : col_get_score(prev, temp);
        if (!ps) { fprintf(stderr, "oracle: missing predecessor score (reference would dereference NULL)\n"); exit(2); }
        score = ps->score;
        count = p->count;
        if (p->kmer == cur->refkmer && cur->count > 1) count--;
        score += count - total * rate;
        SC* q = col_get_score(cur, p->kmer);
        if (q == NULL || q->score < score) col_add_score(cur, p->kmer, score);
    }
}

/* contig.c:456-471 */
static void region_score(Ctg* g, int32_t start, int32_t end, double rate)
{
    Col temp; memset(&temp, 0, sizeof(temp));
    Col* q = &g->col[g->colbase[start]];
    for (int i = 0; i < q->nk; i++) col_add_score(&temp, q->k[i].kmer >> 4, 0);
    Col* p = &temp;
    for (int32_t c = g->colbase[start]; c <= g->colbase[end]; c++) {
        calculate_score(&g->col[c], p, rate);
        p = &g->col[c];
    }
}

/* contig.c:473-496 */
static void region_correct(Ctg* g, int32_t start, int32_t end)
{
    int32_t c = g->colbase[end], c0 = g->colbase[start];
    Col* base = &g->col[c];
    SC* score = col_max_score(base);
    for (;;) {
        base->base = score->base;
        if (base->count == 1) base->flag |= FLAG_ZERO; else base->flag &= (uint8_t)~FLAG_ZERO;
        if (col_coverage(base, base->base) < g->cfg->min_count_ratio_skip) base->flag |= FLAG_COVERAGE;
        else base->flag &= (uint8_t)~FLAG_COVERAGE;
        if (c == c0) break;
        c--;
        base = &g->col[c];
        score = col_get_score(base, score->kmer >> 4);
        if (!score) { fprintf(stderr, "oracle: missing backtrack score\n"); exit(2); }
    }
}

/* contig.c:498-517 */
static void brim(const Ctg* g, int with_ext, uint8_t flag, int32_t bstart, int32_t bend, int32_t* start, int32_t* end)
{
    int32_t ext = g->cfg->ext_len_edge;
    *start = *start >= bstart + ext ? *start - ext : bstart;
    *end = *end <= bend - ext ? *end + ext : bend;
    if (!with_ext) return;
#define PB(i) (g->col[g->colbase[i]].base)
#define PF(i) (g->col[g->colbase[i]].flag)
    while (*start > bstart) {
        int32_t p = *start + 1;
        int eq = (p < g->L) ? (PB(p) == PB(p - 1)) : 0;
        if (eq || (PF(p - 1) & flag) != 0) (*start)--; else break;
    }
    while (*end < bend) {
        int32_t p = *end - 1;
        int eq = (p >= 0) ? (PB(p) == PB(p + 1)) : 0;
        if (eq || (PF(p + 1) & flag) != 0) (*end)++; else break;
    }
#undef PB
#undef PF
}

/* contig.c:519-563 */
static void get_region(Ctg* g, int32_t start, int32_t end, uint16_t gap, uint16_t con, uint8_t flag, int with_ext, IList* result)
{
    int32_t qstart = -1, qend = -1;
    uint16_t pgap = 0, pcon = 0;
    int32_t c = g->colbase[start], cend = g->colbase[end];
    while (c <= cend) {
        int32_t i = g->colpos[c];
        if ((g->col[c].flag & flag) != 0) {
            if (qstart == -1) { qstart = i; pcon = 1; }
            else if (pgap == 0) pcon++;
            else pcon = 1;
            pgap = 0;
            qend = i;
        } else if (qstart != -1) {
            pgap++;
            if (pgap > gap) {
                if (pcon > con) {
                    brim(g, with_ext, flag, start, end, &qstart, &qend);
                    il_push(result, qstart); il_push(result, qend);
                    if (qend > i) { c = g->colbase[qend]; /* i = qend, j = 0 */ }
                }
                qstart = qend = -1;
            }
        }
        /* contig_data_next: from (i,j) to the next column; from (qend,0) after a jump */
        c++;
    }
    if (qstart != -1) {
        brim(g, with_ext, flag, start, end, &qstart, &qend);
        il_push(result, qstart); il_push(result, qend);
    }
}

/* contig.c:595-620 */
static void merge_region(IList* l)
{
    if (l->n == 0) return;
    int32_t *pstart = l->d, *pend = pstart + 1, *qstart = pstart, *qend = pend;
    int length = 2;
    for (int i = 0; i < l->n; i += 2) {
        if (*pstart >= *qend) {
            qstart += 2; qend = qstart + 1;
            if (qstart != pstart) *qstart = *pstart;
            if (qend != pend) *qend = *pend;
            length += 2;
        } else {
            while (*pstart < *qstart) qstart -= 2;
            qend = qstart + 1;
            *qend = *pend;
        }
        pstart += 2; pend = pstart + 1;
    }
    l->n = length;
}

/* contig.c:706-734 */
static void score_correct(Ctg* g, int32_t start, int32_t end, int32_t flag, double rate)
{
    int32_t filterlevel = flag & 0xf, insert = (flag >> 4) & 0xf;
    if ((insert & 0x1) == 0) { create_insert(g, start, end); ctg_layout(g); }
    as_read(g, start, end);
    parse_region(g, start, end, filterlevel);
    region_score(g, start, end, rate);
    region_correct(g, start, end);
    if (filterlevel == 2) {
        IList nd = {0, 0, 0};
        get_region(g, start, end, 0, 0, 1, 0, &nd);
        if (nd.n) {
            merge_region(&nd);
            for (int i = 0; i < nd.n; i += 2) {
                parse_region(g, nd.d[i], nd.d[i + 1], 1);
                region_score(g, nd.d[i], nd.d[i + 1], g->cfg->indel_balance_factor_sgs);
                region_correct(g, nd.d[i], nd.d[i + 1]);
            }
        }
        free(nd.d);
    }
}

/* contig.c:736-799: the sequence and, when g->pts is set (trace_polish_open), the PolishPoint trace */
static int64_t get_contig(Ctg* g, uint8_t flag, uint8_t* out, int64_t cap)
{
    int64_t n = 0; uint8_t sign = 0;
    g->npts = 0;
    for (int32_t c = 0; c <= g->colbase[g->L - 1]; c++) {
        Col* p = &g->col[c];
        const int32_t i = g->colpos[c], j = c - g->colbase[i];
        uint8_t raw = g->seq[i];
        if (raw >= 'a' && raw <= 'z') raw = (uint8_t)(raw - 32);                  /* toupper(seqbase[i]) */
        if (p->base == 3) {
            if ((p->flag & flag) != 0) sign = 1;
            if (g->pts && j == 0) {
                if (g->npts >= g->pts_cap) return -1;
                PolishPoint t = {i, (int16_t)j, '.', (char)raw};
                g->pts[g->npts++] = t;
            }
        } else {
            if (n >= cap) return -1;
            uint8_t ch = (uint8_t)basetostr_[p->base];
            if (g->pts && (j != 0 || ch != raw)) {
                if (g->npts >= g->pts_cap) return -1;
                PolishPoint t = {i, (int16_t)j, (char)ch, j != 0 ? '.' : (char)raw};
                g->pts[g->npts++] = t;
            }
            if (sign || (p->flag & flag) != 0) { ch += 32; sign = 0; }
            out[n++] = ch;
        }
    }
    return n;
}

/* ---- task 2 ----------------------------------------------------------------------------- */
typedef struct { uint8_t* region; int32_t length, qual, mapqual, num; } KS;   /* kmercount.h:6-12 */

/* kmercount.c:128-173 */
static void split_region(Ctg* g, const IList* in, uint8_t flag, uint8_t max, IList* result)
{
    IList temp = {0, 0, 0};
    for (int i = 0; i < in->n; i += 2) {
        int32_t s = in->d[i], e = in->d[i + 1];
        il_push(result, s);
        if (e - s > max) {
            int32_t qstart = -1, qend = -1;
            int32_t c = g->colbase[s], cend = g->colbase[e];
            temp.n = 0;
            while (c <= cend) { if ((g->col[c].flag & flag) != 0) break; c++; }
            while (c <= cend) {
                int32_t j = g->colpos[c];
                if ((g->col[c].flag & flag) == 0) { if (qstart == -1) qstart = j; qend = j; }
                else if (qstart != -1) { il_push(&temp, qstart); il_push(&temp, qend); qstart = qend = -1; }
                c++;
            }
            for (int j = 0; j < temp.n; j += 2) {
                int32_t k = (temp.d[j] + temp.d[j + 1]) >> 1;
                il_push(result, k); il_push(result, k);
            }
        }
        il_push(result, e);
    }
    free(temp.d);
}

/* kmercount.c:365-465 (left = right = -1 form) */
static void parse_read_kmer(Ctg* g, const Rd* rd, int32_t start, int32_t end, KS* ks, int flagzero)
{
    if (!rd->n_cigar) return;
    int32_t pos = rd->pos, qpos = 0, qstart, qend, j, k, len, del = 0;
    int cur, last = BAM_CINS;
    cut_read(g, rd, &qstart, &qend);
    ks->mapqual = rd->mapq;
#define KS_APPEND(b) (ks->region[ks->length++] = (uint8_t)(b))
    for (int i = 0; i < rd->n_cigar; i++) {
        len = cig_len(rd->cigar[i]); cur = cig_op(rd->cigar[i]);
        switch (cur) {
        case BAM_CMATCH: case BAM_CDEL:
            for (j = 0; j < len; j++, pos++) {
                if (pos >= start && pos <= end && qpos >= qstart && qpos <= qend) {
                    if (last != BAM_CINS && pos > start && (qpos > qstart || (qpos == qstart && last == BAM_CDEL))) {
                        int32_t n = nsub(g, pos - 1);
                        for (k = 0; k < n; k++) {
                            KS_APPEND(BASE_DEL);
                            if (flagzero == 0) g->col[g->colbase[pos - 1] + 1 + k].flag &= (uint8_t)~FLAG_ZERO;
                            del++;
                        }
                    }
                    if (cur == BAM_CDEL) KS_APPEND(BASE_DEL);
                    else { KS_APPEND(rdseq(rd, qpos)); ks->qual += rd->qual[qpos]; }
                    if (flagzero == 0) g->col[g->colbase[pos]].flag &= (uint8_t)~FLAG_ZERO;
                }
                if (cur != BAM_CDEL) qpos++;
                last = cur;
            }
            break;
        case BAM_CINS:
            if (pos) {
                int32_t n = pos <= g->L ? nsub(g, pos - 1) : 0;
                for (j = 0; j < len; j++, qpos++) {
                    if (pos > start && pos <= end && qpos >= qstart && qpos <= qend) {
                        if (j >= n) { fprintf(stderr, "oracle: insertion longer than its sub-columns (kmer)\n"); exit(2); }
                        KS_APPEND(rdseq(rd, qpos)); ks->qual += rd->qual[qpos];
                        if (flagzero == 0) g->col[g->colbase[pos - 1] + 1 + j].flag &= (uint8_t)~FLAG_ZERO;
                    }
                }
                if (pos > start && pos <= end && qpos > qstart && qpos <= qend + 1) {
                    for (; j < n; j++) {
                        KS_APPEND(BASE_DEL);
                        if (flagzero == 0) g->col[g->colbase[pos - 1] + 1 + j].flag &= (uint8_t)~FLAG_ZERO;
                        del++;
                    }
                }
                last = cur;
            } else { qpos += len; qstart += len; last = cur; }
            break;
        case BAM_CHARD_CLIP: case BAM_CSOFT_CLIP:
            qpos += len;
            break;
        }
        if (pos > end) break;
    }
#undef KS_APPEND
    if (ks->length > 0 && ks->length != del) ks->qual /= ks->length - del;
    else ks->qual = 0;
}

typedef struct { KS* d; int n, cap; } KSList;

/* kmercount.c:332-363 (count = NULL, left = right = -1, flag = 0) */
static void kmer_get_region(Ctg* g, const Rd* rd, int32_t start, int32_t end, int32_t length, KSList* rl, KS* ks, int flagzero)
{
    parse_read_kmer(g, rd, start, end, ks, flagzero);
    if (ks->length == length) {
        KS* p = NULL;
        for (int i = 0; i < rl->n; i++)
            if (memcmp(rl->d[i].region, ks->region, (size_t)rl->d[i].length) == 0) { p = &rl->d[i]; break; }
        if (!p) {
            ks->num = 1;
            if (rl->n == rl->cap) { rl->cap = rl->cap ? rl->cap * 2 : 16; rl->d = realloc(rl->d, sizeof(KS) * (size_t)rl->cap); }
            rl->d[rl->n] = *ks;
            rl->d[rl->n].region = malloc((size_t)length + 1);
            memcpy(rl->d[rl->n].region, ks->region, (size_t)length);
            rl->n++;
        } else { p->num++; p->mapqual += ks->mapqual; p->qual += ks->qual; }
    } else ks->mapqual = 0;
}

static int ks_compare(const KS* a, const KS* b)              /* kmercount.c:63-88 */
{
    if (a == b) return 0;
    if (a->num != b->num) return a->num > b->num ? 1 : -1;
    if (a->mapqual != b->mapqual) return a->mapqual > b->mapqual ? 1 : -1;
    if (a->qual != b->qual) return a->qual > b->qual ? 1 : -1;
    return 0;
}

/* kmercount.c:175-261: windows without an accepted string go to `nodepth` (may be NULL); flagzero as in the reference */
static void kmer_correct(Ctg* g, const IList* region, IList* nodepth, int flagzero)
{
    KSList rl = {0, 0, 0};
    for (int i = 0; i < region->n; i += 2) {
        int32_t start = region->d[i], end = region->d[i + 1];
        int32_t length = g->colbase[end] - g->colbase[start] + 1;     /* contig.c:801-809 */
        int32_t count = 0;
        KS ks; memset(&ks, 0, sizeof(ks));
        ks.region = calloc((size_t)length + 8 + (size_t)g->max_rlen, 1);
        int64_t term = first_read_at_or_after(g, start);              /* record that ends the swapped iterator */
        int64_t r = first_read_at_or_after(g, start - g->max_rlen);
        int broke = 0; int64_t ncand = 0;
        for (; r < term; r++) {
            Rd rd; get_read(g->v, r, &rd);
            if (!(hts_endpos(&rd) > end + 1)) continue;
            ncand++;
            if (read_filter(g, &rd) == 2) {
                ks.length = 0; ks.qual = 0; ks.mapqual = 0; ks.num = 0;   /* ks_clean */
                kmer_get_region(g, &rd, start, end, length, &rl, &ks, flagzero);
                if (ks.mapqual == MAX_MAPQ) {
                    count++;
                    if (count >= g->cfg->max_count_kmer) { broke = 1; break; }
                }
            }
        }
        if (rl.n == 0 && !broke && term < g->r1 && ncand > 0) {
            /* kmercount.c:209-219: the second loop filters/parses the stale `read` (the record
             * that terminated the first iterator) once per record the second iterator yields */
            Rd st; get_read(g->v, term, &st);
            for (int64_t t = 0; t < ncand; t++) {
                if (read_filter(g, &st) == 1) {
                    ks.length = 0; ks.qual = 0; ks.mapqual = 0; ks.num = 0;
                    kmer_get_region(g, &st, start, end, length, &rl, &ks, flagzero);
                }
            }
        }
        if (rl.n > 0) {
            if (flagzero)                                               /* contig_clean_flag(start, end, FLAG_ZERO_N), contig.c:823-831 */
                for (int32_t c = g->colbase[start]; c <= g->colbase[end]; c++) g->col[c].flag &= (uint8_t)~FLAG_ZERO;
            KS* best = NULL;
            if (count == g->cfg->max_count_kmer) {
                int32_t want = MAX_MAPQ * count;
                for (int k = 0; k < rl.n; k++) if (rl.d[k].mapqual == want) { best = &rl.d[k]; break; }
            }
            if (!best) {
                best = &rl.d[0];
                for (int k = 0; k < rl.n; k++) if (ks_compare(best, &rl.d[k]) < 0) best = &rl.d[k];
            }
            /* contig.c:811-821 */
            for (int32_t c = g->colbase[start], q = 0; c <= g->colbase[end]; c++, q++) g->col[c].base = best->region[q];
        } else if (nodepth) { il_push(nodepth, start); il_push(nodepth, end); }
        for (int k = 0; k < rl.n; k++) free(rl.d[k].region);
        rl.n = 0;
        free(ks.region);
    }
    free(rl.d);
}

/* scorechain.c:3-15 */
static int64_t run_score_chain(Ctg* g, uint8_t* out, int64_t cap)
{
    g->filter_kind = 1;
    if (g->L == 0) return 0;
    score_correct(g, 0, g->L - 1, 0x1, g->cfg->indel_balance_factor_sgs);
    return get_contig(g, FLAG_ZERO | FLAG_COVERAGE, out, cap);
}

/* kmercount.c:93-126 */
static int64_t run_kmer_count(Ctg* g, uint8_t* out, int64_t cap)
{
    g->filter_kind = 0;
    if (g->L == 0) return 0;
    ctg_layout(g);
    IList nodepth = {0, 0, 0}, kmerregion = {0, 0, 0};
    get_region(g, 0, g->L - 1, 0, g->cfg->min_len_ldr, 0x1, 0, &nodepth);
    get_region(g, 0, g->L - 1, g->cfg->min_len_inter_kmer, 0, 0x1, 1, &kmerregion);
    if (kmerregion.n > 0) {
        merge_region(&kmerregion);
        for (int i = 0; i < kmerregion.n; i += 2) create_insert(g, kmerregion.d[i], kmerregion.d[i + 1]);
    }
    if (nodepth.n > 0) {
        merge_region(&nodepth);
        for (int i = 0; i < nodepth.n; i += 2) create_insert(g, nodepth.d[i], nodepth.d[i + 1]);
    }
    ctg_layout(g);
    if (nodepth.n > 0)
        for (int i = 0; i < nodepth.n; i += 2)
            score_correct(g, nodepth.d[i], nodepth.d[i + 1], 0x12, g->cfg->indel_balance_factor_sgs);
    if (kmerregion.n > 0) {
        IList win = {0, 0, 0};
        split_region(g, &kmerregion, 0x1, g->cfg->max_len_kmer, &win);
        kmer_correct(g, &win, NULL, 0);
        free(win.d);
    }
    free(nodepth.d); free(kmerregion.d);
    return get_contig(g, FLAG_ZERO, out, cap);
}

/* snpvalid.c:37-66: cut points of a window that found no string: the middle of every unflagged stretch that a flagged
 * column follows (twice, or once for the stretch the window starts with), then the window's end.  The list is read as
 * (start, end) pairs by ss_kmer_correct: a window that starts on a flagged column, or has no flagged column at all,
 * yields an odd number of points — the reference then reads one int past the list (undefined behaviour); callers of this
 * restatement only pin inputs on which every list is even. */
static void fts_split_region(Ctg* g, int32_t start, int32_t end, uint8_t flag, IList* result)
{
    int32_t qstart = -1, qend = -1;
    for (int32_t c = g->colbase[start]; c <= g->colbase[end]; c++) {
        const int32_t i = g->colpos[c];
        if ((g->col[c].flag & flag) == 0) { if (qstart == -1) qstart = i; qend = i; }
        else if (qstart != -1) {
            int count = 2;
            if (qstart == start) { qend = start; count--; }
            int32_t mid = (qstart + qend) / 2;
            for (int k = 0; k < count; k++) { il_push(result, mid); if (qstart != qend) mid++; }
            qstart = qend = -1;
        }
    }
    il_push(result, end);
}

/* snpvalid.c:3-35.  Returns -2 when a second-pass window list is odd (the reference's behaviour is undefined there). */
static int64_t run_snp_valid(Ctg* g, uint8_t* out, int64_t cap)
{
    g->filter_kind = 0;
    if (g->L == 0) return 0;
    ctg_layout(g);
    IList kmerregion = {0, 0, 0};
    int odd = 0;
    get_region(g, 0, g->L - 1, g->cfg->min_len_inter_kmer, 0, FLAG_ZERO, 1, &kmerregion);
    if (kmerregion.n > 0) {
        merge_region(&kmerregion);
        for (int i = 0; i < kmerregion.n; i += 2) create_insert(g, kmerregion.d[i], kmerregion.d[i + 1]);
    }
    ctg_layout(g);
    if (kmerregion.n > 0) {
        IList win = {0, 0, 0}, failed = {0, 0, 0};
        split_region(g, &kmerregion, FLAG_ZERO, g->cfg->max_len_kmer, &win);
        kmer_correct(g, &win, &failed, 1);
        win.n = 0;
        if (failed.n > 0) {
            for (int i = 0; i < failed.n; i += 2) fts_split_region(g, failed.d[i], failed.d[i + 1], FLAG_ZERO, &win);
            if (win.n & 1) odd = 1;
            else {
                for (int i = 0; i < win.n; i += 2) if (win.d[i] > win.d[i + 1]) odd = 1;
                if (!odd) kmer_correct(g, &win, NULL, 0);
            }
        }
        free(win.d); free(failed.d);
    }
    free(kmerregion.d);
    if (odd) return -2;
    return get_contig(g, 0, out, cap);
}

int np_oracle_run_contig(const np_shard_view* v, int contig, int task, const Configure* cfg,
                         uint8_t* out_seq, int64_t out_cap, int64_t* out_len)
{
    if (!v || contig < 0 || contig >= v->n_contigs || (task != 1 && task != 2 && task != 4)) return -1;
    if (task != 1 && !v->qual) return -1;
    Ctg g;
    ctg_init(&g, v, contig, cfg);
    int64_t n = task == 1 ? run_score_chain(&g, out_seq, out_cap) : task == 2 ? run_kmer_count(&g, out_seq, out_cap) : run_snp_valid(&g, out_seq, out_cap);
    if (n == -2) { ctg_free(&g); return -2; }
    ctg_free(&g);
    if (n < 0) return -1;
    *out_len = n;
    return 0;
}

/* The same run with the PolishPoint trace of contig.c:743-799 (what trace_polish_open / nextpolish1.py -debug returns) */
int np_oracle_run_contig_points(const np_shard_view* v, int contig, int task, const Configure* cfg,
                                uint8_t* out_seq, int64_t out_cap, int64_t* out_len,
                                PolishPoint* pts, int64_t pts_cap, int64_t* n_pts)
{
    if (!v || contig < 0 || contig >= v->n_contigs || (task != 1 && task != 2) || !pts) return -1;
    if (task == 2 && !v->qual) return -1;
    Ctg g;
    ctg_init(&g, v, contig, cfg);
    g.pts = pts; g.pts_cap = pts_cap;
    int64_t n = task == 1 ? run_score_chain(&g, out_seq, out_cap) : run_kmer_count(&g, out_seq, out_cap);
    *n_pts = g.npts;
    ctg_free(&g);
    if (n < 0) return -1;
    *out_len = n;
    return 0;
}

int np_oracle_run(const np_shard_view* v, int task, const Configure* cfg,
                  uint8_t* out_seq, int64_t out_cap, int64_t* out_off)
{
    int64_t o = 0;
    out_off[0] = 0;
    for (int c = 0; c < v->n_contigs; c++) {
        int64_t n = 0;
        if (np_oracle_run_contig(v, c, task, cfg, out_seq + o, out_cap - o, &n) != 0) return -1;
        o += n;
        out_off[c + 1] = o;
    }
    return 0;
}

/* config.c:11-40 */
void np_oracle_default_config(Configure* c)
{
    memset(c, 0, sizeof(*c));
    c->trim_len_edge = 2; c->ext_len_edge = 2; c->min_map_quality = 0;
    c->indel_balance_factor_sgs = 0.5; c->min_count_ratio_skip = 0.8;
    c->min_len_ldr = 3; c->min_len_inter_kmer = 5; c->max_len_kmer = 50; c->max_count_kmer = 50;
    c->min_depth_snp = 3; c->min_count_snp = 5; c->min_count_snp_link = 5; c->ploidy = 2;
    c->indel_balance_factor_lgs = 0.33; c->max_indel_factor_lgs = 0.21; c->max_snp_factor_lgs = 0.53;
    c->min_snp_factor_sgs = 0.34;
    c->region_count = 10000; c->count_read_ins_sgs = 10000; c->max_ins_len_sgs = 10000;
    c->max_ins_fold_sgs = 5; c->max_variant_count_lgs = 150000;
    c->max_clip_ratio_sgs = 0.15; c->max_clip_ratio_lgs = 0.4;
    c->trace_polish_open = 0;
}
