/* np2_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into, loaded by or shipped with the product libraries).
 *
 * Plain-C restatement of the FIRST PASS of the reference's long-read consensus window (nextpolish2.so):
 *   tags      get_align_tags / get_align_tag      source/lib/ctg_cns.c:1213-1255, :303-321
 *   tally     update_msa                          ctg_cns.c:324-368   (link triples (p, pp, ppp) per node, first-seen order)
 *   chain     get_cns_from_align_tags             ctg_cns.c:1876-2128 (integer scores; one rule set per read type)
 *   backtrack generate_cns_from_best_score_fast   ctg_cns.c:1475-1509 (+ qv of generate_cns_from_best_score, :1839-1846)
 * Pinned against the reference itself, entered through oracle/ref2_shim.c (np2_ref_first_pass): tests/test_lgs_first_pass.py
 * compares the two on seeded alignment sets of every read type and on windows cut from the reference's own test_data long
 * reads (tests/golden/lgs_*.json, minted by tests/golden/make_golden_lgs.py).
 *
 * NOT restated (the part of ctg_cns_core that follows this pass): low-quality regions, k-mer ranked candidates, POA and
 * the second alignment round (ctg_cns.c:822-1474, dag.c, align.c), window linking (:3053-3330). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

typedef struct { int32_t t_pos; uint16_t delta; uint8_t q_base; } tag_t;
typedef struct { tag_t pp, ppp; int64_t score; uint16_t link_count; } ent_t;
typedef struct { ent_t *e; uint32_t len, cap, best; } node_t;                    /* best = the reference's reused max_size */
typedef struct { uint16_t coverage, max_size; node_t *nodes; /* [max_size][6] */ } col_t;

static const uint8_t INT_TO_BASE[] = {65, 84, 71, 67, 45, 78, 77};              /* ctg_cns.c:48-50: A T G C - N M */
static uint8_t base_to_int(unsigned char c) {                                   /* ctg_cns.c:58-67 */
	switch (c) {
	case 'A': case 'a': return 0;
	case 'T': case 't': return 1;
	case 'G': case 'g': return 2;
	case 'C': case 'c': return 3;
	case 'N': return 5;
	case 'M': return 6;
	default: return 4;
	}
}
static int tag_eq(const tag_t *a, const tag_t *b) { return a->t_pos == b->t_pos && a->delta == b->delta && a->q_base == b->q_base; }

/* Same signature as np2_ref_first_pass (oracle/ref2_shim.c) plus the production qv (may be NULL).
 * Returns the number of consensus bases (forward order); -1 cap too small; -2 no node at the last column; -3 an alignment
 * that leaves the window or starts with a gap column (the reference indexes msa[] out of range there); -4 the backtrack
 * reached a node without entries (the reference reads past its allocation there). */
int np2_oracle_first_pass_qv(int read_type, int n_reads, const uint32_t *aln_t_s, const uint32_t *aln_len, const uint64_t *str_off,
		const char *t_str, const char *q_str, int len, int min_cov, uint32_t *out_pos, char *out_base, uint8_t *out_qv, int cap) {
	if (len < 1) return -2;
	col_t *msa = calloc((size_t)len + 1, sizeof(col_t));
	int rc = 0;
	/* ---- pass 1 over the alignments: coverage and sub-column count of every window position (get_align_tags) */
	for (int r = 0; r < n_reads && !rc; r++) {
		const char *t = t_str + str_off[r], *q = q_str + str_off[r];
		int64_t te = (int64_t)aln_t_s[r] - 1;
		uint32_t delta = 0;
		if (aln_len[r] == 0 || t[0] == '-') { rc = -3; break; }
		for (uint32_t i = 0; i < aln_len[r]; i++) {
			if (t[i] == '-') delta++;
			else { te++; delta = 0; }
			if (te >= len) { rc = -3; break; }
			if (delta == 0 && q[i] != 'M') msa[te].coverage++;
			if (delta >= msa[te].max_size) msa[te].max_size = (uint16_t)(delta + 1);
		}
	}
	if (!rc && msa[len - 1].max_size == 0) rc = -2;
	if (rc) { free(msa); return rc; }
	for (int p = 0; p < len; p++) msa[p].nodes = calloc((size_t)msa[p].max_size * 6 + 1, sizeof(node_t));
	/* ---- tally (update_msa): tags in stream order; a tag whose own base or whose predecessor is 'M' (6) is not counted */
	for (int r = 0; r < n_reads; r++) {
		const char *t = t_str + str_off[r], *q = q_str + str_off[r];
		tag_t p1 = {0, 0, 0}, pp = {-1, 0, 0}, ppp = {-1, 0, 0};                 /* align_tag_head, ctg_cns.c:52-56 */
		for (uint32_t i = 0; i < aln_len[r]; i++) {
			p1.q_base = base_to_int((unsigned char)q[i]);
			if (i == 0) { p1.t_pos = (int32_t)aln_t_s[r]; p1.delta = 0; }
			else if (t[i] == '-') p1.delta++;
			else { p1.delta = 0; p1.t_pos++; }
			if (p1.q_base != 6 && pp.q_base != 6) {
				node_t *nd = &msa[p1.t_pos].nodes[(size_t)p1.delta * 6 + p1.q_base];
				uint32_t k;
				for (k = 0; k < nd->len; k++)
					if (tag_eq(&nd->e[k].pp, &pp) && tag_eq(&nd->e[k].ppp, &ppp)) { nd->e[k].link_count++; break; }
				if (k == nd->len) {
					if (nd->len == nd->cap) { nd->cap = nd->cap ? nd->cap * 2 : 4; nd->e = realloc(nd->e, nd->cap * sizeof(ent_t)); }
					nd->e[nd->len].pp = pp; nd->e[nd->len].ppp = ppp; nd->e[nd->len].link_count = 1; nd->e[nd->len].score = 0;
					nd->len++;
				}
			}
			ppp = pp; pp = p1;
		}
	}
	/* ---- chain (get_cns_from_align_tags): nodes in (position, sub-column, base) order */
	const int pen = read_type == 3 ? 4 : 3;                                       /* READS_HIFI: 4 x coverage */
	int64_t global_best_score = INT64_MIN;
	tag_t gb = {-1, 0, 0};
	for (int p = 0; p < len; p++)
		for (int d = 0; d < msa[p].max_size; d++)
			for (int b = 0; b < 6; b++) {
				node_t *nd = &msa[p].nodes[(size_t)d * 6 + b];
				nd->best = 0;
				int64_t p_pp_score = INT64_MIN, p_pp_score_ = INT64_MIN;
				int tmp = 0;
				for (uint32_t m = 0; m < nd->len; m++) if (nd->e[m].link_count > tmp) tmp = nd->e[m].link_count;
				for (uint32_t m = 0; m < nd->len; m++) {
					ent_t *em = &nd->e[m];
					if (em->pp.t_pos == -1) em->score = 10 * (int64_t)em->link_count - pen * (int64_t)msa[p].coverage;
					else {
						node_t *pn = &msa[em->pp.t_pos].nodes[(size_t)em->pp.delta * 6 + em->pp.q_base];
						for (uint32_t n = 0; n < pn->len; n++) {
							ent_t *en = &pn->e[n];
							if (!tag_eq(&en->pp, &em->ppp)) continue;
							const int64_t s = en->score + 10 * (int64_t)em->link_count - pen * (int64_t)msa[p].coverage;
							if (s > em->score) { em->score = s; p_pp_score_ = en->score; }
							if (read_type == 2 || read_type == 3) {                    /* CLR :1958-1963, HIFI :2023-2028 */
								if (en->score > p_pp_score || (en->score == p_pp_score && em->pp.q_base != 4)) { nd->best = m; p_pp_score = en->score; }
							} else if (read_type != 4) {                               /* ONT (and anything else) :2086-2094 */
								if (((em->ppp.delta > 1 || em->pp.delta > 0) &&
								     (em->link_count > msa[p].coverage * 0.2 || em->link_count > tmp / 2)) ||
								    (em->link_count > nd->e[nd->best].link_count / 2 && en->score > p_pp_score &&
								     (em->pp.q_base == 4 || em->pp.q_base == b || em->ppp.q_base == b || em->pp.q_base == em->ppp.q_base))) {
									nd->best = m; p_pp_score = en->score;
								}
							}
						}
					}
					if (read_type == 4) {                                              /* RS :1918-1921 */
						if (em->score >= nd->e[nd->best].score) { nd->best = m; p_pp_score = p_pp_score_; }
					} else if (em->score > nd->e[nd->best].score || (em->score == nd->e[nd->best].score && em->pp.q_base != 4)) {
						nd->best = m; p_pp_score = p_pp_score_;
					}
				}
				if (nd->len && p == len - 1 && nd->e[nd->best].score >= global_best_score) {
					gb.t_pos = p; gb.delta = (uint16_t)d; gb.q_base = (uint8_t)b;
					if (nd->e[nd->best].score > global_best_score) global_best_score = nd->e[nd->best].score;
				}
			}
	/* ---- backtrack (generate_cns_from_best_score_fast; qv as in generate_cns_from_best_score) */
	int n = 0;
	if (gb.t_pos < 0) rc = -2;
	while (!rc) {
		const node_t *nd = &msa[gb.t_pos].nodes[(size_t)gb.delta * 6 + gb.q_base];
		if (nd->len == 0) { rc = -4; break; }
		if (gb.q_base != 4) {
			if (n >= cap) { rc = -1; break; }
			out_pos[n] = (uint32_t)gb.t_pos;
			out_base[n] = msa[gb.t_pos].coverage > min_cov ? (char)INT_TO_BASE[gb.q_base] : (char)tolower(INT_TO_BASE[gb.q_base]);
			if (out_qv) out_qv[n] = msa[gb.t_pos].coverage ? (uint8_t)(100 * (int)nd->e[nd->best].link_count / msa[gb.t_pos].coverage) : 0;
			n++;
		}
		gb = nd->e[nd->best].pp;
		if (gb.t_pos == -1) break;
	}
	for (int i = 0, j = n - 1; !rc && i < j; i++, j--) {
		uint32_t tp = out_pos[i]; out_pos[i] = out_pos[j]; out_pos[j] = tp;
		char tb = out_base[i]; out_base[i] = out_base[j]; out_base[j] = tb;
		if (out_qv) { uint8_t tq = out_qv[i]; out_qv[i] = out_qv[j]; out_qv[j] = tq; }
	}
	for (int p = 0; p < len; p++) {
		for (int k = 0; k < msa[p].max_size * 6; k++) free(msa[p].nodes[k].e);
		free(msa[p].nodes);
	}
	free(msa);
	return rc ? rc : n;
}

int np2_oracle_first_pass(int read_type, int n_reads, const uint32_t *aln_t_s, const uint32_t *aln_len, const uint64_t *str_off,
		const char *t_str, const char *q_str, int len, int min_cov, uint32_t *out_pos, char *out_base, int cap) {
	return np2_oracle_first_pass_qv(read_type, n_reads, aln_t_s, aln_len, str_off, t_str, q_str, len, min_cov, out_pos, out_base, NULL, cap);
}
