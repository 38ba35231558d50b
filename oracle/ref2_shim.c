/* ref2_shim.c — TEST INFRASTRUCTURE ONLY.
 *
 * Gives the tests a door into the UNMODIFIED long-read consensus of the reference (source/lib/ctg_cns.c): the file is
 * compiled from where it lies under /root/reference by including it here (nothing of it is copied into this repository),
 * so that its file-local stages can be called one at a time.  Exposed stage = the first pass of a consensus window:
 *
 *   gapped alignment strings  --get_align_tags (ctg_cns.c:1213)-->  packed (base | insert) tags + coverage / max_size
 *                             --update_msa (:324)-->                 link triples (p, pp, ppp) counted per node
 *                             --get_cns_from_align_tags (:1876)-->   integer score chain, per read type
 *                             --generate_cns_from_best_score_fast (:1475)--> backtrack: (pos, base) list
 *
 * i.e. get_cns_from_align_tags(..., fast = 1, ...): the same tally, chain and backtrack the production call (fast = 0)
 * runs before its low-quality-region / POA stage.  Built into oracle/_ref/libnp2_refshim.so by oracle/Makefile (ref2);
 * used by tests/test_lgs_first_pass.py and tests/golden/make_golden_lgs.py to pin oracle/np2_oracle.c and the GPU path. */
#define main np2_ref_unused_main
#include "ctg_cns.c"
#undef main

/* alignment i: columns [str_off[i], str_off[i] + aln_len[i]) of t_str / q_str ('-' = gap; 'M' = masked), first target
 * position aln_t_s[i] (window-relative).  Returns the number of consensus bases written (forward order), -1 when cap is
 * too small, -2 when the last window column has no node (the reference would read msa[-1]). */
int np2_ref_first_pass(int read_type, int n_reads, const uint32_t *aln_t_s, const uint32_t *aln_len, const uint64_t *str_off,
		const char *t_str, const char *q_str, int len, int min_cov, uint32_t *out_pos, char *out_base, int cap){
	READS_TYPE = read_type;
	if (READS_TYPE != READS_ONT){ GAP_MIN_LEN = 5; GAP_MIN_RATIO1 = 0.3; }      /* ctg_cns.c:3435-3442 */
	else { GAP_MIN_LEN = 3; GAP_MIN_RATIO1 = 0.01; }
	msa_p *msa = calloc(len + 1, sizeof(msa_p));
	align_tags_t *tags_list = malloc((n_reads > 0 ? n_reads : 1) * sizeof(align_tags_t));
	for (int i = 0; i < n_reads; i++){
		alignment aln;
		memset(&aln, 0, sizeof(aln));
		aln.shift = 0;
		aln.aln_len = aln_len[i];
		aln.aln_t_s = aln_t_s[i];
		aln.t_aln_str = (char *) t_str + str_off[i];
		aln.q_aln_str = (char *) q_str + str_off[i];
		get_align_tags(&aln, &tags_list[i], msa);
	}
	if (len < 1 || msa[len - 1].max_size == 0){
		for (int i = 0; i < n_reads; i++) free(tags_list[i].align_tags);
		free(tags_list); free(msa);
		return -2;
	}
	consensus_data *c = get_cns_from_align_tags(tags_list, msa, n_reads, len, min_cov, 0, 1, NULL);   /* frees tags + msa */
	int n = (int) c->len;
	if (n > cap) n = -1;
	for (int i = 0; i < n; i++){                /* the fast path leaves the list in backtrack order */
		out_pos[i] = c->cns_bases[c->len - 1 - i].pos;
		out_base[i] = c->cns_bases[c->len - 1 - i].base;
	}
	free(c->cns_bases);
	free(c);
	return n;
}
