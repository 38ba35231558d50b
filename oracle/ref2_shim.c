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

/* ---- the stage in front of the first pass: BAM records -> clipped, anchored alignment strings ---------------------------
 * The record loop of ctg_cns_core (ctg_cns.c:3455-3566) for ONE indexed BAM, written out here with the reference's own
 * functions (cal_l_qseq, cigarint2ul, set_satags, check_indel, bam2aln, clip_aln, get_align_shift, get_align_tags,
 * cal_win_len) and its own window geometry (w = window size, ovl = overlap; ctg_cns.c:3368-3372 fixes 1 Mb overlap and
 * >= 5 Mb windows, the tests also use small ones), restricted to what does not need the large-indel machinery: it
 * returns -10 as soon as a supplementary / secondary record with a split-read gap is seen on a contig longer than
 * INS_MIN_CHECK_LEN (the only way `brk_g` survives the loop, :3567).  Per window it reports the number of alignments, an
 * FNV-1a hash over (start, length, target string, read string) of every alignment in tags_list order, and the first-pass
 * consensus (fast variant, as np2_ref_first_pass).  Returns the number of windows, -1 (caps), -10, -11 (I/O). */
static uint64_t fnv(uint64_t h, const void *p, size_t n){
	const unsigned char *c = p;
	for (size_t i = 0; i < n; i++){ h ^= c[i]; h *= 1099511628211ULL; }
	return h;
}

static int contig_windows_impl(const char *bam_path, const char *ctg, const char *rfseq, int ref_len, int read_type, int w, int ovl,
		int min_cov, int max_windows, int32_t *win_s, int32_t *win_e, int32_t *win_nalns, uint64_t *win_hash,
		int64_t *win_out_off, uint32_t *out_pos, char *out_base, int64_t cap, consensuss_data *keep){
	READS_TYPE = read_type;
	if (READS_TYPE != READS_ONT){ GAP_MIN_LEN = 5; GAP_MIN_RATIO1 = 0.3; }
	else { GAP_MIN_LEN = 3; GAP_MIN_RATIO1 = 0.01; }
	MAX_CLIP_RATIO = READS_TYPE == READS_HIFI ? 0.1 : 0.7;
	/* a path ending in ".list" is a BAM list and goes through the reference's own merge iterator (bam_merge_iter, bsort.c:1202,
	 * 1428: what ctg_cns_core uses, :3475); anything else is one BAM read with htslib's region iterator */
	const size_t plen = strlen(bam_path);
	const int use_merge = plen > 5 && strcmp(bam_path + plen - 5, ".list") == 0;
	samFile *fp = NULL; bam_hdr_t *hdr = NULL; hts_idx_t *idx = NULL;
	bam1_t *brecord = NULL, *own = NULL;
	bam_merge_iter bam_iter;
	if (!use_merge){
		fp = sam_open(bam_path, "r");
		if (!fp) return -11;
		hdr = sam_hdr_read(fp);
		idx = sam_index_load(fp, bam_path);
		if (!hdr || !idx) return -11;
		brecord = own = bam_init1();
	}
	alignment aln_, aln;
	memset(&aln, 0, sizeof(aln));
	aln.max_aln_len = 100000;
	aln.t_aln_str = malloc(aln.max_aln_len);
	aln.q_aln_str = malloc(aln.max_aln_len);
	satags sas; memset(&sas, 0, sizeof(sas));
	int32_t j, l, p, s, e, l_qseq, rege, nw = 0, rc = 0;
	pos rfp1, rdp1, rfp2, rdp2;
	gap g;
	int64_t total = 0;
	int32_t b = cal_win_len(w, ovl, ref_len);
	s = e = 0;
	while (e < ref_len && !rc){
		e = s + b > ref_len ? ref_len : s + b;
		l = e - s;
		if (nw >= max_windows) { rc = -1; break; }
		memset(&aln_, 0, sizeof(aln_));
		aln_.q_aln_str = aln_.t_aln_str = (char *) rfseq + s;
		aln_.aln_q_len = aln_.aln_t_len = l;
		aln_.aln_t_e = aln_.aln_len = l;
		uint32_t seq_count = 0, seq_count_m = max(l / 1000, 2);
		msa_p *msa = calloc(l + 1, sizeof(msa_p));
		align_tags_t *tags_list = malloc(seq_count_m * sizeof(align_tags_t));
		uint64_t h = 14695981039346656037ULL;
		get_align_tags(&aln_, &tags_list[seq_count++], msa);
		h = fnv(h, &aln_.aln_t_s, 4); h = fnv(h, &aln_.aln_len, 4); h = fnv(h, aln_.t_aln_str, l); h = fnv(h, aln_.q_aln_str, l);
		rege = s == 0 ? (e > INS_RADOM_LEN ? e : INS_RADOM_LEN) : e;
		char reg[1024];
		sprintf(reg, "%s:%d-%d", ctg, s, rege);
		hts_itr_t *it = NULL;
		if (use_merge) bam_merge_iter_init(0, NULL, bam_path, reg, &bam_iter);
		else it = sam_itr_querys(idx, hdr, reg);
		p = 0;
		while (use_merge ? bam_merge_iter_core(&bam_iter) > 0 : (it && sam_itr_next(fp, it, brecord) >= 0)){
			if (use_merge) brecord = bam_iter.heap->entry.bam_record;
			p = brecord->core.pos;
			if (p >= e) rege = 0;
			uint32_t *cigar = bam_get_cigar(brecord);
			l_qseq = cal_l_qseq(brecord);
			rfp1.s = brecord->core.pos;
			rfp1.e = bam_endpos(brecord);
			rdp1.s = cigarint2ul(cigar, brecord->core.n_cigar, 0);
			rdp1.e = l_qseq - cigarint2ul(cigar, brecord->core.n_cigar, 1);
			uint8_t *satag_ = bam_aux_get(brecord, "SA");
			g.score = 0;
			if (satag_){
				set_satags(satag_, &sas);
				uint8_t strand = brecord->core.flag & 16 ? 1 : 0;
				for (j = 0; j < sas.i; j++){
					satag *sa = &sas.sa[j];
					if (strcmp(sa->rname, ctg) == 0 && sa->strand == strand){
						rfp2.s = sa->pos;
						rfp2.e = sa->pos + cigarstr2rlen(sa->cigar);
						rdp2.s = cigarstr2ul(sa->cigar, 0);
						rdp2.e = l_qseq - cigarstr2ul(sa->cigar, 1);
						check_indel(&g, l_qseq, &rfp1, &rdp1, &rfp2, &rdp2);
					}
				}
			}
			if (rege && brecord->core.flag & 0xD04 && ref_len > INS_MIN_CHECK_LEN && g.score) { rc = -10; break; }
			if (brecord->core.flag & 0xD04) continue;
			if ((!g.score) && (rdp1.e - rdp1.s) / (double) l_qseq <= MAX_CLIP_RATIO) continue;
			if (!rege) continue;
			aln.aln_t_s = rfp1.s;
			aln.aln_t_e = rfp1.e;
			aln.aln_q_s = rdp1.s;
			aln.aln_q_e = rdp1.e;
			aln.aln_len = aln.shift = 0;
			l = bam2aln(&aln, rfseq, bam_get_seq(brecord), cigar, brecord->core.n_cigar);
			if (l != aln.aln_t_e) { rc = -11; break; }
			if (aln.aln_t_s < s || aln.aln_t_e > e) clip_aln(&aln, s, e, g.score);
			get_align_shift(&aln, 8, g.score);
			if (aln.aln_t_s > aln.aln_t_e - 500) continue;
			aln.aln_t_s -= s;
			aln.aln_t_e -= s;
			if ((msa[aln.aln_t_s].coverage > 3000 && msa[aln.aln_t_e].coverage > 3000) ||
				(msa[aln.aln_t_s].coverage > 500 && msa[aln.aln_t_e].coverage > 500 && rdp1.e - rdp1.s < l_qseq * 0.9)) continue;
			get_align_tags(&aln, &tags_list[seq_count++], msa);
			h = fnv(h, &aln.aln_t_s, 4); h = fnv(h, &aln.aln_len, 4);
			h = fnv(h, aln.t_aln_str + aln.shift, aln.aln_len); h = fnv(h, aln.q_aln_str + aln.shift, aln.aln_len);
			if (seq_count >= seq_count_m){
				seq_count_m += 1000;
				tags_list = realloc(tags_list, seq_count_m * sizeof(align_tags_t));
			}
		}
		if (it) hts_itr_destroy(it);
		if (use_merge) bam_merge_iter_destroy(&bam_iter);
		if (rc) { for (uint32_t i = 0; i < seq_count; i++) free(tags_list[i].align_tags); free(tags_list); free(msa); break; }
		win_s[nw] = s; win_e[nw] = e; win_nalns[nw] = seq_count; win_hash[nw] = h; win_out_off[nw] = total;
		consensus_data *c = get_cns_from_align_tags(tags_list, msa, seq_count, e - s, min_cov, 0, 1, NULL);
		if (total + c->len > cap) rc = -1;
		for (unsigned i = 0; !rc && i < c->len; i++){
			out_pos[total + i] = c->cns_bases[c->len - 1 - i].pos;
			out_base[total + i] = c->cns_bases[c->len - 1 - i].base;
		}
		if (!rc) total += c->len;
		if (keep && !rc){                               /* as ctg_cns_core stores it (:3587-3593) */
			c->uncorrected_len = s;
			keep->consensus[keep->i] = c;
			if (++keep->i >= keep->i_m){ keep->i_m += 5; keep->consensus = realloc(keep->consensus, keep->i_m * sizeof(consensus_data *)); }
		}else { free(c->cns_bases); free(c); }
		nw++;
		s = e - ovl;
	}
	if (!rc) win_out_off[nw] = total;
	free(aln.t_aln_str); free(aln.q_aln_str);
	if (sas.i_m) free(sas.sa);
	if (!use_merge){ bam_destroy1(own); hts_idx_destroy(idx); bam_hdr_destroy(hdr); sam_close(fp); }
	return rc ? rc : nw;
}

/* the draft as ctg_cns_core sees it: read_ref's 2-bit packing (seq2bit1) and bit2seq1 (ctg_cns.c:2283,3446) */
void np2_ref_roundtrip(const char *seq, int len, char *out){
	uint32_t *s = malloc(sizeof(uint32_t) * (len / 16 + 1));
	seq2bit1(s, len, (char *) seq);
	bit2seq1(s, len, out);
	free(s);
}

int np2_ref_contig_windows(const char *bam_path, const char *ctg, const char *rfseq, int ref_len, int read_type, int w, int ovl,
		int min_cov, int max_windows, int32_t *win_s, int32_t *win_e, int32_t *win_nalns, uint64_t *win_hash,
		int64_t *win_out_off, uint32_t *out_pos, char *out_base, int64_t cap){
	return contig_windows_impl(bam_path, ctg, rfseq, ref_len, read_type, w, ovl, min_cov, max_windows, win_s, win_e, win_nalns,
		win_hash, win_out_off, out_pos, out_base, cap, NULL);
}

/* The reference's FAST mode for one contig, end to end: the windows' first-pass consensus linked by link_consensus_fast
 * (ctg_cns.c:3053-3119; what ctg_cns_core returns when its local `fast` is set, :3620 — the shipped code never sets it).
 * Returns the length of the linked sequence written to out_seq, or a negative code of np2_ref_contig_windows. */
int64_t np2_ref_contig_fast(const char *bam_path, const char *ctg, const char *rfseq, int ref_len, int read_type, int w, int ovl,
		char *out_seq, int64_t cap){
	enum { MW = 4096 };
	int32_t *ws = malloc(MW * 4), *we = malloc(MW * 4), *wn = malloc(MW * 4);
	uint64_t *wh = malloc(MW * 8);
	int64_t *woff = malloc((MW + 1) * 8);
	int64_t ocap = (int64_t) ref_len * 4 + 1024;
	uint32_t *pos = malloc(ocap * 4);
	char *base = malloc(ocap);
	consensuss_data ct;
	memset(&ct, 0, sizeof(ct));
	ct.i_m = 5; ct.s = ovl; ct.w = w;
	ct.consensus = malloc(ct.i_m * sizeof(consensus_data *));
	int nw = contig_windows_impl(bam_path, ctg, rfseq, ref_len, read_type, w, ovl, 4, MW, ws, we, wn, wh, woff, pos, base, ocap, &ct);
	int64_t n = nw;
	if (nw > 0){
		consensus_trimed_data *d = link_consensus_fast(&ct, ref_len, 50);
		n = d->data[0].len;
		if (n > cap) n = -1; else memcpy(out_seq, d->data[0].seq, n);
		free_consensus_trimed_data(d);
	}
	free(ct.consensus); free(ws); free(we); free(wn); free(wh); free(woff); free(pos); free(base);
	return n;
}

/* The PRODUCTION variant of the window consensus on alignment strings: get_cns_from_align_tags(..., fast = 0, ...) with no
 * gap clusters — first pass, low-quality regions, candidate ranking, POA, second round (ctg_cns.c:1859-1873 and everything
 * it calls).  Output as np2_ref_first_pass plus the per-base qv; list order as the reference leaves it (forward). */
int np2_ref_window_prod(int read_type, int n_reads, const uint32_t *aln_t_s, const uint32_t *aln_len, const uint64_t *str_off,
		const char *t_str, const char *q_str, int len, int min_cov, uint32_t *out_pos, char *out_base, uint8_t *out_qv, int cap){
	READS_TYPE = read_type;
	if (READS_TYPE != READS_ONT){ GAP_MIN_LEN = 5; GAP_MIN_RATIO1 = 0.3; }
	else { GAP_MIN_LEN = 3; GAP_MIN_RATIO1 = 0.01; }
	msa_p *msa = calloc(len + 1, sizeof(msa_p));
	align_tags_t *tags_list = malloc((n_reads > 0 ? n_reads : 1) * sizeof(align_tags_t));
	for (int i = 0; i < n_reads; i++){
		alignment aln;
		memset(&aln, 0, sizeof(aln));
		aln.aln_len = aln_len[i];
		aln.aln_t_s = aln_t_s[i];
		aln.t_aln_str = (char *) t_str + str_off[i];
		aln.q_aln_str = (char *) q_str + str_off[i];
		get_align_tags(&aln, &tags_list[i], msa);
	}
	if (len < 1 || msa[len - 1].max_size == 0) return -2;
	gap_clusters clusters;
	memset(&clusters, 0, sizeof(clusters));
	consensus_data *c = get_cns_from_align_tags(tags_list, msa, n_reads, len, min_cov, 0, 0, &clusters);
	int n = (int) c->len;
	if (n > cap) n = -1;
	for (int i = 0; i < n; i++){
		out_pos[i] = c->cns_bases[i].pos;
		out_base[i] = c->cns_bases[i].base;
		out_qv[i] = (uint8_t) c->cns_bases[i].qv;
	}
	free(c->cns_bases);
	free(c);
	return n;
}
