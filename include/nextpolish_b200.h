/*
 * nextpolish_b200.h — C ABI of the B200-native polishing engine.
 *
 * Two groups of entry points:
 *
 *  (1) The reference ABI: exactly the symbols lib/nextpolish1.py binds through ctypes
 *      (reference source/lib/nextpolish1.py:84-100) and source/lib/main.c calls, with
 *      byte-identical struct layouts.  The shared object is built as
 *      nextpolish_b200/lib/nextpolish1.so so that it can be dropped next to an unmodified
 *      nextpolish1.py.
 *
 *  (2) The batch ABI (np_*): plain pointers and sizes over "packed shards" (many contigs and
 *      their coordinate-sorted read blocks) for callers that already hold decoded reads.
 *      bench.py, the tests and the native CLI use it; score_chain()/kmer_count() are thin
 *      per-contig wrappers over the same kernels.
 *
 * There is no CPU implementation behind these symbols: every compute entry point launches the
 * sm_100a kernels and returns an error (or exits like the reference does) when no CUDA
 * device is usable.
 */
#ifndef NEXTPOLISH_B200_H
#define NEXTPOLISH_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * (1) Reference ABI
 * ---------------------------------------------------------------------------------------- */

/* reference source/lib/config.h:25-67 (mirrored by nextpolish1.py:27-65); sizeof == 152 */
typedef struct {
    uint8_t  trim_len_edge;
    uint8_t  ext_len_edge;
    uint8_t  min_map_quality;
    double   indel_balance_factor_sgs;
    double   min_count_ratio_skip;

    uint8_t  min_len_ldr;
    uint8_t  min_len_inter_kmer;
    uint8_t  max_len_kmer;
    uint8_t  max_count_kmer;

    uint8_t  min_depth_snp;
    uint8_t  min_count_snp;
    int8_t   min_count_snp_link;
    double   ploidy;
    double   indel_balance_factor_lgs;
    double   max_indel_factor_lgs;
    double   max_snp_factor_lgs;
    double   min_snp_factor_sgs;

    int32_t  region_count;
    uint32_t count_read_ins_sgs;
    uint32_t max_ins_len_sgs;
    int32_t  max_ins_fold_sgs;
    int32_t  max_variant_count_lgs;

    double   max_clip_ratio_sgs;
    double   max_clip_ratio_lgs;

    int32_t  trace_polish_open;
    int32_t  read_tlen;
    int32_t  read_len;

    char*    fastafn;
    char*    bamfn;
    char*    thirdbamfn;
} Configure;

/* reference source/lib/contig.h:10-22 (nextpolish1.py:67-81) */
typedef struct {
    int32_t pos;
    int16_t index;
    char    curbase;
    char    base;
} PolishPoint;

typedef struct {
    char*        contig;     /* NUL-terminated polished sequence, malloc()ed */
    PolishPoint* data;       /* change trace when trace_polish_open, else NULL */
    int32_t      length;     /* strlen(contig) */
    int32_t      datalength;
} PolishResult;

/* replaces config.c:8-56 (config_init), config.c:58-68 (config_destory) */
Configure* config_init(const char* fastafn, const char* bamfn, const char* thirdbamfn);
void       config_destory(Configure* config);

/* replaces scorechain.c:3-15 — task 1, whole-contig pileup + k-mer score chain */
PolishResult* score_chain(const char* tigname, Configure* configure);
/* replaces kmercount.c:93-126 — task 2, low-depth re-score + spanning-read k-mer vote */
PolishResult* kmer_count(const char* tigname, Configure* configure);
/* bound by nextpolish1.py:95-100 but outside this engine's scope:
 * they print a diagnostic and exit(1), the reference's own fatal-error convention. */
PolishResult* snp_phase(const char* tigname, Configure* configure);
PolishResult* snp_valid(const char* tigname, Configure* configure);
PolishResult* lgspolish(const char* tigname, Configure* configure);

/* replaces contig.c:20-30 */
PolishResult* polishresult_init(void);
void          polishresult_destory(PolishResult* polishresult);

/* ------------------------------------------------------------------------------------------
 * (2) Batch ABI over packed shards
 * ---------------------------------------------------------------------------------------- */

#define NP_OK            0
#define NP_ERR_CUDA     -1   /* no device / CUDA runtime failure (message via np_last_error) */
#define NP_ERR_ARG      -2
#define NP_ERR_IO       -3
#define NP_ERR_LIMIT    -4   /* shard exceeds a documented limit (columns, depth >= 65535) */
#define NP_ERR_RATE     -5   /* reserved */

#define NP_TASK_SCORE_CHAIN 1
#define NP_TASK_KMER_COUNT  2
#define NP_TASK_SNP_VALID   4   /* snp_valid, snpvalid.c:3-35 (task 4 of nextpolish1.py) */

/*
 * One packed read block ("record"), 16-byte aligned, little endian:
 *   int32  pos        0-based leftmost reference position   (bam1_core_t.pos)
 *   uint16 flag       BAM FLAG
 *   uint8  mapq       BAM MAPQ
 *   uint8  enc        encoding of seq[]: 0 = BAM 4-bit nt16 codes, 1 = 2 bits per base (see below)
 *   int32  isize      BAM TLEN
 *   uint16 l_qseq     read length
 *   uint16 n_cigar    number of CIGAR ops (>= 1)
 *   uint32 cigar[n_cigar]          BAM encoding, len<<4 | op
 *   uint8  seq[]                   enc 0: (l_qseq+1)/2 bytes, BAM 4-bit nt16 codes, high nibble first;
 *                                  enc 1: (l_qseq+3)/4 bytes, A=0 C=1 G=2 T=3 (nt16 code = 1 << v), first base in
 *                                  the two highest bits — chosen by the packer for reads made of A/C/G/T only
 *                                  (halves the bytes that cross PCIe; set NEXTPOLISH_B200_4BIT=1 to disable)
 *   zero padding to a multiple of 16 bytes
 * Qualities (needed by task 2 only) live in a second stream: l_qseq bytes per read, each
 * read padded to 16 bytes.
 * Records are in BAM file order (coordinate sorted) and grouped by contig.
 */
typedef struct {
    int32_t         n_contigs;
    int64_t         n_reads;
    const int64_t*  ctg_off;       /* [n_contigs+1] offsets of each contig in ctg_seq           */
    const uint8_t*  ctg_seq;       /* draft bases as in the FASTA (case preserved), concatenated */
    const int64_t*  ctg_read_off;  /* [n_contigs+1] read index range of each contig              */
    const uint32_t* rec_off;       /* [n_reads+1] record offsets in 16-byte units                */
    const uint8_t*  rec;           /* packed records                                              */
    const uint32_t* qual_off;      /* [n_reads+1] quality offsets in 16-byte units, or NULL      */
    const uint8_t*  qual;          /* qualities, or NULL                                          */
} np_shard_view;

/* ---- host side: FASTA/BAM -> packed shard (own BGZF/BAM/BAI/FASTA reader; zlib only) ---- */
typedef struct np_shard np_shard;
/* names == NULL or n_names == 0: every contig of the FASTA, in FASTA order.
 * with_qual: 0 = no quality stream (task 1), 1 = qualities of every read, 2 = qualities only of reads
 * whose reference span contains a lowercase draft base — the only reads task 2 can consult (every
 * k-mer window contains a lowercase column and its candidate reads span it); the engine reports an
 * error if a window candidate lacks its qualities. */
np_shard* np_shard_load(const char* fasta, const char* bam, const char* const* names,
                        int32_t n_names, int32_t with_qual, int32_t threads);
void      np_shard_view_of(const np_shard* shard, np_shard_view* out);
const char* np_shard_contig_name(const np_shard* shard, int32_t i);
void      np_shard_free(np_shard* shard);
/* position of shard contig i in the requested (or FASTA) order; shards are in BAM tid order */
int32_t   np_shard_contig_rank(const np_shard* shard, int32_t i);
/* Algorithmic bytes of one task step over this shard:
 * sum over reads of (16 + 4*n_cigar + ceil(l_qseq/2)) [+ l_qseq for task 2] + 2 * sum(L). */
int64_t   np_shard_algorithmic_bytes(const np_shard* shard, int32_t task);

/* ---- device engine ---------------------------------------------------------------------- */
typedef struct np_engine np_engine;

np_engine*  np_engine_create(int32_t device);   /* NULL + np_last_error() when CUDA is unusable */
void        np_engine_destroy(np_engine* e);
const char* np_last_error(void);

/* Copy a shard (host pointers, pageable or pinned) to HBM; replaces any resident shard. */
int32_t np_engine_upload(np_engine* e, const np_shard_view* host_shard);
/* Adopt a shard whose arrays already live in device memory (pointers are device pointers;
 * they must stay valid until the next upload/adopt). ctg_off / ctg_read_off are host arrays. */
int32_t np_engine_adopt_device(np_engine* e, const np_shard_view* dev_shard);

/* Run one task step on the resident shard (kernels only; asynchronous on the engine stream
 * until np_engine_sync / np_engine_download). */
int32_t np_engine_run(np_engine* e, int32_t task, const Configure* cfg);
int32_t np_engine_sync(np_engine* e);
/* Total polished length of the last run (sum over contigs), valid after sync. */
int64_t np_engine_result_bytes(np_engine* e);
/* Copy the polished sequences to host: out_seq receives the concatenation (no separators),
 * out_off[n_contigs+1] the per-contig offsets. */
int32_t np_engine_download(np_engine* e, uint8_t* out_seq, int64_t out_cap, int64_t* out_off);
/* Copy the concatenated result into another DEVICE buffer (asynchronous on the engine stream);
 * used for the on-device gather of the corrected FASTA across GPUs. */
int32_t np_engine_copy_result(np_engine* e, void* dst_device, int64_t dst_cap);
/* The same with a 16-byte header in front (int64 byte count + padding), all written from device memory on the engine
 * stream: the per-rank buffer of the multi-GPU gather.  dst_cap >= result bytes + 16. */
int32_t np_engine_pack_result(np_engine* e, void* dst_device, int64_t dst_cap);
/* Device pointer of the concatenated result (for an on-device gather) and its offsets. */
const uint8_t* np_engine_result_device(np_engine* e);

/* PolishPoint trace of the last run (contig_get_contig's change list, contig.c:743-799; what nextpolish1.py -debug
 * prints): produced only when cfg->trace_polish_open is set.  Positions are contig-relative; off[n_contigs+1]. */
int64_t np_engine_point_count(np_engine* e);
int32_t np_engine_points(np_engine* e, PolishPoint* out, int64_t cap, int64_t* off);

/* Per-kernel device time (ms) of the last run, measured with CUDA events on the engine
 * stream; names[i] points to static strings. Returns the number of entries written. */
int32_t np_engine_kernel_times(np_engine* e, const char** names, float* ms, int32_t cap);
/* Per-kernel CUDA-event timing is off by default (two event records per launch cost ~0.2 ms per step);
 * switch it on before a run whose np_engine_kernel_times() you want to read. */
void    np_engine_set_timing(np_engine* e, int32_t on);
/* Number of kernel launches issued by the last np_engine_run. */
int32_t np_engine_launch_count(np_engine* e);
int64_t np_engine_launch_total(np_engine* e);      /* kernel launches of every successful run of this engine */
/* Fused pileup-scan kernel statistics of the last task-1 run:
 * {window size W, windows, shared-memory bytes per CTA, windows with unresolved stretches,
 *  columns handled by the general (global-memory) kernels}. */
int32_t np_engine_window_stats(np_engine* e, int32_t* out5);
/* The stream the engine launches on (cudaStream_t as void*). */
void*   np_engine_stream(np_engine* e);

/* End-to-end convenience: upload + run + download in one call (host buffers in and out). */
int32_t np_polish_host(np_engine* e, int32_t task, const np_shard_view* host_shard,
                       const Configure* cfg, uint8_t* out_seq, int64_t out_cap, int64_t* out_off);


/* Streaming front end (double-buffered): jobs are submitted in order; a job's host-to-device copy is
 * enqueued at submission on its own slot (engine, stream, buffers) and overlaps the kernels of the jobs
 * submitted before it.  `depth` = slots (2 = one job computing while the next one uploads).
 * host_shard / out_seq / out_off must stay valid until np_stream_wait(ticket) returned; cfg is copied.
 * np_stream_submit returns a ticket >= 0 or a negative NP_ERR_*; when every slot is busy it first
 * finishes the oldest job.  np_stream_wait finishes every job up to `ticket` and returns its status. */
typedef struct np_stream np_stream;
np_stream* np_stream_create(int32_t device, int32_t depth);
void       np_stream_destroy(np_stream* s);
int64_t    np_stream_submit(np_stream* s, int32_t task, const np_shard_view* host_shard, const Configure* cfg,
                            uint8_t* out_seq, int64_t out_cap, int64_t* out_off);
int32_t    np_stream_wait(np_stream* s, int64_t ticket);
int64_t    np_stream_launch_count(np_stream* s);   /* kernel launches of every finished job */

/* Shards that already sit in HBM, polished by `slots` engines at once (resident_slots.cu): contigs are the reference's
 * parallel unit (one worker process per contig, nextpolish1.py:219-224), so shards are independent and a slot — engine,
 * stream, scratch, host thread — per job in flight fills the GPU that one chain of short kernels leaves idle.
 * np_resident_submit gives the job to slot `ticket % slots` (which must be free: wait for ticket - slots first) and
 * returns; the result lands in dst_device in the gather form of np_engine_pack_result (16-byte header holding the byte
 * count, then the polished bytes).  dev_shard's arrays must stay valid until np_resident_wait(ticket) returned. */
typedef struct np_resident np_resident;
np_resident* np_resident_create(int32_t device, int32_t slots);
void         np_resident_destroy(np_resident* p);
int64_t      np_resident_submit(np_resident* p, int32_t task, const np_shard_view* dev_shard, const Configure* cfg,
                                void* dst_device, int64_t dst_cap);
int32_t      np_resident_wait(np_resident* p, int64_t ticket, int64_t* out_bytes);
int64_t      np_resident_launch_count(np_resident* p);   /* kernel launches of every job so far */


/* ---- GPU inflate of BGZF (replaces htslib bgzf.c -> zlib inflate on the host, the step
 * immediately before the polishing path: contig.c:170-180,688-704 decode the BAM region twice per contig) ----
 * Inflates a BGZF byte range (concatenated blocks: a whole BAM file or one contig's chunk) with one warp per
 * block: headers are parsed on the host, compressed bytes go to HBM, the inflated bytes come back.
 * out == NULL: size query (out_bytes, n_blocks only).  kernel_ms (optional): device time of the kernel alone.
 * CRC32 of the blocks is not verified on the device (ISIZE is). */
int32_t np_bgzf_inflate(int32_t device, const uint8_t* comp, int64_t comp_bytes, uint8_t* out, int64_t out_cap,
                        int64_t* out_bytes, int32_t* n_blocks, float* kernel_ms);


/* ---- device-side shard construction: FASTA + BAM (+ .bai) -> packed shard built in HBM ----
 * Replaces np_shard_load's host inflate + packing (the reference's contig_init + bam_itr_queryi / sam_itr_next /
 * bam_read1 loop, contig.c:32-79,170-180,688-704): the compressed byte range of the wanted contigs is copied to
 * the GPU as it is on disk; BGZF inflate, record-boundary discovery (chains between the record starts the .bai
 * index knows, each verified to land on the next one), field extraction and packing run as kernels.  The shard
 * is byte-identical to np_shard_load's.  Needs <bam>.bai; returns NULL + np_last_error() otherwise (callers fall
 * back to np_shard_load).  The view holds DEVICE pointers (ctg_off / ctg_read_off are host arrays): pass it to
 * np_engine_adopt_device; it stays valid until np_dev_shard_free. */
typedef struct np_dev_shard np_dev_shard;
np_dev_shard* np_shard_load_gpu(int32_t device, const char* fasta, const char* bam, const char* const* names,
                                int32_t n_names, int32_t with_qual);
/* the same load for callers that already hold the draft in host memory: names[i] has the bases seq[i][0 .. len[i]) */
np_dev_shard* np_shard_load_gpu_seqs(int32_t device, const char* bam, const char* const* names, const uint8_t* const* seq,
                                     const int64_t* len, int32_t n_names, int32_t with_qual);
void        np_dev_shard_view(const np_dev_shard* shard, np_shard_view* out);
const char* np_dev_shard_contig_name(const np_dev_shard* shard, int32_t i);
int32_t     np_dev_shard_contig_rank(const np_dev_shard* shard, int32_t i);
/* sizes5: {record bytes, quality bytes, draft bytes, compressed bytes shipped, inflated bytes}; ms2: {inflate kernel,
 * whole load on the device} */
void        np_dev_shard_stats(const np_dev_shard* shard, int64_t* sizes5, float* ms2);
int32_t     np_dev_shard_download(const np_dev_shard* shard, uint8_t* ctg_seq, uint32_t* rec_off, uint8_t* rec,
                                  uint32_t* qual_off, uint8_t* qual);
void        np_dev_shard_free(np_dev_shard* shard);


/* ---- from-files front end (pipelined): the batch form of the reference ABI's per-contig calls (score_chain / kmer_count,
 * scorechain.c:3-15, kmercount.c:93-126) and of main.c:12-26 — draft FASTA + coordinate-sorted BGZF BAM (+ .bai) in,
 * polished sequences in FASTA order out.  A pipeline owns `depth` workers (engine + host thread) that pull jobs from a
 * queue: a job's file reads, upload, inflate and unpack overlap the kernels of the others.  np_files_submit returns a
 * ticket >= 0 or a negative NP_ERR_*; at most 2 x `depth` tickets may be outstanding (submitted and not yet waited for).
 * The arrays of a result stay valid until the np_files_submit that reuses its record (ticket + 2 x depth) or
 * np_files_destroy. */
typedef struct np_files np_files;
typedef struct {
    int32_t            task;
    int32_t            n_contigs;
    const char* const* names;      /* [n_contigs], FASTA order                                   */
    const uint8_t*     seq;        /* polished bytes of all contigs (pinned host memory)         */
    const int64_t*     start;      /* [n_contigs] offset of contig i in seq                      */
    const int64_t*     len;        /* [n_contigs] polished length of contig i                    */
    int64_t            h2d_bytes;  /* compressed BAM range + draft bases copied host -> device   */
    int64_t            d2h_bytes;  /* polished bytes + offsets copied device -> host             */
    float              load_ms, polish_ms;   /* host wall clock: shard construction / adopt + kernels + download */
} np_files_result;
np_files* np_files_create(int32_t device, int32_t depth);
void      np_files_destroy(np_files* p);
int64_t   np_files_submit(np_files* p, int32_t task, const char* fasta, const char* bam, const Configure* cfg);
int32_t   np_files_wait(np_files* p, int64_t ticket, np_files_result* out);


/* ---- one input on several GPUs of one box: the contig list of ONE draft is cut into contiguous blocks of
 * balanced cumulative length (the reference's driver cuts its worker jobs the same way: blc_genome,
 * source/nextPolish:93-117, consumed by nextpolish1.py -b/-i:148-161); every GPU owns a contiguous range of blocks and
 * works through it with a few pipelined slots (NEXTPOLISH_B200_SLOTS, default 3: load / inflate of one block overlap the
 * kernels of the others), appending the polished bytes to its result buffer in HBM; at the end the path's only collective
 * gathers the result buffers on the first GPU (grouped ncclSend / ncclRecv, exact sizes) and they are downloaded once.
 * A block never exceeds NEXTPOLISH_B200_BLOCK_MBP million draft bases (default 8), so genome size is bounded by the
 * result buffers (1 B per base), not by one shard's 2^31 limits or by HBM.  Single process, ncclCommInitAll; NCCL is bound
 * at run time.  In the result, load_ms = wall clock of parse + all blocks, polish_ms = gather + download.
 * devices == NULL: GPUs 0 .. n_devices-1.  The result arrays (FASTA order) stay valid until the next np_multi_run. */
typedef struct np_multi np_multi;
np_multi* np_multi_create(const int32_t* devices, int32_t n_devices);
int32_t   np_multi_run(np_multi* m, int32_t task, const char* fasta, const char* bam, const Configure* cfg, np_files_result* out);
/* the same run restricted to the contigs `names` (a worker's block, nextpolish1.py -b/-i :148-161); n_names < 0 = all.
 * A name the draft does not hold is NP_ERR_ARG.  The result lists the selected contigs in FASTA order. */
int32_t   np_multi_run_names(np_multi* m, int32_t task, const char* fasta, const char* bam, const Configure* cfg,
                             const char* const* names, int32_t n_names, np_files_result* out);
void      np_multi_destroy(np_multi* m);
/* part[i] = block (0 .. n_parts-1) of contig i: contiguous blocks with balanced cumulative length */
void      np_partition_contiguous(const int64_t* lengths, int32_t n_contigs, int32_t n_parts, int32_t* part);
int32_t   np_engine_result_offsets(np_engine* e, int64_t* out_off);


/* ---- the worker's file conventions either side of the path (host only, no GPU needed) --------------
 * Which contigs one worker job polishes and where its output part resumes — nextpolish1.py:148-179,203-210:
 *   block == NULL / "" or index == NULL / "all": every header of `genome`; otherwise the lines "name<TAB>index" of the
 *   block file (written by the driver, source/nextPolish:93-117) whose index field equals `index`;
 *   an existing out_path (not NULL / "stdout") is scanned: contigs of finished records are dropped from the plan, the
 *   last record (possibly partial) is polished again and resume_offset is the byte offset at which it starts. */
typedef struct np_part_plan np_part_plan;
np_part_plan* np_part_plan_create(const char* genome, const char* block, const char* index, const char* out_path);
void          np_part_plan_destroy(np_part_plan* p);
int32_t       np_part_plan_count(const np_part_plan* p);            /* contigs still to polish                  */
const char*   np_part_plan_name(const np_part_plan* p, int32_t i);  /* block-file (or FASTA) order              */
int32_t       np_part_plan_finished(const np_part_plan* p);         /* finished contigs found in out_path       */
int64_t       np_part_plan_resume_offset(const np_part_plan* p);
/* record name of nextpolish1.py:227: name + "_np<task>", or name + "<task>" when its last '_' field starts with "np".
 * Returns the length written, -1 when cap is too small. */
int32_t       np_part_record_name(const char* name, int32_t task, char* out, int32_t cap);
/* the output part: NULL / "stdout" = stdout; an existing file is cut at resume_offset and appended to; else created.
 * np_part_write prints one record as nextpolish1.py:228 does: ">name_np<task> <len>\nSEQ\n" (uppercase != 0: -u). */
typedef struct np_part_file np_part_file;
np_part_file* np_part_open(const char* out_path, int64_t resume_offset);
int32_t       np_part_write(np_part_file* f, const char* name, int32_t task, const uint8_t* seq, int64_t len, int32_t uppercase);
int32_t       np_part_close(np_part_file* f);


/* ---- seeded synthetic inputs (draft FASTA + coordinate-sorted BAM), for bench and tests -- */
typedef struct {
    uint64_t seed;
    int32_t  n_contigs;
    int64_t  contig_len;        /* every contig has this length (min_len==0) ...            */
    int64_t  min_len, max_len;  /* ... or log-uniform lengths in [min_len, max_len]          */
    double   depth;             /* e.g. 30                                                    */
    int32_t  read_len;          /* e.g. 150                                                   */
    double   draft_snv, draft_indel;   /* draft error rates vs truth (0.001 / 0.003)         */
    double   read_sub, read_indel;     /* sequencing error rates (0.002 / 0.0001)            */
    double   lowercase_frac;    /* fraction of draft bases written lowercase (task-2 inputs) */
    int32_t  compress_level;    /* BGZF zlib level (0 = stored blocks)                       */
} np_synth_params;
int32_t np_synth_write(const np_synth_params* p, const char* fasta_path, const char* bam_path);
/* The same reads as np_synth_write, packed directly (contigs [contig_lo, contig_hi)). */
np_shard* np_synth_shard(const np_synth_params* p, int32_t contig_lo, int32_t contig_hi,
                         int32_t with_qual, int32_t threads);

#ifdef __cplusplus
}
#endif
#endif /* NEXTPOLISH_B200_H */
