// cli_main.cpp — native CLI with the argument grammar of the reference's main.c:12-75:
//   nextpolish1 scorechain <fasta> <sgs.bam>   > out.fa
//   nextpolish1 kmercount  <fasta> <sgs.bam>   > out.fa
//   nextpolish1 snpvalid   <fasta> <sgs.bam>   > out.fa
// Output format of contig_write_to_file (contig.c:1050): ">name_<step>\nSEQ\n", contigs in FASTA
// order; "total time" trace on stderr (contig.c:1116).  Unlike the reference, all contigs are
// polished in ONE batch on the GPU (np_* batch ABI) instead of one call per contig.
// Extra command (not in the reference): simulate — seeded synthetic draft + BAM.
//
// Worker grammar (argv[1] starts with '-'): the command line of the reference's per-step worker, lib/nextpolish1.py
// (:252-334), so that the driver's job lines (source/nextPolish:87-90) can call this binary directly:
//   nextpolish1 -g genome.fa -t 1 -s sgs.sort.bam [-b input.genome.fasta.blc -i 0] [-o part000.fasta] [-u] [-debug] [flags]
// (--plan, ours: print what the job would polish and where its output part resumes, then exit without touching a device)
// with the block file / resume / ">name_np<task> <len>" conventions of nextpolish1.py:148-179,226-229 (part_writer.cpp).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <unistd.h>
#include <string>
#include <vector>
#include "../../include/nextpolish_b200.h"

static int usage(const char* a0) {
    printf("Usage: %s <command> [options]\n\nCommands:\n"
           "\tscorechain\t\tscore chain run\n\t\t\t\teg. scorechain fastafn sgsbamf > output.fa\n"
           "\tkmercount\t\tkmer count run\n\t\t\t\teg. kmercount fastafn sgsbamf > output.fa\n"
           "\tsnpvalid\t\tsnp valid run\n\t\t\t\teg. snpvalid fastafn sgsbamf > output.fa\n"
           "\tsimulate\t\twrite a seeded synthetic draft + sorted BAM\n"
           "\t\t\t\teg. simulate out.fa out.bam n_contigs contig_len depth seed [lowercase_frac]\n\n", a0);
    return 0;
}

// ---- worker grammar -----------------------------------------------------------------------------------------------
static long long parse_count(const char* v) {          // parse_num_unit of the reference's kit.py: 150k, 2m, 1g
    char* end = nullptr;
    double x = strtod(v, &end);
    if (end && (*end == 'k' || *end == 'K')) x *= 1e3;
    else if (end && (*end == 'm' || *end == 'M')) x *= 1e6;
    else if (end && (*end == 'g' || *end == 'G')) x *= 1e9;
    return (long long)x;
}

static int worker_main(int argc, char* argv[]) {
    const char *genome = nullptr, *sgs = nullptr, *lgs = nullptr, *block = nullptr, *index = "all", *outp = "stdout";
    int task = 0, upper = 0, debug = 0, plan_only = 0;
    struct Opt { const char* name; double val; bool set; };
    Opt opts[] = {{"count_read_ins_sgs", 0, false}, {"min_map_quality", 0, false}, {"max_ins_len_sgs", 0, false}, {"max_ins_fold_sgs", 0, false},
                  {"max_clip_ratio_sgs", 0, false}, {"max_clip_ratio_lgs", 0, false}, {"trim_len_edge", 0, false}, {"ext_len_edge", 0, false},
                  {"indel_balance_factor_sgs", 0, false}, {"min_count_ratio_skip", 0, false}, {"min_len_ldr", 0, false}, {"max_len_kmer", 0, false},
                  {"min_len_inter_kmer", 0, false}, {"max_count_kmer", 0, false}, {"ploidy", 0, false}, {"max_variant_count_lgs", 0, false},
                  {"indel_balance_factor_lgs", 0, false}, {"min_depth_snp", 0, false}, {"min_count_snp", 0, false}, {"min_count_snp_link", 0, false},
                  {"max_indel_factor_lgs", 0, false}, {"max_snp_factor_lgs", 0, false}, {"min_snp_factor_sgs", 0, false}};
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "%s: option %s needs a value\n", argv[0], a.c_str()); exit(2); } return argv[++i]; };
        if (a == "-g" || a == "--genome") genome = val();
        else if (a == "-s" || a == "--bam_sgs") sgs = val();
        else if (a == "-l" || a == "--bam_lgs") lgs = val();
        else if (a == "-b" || a == "--block") block = val();
        else if (a == "-i" || a == "--block_index") index = val();
        else if (a == "-o" || a == "--out") outp = val();
        else if (a == "-t" || a == "--task") task = atoi(val());
        else if (a == "-p" || a == "--process") (void)val();        // host processes of the reference's Pool: one GPU batch here
        else if (a == "-u" || a == "--uppercase") upper = 1;
        else if (a == "-debug") debug = 1;
        else if (a == "--plan") plan_only = 1;                       // not in the reference: print the job's plan, touch no device
        else {
            bool known = false;
            for (Opt& o : opts)
                if (a == std::string("-") + o.name) { const char* v = val(); o.val = strcmp(o.name, "max_variant_count_lgs") == 0 ? (double)parse_count(v) : atof(v); o.set = true; known = true; break; }
            if (!known) { fprintf(stderr, "%s: unrecognized argument %s\n", argv[0], a.c_str()); return 2; }
        }
    }
    if (!genome || task < 1 || task > 5) { fprintf(stderr, "usage: %s -g genome.fa -t {1,2,3,4,5} -s sgs.sort.bam [-b blockfile -i index] [-o out] [-u] [-debug]\n", argv[0]); return 2; }
    if (task == 3 || task == 5 || !sgs) {                            // nextpolish1.py:338-340 refuses task 5 the same way
        fprintf(stderr, "task %d is outside this engine (score_chain, kmer_count and snp_valid run on the GPU): use the reference worker for it\n", task);
        return 1;
    }
    time_t t0 = time(nullptr);
    np_part_plan* plan = np_part_plan_create(genome, block, index, outp);
    if (!plan) { fprintf(stderr, "%s\n", np_last_error()); return 1; }
    if (np_part_plan_finished(plan)) fprintf(stderr, "skip %d polished seqs found in %s\n", np_part_plan_finished(plan), outp);
    if (plan_only) {
        printf("task\t%d\nfinished\t%d\nresume_offset\t%lld\n", task, np_part_plan_finished(plan), (long long)np_part_plan_resume_offset(plan));
        for (int i = 0; i < np_part_plan_count(plan); i++) printf("polish\t%s\n", np_part_plan_name(plan, i));
        np_part_plan_destroy(plan);
        return 0;
    }
    np_part_file* out = np_part_open(outp, np_part_plan_resume_offset(plan));
    if (!out) { fprintf(stderr, "%s\n", np_last_error()); return 1; }
    const int n_names = np_part_plan_count(plan);
    // config_init first, the option values afterwards: read_tlen keeps the estimate made with the defaults, exactly like
    // update_cfg after config_init in nextpolish1.py:219-221
    Configure* cfg = config_init(genome, sgs, lgs);
    for (const Opt& o : opts) {
        if (!o.set) continue;
        const std::string k = o.name;
#define NP_SET(field, type) if (k == #field) cfg->field = (type)o.val;
        NP_SET(count_read_ins_sgs, uint32_t) NP_SET(min_map_quality, uint8_t) NP_SET(max_ins_len_sgs, uint32_t) NP_SET(max_ins_fold_sgs, int32_t)
        NP_SET(max_clip_ratio_sgs, double) NP_SET(max_clip_ratio_lgs, double) NP_SET(trim_len_edge, uint8_t) NP_SET(ext_len_edge, uint8_t)
        NP_SET(indel_balance_factor_sgs, double) NP_SET(min_count_ratio_skip, double) NP_SET(min_len_ldr, uint8_t) NP_SET(max_len_kmer, uint8_t)
        NP_SET(min_len_inter_kmer, uint8_t) NP_SET(max_count_kmer, uint8_t) NP_SET(ploidy, double) NP_SET(max_variant_count_lgs, int32_t)
        NP_SET(indel_balance_factor_lgs, double) NP_SET(min_depth_snp, uint8_t) NP_SET(min_count_snp, uint8_t) NP_SET(min_count_snp_link, int8_t)
        NP_SET(max_indel_factor_lgs, double) NP_SET(max_snp_factor_lgs, double) NP_SET(min_snp_factor_sgs, double)
#undef NP_SET
    }
    cfg->region_count = 10000;                                       // nextpolish1.py:124
    cfg->trace_polish_open = debug ? 1 : 0;                          // nextpolish1.py:133
    if (n_names > 0) {
        std::vector<const char*> names((size_t)n_names);
        for (int i = 0; i < n_names; i++) names[(size_t)i] = np_part_plan_name(plan, i);
        const int dev = getenv("NEXTPOLISH_B200_DEVICE") ? atoi(getenv("NEXTPOLISH_B200_DEVICE")) : 0;
        int gpus = getenv("NEXTPOLISH_B200_GPUS") ? atoi(getenv("NEXTPOLISH_B200_GPUS")) : 1;
        if (gpus < 1) gpus = 1;
        const std::string bai = std::string(sgs) + ".bai";
        const char* hl = getenv("NEXTPOLISH_B200_HOST_LOAD");
        const bool indexed = access(bai.c_str(), R_OK) == 0 && !(hl && hl[0] == '1');
        if (indexed && !debug) {
            // block rounds under the shard budget, pipelined slots per GPU, one gather (multi_gpu.cu)
            const int32_t one[1] = {dev};
            fflush(stdout);
            const int saved = dup(1);                                // NCCL's banner must not land in a FASTA on stdout
            dup2(2, 1);
            np_multi* m = np_multi_create(gpus == 1 ? one : nullptr, gpus);
            np_files_result r;
            const bool ok = m && np_multi_run_names(m, task, genome, sgs, cfg, names.data(), n_names, &r) == NP_OK;
            fflush(stdout);
            dup2(saved, 1);
            close(saved);
            if (!ok) { fprintf(stderr, "%s\n", np_last_error()); return 1; }
            for (int i = 0; i < r.n_contigs; i++)
                if (np_part_write(out, r.names[i], task, r.seq + r.start[i], r.len[i], upper) != NP_OK) { fprintf(stderr, "%s\n", np_last_error()); return 1; }
            np_multi_destroy(m);
        } else {
            // one shard (the PolishPoint trace of -debug lives in the engine; BAMs without an index go through the host packer)
            const int wq = task == 2 ? 2 : task == 4 ? 1 : 0;
            np_dev_shard* ds = indexed ? np_shard_load_gpu(dev, genome, sgs, names.data(), n_names, wq) : nullptr;
            np_shard* sh = nullptr;
            np_shard_view v;
            if (ds) np_dev_shard_view(ds, &v);
            else {
                sh = np_shard_load(genome, sgs, names.data(), n_names, wq, 8);
                if (!sh) { fprintf(stderr, "%s\n", np_last_error()); return 1; }
                np_shard_view_of(sh, &v);
            }
            np_engine* e = np_engine_create(dev);
            if (!e || (ds ? np_engine_adopt_device(e, &v) : np_engine_upload(e, &v)) != NP_OK || np_engine_run(e, task, cfg) != NP_OK) {
                fprintf(stderr, "%s\n", np_last_error());
                return 1;
            }
            const int64_t n = np_engine_result_bytes(e);
            std::vector<uint8_t> seq((size_t)n + 1);
            std::vector<int64_t> off((size_t)v.n_contigs + 1), poff((size_t)v.n_contigs + 1, 0);
            if (np_engine_download(e, seq.data(), n + 1, off.data()) != NP_OK) { fprintf(stderr, "%s\n", np_last_error()); return 1; }
            std::vector<PolishPoint> pts;
            if (debug) {
                pts.resize((size_t)np_engine_point_count(e) + 1);
                if (np_engine_points(e, pts.data(), (int64_t)pts.size(), poff.data()) != NP_OK) { fprintf(stderr, "%s\n", np_last_error()); return 1; }
            }
            std::vector<int> slot_of_rank((size_t)n_names, -1);
            for (int i = 0; i < v.n_contigs; i++) slot_of_rank[(size_t)(ds ? np_dev_shard_contig_rank(ds, i) : np_shard_contig_rank(sh, i))] = i;
            for (int r = 0; r < n_names; r++) {
                const int i = slot_of_rank[(size_t)r];
                if (i < 0) { fprintf(stderr, "contig %s is not in %s\n", names[(size_t)r], genome); return 1; }
                if (np_part_write(out, names[(size_t)r], task, seq.data() + off[(size_t)i], off[(size_t)i + 1] - off[(size_t)i], upper) != NP_OK) { fprintf(stderr, "%s\n", np_last_error()); return 1; }
                for (int64_t k = poff[(size_t)i]; debug && k < poff[(size_t)i + 1]; k++)          // nextpolish1.py:230-231
                    fprintf(stderr, "%s %d %d %c %c\n", names[(size_t)r], pts[(size_t)k].pos, pts[(size_t)k].index, pts[(size_t)k].curbase, pts[(size_t)k].base);
            }
            np_engine_destroy(e);
            if (ds) np_dev_shard_free(ds);
            if (sh) np_shard_free(sh);
        }
    }
    if (np_part_close(out) != NP_OK) { fprintf(stderr, "%s\n", np_last_error()); return 1; }
    np_part_plan_destroy(plan);
    config_destory(cfg);
    fprintf(stderr, "total time:%lds\n", (long)(time(nullptr) - t0));
    return 0;
}

int main(int argc, char* argv[]) {
    if (argc < 2) return usage(argv[0]);
    if (argv[1][0] == '-' && strcmp(argv[1], "-h") != 0 && strcmp(argv[1], "--help") != 0) return worker_main(argc, argv);
    if (strcmp(argv[1], "simulate") == 0) {
        if (argc < 8) return usage(argv[0]);
        np_synth_params p; memset(&p, 0, sizeof p);
        p.n_contigs = atoi(argv[4]); p.contig_len = atoll(argv[5]); p.depth = atof(argv[6]);
        p.seed = strtoull(argv[7], nullptr, 10);
        p.read_len = 150; p.draft_snv = 0.001; p.draft_indel = 0.003; p.read_sub = 0.002; p.read_indel = 0.0001;
        p.lowercase_frac = argc > 8 ? atof(argv[8]) : 0; p.compress_level = 1;
        if (np_synth_write(&p, argv[2], argv[3]) != NP_OK) { fprintf(stderr, "%s\n", np_last_error()); return 1; }
        return 0;
    }
    int step = 0;
    if (strcmp(argv[1], "scorechain") == 0) step = 1;
    else if (strcmp(argv[1], "kmercount") == 0) step = 2;
    else if (strcmp(argv[1], "snpvalid") == 0) step = 4;
    else return usage(argv[0]);
    if (argc != 4) { printf("%s %s fastafn lgsbam\n", argv[0], argv[1]); return 0; }
    time_t t0 = time(nullptr);
    Configure* cfg = config_init(argv[2], argv[3], nullptr);
    const int dev = getenv("NEXTPOLISH_B200_DEVICE") ? atoi(getenv("NEXTPOLISH_B200_DEVICE")) : 0;
    int gpus = getenv("NEXTPOLISH_B200_GPUS") ? atoi(getenv("NEXTPOLISH_B200_GPUS")) : 1;
    if (gpus < 1) gpus = 1;
    const char* hl0 = getenv("NEXTPOLISH_B200_HOST_LOAD");
    std::string bai = std::string(argv[3]) + ".bai";
    FILE* fb = fopen(bai.c_str(), "rb");
    if (fb) fclose(fb);
    if (fb && !(hl0 && hl0[0] == '1')) {
        // the indexed-BAM path: contiguous contig blocks under the block budget (genomes larger than one shard are polished
        // block by block, like the reference polishes contig by contig), pipelined slots per GPU, one NCCL gather at the
        // end when several GPUs are used (multi_gpu.cu)
        const int32_t one[1] = {dev};
        // stdout is the FASTA stream: whatever a library prints there while the GPUs work (NCCL's version banner) goes to stderr
        fflush(stdout);
        const int saved = dup(1);
        dup2(2, 1);
        np_multi* m = np_multi_create(gpus == 1 ? one : nullptr, gpus);
        np_files_result r;
        const bool ok = m && np_multi_run(m, step, argv[2], argv[3], cfg, &r) == NP_OK;
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
        if (!ok) { fprintf(stderr, "%s\n", np_last_error()); return 1; }
        for (int i = 0; i < r.n_contigs; i++) {
            printf(">%s_%d\n", r.names[i], step);
            fwrite(r.seq + r.start[i], 1, (size_t)r.len[i], stdout);
            fputc('\n', stdout);
        }
        np_multi_destroy(m);
        config_destory(cfg);
        fprintf(stderr, "total time:%lds\n", (long)(time(nullptr) - t0));
        return 0;
    }
    // the shard is built on the GPU (inflate + record unpack + packing, devload.cu) when the BAM has a .bai index;
    // otherwise (or with NEXTPOLISH_B200_HOST_LOAD=1) the host packer builds it and it is uploaded
    np_dev_shard* ds = nullptr;
    np_shard* sh = nullptr;
    const char* hl = getenv("NEXTPOLISH_B200_HOST_LOAD");
    if (!(hl && hl[0] == '1')) ds = np_shard_load_gpu(dev, argv[2], argv[3], nullptr, 0, step == 2 ? 2 : step == 4 ? 1 : 0);
    np_shard_view v;
    if (ds) np_dev_shard_view(ds, &v);
    else {
        sh = np_shard_load(argv[2], argv[3], nullptr, 0, step == 2 ? 2 : step == 4 ? 1 : 0, 8);
        if (!sh) { fprintf(stderr, "%s\n", np_last_error()); return 1; }
        np_shard_view_of(sh, &v);
    }
    np_engine* e = np_engine_create(dev);
    if (!e) { fprintf(stderr, "%s\n", np_last_error()); return 1; }
    if ((ds ? np_engine_adopt_device(e, &v) : np_engine_upload(e, &v)) != NP_OK || np_engine_run(e, step, cfg) != NP_OK) {
        fprintf(stderr, "%s\n", np_last_error());
        return 1;
    }
    int64_t n = np_engine_result_bytes(e);
    std::vector<uint8_t> out((size_t)n + 1);
    std::vector<int64_t> off((size_t)v.n_contigs + 1);
    if (np_engine_download(e, out.data(), n + 1, off.data()) != NP_OK) { fprintf(stderr, "%s\n", np_last_error()); return 1; }
    std::vector<int> slot_of_rank((size_t)v.n_contigs, -1);
    for (int i = 0; i < v.n_contigs; i++) slot_of_rank[(size_t)(ds ? np_dev_shard_contig_rank(ds, i) : np_shard_contig_rank(sh, i))] = i;
    for (int r = 0; r < v.n_contigs; r++) {
        int i = slot_of_rank[(size_t)r];
        printf(">%s_%d\n", ds ? np_dev_shard_contig_name(ds, i) : np_shard_contig_name(sh, i), step);
        fwrite(out.data() + off[(size_t)i], 1, (size_t)(off[(size_t)i + 1] - off[(size_t)i]), stdout);
        fputc('\n', stdout);
    }
    np_engine_destroy(e);
    if (ds) np_dev_shard_free(ds);
    if (sh) np_shard_free(sh);
    config_destory(cfg);
    fprintf(stderr, "total time:%lds\n", (long)(time(nullptr) - t0));
    return 0;
}
