// cli_main.cpp — native CLI with the argument grammar of the reference's main.c:12-75:
//   nextpolish1 scorechain <fasta> <sgs.bam>   > out.fa
//   nextpolish1 kmercount  <fasta> <sgs.bam>   > out.fa
//   nextpolish1 snpvalid   <fasta> <sgs.bam>   > out.fa
// Output format of contig_write_to_file (contig.c:1050): ">name_<step>\nSEQ\n", contigs in FASTA
// order; "total time" trace on stderr (contig.c:1116).  Unlike the reference, all contigs are
// polished in ONE batch on the GPU (np_* batch ABI) instead of one call per contig.
// Extra command (not in the reference): simulate — seeded synthetic draft + BAM.
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

int main(int argc, char* argv[]) {
    if (argc < 2) return usage(argv[0]);
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
