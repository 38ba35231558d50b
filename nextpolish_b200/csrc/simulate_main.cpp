// simulate_main.cpp — standalone writer of the seeded synthetic inputs (draft FASTA + coordinate-sorted BAM).
// Linked from synth.o + hostio.o only: it does NOT load nextpolish1.so, so that bench.py's reference arm can
// produce its inputs without the product library ever being mapped into that process tree.
//   np_simulate out.fa out.bam key=value ...   (keys = fields of np_synth_params; defaults = SURVEY.md 8d profile)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "errors.h"
#include "../../include/nextpolish_b200.h"

namespace np {
static thread_local std::string g_err;
void set_error(const std::string& e) { g_err = e; }
const std::string& get_error() { return g_err; }
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s out.fa out.bam [seed=N n_contigs=N contig_len=N min_len=N max_len=N depth=F read_len=N "
                        "draft_snv=F draft_indel=F read_sub=F read_indel=F lowercase_frac=F compress_level=N]\n", argv[0]);
        return 2;
    }
    np_synth_params p; memset(&p, 0, sizeof p);
    p.seed = 1; p.n_contigs = 1; p.contig_len = 100000; p.depth = 30; p.read_len = 150;
    p.draft_snv = 0.001; p.draft_indel = 0.003; p.read_sub = 0.002; p.read_indel = 0.0001; p.compress_level = 1;
    for (int i = 3; i < argc; i++) {
        const char* eq = strchr(argv[i], '=');
        if (!eq) { fprintf(stderr, "bad argument %s\n", argv[i]); return 2; }
        std::string k(argv[i], eq - argv[i]); const char* v = eq + 1;
        if (k == "seed") p.seed = strtoull(v, nullptr, 10);
        else if (k == "n_contigs") p.n_contigs = atoi(v);
        else if (k == "contig_len") p.contig_len = atoll(v);
        else if (k == "min_len") p.min_len = atoll(v);
        else if (k == "max_len") p.max_len = atoll(v);
        else if (k == "depth") p.depth = atof(v);
        else if (k == "read_len") p.read_len = atoi(v);
        else if (k == "draft_snv") p.draft_snv = atof(v);
        else if (k == "draft_indel") p.draft_indel = atof(v);
        else if (k == "read_sub") p.read_sub = atof(v);
        else if (k == "read_indel") p.read_indel = atof(v);
        else if (k == "lowercase_frac") p.lowercase_frac = atof(v);
        else if (k == "compress_level") p.compress_level = atoi(v);
        else { fprintf(stderr, "unknown key %s\n", k.c_str()); return 2; }
    }
    if (np_synth_write(&p, argv[1], argv[2]) != NP_OK) { fprintf(stderr, "np_simulate: %s\n", np::get_error().c_str()); return 1; }
    return 0;
}
