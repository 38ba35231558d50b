// abi_host.cpp — host-only part of the C ABI: Configure / PolishResult lifetime, packed-shard
// loading, error string.  No compute lives here; the compute entry points are in engine.cu.
#include "hostio.h"
#include "errors.h"

#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace np {
static thread_local std::string g_err;
void set_error(const std::string& e) { g_err = e; }
const std::string& get_error() { return g_err; }
}  // namespace np

struct np_shard { np::Shard s; };

// NEXTPOLISH_B200_BACKTRACE=1: print a native backtrace on SIGSEGV / SIGABRT (debugging aid; addresses resolve with
// addr2line -e nextpolish1.so since the library is built with -lineinfo / symbols)
#include <execinfo.h>
#include <signal.h>
static void np_crash_handler(int sig) {
    void* frames[64];
    int n = backtrace(frames, 64);
    const char msg[] = "nextpolish_b200: fatal signal, native backtrace:\n";
    if (write(2, msg, sizeof msg - 1) < 0) {}
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}
__attribute__((constructor)) static void np_install_crash_handler() {
    const char* e = getenv("NEXTPOLISH_B200_BACKTRACE");
    if (e && e[0] == '1') { signal(SIGSEGV, np_crash_handler); signal(SIGABRT, np_crash_handler); signal(SIGBUS, np_crash_handler); }
}

extern "C" {

const char* np_last_error(void) { return np::get_error().c_str(); }

// ---- config.c:8-56 -----------------------------------------------------------------------
Configure* config_init(const char* fastafn, const char* bamfn, const char* thirdbamfn) {
    Configure* r = (Configure*)calloc(sizeof(Configure), 1);
    r->trim_len_edge = 2;
    r->ext_len_edge = 2;
    r->min_map_quality = 0;
    r->indel_balance_factor_sgs = 0.5;
    r->min_count_ratio_skip = 0.8;
    r->min_len_ldr = 3;
    r->min_len_inter_kmer = 5;
    r->max_len_kmer = 50;
    r->max_count_kmer = 50;
    r->min_depth_snp = 3;
    r->min_count_snp = 5;
    r->min_count_snp_link = 5;
    r->ploidy = 2;
    r->indel_balance_factor_lgs = 0.33;
    r->max_indel_factor_lgs = 0.21;
    r->max_snp_factor_lgs = 0.53;
    r->min_snp_factor_sgs = 0.34;
    r->region_count = 10000;
    r->count_read_ins_sgs = 10000;
    r->max_ins_len_sgs = 10000;
    r->max_ins_fold_sgs = 5;
    r->max_variant_count_lgs = 150000;
    r->max_clip_ratio_sgs = 0.15;
    r->max_clip_ratio_lgs = 0.4;
    r->trace_polish_open = 0;
    r->fastafn = fastafn ? strdup(fastafn) : nullptr;
    r->bamfn = (bamfn && access(bamfn, F_OK) == 0) ? strdup(bamfn) : nullptr;
    if (r->bamfn) {
        uint32_t mean = 0; int32_t rl = 0; std::string err;
        if (!np::bam_insert_estimate(r->bamfn, r->count_read_ins_sgs, r->max_ins_len_sgs, &mean, &rl, err)) {
            fprintf(stderr, "config_init: %s\n", err.c_str());
            exit(1);
        }
        r->read_len = rl;
        r->read_tlen = (int32_t)(mean * (uint32_t)r->max_ins_fold_sgs);   // config.c:46
    } else {
        r->read_tlen = 0;
    }
    r->thirdbamfn = (thirdbamfn && access(thirdbamfn, F_OK) == 0) ? strdup(thirdbamfn) : nullptr;
    return r;
}

void config_destory(Configure* c) {   // config.c:58-68
    if (!c) return;
    if (c->fastafn) free(c->fastafn);
    if (c->bamfn) free(c->bamfn);
    if (c->thirdbamfn) free(c->thirdbamfn);
    free(c);
}

PolishResult* polishresult_init(void) { return (PolishResult*)calloc(sizeof(PolishResult), 1); }   // contig.c:20-23
void polishresult_destory(PolishResult* p) {   // contig.c:25-30
    if (!p) return;
    if (p->contig) free(p->contig);
    if (p->data) free(p->data);
    free(p);
}

// ---- packed shards -------------------------------------------------------------------------
np_shard* np_shard_load(const char* fasta, const char* bam, const char* const* names,
                        int32_t n_names, int32_t with_qual, int32_t threads) {
    if (!fasta) { np::set_error("np_shard_load: fasta is NULL"); return nullptr; }
    std::vector<std::string> nm;
    for (int32_t i = 0; names && i < n_names; i++) nm.emplace_back(names[i]);
    np_shard* sh = new np_shard();
    std::string err;
    if (!np::shard_load(fasta, bam ? bam : "", nm, with_qual, threads > 0 ? threads : 1, sh->s, err)) {
        np::set_error("np_shard_load: " + err);
        delete sh;
        return nullptr;
    }
    return sh;
}
void np_shard_view_of(const np_shard* shard, np_shard_view* out) { shard->s.view(out); }
const char* np_shard_contig_name(const np_shard* shard, int32_t i) {
    if (i < 0 || (size_t)i >= shard->s.names.size()) return nullptr;
    return shard->s.names[(size_t)i].c_str();
}
void np_shard_free(np_shard* shard) { delete shard; }
int32_t np_shard_contig_rank(const np_shard* shard, int32_t i) {
    if (i < 0 || (size_t)i >= shard->s.fasta_rank.size()) return -1;
    return shard->s.fasta_rank[(size_t)i];
}
int64_t np_shard_algorithmic_bytes(const np_shard* shard, int32_t task) {
    int64_t G = shard->s.ctg_off.empty() ? 0 : shard->s.ctg_off.back();
    return shard->s.alg_bytes + 2 * G + (task == NP_TASK_KMER_COUNT ? shard->s.qual_bytes : 0);
}
np_shard* np_synth_shard(const np_synth_params* p, int32_t contig_lo, int32_t contig_hi,
                         int32_t with_qual, int32_t threads) {
    if (!p) { np::set_error("np_synth_shard: params is NULL"); return nullptr; }
    np_shard* sh = new np_shard();
    std::string err;
    if (!np::synth_shard(*p, contig_lo, contig_hi, with_qual, threads > 0 ? threads : 1, sh->s, err)) {
        np::set_error("np_synth_shard: " + err);
        delete sh;
        return nullptr;
    }
    return sh;
}

}  // extern "C"
