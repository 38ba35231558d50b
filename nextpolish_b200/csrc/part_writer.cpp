// part_writer.cpp — the worker's file conventions either side of the hot path, host only:
//   * which contigs one worker job polishes: the block file written by the driver ("name<TAB>index" lines,
//     source/nextPolish:93-117) filtered by -i, or every header of the draft (nextpolish1.py:148-161);
//   * resume: an existing output part is scanned, every finished record's contig is skipped and the file is cut at the
//     start of its last (possibly partial) record, which is polished again (nextpolish1.py:163-179,203-210);
//   * the record header: ">name_np<task> <length>" — a name that already ends in an "_np…" field gets the task digit
//     appended instead (nextpolish1.py:226-229).
// Used by the native CLI's worker grammar (cli_main.cpp) and pinned against the Python rules in tests/test_part_writer.py.
#include <cstdio>
#include <cstring>
#include <unistd.h>
#include <string>
#include <unordered_set>
#include <vector>

#include "errors.h"
#include "../../include/nextpolish_b200.h"

struct np_part_plan {
    std::vector<std::string> names;      // contigs still to polish, block-file / FASTA order
    std::unordered_set<std::string> done;
    int64_t resume_offset = 0;           // bytes of the output part that hold finished records
    int32_t n_done = 0;
};

namespace {
// line reader without a length limit (FASTA lines of unwrapped contigs are hundreds of megabytes)
bool next_line(FILE* f, std::string& line) {
    line.clear();
    char buf[1 << 16];
    while (fgets(buf, sizeof buf, f)) {
        const size_t n = strlen(buf);
        line.append(buf, n);
        if (n && buf[n - 1] == '\n') return true;
    }
    return !line.empty();
}
// Python's str.split()[k] on ASCII whitespace
std::vector<std::string> fields(const std::string& s, size_t max_fields) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size() && out.size() < max_fields) {
        while (i < s.size() && strchr(" \t\r\n\v\f", s[i])) i++;
        size_t j = i;
        while (j < s.size() && !strchr(" \t\r\n\v\f", s[j])) j++;
        if (j > i) out.emplace_back(s, i, j - i);
        i = j;
    }
    return out;
}
std::string before_np(const std::string& name) {      // name.split('_np')[0]
    const size_t p = name.find("_np");
    return p == std::string::npos ? name : name.substr(0, p);
}
}  // namespace

extern "C" {

np_part_plan* np_part_plan_create(const char* genome, const char* block, const char* index, const char* out_path) {
    if (!genome) { np::set_error("np_part_plan_create: no draft"); return nullptr; }
    np_part_plan* p = new np_part_plan();
    std::string line;
    // ---- resume scan (read_polished_seqs, nextpolish1.py:163-179)
    if (out_path && strcmp(out_path, "stdout") != 0) {
        if (FILE* f = fopen(out_path, "rb")) {
            std::string last;
            bool any = false;
            int64_t cur = 0;
            while (next_line(f, line)) {
                if (line[0] == '>') {
                    p->resume_offset += cur;
                    cur = (int64_t)line.size();
                    auto fs = fields(line, 1);
                    last = before_np(fs.empty() ? std::string(">") : fs[0]).substr(1);
                    any = true;
                    p->done.insert(last);
                } else cur += (int64_t)line.size();
            }
            fclose(f);
            if (any) p->done.erase(last);              // the last record may be partial: polished again
            p->n_done = (int32_t)p->done.size();
        }
    }
    // ---- the job's contigs (read_unpolished_seqs, nextpolish1.py:148-161)
    const bool all = !block || !block[0] || !index || strcmp(index, "all") == 0;
    const char* path = all ? genome : block;
    FILE* f = fopen(path, "rb");
    if (!f) { np::set_error(std::string("np_part_plan_create: cannot open ") + path); delete p; return nullptr; }
    std::unordered_set<std::string> seen;              // the reference collects a set
    while (next_line(f, line)) {
        if (all) {
            if (line[0] != '>') continue;
            auto fs = fields(line, 1);
            if (fs.empty() || fs[0].size() < 2) continue;
            std::string nm = fs[0].substr(1);
            // the reference's "all" branch does not consult the finished set (nextpolish1.py:157-160); skipping finished
            // contigs here as well is what makes -o resumable without a block file (a rerun would otherwise append
            // duplicates) — the mirror in nextpolish_b200/nextpolish1.py does the same
            if (p->done.count(before_np(nm)) || !seen.insert(nm).second) continue;
            p->names.push_back(std::move(nm));
        } else {
            auto fs = fields(line, 3);
            if (fs.size() < 2 || fs[1] != index) continue;
            if (p->done.count(before_np(fs[0])) || !seen.insert(fs[0]).second) continue;
            p->names.push_back(fs[0]);
        }
    }
    fclose(f);
    return p;
}

void np_part_plan_destroy(np_part_plan* p) { delete p; }
int32_t np_part_plan_count(const np_part_plan* p) { return p ? (int32_t)p->names.size() : 0; }
const char* np_part_plan_name(const np_part_plan* p, int32_t i) {
    return (p && i >= 0 && (size_t)i < p->names.size()) ? p->names[(size_t)i].c_str() : nullptr;
}
int32_t np_part_plan_finished(const np_part_plan* p) { return p ? p->n_done : 0; }
int64_t np_part_plan_resume_offset(const np_part_plan* p) { return p ? p->resume_offset : 0; }

// ">name_np<task>" naming rule; returns the length written (without the NUL) or -1 when cap is too small
int32_t np_part_record_name(const char* name, int32_t task, char* out, int32_t cap) {
    if (!name || !out) return -1;
    const std::string nm(name);
    const size_t us = nm.rfind('_');
    const std::string lastf = us == std::string::npos ? nm : nm.substr(us + 1);      // name.split('_')[-1]
    std::string r = nm + (lastf.compare(0, 2, "np") == 0 ? std::to_string(task) : "_np" + std::to_string(task));
    if ((int32_t)r.size() + 1 > cap) return -1;
    memcpy(out, r.c_str(), r.size() + 1);
    return (int32_t)r.size();
}

// The output part: stdout, or the file cut at resume_offset (records after it are rewritten), or a new file.
struct np_part_file { FILE* f = nullptr; bool own = false; };

np_part_file* np_part_open(const char* out_path, int64_t resume_offset) {
    np_part_file* pf = new np_part_file();
    if (!out_path || strcmp(out_path, "stdout") == 0) { pf->f = stdout; return pf; }
    FILE* f = fopen(out_path, "r+b");
    if (f) {
        fflush(f);
        if (ftruncate(fileno(f), (off_t)resume_offset) != 0 || fseeko(f, (off_t)resume_offset, SEEK_SET) != 0) {
            np::set_error(std::string("np_part_open: cannot cut ") + out_path); fclose(f); delete pf; return nullptr;
        }
    } else f = fopen(out_path, "wb");
    if (!f) { np::set_error(std::string("np_part_open: cannot open ") + out_path); delete pf; return nullptr; }
    pf->f = f; pf->own = true;
    return pf;
}

int32_t np_part_close(np_part_file* pf) {
    if (!pf) return NP_OK;
    int rc = pf->own ? fclose(pf->f) : fflush(pf->f);
    delete pf;
    if (rc != 0) { np::set_error("np_part_close: write failed"); return NP_ERR_IO; }
    return NP_OK;
}

// One record as nextpolish1.py:228 prints it: ">%s %d\n%s\n"; uppercase != 0 is -u
int32_t np_part_write(np_part_file* pf, const char* name, int32_t task, const uint8_t* seq, int64_t len, int32_t uppercase) {
    char tag[4096];
    FILE* f = pf ? pf->f : nullptr;
    if (!f || (len > 0 && !seq) || np_part_record_name(name, task, tag, (int32_t)sizeof tag) < 0) { np::set_error("np_part_write: bad arguments"); return NP_ERR_ARG; }
    bool ok = fprintf(f, ">%s %lld\n", tag, (long long)len) >= 0;
    if (!uppercase) ok = ok && (len == 0 || fwrite(seq, 1, (size_t)len, f) == (size_t)len);
    else {
        char buf[1 << 16];
        for (int64_t i = 0; ok && i < len;) {
            const int64_t n = len - i < (int64_t)sizeof buf ? len - i : (int64_t)sizeof buf;
            for (int64_t k = 0; k < n; k++) { const uint8_t c = seq[i + k]; buf[k] = (char)((c >= 'a' && c <= 'z') ? c - 32 : c); }
            ok = fwrite(buf, 1, (size_t)n, f) == (size_t)n;
            i += n;
        }
    }
    ok = ok && fputc('\n', f) != EOF;
    if (!ok) { np::set_error("np_part_write: write failed"); return NP_ERR_IO; }
    return NP_OK;
}

}  // extern "C"
