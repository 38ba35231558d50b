// hostio.h — host-side FASTA / BGZF / BAM / BAI readers and the packed-shard builder.
//
// This is the engine's own I/O layer (zlib is the only dependency); it stands where the
// reference links its vendored htslib 1.9 (fai_load/fai_fetch: contig.c:35-36,1119-1128;
// bam_itr_queryi + sam_itr_next: contig.c:172-174,692-694; bam_read1: config.c:87).
// Only the iteration ORDER of htslib matters to the algorithm (first-seen tie-breaks):
// records of one contig are delivered in file order, which is what this reader does.
#pragma once
#include "bgzf_inflate.h"
#include <stdlib.h>
#include <cstdint>
#include <string>
#include <vector>
#include <functional>
#include "../../include/nextpolish_b200.h"

namespace np {

struct FastaRecord {
    std::string name;
    std::string seq;   // whitespace removed, case preserved
};

// Every sequence of a FASTA file, concatenated without line breaks into `dst` (grown through `grow(bytes)`, which returns
// the — possibly moved — buffer of at least that many bytes; contents up to the previous size are preserved by the caller's
// allocator or re-parsed): names[i] owns dst[off[i], off[i+1]).  One pass over an mmap of the file, one copy.
// Same character rules as fasta_load (isgraph bytes of non-header lines; junk before the first header is skipped).
bool fasta_load_flat(const std::string& path, std::vector<std::string>& names, std::vector<int64_t>& off,
                     uint8_t* (*grow)(void* ctx, size_t bytes), void* ctx, std::string& err);

// Reads all (names == empty) or the named sequences. Uses <fasta>.fai when present to seek.
bool fasta_load(const std::string& path, const std::vector<std::string>& names,
                std::vector<FastaRecord>& out, std::string& err);
// Names in file order (from .fai when present, else by scanning).
bool fasta_names(const std::string& path, std::vector<std::string>& names,
                 std::vector<int64_t>& lengths, std::string& err);

struct BamHeader {
    std::vector<std::string> names;
    std::vector<int64_t>     lengths;
    uint64_t                 first_record_voffset = 0;
};

// A decoded BAM alignment record (pointers into a transient buffer).
struct BamRec {
    int32_t  tid, pos;
    uint8_t  mapq;
    uint16_t flag;
    uint32_t n_cigar;
    int32_t  l_qseq, isize;
    const uint32_t* cigar;   // may be unaligned: use memcpy
    const uint8_t*  seq;
    const uint8_t*  qual;
    const uint8_t*  aux;     // optional fields (tag, type, value ...), l_aux bytes
    int32_t  l_aux;
};

class BamFile {
public:
    ~BamFile();
    bool open(const std::string& path, std::string& err);
    const BamHeader& header() const { return hdr_; }
    // Visit records in file order starting at virtual offset `voff` (0 = first record) with
    // `threads` inflate threads. The visitor returns false to stop.
    bool scan(uint64_t voff, int threads, const std::function<bool(const BamRec&)>& visit,
              std::string& err);
    // Smallest chunk-begin virtual offset of `tid` from <bam>.bai; returns false if the index
    // is missing/unreadable; *has_reads=false if the contig has no chunks.
    bool bai_first_offset(int tid, uint64_t* voff, bool* has_reads, std::string& err) const;
    // Every record-start virtual offset the index knows per reference: chunk begins of all bins and the linear
    // index entries (16 kb windows) — anchors for the device-side record walk (devload.cu).  sorted, unique.
    bool bai_record_starts(std::vector<std::vector<uint64_t>>& starts, std::string& err) const;
    const uint8_t* data() const { return data_; }
    size_t size() const { return size_; }
private:
    bool inflate_batch(size_t first_block, size_t n_blocks, int threads,
                       std::vector<uint8_t>& out, std::vector<size_t>& block_uoff, std::string& err);
    bool index_blocks(std::string& err);
    std::string path_;
    int fd_ = -1;
    const uint8_t* data_ = nullptr;
    size_t size_ = 0;
    std::vector<uint64_t> block_coff_;   // compressed offset of every BGZF block
    BamHeader hdr_;
    bool read_header(std::string& err);
};

// Host-owned packed shard (see np_shard_view in nextpolish_b200.h for the layout).
struct Shard {
    std::vector<std::string> names;
    std::vector<int64_t>  ctg_off;
    std::vector<uint8_t>  ctg_seq;
    std::vector<int64_t>  ctg_read_off;
    std::vector<uint32_t> rec_off;
    std::vector<uint8_t>  rec;
    std::vector<uint32_t> qual_off;
    std::vector<uint8_t>  qual;
    std::vector<int32_t>  fasta_rank;   // position of each shard contig in the requested / FASTA order
    int64_t alg_bytes = 0;              // sum over reads of 16 + 4*n_cigar + ceil(l_qseq/2)
    int64_t qual_bytes = 0;             // sum over reads of l_qseq
    bool with_qual = false;
    // quality mode: 1 = every read, 2 = only reads whose reference span contains a lowercase draft base
    // (the only reads whose qualities task 2 can consult: every k-mer window contains a lowercase column
    // and its candidates span the whole window, kmercount.c:128-173,196-199; contig.c:1027)
    int qual_mode = 0;
    // bases of A/C/G/T-only reads are packed with 2 bits each unless NEXTPOLISH_B200_4BIT=1 (A/B tests)
    bool two_bit = !(getenv("NEXTPOLISH_B200_4BIT") && getenv("NEXTPOLISH_B200_4BIT")[0] == '1');
    std::vector<int32_t> cur_lc;        // lowercase prefix counts of the contig being packed (mode 2)
    void begin_contig(const uint8_t* seq, size_t len);
    void view(np_shard_view* v) const;
};
bool shard_pack_record(const BamRec& r, Shard& s, std::string& err);
bool synth_shard(const np_synth_params& P, int32_t lo, int32_t hi, int with_qual, int threads,
                 Shard& out, std::string& err);

bool shard_load(const std::string& fasta, const std::string& bam,
                const std::vector<std::string>& names, int with_qual, int threads,
                Shard& out, std::string& err);

// Walks the BGZF blocks of a byte range: deflate payload location and output placement of every block
// (for np_bgzf_inflate, bgzf_inflate.cu); total = sum of the blocks' ISIZE.
bool bgzf_scan(const uint8_t* data, size_t size, std::vector<npz::Block>& blocks, int64_t& total, std::string& err,
               std::vector<uint64_t>* coffs = nullptr, size_t begin = 0, size_t end = (size_t)-1);

// config.c:80-101 (bam_tlen): mean insert size estimate over the head of the BAM.
bool bam_insert_estimate(const std::string& bam, uint32_t count_read_ins, uint32_t max_ins_len,
                         uint32_t* mean, int32_t* read_len, std::string& err);

}  // namespace np
