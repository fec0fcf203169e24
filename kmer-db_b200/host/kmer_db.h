// Incremental database construction: the host-side restatement of PrefixKmerDb::addKmers
// (src/prefix_kmer_db.cpp:244-434).  One sample at a time, for every k-mer of
// the sample: find-or-insert in the prefix bucket's table (new k-mers start on pattern 0), group
// the sample's k-mers by their current pattern, then per group either append the sample to the
// pattern in place (the group covers all of the pattern's k-mers and it has no children) or
// split off a child pattern and repoint the group's table slots.  Tree shape, n_p, l_p, U and
// every CSV derived from the result equal the reference's; pattern numbering inside one
// sample's batch is not deterministic in the reference either (atomic id counter, :219).
#pragma once
#include <string>
#include <vector>

#include "trie.h"

namespace kdbx {

class DbBuilder {
public:
    explicit DbBuilder(int threads);
    ~DbBuilder();
    DbBuilder(const DbBuilder&) = delete;
    DbBuilder& operator=(const DbBuilder&) = delete;

    // continue an existing database (build -extend, src/console_build.cpp:48-57); consumes `db`
    void adopt(Trie&& db);
    bool initialized() const { return hdr_.is_initialized != 0 && !tables_.empty(); }
    const DbHeader& header() const { return hdr_; }

    // kmers: ascending, unique, already shifted/filtered (kmers.h)
    uint32_t add_sample(const std::string& name, const uint64_t* kmers, size_t count, uint32_t k, double fraction,
                        int32_t alphabet_id, int bits_per_symbol);
    // moves everything into `out` (SoA trie + tables)
    void finish(Trie& out);
    uint64_t num_patterns() const { return pats_.size(); }

private:
    struct Pattern {
        int64_t num_kmers = 0;
        int64_t parent = -1;
        uint32_t n = 0, l = 0, last = 0, bits = 0;
        uint32_t cap_words = 0;
        bool is_parent = false;
        uint64_t* data = nullptr;
    };
    void append_sample(Pattern& p, uint32_t sample);
    int threads_;
    DbHeader hdr_;
    std::vector<std::string> names_;
    std::vector<uint64_t> sample_kmers_;
    std::vector<HashTable> tables_;
    std::vector<Pattern> pats_;
    std::vector<std::pair<int32_t, uint64_t*>> sample_patterns_, sorted_;  // (pattern id, table slot) per k-mer
    std::vector<uint64_t> sort_a_, sort_b_;
};

}  // namespace kdbx
