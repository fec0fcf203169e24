// Sample ingest shared by `build` and `new2all`: sample list -> FASTA records -> sorted,
// de-duplicated k-mers per sample, delivered strictly in list order (the reference's ordered
// output queue, src/loader_ex.h:42-48).  Files are parsed and their k-mers extracted and
// sorted by a small pool of reader threads running ahead of the consumer.
#pragma once
#include <condition_variable>
#include <deque>
#include <future>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "kmers.h"

namespace kdbx {

struct FastaRecord {
    std::string header;  // up to the first space
    const char* seq = nullptr;
    size_t len = 0;
};

struct SampleKmers {
    std::string name;
    std::vector<uint64_t> kmers;  // ascending, unique
    std::string entry;            // the list entry the sample came from (where `minhash` stores <entry>.minhash)
    uint32_t k = 0;               // from a .minhash file: the k-mer length and the fraction recorded there (0 otherwise)
    double fraction = 0.0;
};

// <entry>.minhash, the reference's minhashed-sample file (MihashedInputFile::store / open, src/minhashed_input_file.h:56-121):
// u32 0xfedcba98, u64 count, u64 kmers[count] (ascending, unique, already filtered), u32 k-mer length, f64 fraction.
void store_minhash(const std::string& entry, const uint64_t* kmers, size_t count, uint32_t k, double fraction);
bool load_minhash(const std::string& entry, SampleKmers& out);

std::vector<std::string> read_sample_list(const std::string& arg);
bool load_sequence_file(const std::string& entry, std::string& data);  // plain or gzip
void split_fasta(std::string& data, std::vector<FastaRecord>& out);    // edits `data` in place
std::string sample_name_of(const std::string& entry);

// One sample as the device builder takes it: the symbols of its records back to back, each record
// followed by a NUL byte (outside every alphabet, so no k-mer spans two records).
struct SampleSeq {
    std::string name;
    std::string symbols;
};

// Like SampleStream, but stops before k-mer extraction: the sequences go to the GPU
// (kdbx_builder_add_sequence).  Reader threads load, gunzip and split the files ahead of the consumer.
class SequenceStream {
public:
    SequenceStream(const std::string& list_arg, bool multisample, int threads);
    // the entries themselves (one2all: a single sample file, with or without its extension)
    SequenceStream(std::vector<std::string> entries, bool multisample, int threads);
    bool next(SampleSeq& out);
    size_t num_files() const { return files_.size(); }

private:
    std::vector<SampleSeq> load_file(size_t idx) const;
    void refill();
    std::vector<std::string> files_;
    bool multisample_;
    size_t ahead_;
    size_t next_file_ = 0;
    std::deque<std::future<std::vector<SampleSeq>>> inflight_;
    std::deque<SampleSeq> ready_;
};

class SampleStream {
public:
    SampleStream(const std::string& list_arg, const Alphabet& alphabet, const MinHash& filter, uint32_t k, bool multisample,
                 int threads);
    // -from-minhash: every list entry names <entry>.minhash (k-mer length and fraction come from the files)
    struct FromMinhash {};
    SampleStream(const std::string& list_arg, FromMinhash, int threads);
    SampleStream(std::vector<std::string> entries, FromMinhash, int threads);
    // Next sample in input order; false when the input is exhausted.
    bool next(SampleKmers& out);
    size_t num_files() const { return files_.size(); }

private:
    std::vector<SampleKmers> load_file(size_t idx) const;
    void refill();
    std::vector<std::string> files_;
    Alphabet alphabet_;
    MinHash filter_;
    uint32_t k_;
    bool multisample_;
    bool from_minhash_ = false;
    size_t ahead_;
    size_t next_file_ = 0;
    std::deque<std::future<std::vector<SampleKmers>>> inflight_;
    std::deque<SampleKmers> ready_;
};

}  // namespace kdbx
