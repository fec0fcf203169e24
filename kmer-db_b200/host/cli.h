// Command-line surface of kmer-db-b200: the reference's modes and switches for the path this
// repository covers (src/params.h:28-66, src/params.cpp:60-136,418-709).  Options may appear
// anywhere after the mode; whatever is left over is the positional file list, exactly like the
// reference's findSwitch / findOption scheme (src/params.h:115-156).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "kmers.h"
#include "metrics.h"

namespace kdbx {

struct usage_error : std::runtime_error {
    explicit usage_error(const std::string& mode) : std::runtime_error(mode) {}
};

struct Params {
    std::string mode;
    double fraction = 1.0, fraction_start = 0.0;
    uint32_t kmer_length = 18;
    int num_threads = 0, num_reader_threads = 0;
    int cache_buffer_mb = 8;   // -buffer: accepted, no meaning on the GPU
    int bubble_size = 8000;    // -bubble-size: accepted, result-neutral (src/bubble_helper.h)
    bool multisample_fasta = false, sparse_out = false, extend_db = false, phylip_out = false;
    int gpu = -1;              // -gpu <ordinal> (ours)
    int num_gpus = 1;          // -gpus <n> (ours): row-block sharding over n devices
    bool from_minhash = false; // -from-minhash: samples / queries are <entry>.minhash files written by the `minhash` mode
    bool fraction_given = false;
    bool host_build = false;   // build -host-build (ours): run the host builder explicitly (machines without a GPU)
    bool device_distance = false;   // distance -device (ours): measure + six-decimal text on the GPU (explicit; no fallback)
    bool host_csv = false;     // all2all -host-csv (ours): format the dense table on the host instead of on the device
    Alphabet alphabet = Alphabet::make(kNt);
    OutputFilters filters;
    int sampling_size = 0;               // -sample-rows [criterion:]count of all2all-sp / all2all-parts (0 = off)
    metric_fn sampling_criterion = nullptr;
    std::string metric_name;
    std::vector<std::string> files;
};

// Throws usage_error / std::runtime_error like the reference; returns false when only help or
// the version was requested.
bool parse_params(int argc, char** argv, Params& out);
void print_usage(const std::string& mode);

void run_build(const Params& p);
void run_minhash(const Params& p);
void run_all2all(const Params& p);
void run_all2all_sparse(const Params& p);
void run_all2all_parts(const Params& p);
void run_new2all(const Params& p);
void run_one2all(const Params& p);
void run_distance(const Params& p);

}  // namespace kdbx
