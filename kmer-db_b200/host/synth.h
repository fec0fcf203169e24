// Parameters of the pattern-level synthetic database generator (synth.cpp).
#pragma once
#include "trie.h"

namespace kdbx {

struct SynthParams {
    uint32_t num_samples = 1000;
    uint32_t num_clusters = 4;
    uint64_t genome_kmers = 5000000;  // distinct k-mers per genome (~ genome length in bp)
    uint32_t k = 18;
    double mutation_rate = 0.005;     // substitutions per base between a genome and its template
    uint64_t seed = 2;
    int interleaved = 0;              // 0: clusters contiguous in sample order; 1: round-robin
    int threads = 0;                  // 0 = hardware concurrency (one cluster per thread)
    double cluster_skew = 0.0;        // 0: equal clusters; s in (0,1): cluster sizes spread over [1-s, 1+s] x the mean (contiguous order only)
};

void synth_generate(const SynthParams& sp, Trie& out);

}  // namespace kdbx
