// kdbx-synth — writes a generated kmer-db database (host/synth.cpp) and a JSON side file with its totals.
//
// A stand-alone host program: no CUDA, no libkdbx.so.  bench.py calls it from BOTH arms, so that the reference arm
// (the unmodified kmer-db binary) and the GPU arm read the very same .db file and neither needs the other's code to
// produce it.
//   kdbx-synth -o out.db [-n samples] [-c clusters] [-L kmers per genome] [-k k] [-mu rate] [-seed s] [-skew s]
//              [-interleaved] [-t threads]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../host/synth.h"
#include "../host/trie.h"

// trie.h can place its arrays in page-locked memory through libkdbx.so; this tool never asks for that
extern "C" int kdbx_host_alloc(void** out, size_t) { if (out) *out = nullptr; return KDBX_ERR_CUDA; }
extern "C" void kdbx_host_free(void*) {}

int main(int argc, char** argv) {
    kdbx::SynthParams sp;
    std::string out;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&]() -> const char* { if (i + 1 >= argc) { std::fprintf(stderr, "kdbx-synth: %s needs a value\n", a.c_str()); std::exit(2); } return argv[++i]; };
        if (a == "-o") out = val();
        else if (a == "-n") sp.num_samples = (uint32_t)std::strtoul(val(), nullptr, 10);
        else if (a == "-c") sp.num_clusters = (uint32_t)std::strtoul(val(), nullptr, 10);
        else if (a == "-L") sp.genome_kmers = std::strtoull(val(), nullptr, 10);
        else if (a == "-k") sp.k = (uint32_t)std::strtoul(val(), nullptr, 10);
        else if (a == "-mu") sp.mutation_rate = std::strtod(val(), nullptr);
        else if (a == "-seed") sp.seed = std::strtoull(val(), nullptr, 10);
        else if (a == "-skew") sp.cluster_skew = std::strtod(val(), nullptr);
        else if (a == "-t") sp.threads = std::atoi(val());
        else if (a == "-interleaved") sp.interleaved = 1;
        else { std::fprintf(stderr, "kdbx-synth: unknown option %s\n", a.c_str()); return 2; }
    }
    if (out.empty()) { std::fprintf(stderr, "usage: kdbx-synth -o out.db [-n N] [-c C] [-L kmers] [-k k] [-mu r] [-seed s] [-skew s] [-interleaved] [-t T]\n"); return 2; }
    try {
        const auto t0 = std::chrono::steady_clock::now();
        kdbx::Trie t(false);
        kdbx::synth_generate(sp, t);
        const auto t1 = std::chrono::steady_clock::now();
        const std::string tmp = out + ".tmp";
        kdbx::write_db(tmp, t);
        if (std::rename(tmp.c_str(), out.c_str()) != 0) throw std::runtime_error("cannot rename " + tmp);
        const auto t2 = std::chrono::steady_clock::now();
        const kdbx::Trie::Totals tot = t.totals();
        const std::string js = out + ".json";
        FILE* f = std::fopen((js + ".tmp").c_str(), "w");
        if (!f) throw std::runtime_error("cannot write " + js);
        std::fprintf(f, "{\"num_samples\": %u, \"num_patterns\": %llu, \"updates\": %llu, \"sum_n\": %llu, \"sum_l\": %llu, "
                        "\"payload_bytes\": %llu, \"kmers_count\": %llu, \"k\": %u, \"clusters\": %u, \"genome_kmers\": %llu, "
                        "\"mutation_rate\": %.10g, \"seed\": %llu, \"cluster_skew\": %.10g, \"generate_s\": %.2f, \"write_s\": %.2f}\n",
                     t.num_samples(), (unsigned long long)t.num_patterns(), (unsigned long long)tot.U, (unsigned long long)tot.sum_n,
                     (unsigned long long)tot.sum_l, (unsigned long long)tot.payload_bytes, (unsigned long long)t.hdr.kmers_count, sp.k,
                     sp.num_clusters, (unsigned long long)sp.genome_kmers, sp.mutation_rate, (unsigned long long)sp.seed, sp.cluster_skew,
                     std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count());
        std::fclose(f);
        if (std::rename((js + ".tmp").c_str(), js.c_str()) != 0) throw std::runtime_error("cannot rename " + js);
        std::fprintf(stderr, "kdbx-synth: %s  N=%u P=%llu U=%llu  (%.1f s + %.1f s)\n", out.c_str(), t.num_samples(),
                     (unsigned long long)t.num_patterns(), (unsigned long long)tot.U, std::chrono::duration<double>(t1 - t0).count(),
                     std::chrono::duration<double>(t2 - t1).count());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "kdbx-synth: %s\n", e.what());
        return 1;
    }
    return 0;
}
