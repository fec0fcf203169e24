// kmer-db-b200 — command-line front end for the B200-native path.  Mirrors the reference's
// mode drivers for the modes this repository covers (src/main.cpp:22-62, src/console.h:86-108):
//     kmer-db-b200 all2all [-sparse] [-t n] [-buffer mb] [-gpu id] <db> <out.csv>
//         (All2AllConsole::run, src/console_all2all.cpp:7-89)
//     kmer-db-b200 synth [-n N] [-clusters C] [-len L] [-k K] [-mu r] [-seed s] [-interleaved] <out.db>
//         (ours: writes a synthetic database the reference's all2all also accepts)
//     kmer-db-b200 info <db>
// Errors: "ERROR: <text>" on stderr and exit code -1, like the reference (src/main.cpp:56-59).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <iostream>

#include "similarity_calculator.h"
#include "synth.h"

using namespace kdbx;

namespace {
struct usage_error : std::runtime_error { using std::runtime_error::runtime_error; };
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int run_all2all(std::vector<std::string> args) {
    bool sparse = false; int threads = 0, gpu = -1; size_t buffer_mb = 8;
    std::vector<std::string> files;
    for (size_t i = 0; i < args.size(); ++i) {
        if (args[i] == "-sparse") sparse = true;
        else if (args[i] == "-t" && i + 1 < args.size()) threads = std::atoi(args[++i].c_str());
        else if (args[i] == "-buffer" && i + 1 < args.size()) buffer_mb = (size_t)std::atoll(args[++i].c_str());
        else if (args[i] == "-gpu" && i + 1 < args.size()) gpu = std::atoi(args[++i].c_str());
        else files.push_back(args[i]);
    }
    if (files.size() != 2) throw usage_error("all2all [-sparse] [-t n] [-gpu id] <db> <out.csv>");
    std::cerr << "All versus all comparison" << std::endl;
    SimilarityCalculator calculator(threads, buffer_mb, gpu);
    Trie db(true);
    std::cerr << "Loading k-mer database " << files[0] << "..." << std::endl;
    double t0 = now();
    read_db(files[0], db);
    std::cerr << "OK (" << now() - t0 << " seconds)" << std::endl;
    std::cerr << "Calculating matrix of common k-mers..." << std::endl;
    t0 = now();
    LowerTriangularMatrix<uint32_t> matrix;
    calculator.all2all(db, matrix);
    const double dt = now() - t0;
    const kdbx_stats& s = calculator.last_stats();
    std::cerr << "OK (" << dt << " seconds)" << std::endl;
    std::fprintf(stderr,
                 "{\"updates\": %llu, \"seconds\": %.6f, \"updates_per_s\": %.4g, \"ms_upload\": %.3f, \"ms_prepare\": %.3f, "
                 "\"ms_expand\": %.3f, \"ms_bucket\": %.3f, \"ms_scatter\": %.3f, \"ms_total\": %.3f, \"ms_download\": %.3f, "
                 "\"chunks\": %u, \"jobs\": %llu}\n",
                 (unsigned long long)s.updates, dt, dt > 0 ? (double)s.updates / dt : 0.0, s.ms_upload, s.ms_prepare, s.ms_expand,
                 s.ms_bucket, s.ms_scatter, s.ms_total, s.ms_download, s.chunks, (unsigned long long)s.jobs);
    std::cerr << "Storing matrix of common k-mers in " << files[1] << "...";
    t0 = now();
    write_all2all_csv(files[1], db, matrix.data(), sparse);
    std::cerr << "OK (" << now() - t0 << " seconds)" << std::endl;
    return 0;
}

int run_synth(std::vector<std::string> args) {
    SynthParams sp; std::vector<std::string> files;
    for (size_t i = 0; i < args.size(); ++i) {
        auto next = [&]() -> const char* { if (i + 1 >= args.size()) throw usage_error("synth: missing value"); return args[++i].c_str(); };
        if (args[i] == "-n") sp.num_samples = (uint32_t)std::atoll(next());
        else if (args[i] == "-clusters") sp.num_clusters = (uint32_t)std::atoll(next());
        else if (args[i] == "-len") sp.genome_kmers = (uint64_t)std::atoll(next());
        else if (args[i] == "-k") sp.k = (uint32_t)std::atoll(next());
        else if (args[i] == "-mu") sp.mutation_rate = std::atof(next());
        else if (args[i] == "-seed") sp.seed = (uint64_t)std::atoll(next());
        else if (args[i] == "-t") sp.threads = std::atoi(next());
        else if (args[i] == "-interleaved") sp.interleaved = 1;
        else files.push_back(args[i]);
    }
    if (files.size() != 1) throw usage_error("synth [-n N] [-clusters C] [-len L] [-k K] [-mu r] [-seed s] [-interleaved] <out.db>");
    Trie t;
    double t0 = now();
    synth_generate(sp, t);
    const auto tt = t.totals();
    std::fprintf(stderr, "generated N=%u P=%llu sum_n=%llu sum_l=%llu U=%llu payload=%llu B in %.2f s\n", t.num_samples(),
                 (unsigned long long)t.num_patterns(), (unsigned long long)tt.sum_n, (unsigned long long)tt.sum_l,
                 (unsigned long long)tt.U, (unsigned long long)tt.payload_bytes, now() - t0);
    write_db(files[0], t);
    return 0;
}

int run_info(std::vector<std::string> args) {
    if (args.size() != 1) throw usage_error("info <db>");
    Trie t; read_db(args[0], t);
    const auto tt = t.totals();
    std::printf("k=%u fraction=%g samples=%u patterns=%llu sum_n=%llu sum_l=%llu U=%llu payload_bytes=%llu\n", t.hdr.kmer_length,
                t.hdr.fraction, t.num_samples(), (unsigned long long)t.num_patterns(), (unsigned long long)tt.sum_n,
                (unsigned long long)tt.sum_l, (unsigned long long)tt.U, (unsigned long long)tt.payload_bytes);
    return 0;
}
}  // namespace

int main(int argc, char** argv) {
    try {
        if (argc < 2) throw usage_error("<mode> ...  (modes: all2all, synth, info)");
        const std::string mode = argv[1];
        std::vector<std::string> args(argv + 2, argv + argc);
        if (mode == "all2all") return run_all2all(args);
        if (mode == "synth") return run_synth(args);
        if (mode == "info") return run_info(args);
        throw usage_error("unknown mode " + mode);
    } catch (const usage_error& e) {
        std::cerr << "USAGE: kmer-db-b200 " << e.what() << std::endl;
        return -1;
    } catch (const std::runtime_error& e) {
        std::cerr << "ERROR: " << e.what() << std::endl;
        return -1;
    }
}
