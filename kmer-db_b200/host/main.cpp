// kmer-db-b200 — command-line front end of the B200-native path.  Same modes, switches, files and
// CSV bytes as the reference's mode drivers for what this repository covers
// (src/main.cpp:22-62, src/console.h:86-108):
//     build       BuildConsole::run            src/console_build.cpp:33-157     (host)
//     all2all     All2AllConsole::run          src/console_all2all.cpp:7-89     (GPU)
//     all2all-sp  All2AllSparseConsole::run    src/console_all2all_sparse.cpp:13-111 (GPU)
//     new2all     New2AllConsole::run          src/console_new2all.cpp:12-174   (GPU)
//     distance    DistanceConsole::run         src/console_distance.cpp:7-213   (host)
// plus two tools of ours: `synth` (pattern-level synthetic database) and `info`.
// Errors: "ERROR: <text>" on stderr and exit code -1, like the reference (src/main.cpp:51-59).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <iostream>

#include "build_device.h"
#include "cli.h"
#include "csv_out.h"
#include "ingest.h"
#include "kmer_db.h"
#include "similarity_calculator.h"
#include "synth.h"

namespace kdbx {
namespace {
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void print_stats_json(const kdbx_stats& s, double dt) {
    std::fprintf(stderr,
                 "{\"updates\": %llu, \"seconds\": %.6f, \"updates_per_s\": %.4g, \"ms_upload\": %.3f, \"ms_prepare\": %.3f, "
                 "\"ms_expand\": %.3f, \"ms_bucket\": %.3f, \"ms_scatter\": %.3f, \"ms_compact\": %.3f, \"ms_probe\": %.3f, "
                 "\"ms_total\": %.3f, \"ms_download\": %.3f, \"chunks\": %u, \"probes\": %llu, \"hits\": %llu}\n",
                 (unsigned long long)s.updates, dt, dt > 0 ? (double)s.updates / dt : 0.0, s.ms_upload, s.ms_prepare, s.ms_expand,
                 s.ms_bucket, s.ms_scatter, s.ms_compact, s.ms_probe, s.ms_total, s.ms_download, s.chunks,
                 (unsigned long long)s.probes, (unsigned long long)s.hits);
}
}  // namespace

// build on the device (the default): the host only reads, gunzips and splits the FASTA files; k-mer
// extraction, minhash, sort/unique, the k-mer table and the pattern trie live on the GPU (csrc/build.cuh)
static void run_build_device(const Params& p) {
    Alphabet alphabet = p.alphabet;
    uint32_t k = p.kmer_length;
    double fraction = p.fraction, start = p.fraction_start;
    Trie old;
    if (p.extend_db) {
        std::cerr << "Loading k-mer database " << p.files[1] << "..." << std::endl;
        read_db(p.files[1], old, true);
        alphabet = Alphabet::make(old.hdr.alphabet_type);
        k = old.hdr.kmer_length;
        fraction = old.hdr.fraction; start = old.hdr.start_fraction;
        // the alphabet recorded with every sample is the command line's one (src/console_build.cpp:117)
        if (old.hdr.alphabet_type != p.alphabet.id)
            throw std::runtime_error("Error in AbstractKmerDb::addKmers(): adding samples from different alphabet");
    }
    DeviceDbBuilder builder(p.gpu, alphabet, k, fraction, start);
    if (p.extend_db) { builder.adopt(old); old = Trie(); }
    std::cerr << "Processing samples..." << std::endl;
    const double t0 = now();
    SequenceStream stream(p.files[0], p.multisample_fasta, p.num_reader_threads);
    SampleSeq s;
    size_t n = 0;
    while (stream.next(s)) {
        if (builder.add_sample(s.name, s.symbols.data(), s.symbols.size()) == 0) std::cerr << "Empty sample: " << s.name << std::endl;
        if (++n % 10 == 0) std::cerr << "\r" << n << "/" << stream.num_files() << "..." << std::flush;
    }
    std::cerr << "\r" << n << "/" << n << "                      " << std::endl;
    std::cerr << "Database update time: " << now() - t0 << std::endl;
    std::cerr << "Serializing database..." << std::endl;
    Trie db;
    builder.finish(db);
    const kdbx_build_result& r = builder.result();
    std::fprintf(stderr, "{\"samples\": %u, \"patterns\": %llu, \"kmers\": %llu, \"table_slots\": %llu, \"table_growths\": %u, "
                         "\"kernel_launches\": %u, \"ms_finish\": %.3f, \"seconds\": %.6f}\n",
                 r.num_samples, (unsigned long long)r.num_patterns, (unsigned long long)r.kmers_count,
                 (unsigned long long)r.table_capacity, r.table_growths, r.kernel_launches, r.ms_finish, now() - t0);
    write_db(p.files[1], db);
}

void run_build(const Params& p) {
    if (p.files.size() != 2) throw usage_error(p.mode);
    std::cerr << "Building database (from genomes)" << std::endl;
    if (!p.host_build) { run_build_device(p); return; }
    DbBuilder builder(p.num_threads);
    Alphabet alphabet = p.alphabet;
    MinHash filter(p.fraction, p.fraction_start, p.kmer_length);
    uint32_t k = p.kmer_length;
    double fraction = p.fraction;
    if (p.extend_db) {
        std::cerr << "Loading k-mer database " << p.files[1] << "..." << std::endl;
        Trie old;
        read_db(p.files[1], old, true);
        alphabet = Alphabet::make(old.hdr.alphabet_type);
        k = old.hdr.kmer_length;
        fraction = old.hdr.fraction;
        filter = MinHash(old.hdr.fraction, old.hdr.start_fraction, k);
        builder.adopt(std::move(old));
    }
    std::cerr << "Processing samples..." << std::endl;
    const double t0 = now();
    SampleStream stream(p.files[0], alphabet, filter, k, p.multisample_fasta, p.num_reader_threads);
    SampleKmers s;
    size_t n = 0;
    // the alphabet recorded with every sample is the command line's one (src/console_build.cpp:117)
    const int32_t alphabet_id = p.extend_db ? p.alphabet.id : alphabet.id;
    while (stream.next(s)) {
        if (s.kmers.empty()) std::cerr << "Empty sample: " << s.name << std::endl;
        builder.add_sample(s.name, s.kmers.data(), s.kmers.size(), k, fraction, alphabet_id, alphabet.bits_per_symbol);
        if (++n % 10 == 0) std::cerr << "\r" << n << "/" << stream.num_files() << "..." << std::flush;
    }
    std::cerr << "\r" << n << "/" << n << "                      " << std::endl;
    std::cerr << "Database update time: " << now() - t0 << std::endl;
    std::cerr << "Serializing database..." << std::endl;
    Trie db;
    builder.finish(db);
    write_db(p.files[1], db);
}

void run_all2all(const Params& p) {
    if (p.files.size() != 2) throw usage_error(p.mode);
    std::cerr << "All versus all comparison" << std::endl;
    SimilarityCalculator calculator(p.num_threads, (size_t)p.cache_buffer_mb, p.gpu);
    Trie db(true);
    std::cerr << "Loading k-mer database " << p.files[0] << "..." << std::endl;
    double t0 = now();
    read_db(p.files[0], db);
    std::cerr << "OK (" << now() - t0 << " seconds)" << std::endl;
    std::cerr << "Calculating matrix of common k-mers..." << std::endl;
    t0 = now();
    LowerTriangularMatrix<uint32_t> matrix;
    if (p.num_gpus > 1) calculator.all2all_multi(db, matrix, p.num_gpus);
    else calculator.all2all(db, matrix);
    const double dt = now() - t0;
    std::cerr << "OK (" << dt << " seconds)" << std::endl;
    print_stats_json(calculator.last_stats(), dt);
    std::cerr << "Storing matrix of common k-mers in " << p.files[1] << "...";
    t0 = now();
    // the dense table's numbers are formatted on the device, from the matrix the call above left there (one GPU; the
    // -sparse table and the multi-GPU blocks go through the host emitter)
    if (!p.sparse_out && p.num_gpus <= 1 && !p.host_csv) write_all2all_csv_device(p.files[1], db, calculator.context());
    else write_all2all_csv(p.files[1], db, matrix.data(), p.sparse_out, p.sparse_out ? &p.filters : nullptr);
    std::cerr << "OK (" << now() - t0 << " seconds)" << std::endl;
}

void run_all2all_sparse(const Params& p) {
    if (p.files.size() != 2) throw usage_error(p.mode);
    std::cerr << "All versus all comparison (sparse computation)" << std::endl;
    SimilarityCalculator calculator(p.num_threads, (size_t)p.cache_buffer_mb, p.gpu);
    Trie db(true);
    std::cerr << "Loading k-mer database " << p.files[0] << "..." << std::endl;
    double t0 = now();
    read_db(p.files[0], db);
    std::cerr << "OK (" << now() - t0 << " seconds)" << std::endl;
    std::cerr << "Calculating matrix of common k-mers...";
    t0 = now();
    SparseMatrix<uint32_t> matrix;
    if (p.num_gpus > 1) calculator.all2all_sp_multi(db, matrix, p.filters, p.num_gpus);
    else calculator.all2all_sp(db, matrix, p.filters);
    const double dt = now() - t0;
    std::cerr << "OK (" << dt << " seconds)" << std::endl;
    print_stats_json(calculator.last_stats(), dt);
    std::cerr << "Storing matrix of common k-mers in " << p.files[1] << "...";
    t0 = now();
    const uint64_t saved = write_sparse_csv(p.files[1], db, *matrix.raw(), &p.filters);
    std::cerr << "OK (" << now() - t0 << " seconds)" << std::endl;
    std::cerr << "No. saved pairs: " << saved << std::endl;
}

void run_new2all(const Params& p) {
    if (p.files.size() != 3) throw usage_error(p.mode);
    std::cerr << "Set of new samples  (from genomes) versus entire database comparison" << std::endl;
    SimilarityCalculator calculator(p.num_threads, (size_t)p.cache_buffer_mb, p.gpu);
    Trie db(true);
    std::cerr << "Loading k-mer database " << p.files[0] << "..." << std::endl;
    double t0 = now();
    read_db(p.files[0], db, true);
    std::cerr << "OK (" << now() - t0 << " seconds)" << std::endl;
    calculator.load_database(db);
    SequenceStream stream(p.files[1], p.multisample_fasta, p.num_reader_threads);
    std::cerr << "Processing queries..." << std::endl;
    t0 = now();
    QueryTableWriter writer(p.files[2], db, p.sparse_out, &p.filters);
    // the queries' SEQUENCES go to the device in batches (bounded symbols); k-mer extraction, sort and unique
    // happen there; rows are written in input order
    const uint64_t batch_symbols = (uint64_t)1 << 27;
    const size_t N = db.num_samples();
    std::vector<std::string> names;
    Buf<char> symbols;
    symbols.set_pinned(true);
    symbols.reserve(batch_symbols + ((uint64_t)1 << 24));
    std::vector<uint64_t> q_off, unique;
    std::vector<uint32_t> sims;
    kdbx_stats total{};
    size_t done = 0;
    bool more = true;
    SampleSeq s;
    bool pending = false;   // `s` holds a query that did not fit the previous batch
    while (more || pending) {
        names.clear(); symbols.clear(); q_off.assign(1, 0);
        while (names.size() < 4096) {
            if (!pending) { more = stream.next(s); if (!more) break; }
            pending = false;
            if (!names.empty() && symbols.size() + s.symbols.size() > batch_symbols) { pending = true; break; }
            const size_t at = symbols.size();
            symbols.resize(at + s.symbols.size());
            std::copy(s.symbols.begin(), s.symbols.end(), symbols.data() + at);
            q_off.push_back(symbols.size());
            names.push_back(std::move(s.name));
            s = SampleSeq();
        }
        if (names.empty()) break;
        calculator.one2all_sequences(db.hdr, symbols.data(), q_off, sims, unique);
        const kdbx_stats& st = calculator.last_stats();
        total.probes += st.probes; total.hits += st.hits; total.ms_probe += st.ms_probe; total.ms_scatter += st.ms_scatter;
        total.ms_total += st.ms_total; total.ms_prepare += st.ms_prepare; total.ms_download += st.ms_download; total.ms_expand += st.ms_expand;
        for (size_t q = 0; q < names.size(); ++q) writer.write_row(names[q], unique[q], sims.data() + q * N);
        done += names.size();
        if (done % 10 == 0) std::cerr << "\r" << done << "...                      " << std::flush;
    }
    writer.close();
    const double dt = now() - t0;
    std::cerr << std::endl << std::endl << "EXECUTION TIMES" << std::endl << "Total: " << dt << std::endl;
    print_stats_json(total, dt);
}

namespace {
int run_synth(std::vector<std::string> args) {
    SynthParams sp; std::vector<std::string> files;
    for (size_t i = 0; i < args.size(); ++i) {
        auto next = [&]() -> const char* { if (i + 1 >= args.size()) throw std::runtime_error("synth: missing value"); return args[++i].c_str(); };
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
    if (files.size() != 1) throw std::runtime_error("usage: synth [-n N] [-clusters C] [-len L] [-k K] [-mu r] [-seed s] [-interleaved] <out.db>");
    Trie t;
    const double t0 = now();
    synth_generate(sp, t);
    const auto tt = t.totals();
    std::fprintf(stderr, "generated N=%u P=%llu sum_n=%llu sum_l=%llu U=%llu payload=%llu B in %.2f s\n", t.num_samples(),
                 (unsigned long long)t.num_patterns(), (unsigned long long)tt.sum_n, (unsigned long long)tt.sum_l,
                 (unsigned long long)tt.U, (unsigned long long)tt.payload_bytes, now() - t0);
    write_db(files[0], t);
    return 0;
}

int run_info(std::vector<std::string> args) {
    if (args.size() != 1) throw std::runtime_error("usage: info <db>");
    Trie t; read_db(args[0], t);
    const auto tt = t.totals();
    std::printf("k=%u fraction=%g samples=%u patterns=%llu sum_n=%llu sum_l=%llu U=%llu payload_bytes=%llu\n", t.hdr.kmer_length,
                t.hdr.fraction, t.num_samples(), (unsigned long long)t.num_patterns(), (unsigned long long)tt.sum_n,
                (unsigned long long)tt.sum_l, (unsigned long long)tt.U, (unsigned long long)tt.payload_bytes);
    return 0;
}
}  // namespace
}  // namespace kdbx

int main(int argc, char** argv) {
    using namespace kdbx;
    try {
        if (argc >= 2 && std::string(argv[1]) == "synth") return run_synth(std::vector<std::string>(argv + 2, argv + argc));
        if (argc >= 2 && std::string(argv[1]) == "info") return run_info(std::vector<std::string>(argv + 2, argv + argc));
        Params p;
        if (!parse_params(argc, argv, p)) return 0;
        if (p.mode == "build") run_build(p);
        else if (p.mode == "all2all") run_all2all(p);
        else if (p.mode == "all2all-sp") run_all2all_sparse(p);
        else if (p.mode == "new2all") run_new2all(p);
        else if (p.mode == "distance") run_distance(p);
    } catch (const usage_error& e) {
        print_usage(e.what());
        return -1;
    } catch (const std::runtime_error& e) {
        std::cerr << "ERROR: " << e.what() << std::endl;
        return -1;
    }
    return 0;
}
