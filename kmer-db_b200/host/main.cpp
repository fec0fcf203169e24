// kmer-db-b200 — command-line front end of the B200-native path.  Same modes, switches, files and
// CSV bytes as the reference's mode drivers for what this repository covers
// (src/main.cpp:22-62, src/console.h:86-108):
//     build       BuildConsole::run            src/console_build.cpp:33-157     (host)
//     all2all     All2AllConsole::run          src/console_all2all.cpp:7-89     (GPU)
//     all2all-sp  All2AllSparseConsole::run    src/console_all2all_sparse.cpp:13-111 (GPU)
//     all2all-parts All2AllPartsConsole::run   src/console_all2all_parts.cpp:11-371 (GPU)
//     new2all     New2AllConsole::run          src/console_new2all.cpp:12-174   (GPU)
//     distance    DistanceConsole::run         src/console_distance.cpp:7-213   (host)
// plus two tools of ours: `synth` (pattern-level synthetic database) and `info`.
// Errors: "ERROR: <text>" on stderr and exit code -1, like the reference (src/main.cpp:51-59).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <fstream>
#include <iostream>
#include <iterator>
#include <memory>
#include <mutex>
#include <thread>

#include "build_device.h"
#include "cli.h"
#include "csv_out.h"
#include "ingest.h"
#include "kmer_db.h"
#include "similarity_calculator.h"
#include "numfmt.h"
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
    // -from-minhash: the k-mer length and the fraction are the ones the `minhash` mode recorded in the samples' files
    // (src/minhashed_input_file.h:62-84, src/console_build.cpp:105-118), so the first sample is read before the builder opens
    std::unique_ptr<SampleStream> minhashed;
    SampleKmers first;
    bool have_first = false;
    if (p.from_minhash) {
        minhashed = std::make_unique<SampleStream>(p.files[0], SampleStream::FromMinhash{}, p.num_reader_threads);
        have_first = minhashed->next(first);
        if (have_first && !p.extend_db) { k = first.k; fraction = first.fraction; start = 0.0; }
    }
    DeviceDbBuilder builder(p.gpu, alphabet, k, fraction, start);
    if (p.extend_db) { builder.adopt(old); old = Trie(); }
    std::cerr << "Processing samples..." << std::endl;
    const double t0 = now();
    size_t n = 0;
    size_t num_files = 0;
    if (p.from_minhash) {
        num_files = minhashed->num_files();
        for (bool more = have_first; more; more = minhashed->next(first)) {
            if (first.k != k) throw std::runtime_error("Error in AbstractKmerDb::addKmers(): adding kmers of different length");
            if (first.fraction != fraction) throw std::runtime_error("Error in AbstractKmerDb::addKmers(): adding kmers of different minhash fraction");
            if (builder.add_sample_kmers(first.name, first.kmers.data(), first.kmers.size()) == 0) std::cerr << "Empty sample: " << first.name << std::endl;
            if (++n % 10 == 0) std::cerr << "\r" << n << "/" << num_files << "..." << std::flush;
        }
    } else {
        SequenceStream stream(p.files[0], p.multisample_fasta, p.num_reader_threads);
        SampleSeq s;
        while (stream.next(s)) {
            if (builder.add_sample(s.name, s.symbols.data(), s.symbols.size()) == 0) std::cerr << "Empty sample: " << s.name << std::endl;
            if (++n % 10 == 0) std::cerr << "\r" << n << "/" << stream.num_files() << "..." << std::flush;
        }
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
    SampleStream stream = p.from_minhash ? SampleStream(p.files[0], SampleStream::FromMinhash{}, p.num_reader_threads)
                                         : SampleStream(p.files[0], alphabet, filter, k, p.multisample_fasta, p.num_reader_threads);
    SampleKmers s;
    size_t n = 0;
    // the alphabet recorded with every sample is the command line's one (src/console_build.cpp:117)
    const int32_t alphabet_id = p.extend_db ? p.alphabet.id : alphabet.id;
    while (stream.next(s)) {
        if (s.kmers.empty()) std::cerr << "Empty sample: " << s.name << std::endl;
        if (p.from_minhash) { k = s.k; fraction = s.fraction; }   // (the builder refuses a sample whose k or fraction differs from the database's)
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

// minhash (src/console_minhash.cpp:6-60): every sample's k-mers — extracted, filtered to the fraction, sorted, unique —
// are stored next to the sample as <entry>.minhash, to be read back by build / new2all / one2all -from-minhash.
// Host work (file in, file out); the device enters where those modes consume the files.
void run_minhash(const Params& p) {
    if (p.files.size() != 1) throw usage_error(p.mode);
    std::cerr << "Minhashing samples..." << std::endl;
    const double t0 = now();
    SampleStream stream(p.files[0], p.alphabet, MinHash(p.fraction, 0.0, p.kmer_length), p.kmer_length, false, p.num_reader_threads);
    SampleKmers s;
    size_t n = 0;
    while (stream.next(s)) {
        store_minhash(s.entry, s.kmers.data(), s.kmers.size(), p.kmer_length, p.fraction);
        ++n;
    }
    if (n != stream.num_files()) throw std::runtime_error("Cannot open some of the samples of " + p.files[0]);
    std::cerr << n << " samples, " << now() - t0 << " seconds" << std::endl;
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
    const uint64_t saved = p.sampling_size > 0
        ? write_sparse_csv_sampled(p.files[1], db, *matrix.raw(), &p.filters, (uint32_t)p.sampling_size, p.sampling_criterion)
        : write_sparse_csv(p.files[1], db, *matrix.raw(), &p.filters);
    std::cerr << "OK (" << now() - t0 << " seconds)" << std::endl;
    std::cerr << "No. saved pairs: " << saved << std::endl;
}

// all2all-parts (src/console_all2all_parts.cpp:11-371): the input is a LIST of databases — the parts of one sample
// collection, built separately with the same k and fraction — and the output is the sparse table of all their samples,
// as all2all-sp would write it for one database holding all of them.  Grid row i = the samples of part i; its cells are
// db2db_sp(part i, part j) for j < i and all2all_sp(part i) on the diagonal; a row of the table concatenates the row's
// cells with the column ids shifted by the samples of the parts before (SparseMatrix::saveRowSparse(row, out, idx_shift),
// src/array.h:625-637).  The reference keeps two parts in host memory and reads every column part again for every grid
// row; here a part is read and staged on the device ONCE and stays there while the parts fit the HBM budget (the parts
// with the lowest numbers first: part j is a column of every row after it), so the grid costs O(parts) loads instead of
// O(parts^2).  With -gpus n the grid rows are dealt to n devices (they are independent) and written in order.
namespace {
struct PartsGrid {
    std::vector<std::string> files;
    std::vector<uint32_t> part_samples, first_sample;   // per part
    std::vector<uint64_t> resident_bytes;                // per part: estimate of what it occupies on the device
    Trie all;                                            // header + names + k-mer counts of every sample, in order
};

// Staged parts of one device: handle per part (-1 = not on the device).
struct PartsOnDevice {
    const PartsGrid& g;
    const SimilarityCalculator& calc;
    std::vector<int> handle;
    uint64_t budget, used = 0;
    PartsOnDevice(const PartsGrid& grid, const SimilarityCalculator& c, uint64_t budget_bytes)
        : g(grid), calc(c), handle(grid.files.size(), -1), budget(budget_bytes) {}
    // stages part i if it is not there; `keep`: it may stay (it fits the budget)
    int acquire(uint32_t i, bool& keep) {
        keep = true;
        if (handle[i] >= 0) return handle[i];
        std::cerr << "Deserializing database " << i + 1 << " (" << g.files[i] << ")" << std::endl;
        Trie db;
        read_db(g.files[i], db, true);
        const int h = calc.stage_part(db);
        if (used + g.resident_bytes[i] <= budget) { handle[i] = h; used += g.resident_bytes[i]; }
        else keep = false;
        return h;
    }
};

// -sample-rows: the rows are chosen from the whole grid (a sample's neighbours sit in its row's cells AND in the cells of
// the rows below it), so the cells go into one sampler and the table is written after the last of them
// (src/console_all2all_parts.cpp:188-191,272-275,333-345)
struct GridSampler {
    RowSampler sampler;
    std::mutex mu;   // (-gpus n: the devices' rows arrive concurrently)
    GridSampler(size_t samples, uint32_t count, metric_fn criterion) : sampler(samples, count, criterion) {}
};

// the text of grid row i_row (all lines of its samples); returns the number of pairs written.  With a sampler the cells
// are handed to it instead and the text stays empty.
uint64_t parts_grid_row(const Params& p, const PartsGrid& g, const SimilarityCalculator& calc, PartsOnDevice& dev, uint32_t i_row,
                        std::string& text, kdbx_stats& total, GridSampler* sampling) {
    bool keep_row = true;
    const int row = dev.acquire(i_row, keep_row);
    const uint32_t rows = g.part_samples[i_row];
    std::vector<std::unique_ptr<SparseMatrix<uint32_t>>> cells(i_row + 1);
    auto add_stats = [&](const kdbx_stats& st) {
        total.updates += st.updates; total.probes += st.probes; total.hits += st.hits; total.ms_probe += st.ms_probe;
        total.ms_scatter += st.ms_scatter; total.ms_compact += st.ms_compact; total.ms_total += st.ms_total;
        total.ms_prepare += st.ms_prepare; total.ms_download += st.ms_download; total.kernel_launches += st.kernel_launches;
    };
    for (uint32_t i_col = 0; i_col < i_row; ++i_col) {
        bool keep_col = true;
        const int col = dev.acquire(i_col, keep_col);
        std::cerr << "Processing cell (" << i_row + 1 << "," << i_col + 1 << ")" << std::endl;
        cells[i_col] = std::make_unique<SparseMatrix<uint32_t>>();
        calc.db2db_sp(row, col, *cells[i_col], p.filters);
        add_stats(calc.last_stats());
        if (!keep_col) calc.drop_part(col);
    }
    std::cerr << "Processing cell (" << i_row + 1 << "," << i_row + 1 << ")" << std::endl;
    cells[i_row] = std::make_unique<SparseMatrix<uint32_t>>();
    calc.all2all_sp_part(row, *cells[i_row], p.filters);
    add_stats(calc.last_stats());
    if (!keep_row) calc.drop_part(row);

    const OutputFilters* filters = p.filters.trivial() ? nullptr : &p.filters;   // (again: the bounds the device left to the host)
    const int k = (int)g.all.hdr.kmer_length;
    uint64_t saved = 0;
    text.clear();
    if (sampling) {
        std::lock_guard<std::mutex> lk(sampling->mu);
        for (uint32_t c = 0; c <= i_row; ++c)
            sampling->sampler.add_cell(*cells[c]->raw(), filters, g.all.sample_kmers.data() + g.first_sample[i_row],
                                       g.all.sample_kmers.data() + g.first_sample[c], g.first_sample[i_row], g.first_sample[c], k);
        return 0;
    }
    std::string line;
    for (uint32_t r = 0; r < rows; ++r) {
        const uint32_t s = g.first_sample[i_row] + r;
        size_t pairs = 0;
        for (uint32_t c = 0; c <= i_row; ++c) pairs += cells[c]->getNoInRow(r);
        line.resize(g.all.sample_names[s].size() + 32 + pairs * 22);
        char* q = line.data();
        std::memcpy(q, g.all.sample_names[s].data(), g.all.sample_names[s].size()); q += g.all.sample_names[s].size();
        *q++ = ',';
        q = put_u64(q, g.all.sample_kmers[s]);
        *q++ = ',';
        for (uint32_t c = 0; c <= i_row; ++c) {
            const SparseMatrix<uint32_t>& m = *cells[c];
            const uint32_t* cols = m.cols(r);
            const uint32_t* vals = m.vals(r);
            const size_t cnt = m.getNoInRow(r);
            for (size_t i = 0; i < cnt; ++i) {
                const uint32_t sc = g.first_sample[c] + cols[i];
                if (filters && !filters->pass(vals[i], (uint32_t)g.all.sample_kmers[s], (uint32_t)g.all.sample_kmers[sc], k)) continue;
                q = put_u64(q, (uint64_t)sc + 1); *q++ = ':'; q = put_u64(q, vals[i]); *q++ = ',';
                ++saved;
            }
        }
        *q++ = '\n';
        text.append(line.data(), (size_t)(q - line.data()));
    }
    return saved;
}

uint64_t file_bytes(const std::string& path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    return f ? (uint64_t)f.tellg() : 0;
}
}  // namespace

void run_all2all_parts(const Params& p) {
    if (p.files.size() != 2) throw usage_error(p.mode);
    std::cerr << "All versus all comparison (sparse computation)" << std::endl;
    PartsGrid g;
    {
        std::ifstream ifs(p.files[0]);
        if (!ifs) throw std::runtime_error("Cannot open: " + p.files[0]);
        g.files.assign(std::istream_iterator<std::string>(ifs), std::istream_iterator<std::string>());
    }
    const uint32_t parts = (uint32_t)g.files.size();
    std::cerr << "Processing database grid of size " << parts << " by " << parts << std::endl;
    const double t0 = now();
    for (uint32_t i = 0; i < parts; ++i) {   // names and k-mer counts of all samples (DeserializationMode::SamplesOnly)
        Trie t;
        try { read_db(g.files[i], t, false); }
        catch (const std::runtime_error&) { throw std::runtime_error("Cannot open k-mer database: " + g.files[i]); }
        if (i == 0) g.all.hdr = t.hdr;
        else {
            if (t.hdr.kmer_length != g.all.hdr.kmer_length) throw std::runtime_error("Different k - mer lengths");
            if (t.hdr.fraction != g.all.hdr.fraction) throw std::runtime_error("Different fractions");
        }
        g.first_sample.push_back((uint32_t)g.all.sample_names.size());
        g.part_samples.push_back(t.num_samples());
        // on the device: the raw tables (allocated slots, up to 2.5x the filled ones the file holds), the trie, its decoded
        // local lists and the working buffers of the cells — a generous multiple of the file
        g.resident_bytes.push_back(file_bytes(g.files[i]) * 8 + ((uint64_t)64 << 20));
        g.all.sample_names.insert(g.all.sample_names.end(), t.sample_names.begin(), t.sample_names.end());
        g.all.sample_kmers.insert(g.all.sample_kmers.end(), t.sample_kmers.begin(), t.sample_kmers.end());
    }
    FILE* f = std::fopen(p.files[1].c_str(), "wb");
    if (!f) throw std::runtime_error("Cannot open output file " + p.files[1]);
    const std::string head = table_header(g.all);
    std::fwrite(head.data(), 1, head.size(), f);

    const int num_gpus = std::max(1, std::min<int>(p.num_gpus, (int)std::max<uint32_t>(1, parts)));
    if (num_gpus > kdbx_device_count()) { std::fclose(f); throw std::runtime_error("-gpus " + std::to_string(num_gpus) + " requested but only " + std::to_string(kdbx_device_count()) + " B200 device(s) are visible"); }
    const uint64_t budget = (uint64_t)(p.cache_buffer_mb > 8 ? p.cache_buffer_mb : 100 * 1024) << 20;   // -buffer <mb> bounds the resident parts (default 100 GB)
    uint64_t saved = 0;
    kdbx_stats total{};
    std::unique_ptr<GridSampler> sampling;
    if (p.sampling_size > 0) sampling = std::make_unique<GridSampler>(g.all.sample_names.size(), (uint32_t)p.sampling_size, p.sampling_criterion);
    if (num_gpus == 1) {
        SimilarityCalculator calc(p.num_threads, (size_t)p.cache_buffer_mb, p.gpu);
        PartsOnDevice dev(g, calc, budget);
        std::string text;
        for (uint32_t i = 0; i < parts; ++i) {
            saved += parts_grid_row(p, g, calc, dev, i, text, total, sampling.get());
            std::cerr << "Saving output matrix..." << std::endl;
            std::fwrite(text.data(), 1, text.size(), f);
            std::cerr << " OK (no. currently saved pairs: " << saved << ")" << std::endl;
        }
    } else {
        // grid rows dealt round-robin; every device keeps its own staged parts; the writer takes the rows in order
        std::vector<std::string> texts(parts);
        std::vector<int> ready(parts, 0);
        std::vector<std::string> errors((size_t)num_gpus);
        std::vector<uint64_t> saved_g((size_t)num_gpus, 0);
        std::vector<kdbx_stats> stats_g((size_t)num_gpus, kdbx_stats{});
        std::mutex mu;
        std::condition_variable cv;
        std::vector<std::thread> workers;
        const int base = p.gpu < 0 ? 0 : p.gpu;
        for (int d = 0; d < num_gpus; ++d)
            workers.emplace_back([&, d] {
                try {
                    SimilarityCalculator calc(p.num_threads, (size_t)p.cache_buffer_mb, base + d);
                    PartsOnDevice dev(g, calc, budget);
                    for (uint32_t i = (uint32_t)d; i < parts; i += (uint32_t)num_gpus) {
                        std::string text;
                        saved_g[(size_t)d] += parts_grid_row(p, g, calc, dev, i, text, stats_g[(size_t)d], sampling.get());
                        std::lock_guard<std::mutex> lk(mu);
                        texts[i] = std::move(text); ready[i] = 1;
                        cv.notify_all();
                    }
                } catch (const std::exception& e) {
                    std::lock_guard<std::mutex> lk(mu);
                    errors[(size_t)d] = e.what();
                    for (uint32_t i = (uint32_t)d; i < parts; i += (uint32_t)num_gpus) if (!ready[i]) ready[i] = -1;
                    cv.notify_all();
                }
            });
        bool failed = false;
        for (uint32_t i = 0; i < parts && !failed; ++i) {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return ready[i] != 0; });
            if (ready[i] < 0) { failed = true; break; }
            std::string text = std::move(texts[i]);
            lk.unlock();
            std::fwrite(text.data(), 1, text.size(), f);
        }
        for (auto& w : workers) w.join();
        for (const std::string& e : errors) if (!e.empty()) { std::fclose(f); throw std::runtime_error(e); }
        for (int d = 0; d < num_gpus; ++d) {
            saved += saved_g[(size_t)d];
            total.updates += stats_g[(size_t)d].updates; total.probes += stats_g[(size_t)d].probes; total.hits += stats_g[(size_t)d].hits;
            total.kernel_launches += stats_g[(size_t)d].kernel_launches;
            total.ms_total = std::max(total.ms_total, stats_g[(size_t)d].ms_total);
        }
    }
    if (sampling) saved = sampling->sampler.write_rows(f, g.all.sample_names, g.all.sample_kmers);
    if (std::fclose(f) != 0) throw std::runtime_error("Cannot write output file " + p.files[1]);
    const double dt = now() - t0;
    std::cerr << "Database grid procesed successfully" << std::endl << "No. saved pairs: " << saved << std::endl;
    print_stats_json(total, dt);
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
    std::cerr << "Processing queries..." << std::endl;
    t0 = now();
    QueryTableWriter writer(p.files[2], db, p.sparse_out, &p.filters);
    if (p.from_minhash) {
        // the queries are k-mer sets already (<entry>.minhash): batches of them go to kdbx_new2all_batch as they are
        SampleStream queries(p.files[1], SampleStream::FromMinhash{}, p.num_reader_threads);
        const size_t N = db.num_samples();
        std::vector<std::string> names;
        std::vector<uint64_t> kmers, q_off;
        std::vector<uint32_t> sims;
        kdbx_stats total{};
        SampleKmers q;
        bool more = true;
        size_t done = 0;
        while (more) {
            names.clear(); kmers.clear(); q_off.assign(1, 0);
            while (names.size() < 4096 && kmers.size() < ((size_t)1 << 27) && (more = queries.next(q))) {
                if (q.k != db.hdr.kmer_length) throw std::runtime_error("Sample and database k-mer length differ");
                kmers.insert(kmers.end(), q.kmers.begin(), q.kmers.end());
                q_off.push_back(kmers.size());
                names.push_back(std::move(q.name));
            }
            if (names.empty()) break;
            calculator.one2all_batch(kmers.data(), q_off, sims);
            const kdbx_stats& st = calculator.last_stats();
            total.probes += st.probes; total.hits += st.hits; total.ms_probe += st.ms_probe; total.ms_scatter += st.ms_scatter; total.ms_total += st.ms_total;
            for (size_t i = 0; i < names.size(); ++i) writer.write_row(names[i], q_off[i + 1] - q_off[i], sims.data() + i * N);
            done += names.size();
        }
        writer.close();
        const double dt = now() - t0;
        std::cerr << std::endl << "EXECUTION TIMES" << std::endl << "Total: " << dt << std::endl;
        print_stats_json(total, dt);
        return;
    }
    SequenceStream stream(p.files[1], p.multisample_fasta, p.num_reader_threads);
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

// one2all (src/console_one2all.cpp:12-96): ONE sample file against the database — new2all's device path with a single
// query (k-mer extraction, minhash, sort/unique, probe, scatter), the table in the mode's own layout (csv_out.h).
void run_one2all(const Params& p) {
    if (p.files.size() != 3) throw usage_error(p.mode);
    std::cerr << "One new sample  (from genomes) versus entire database comparison" << std::endl;
    SimilarityCalculator calculator(p.num_threads, (size_t)p.cache_buffer_mb, p.gpu);
    Trie db(true);
    std::cerr << "Loading k-mer database " << p.files[0] << ":" << std::endl;
    double t0 = now();
    read_db(p.files[0], db, true);
    std::cerr << "OK (" << now() - t0 << " seconds)" << std::endl;
    calculator.load_database(db);
    std::cerr << "Loading sample kmers..." << std::endl;
    if (p.from_minhash) {
        SampleKmers q;
        if (!load_minhash(p.files[1], q)) throw std::runtime_error("Cannot open sample file: " + p.files[1]);
        if (q.k != db.hdr.kmer_length) throw std::runtime_error("Sample and database k-mer length differ");
        std::vector<uint32_t> sims;
        std::cerr << "Calculating similarity vector..." << std::endl;
        t0 = now();
        calculator.one2all_batch(q.kmers.data(), {0, (uint64_t)q.kmers.size()}, sims);
        const double dt = now() - t0;
        std::cerr << "OK (" << dt << " seconds)" << std::endl << "Number of k-mers: " << q.kmers.size() << std::endl;
        print_stats_json(calculator.last_stats(), dt);
        write_one2all_csv(p.files[2], db, p.files[1], q.kmers.size(), sims.data());
        return;
    }
    SequenceStream stream(std::vector<std::string>{p.files[1]}, false, 1);
    SampleSeq s;
    if (!stream.next(s)) throw std::runtime_error("Cannot open sample file: " + p.files[1]);
    Buf<char> symbols;
    symbols.set_pinned(true);
    symbols.resize(s.symbols.size());
    std::copy(s.symbols.begin(), s.symbols.end(), symbols.data());
    const std::vector<uint64_t> q_off = {0, (uint64_t)symbols.size()};
    std::vector<uint64_t> unique;
    std::vector<uint32_t> sims;
    std::cerr << "Calculating similarity vector..." << std::endl;
    t0 = now();
    calculator.one2all_sequences(db.hdr, symbols.data(), q_off, sims, unique);
    const double dt = now() - t0;
    std::cerr << "OK (" << dt << " seconds)" << std::endl << "Number of k-mers: " << unique[0] << std::endl
              << "Minhash fraction: " << db.hdr.fraction << std::endl;
    print_stats_json(calculator.last_stats(), dt);
    std::cerr << "Storing similarity vector in " << p.files[2] << "..." << std::endl;
    write_one2all_csv(p.files[2], db, p.files[1], unique[0], sims.data());
    std::cerr << "OK" << std::endl;
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
        else if (p.mode == "minhash") run_minhash(p);
        else if (p.mode == "all2all") run_all2all(p);
        else if (p.mode == "all2all-sp") run_all2all_sparse(p);
        else if (p.mode == "all2all-parts") run_all2all_parts(p);
        else if (p.mode == "new2all") run_new2all(p);
        else if (p.mode == "one2all") run_one2all(p);
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
