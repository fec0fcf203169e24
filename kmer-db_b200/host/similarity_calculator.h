// Host-side C++ mirror of the reference's similarity engine interface:
//     class SimilarityCalculator { SimilarityCalculator(int num_threads, size_t cacheBufferMb);
//         void all2all(PrefixKmerDb& db, LowerTriangularMatrix<uint32_t>& matrix) const;
//         void all2all_sp(PrefixKmerDb& db, SparseMatrix<uint32_t>& matrix, CBubbleHelper& bubbles) const;
//         void db2db_sp(PrefixKmerDb& db1, PrefixKmerDb& db2, SparseMatrix<uint32_t>& matrix, CBubbleHelper& bubbles) const;
//         template <bool parallel> void one2all(const PrefixKmerDb& db, const kmer_t* kmers, size_t kmersCount,
//                                               std::vector<uint32_t>& similarities) const; ... }
// (src/similarity_calculator.h:4-16).  Same names and argument meaning; the work happens in
// libkdbx.so through the C ABI (include/kdbx.h).  Differences, all deliberate:
//   * `num_threads` / `cacheBufferMb` are accepted for CLI compatibility and ignored
//     ("-buffer stays accepted but is a no-op on GPU");
//   * the database is NOT mutated (the reference adds children's num_kmers into their
//     parents in place, src/similarity_calculator.cpp:64-72, so a second call double counts);
//   * errors surface as std::runtime_error, like every error in the reference
//     (src/main.cpp:56-59); there is no CPU fallback.
#pragma once
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <thread>
#include <string>
#include <vector>

#include "../../include/kdbx.h"
#include "kmers.h"
#include "metrics.h"
#include "trie.h"

namespace kdbx {

// Packed lower-triangular matrix with the reference's layout (src/array.h:120-269):
// row i starts at i(i-1)/2 and has i cells.
template <class T>
class LowerTriangularMatrix {
public:
    void resize(size_t size) { size_ = size; data_.assign(size * (size > 0 ? size - 1 : 0) / 2, T()); }
    size_t getSize() const { return size_; }
    T* operator[](size_t i) { return data_.data() + i * (i - 1) / 2; }
    const T* operator[](size_t i) const { return data_.data() + i * (i - 1) / 2; }
    T* data() { return data_.data(); }
    const T* data() const { return data_.data(); }
    size_t cells() const { return data_.size(); }
private:
    size_t size_ = 0;
    std::vector<T> data_;
};

// Rows of ascending (col, val) pairs: what the reference's SparseMatrix<T> holds after compact2
// (src/array.h:275-650, data_compacted).  Owns the library-allocated CSR arrays.
template <class T>
class SparseMatrix {
public:
    SparseMatrix() { csr_ = kdbx_csr{}; }
    ~SparseMatrix() { kdbx_free_csr(&csr_); }
    SparseMatrix(const SparseMatrix&) = delete;
    SparseMatrix& operator=(const SparseMatrix&) = delete;
    size_t getSize() const { return csr_.num_rows; }
    size_t getNoInRow(size_t row) const { return (size_t)(csr_.row_ptr[row + 1] - csr_.row_ptr[row]); }
    const uint32_t* cols(size_t row) const { return csr_.col + csr_.row_ptr[row]; }
    const T* vals(size_t row) const { return csr_.val + csr_.row_ptr[row]; }
    uint64_t nnz() const { return csr_.nnz; }
    kdbx_csr* raw() { return &csr_; }
private:
    kdbx_csr csr_;
};

class SimilarityCalculator {
public:
    SimilarityCalculator(int num_threads, size_t cacheBufferMb, int device = -1)
        : num_threads_(num_threads), cache_buffer_mb_(cacheBufferMb), device_(device) {
        kdbx_config cfg{};
        cfg.device = device;
        if (kdbx_open(&cfg, &ctx_) != KDBX_OK) throw std::runtime_error(kdbx_last_error(nullptr));
    }
    ~SimilarityCalculator() { kdbx_close(ctx_); for (Part& p : parts_) if (p.ctx) kdbx_close(p.ctx); }
    SimilarityCalculator(const SimilarityCalculator&) = delete;
    SimilarityCalculator& operator=(const SimilarityCalculator&) = delete;

    // src/similarity_calculator.cpp:42: matrix is resized to N and filled with the numbers
    // of shared k-mers.
    void all2all(const Trie& db, LowerTriangularMatrix<uint32_t>& matrix) const {
        matrix.resize(db.num_samples());
        const kdbx_trie_view v = db.view();
        check(kdbx_load_patterns(ctx_, &v));
        check(kdbx_all2all_dense(ctx_, matrix.data(), &stats_));
    }
    // Contiguous row blocks with (nearly) equal numbers of updates: boundaries[g] .. boundaries[g+1]
    // is GPU g's share.  Same rule as the reference's row-aligned ranges between threads
    // (src/similarity_calculator.cpp:371-395), balanced on the per-row update counts.
    static std::vector<uint32_t> shard_rows_by_work(const std::vector<uint64_t>& row_updates, int parts) {
        const uint32_t N = (uint32_t)row_updates.size();
        long double total = 0;
        for (uint64_t u : row_updates) total += (long double)u;
        std::vector<uint32_t> bounds(1, 0);
        long double run = 0;
        uint32_t r = 0;
        for (int g = 1; g < parts; ++g) {
            const long double target = total * g / parts;
            while (r < N && run < target) run += (long double)row_updates[r++];
            bounds.push_back(r);
        }
        bounds.push_back(N);
        return bounds;
    }

    // The dense matrix on several GPUs of one node.  The DATABASE is sharded: the trie is cut into one sub-trie per
    // device (TriePartitioner: pieces of the depth-first preorder plus the ancestor chain of each piece with
    // num_kmers = 0 — the matrix is linear in num_kmers, so the parts' matrices add up), every device stages only its
    // own part, declares the band of sample ids it covers and runs the whole single-GPU pipeline on it; ONE
    // ncclReduceScatter inside the library adds the partial matrices and leaves device g with block g of the packed
    // triangle, which it copies straight into the caller's matrix.  One host thread per device.
    // (The reference's only sharding template is the grid of partial databases of all2all-parts,
    // src/console_all2all_parts.cpp:143-331; between threads it uses row ownership, src/similarity_calculator.cpp:371-395.)
    void all2all_multi(const Trie& db, LowerTriangularMatrix<uint32_t>& matrix, int num_gpus) const {
        const int avail = kdbx_device_count();
        if (num_gpus > avail) throw std::runtime_error("-gpus " + std::to_string(num_gpus) + " requested but only " + std::to_string(avail) + " B200 device(s) are visible");
        matrix.resize(db.num_samples());
        const int base = device_ < 0 ? 0 : device_;
        std::vector<kdbx_ctx*> ctxs((size_t)num_gpus, nullptr);
        std::vector<std::string> errors((size_t)num_gpus);
        ctxs[0] = ctx_;
        auto close_extra = [&]() { kdbx_comm_destroy(ctx_); for (int g = 1; g < num_gpus; ++g) kdbx_close(ctxs[(size_t)g]); };
        auto on_all = [&](auto&& fn) {
            std::vector<std::thread> th;
            for (int g = 0; g < num_gpus; ++g)
                th.emplace_back([&, g] { if (errors[(size_t)g].empty()) { try { fn(g); } catch (const std::exception& e) { errors[(size_t)g] = e.what(); } } });
            for (auto& t : th) t.join();
            for (const std::string& e : errors) if (!e.empty()) { close_extra(); throw std::runtime_error(e); }
        };
        for (int g = 1; g < num_gpus; ++g) {   // (serially: kdbx_open reports its errors through one string per thread)
            kdbx_config cfg{};
            cfg.device = base + g;
            if (kdbx_open(&cfg, &ctxs[(size_t)g]) != KDBX_OK) { const std::string e = kdbx_last_error(nullptr); close_extra(); throw std::runtime_error(e); }
        }
        if (kdbx_comm_init_all(ctxs.data(), num_gpus) != KDBX_OK) { const std::string e = kdbx_last_error(ctx_); close_extra(); throw std::runtime_error(e); }
        const TriePartitioner cut(db, (uint32_t)num_gpus);
        std::vector<kdbx_stats> st((size_t)num_gpus);
        on_all([&](int g) {
            Trie part(true);
            uint32_t window[2] = {0, 0};
            cut.extract((uint32_t)g, part, nullptr, window);
            const kdbx_trie_view v = part.view();
            kdbx_ctx* c = ctxs[(size_t)g];
            if (kdbx_load_patterns(c, &v) != KDBX_OK || kdbx_set_sample_window(c, window[0], window[1]) != KDBX_OK)
                throw std::runtime_error(kdbx_last_error(c));
            const uint64_t cells = matrix.cells(), B = (cells + (uint64_t)num_gpus - 1) / (uint64_t)num_gpus;
            uint64_t first = 0, count = 0;
            uint32_t* dst = matrix.data() + std::min<uint64_t>(cells, (uint64_t)g * B);
            if (kdbx_all2all_dense_reduce_scatter(c, dst, &first, &count, &st[(size_t)g]) != KDBX_OK)
                throw std::runtime_error(kdbx_last_error(c));
        });
        stats_ = st[0];
        for (int g = 1; g < num_gpus; ++g) {  // totals; times are the slowest device's
            stats_.updates += st[(size_t)g].updates;
            stats_.physical_updates += st[(size_t)g].physical_updates;
            stats_.kernel_launches += st[(size_t)g].kernel_launches;
            stats_.ms_total = std::max(stats_.ms_total, st[(size_t)g].ms_total);
            stats_.ms_scatter = std::max(stats_.ms_scatter, st[(size_t)g].ms_scatter);
            stats_.ms_collective = std::max(stats_.ms_collective, st[(size_t)g].ms_collective);
        }
        close_extra();
    }

    // src/similarity_calculator.cpp:442 + SparseMatrix::compact2 (src/array.h:391-446): the same
    // matrix as sparse rows.  `filters` are the -min/-max bounds; the ones whose arithmetic is
    // exactly reproducible on the device run there (include/kdbx.h), the log-based ones are left
    // to the CSV emitter, which applies `filters` again on the host.
    void all2all_sp(const Trie& db, SparseMatrix<uint32_t>& matrix, const OutputFilters& filters) const {
        const kdbx_trie_view v = db.view();
        check(kdbx_load_patterns(ctx_, &v));
        kdbx_filter f{};
        std::vector<uint32_t> counts(db.sample_kmers.begin(), db.sample_kmers.end());
        make_filter(filters, counts, f);
        kdbx_free_csr(matrix.raw());
        check(kdbx_all2all_sparse(ctx_, &f, matrix.raw(), &stats_));
    }

    // all2all_sp on several GPUs of one node: every device stages the whole trie (minhashed databases are small) and takes a
    // block of rows balanced on the per-row update counts; the rows are independent, so there is no exchange, and the blocks'
    // CSR rows are concatenated here.  The grid of the reference's all2all-parts (src/console_all2all_parts.cpp:143-331) with
    // the database replicated instead of split.
    void all2all_sp_multi(const Trie& db, SparseMatrix<uint32_t>& matrix, const OutputFilters& filters, int num_gpus) const {
        const int avail = kdbx_device_count();
        if (num_gpus > avail) throw std::runtime_error("-gpus " + std::to_string(num_gpus) + " requested but only " + std::to_string(avail) + " B200 device(s) are visible");
        const uint32_t N = db.num_samples();
        const kdbx_trie_view v = db.view();
        kdbx_filter f{};
        std::vector<uint32_t> counts(db.sample_kmers.begin(), db.sample_kmers.end());
        make_filter(filters, counts, f);
        const int base = device_ < 0 ? 0 : device_;
        std::vector<kdbx_ctx*> ctxs((size_t)num_gpus, nullptr);
        ctxs[0] = ctx_;
        auto close_extra = [&]() { for (int g = 1; g < num_gpus; ++g) kdbx_close(ctxs[(size_t)g]); };
        for (int g = 1; g < num_gpus; ++g) {
            kdbx_config cfg{};
            cfg.device = base + g;
            if (kdbx_open(&cfg, &ctxs[(size_t)g]) != KDBX_OK) { const std::string e = kdbx_last_error(nullptr); close_extra(); throw std::runtime_error(e); }
        }
        std::vector<std::string> errors((size_t)num_gpus);
        auto on_all = [&](auto&& fn) {
            std::vector<std::thread> th;
            for (int g = 0; g < num_gpus; ++g)
                th.emplace_back([&, g] { try { fn(g); } catch (const std::exception& e) { errors[(size_t)g] = e.what(); } });
            for (auto& t : th) t.join();
            for (const std::string& e : errors) if (!e.empty()) { close_extra(); throw std::runtime_error(e); }
        };
        on_all([&](int g) { if (kdbx_load_patterns(ctxs[(size_t)g], &v) != KDBX_OK) throw std::runtime_error(kdbx_last_error(ctxs[(size_t)g])); });
        std::vector<uint64_t> upd(N, 0);
        if (N && kdbx_row_updates(ctx_, upd.data()) != KDBX_OK) { const std::string e = kdbx_last_error(ctx_); close_extra(); throw std::runtime_error(e); }
        const std::vector<uint32_t> bounds = shard_rows_by_work(upd, num_gpus);
        std::vector<kdbx_csr> parts((size_t)num_gpus, kdbx_csr{});
        std::vector<kdbx_stats> st((size_t)num_gpus);
        on_all([&](int g) {
            if (kdbx_all2all_sparse_rows(ctxs[(size_t)g], bounds[(size_t)g], bounds[(size_t)g + 1], &f, &parts[(size_t)g], &st[(size_t)g]) != KDBX_OK)
                throw std::runtime_error(kdbx_last_error(ctxs[(size_t)g]));
        });
        // concatenate: block g holds the rows [bounds[g], bounds[g+1]) and nothing else
        kdbx_free_csr(matrix.raw());
        kdbx_csr& out = *matrix.raw();
        uint64_t nnz = 0;
        for (const kdbx_csr& c : parts) nnz += c.nnz;
        out.num_rows = N; out.nnz = nnz; out._pad = 0;
        out.row_ptr = static_cast<uint64_t*>(std::malloc(((size_t)N + 1) * 8));
        out.col = static_cast<uint32_t*>(std::malloc(std::max<uint64_t>(1, nnz) * 4));
        out.val = static_cast<uint32_t*>(std::malloc(std::max<uint64_t>(1, nnz) * 4));
        if (!out.row_ptr || !out.col || !out.val) { close_extra(); throw std::runtime_error("host allocation failed"); }
        uint64_t at = 0;
        for (int g = 0; g < num_gpus; ++g) {
            const kdbx_csr& c = parts[(size_t)g];
            for (uint32_t r = bounds[(size_t)g]; r < bounds[(size_t)g + 1]; ++r) out.row_ptr[r] = at + c.row_ptr[r];
            if (c.nnz) { std::memcpy(out.col + at, c.col, c.nnz * 4); std::memcpy(out.val + at, c.val, c.nnz * 4); }
            at += c.nnz;
        }
        out.row_ptr[N] = at;
        for (kdbx_csr& c : parts) kdbx_free_csr(&c);
        stats_ = st[0];
        for (int g = 1; g < num_gpus; ++g) {
            stats_.updates += st[(size_t)g].updates;
            stats_.kernel_launches += st[(size_t)g].kernel_launches;
            stats_.ms_total = std::max(stats_.ms_total, st[(size_t)g].ms_total);
        }
        close_extra();
    }

    // Stage the database for queries: patterns + k-mer tables (PrefixKmerDb::deserialize with
    // DeserializationMode::Everything, src/console_new2all.cpp:32).
    void load_database(const Trie& db) const {
        stage(ctx_, db);
        num_samples_ = db.num_samples();
    }

    // ---- all2all-parts: partial databases staged side by side on this calculator's device -----------------------------
    // The reference holds two parts in host memory at a time and reads every column part again for every grid row
    // (src/console_all2all_parts.cpp:209-221); 180 GB of HBM hold many parts at once, so a part is staged ONCE — patterns
    // and raw k-mer tables, on a context of its own — and stays until it is dropped.  Returns the part's handle.
    int stage_part(const Trie& db) const {
        kdbx_ctx* c = nullptr;
        kdbx_config cfg{};
        cfg.device = device_;
        if (kdbx_open(&cfg, &c) != KDBX_OK) throw std::runtime_error(kdbx_last_error(nullptr));
        try { stage(c, db); } catch (...) { kdbx_close(c); throw; }
        Part staged;
        staged.ctx = c;
        staged.counts.assign(db.sample_kmers.begin(), db.sample_kmers.end());
        for (size_t h = 0; h < parts_.size(); ++h) if (!parts_[h].ctx) { parts_[h] = std::move(staged); return (int)h; }
        parts_.push_back(std::move(staged));
        return (int)parts_.size() - 1;
    }
    void drop_part(int h) const {
        if (h < 0 || (size_t)h >= parts_.size() || !parts_[(size_t)h].ctx) return;
        kdbx_close(parts_[(size_t)h].ctx);
        parts_[(size_t)h] = Part();
    }
    // src/similarity_calculator.cpp:1225 + SparseMatrix::compact2: matrix row s1 (a sample of the row part) = ascending
    // (s2, common k-mers) pairs over the samples s2 of the column part.
    void db2db_sp(int row_part, int col_part, SparseMatrix<uint32_t>& matrix, const OutputFilters& filters) const {
        const Part& r = part(row_part);
        const Part& c = part(col_part);
        kdbx_filter f{};
        make_filter(filters, r.counts, f);
        kdbx_free_csr(matrix.raw());
        if (kdbx_db2db_sparse(r.ctx, c.ctx, &f, c.counts.data(), matrix.raw(), &stats_) != KDBX_OK) throw std::runtime_error(kdbx_last_error(r.ctx));
    }
    // the diagonal cell: all2all_sp of a staged part
    void all2all_sp_part(int h, SparseMatrix<uint32_t>& matrix, const OutputFilters& filters) const {
        const Part& r = part(h);
        kdbx_filter f{};
        make_filter(filters, r.counts, f);
        kdbx_free_csr(matrix.raw());
        if (kdbx_all2all_sparse(r.ctx, &f, matrix.raw(), &stats_) != KDBX_OK) throw std::runtime_error(kdbx_last_error(r.ctx));
    }

    // A batch of one2all<false> calls (src/similarity_calculator.cpp:810-925): query q owns
    // kmers[q_off[q] .. q_off[q+1]) (ascending, unique); similarities is resized to
    // n_queries x N and row q receives the numbers of k-mers shared with every database sample.
    void one2all_batch(const uint64_t* kmers, const std::vector<uint64_t>& q_off, std::vector<uint32_t>& similarities) const {
        const uint32_t nq = q_off.empty() ? 0 : (uint32_t)(q_off.size() - 1);
        similarities.assign((size_t)nq * num_samples_, 0);
        check(kdbx_new2all_batch(ctx_, kmers, q_off.data(), nq, similarities.data(), &stats_));
    }
    // The same from the queries' sequences: k-mer extraction, minhash, sort and unique run on the device with
    // the database's own parameters (what the reference's loader threads do per query on the host,
    // src/console_new2all.cpp:64-94).  Query q = symbols[q_off[q] .. q_off[q+1]) (ingest.h: SampleSeq);
    // unique_kmers receives each query's number of distinct k-mers.
    void one2all_sequences(const DbHeader& hdr, const char* symbols, const std::vector<uint64_t>& q_off,
                           std::vector<uint32_t>& similarities, std::vector<uint64_t>& unique_kmers) const {
        const uint32_t nq = q_off.empty() ? 0 : (uint32_t)(q_off.size() - 1);
        const Alphabet al = Alphabet::make(hdr.alphabet_type);
        kdbx_build_params bp{};
        bp.kmer_length = hdr.kmer_length; bp.bits_per_symbol = (uint32_t)al.bits_per_symbol; bp.alphabet_size = (uint32_t)al.size;
        bp.preserve_strand = al.preserve_strand ? 1u : 0u; bp.fraction = hdr.fraction; bp.fraction_start = hdr.start_fraction;
        for (int i = 0; i < 256; ++i) bp.symbol_map[i] = al.map[i];
        similarities.assign((size_t)nq * num_samples_, 0);
        unique_kmers.assign(nq, 0);
        check(kdbx_new2all_sequences(ctx_, &bp, symbols, q_off.data(), nq, similarities.data(), unique_kmers.data(), &stats_));
    }
    const kdbx_stats& last_stats() const { return stats_; }
    kdbx_ctx* context() const { return ctx_; }   // for the emitters that format on the device (csv_out.h)

private:
    // the -min/-max bounds whose arithmetic is exactly reproducible on the device (include/kdbx.h)
    static void make_filter(const OutputFilters& filters, const std::vector<uint32_t>& counts, kdbx_filter& f) {
        f.min_common = filters.kmers_lo; f.max_common = filters.kmers_hi;
        for (const auto& m : filters.metrics) {
            int id = -1;
            if (m.first == "jaccard") id = KDBX_METRIC_JACCARD;
            else if (m.first == "min") id = KDBX_METRIC_MIN;
            else if (m.first == "max") id = KDBX_METRIC_MAX;
            else if (m.first == "cosine") id = KDBX_METRIC_COSINE;
            if (id < 0 || f.num_metric_bounds == 4) continue;
            kdbx_metric_bound& b = f.metric_bounds[f.num_metric_bounds++];
            b.metric = id; b.lo = m.second.lo; b.hi = m.second.hi;
        }
        f.sample_kmers = counts.data();
    }
    void check(int rc) const { if (rc != KDBX_OK) throw std::runtime_error(kdbx_last_error(ctx_)); }
    // patterns + k-mer tables of `db` to the device of context c.  (The tables go through one contiguous pageable copy:
    // page-locking a buffer that is used once costs more than the slower copy.)
    static void stage(kdbx_ctx* c, const Trie& db) {
        if (db.tables.empty()) throw std::runtime_error("database was loaded without its k-mer tables");
        const kdbx_trie_view v = db.view();
        if (kdbx_load_patterns(c, &v) != KDBX_OK) throw std::runtime_error(kdbx_last_error(c));
        std::vector<uint64_t> off(db.tables.size() + 1, 0);
        for (size_t t = 0; t < db.tables.size(); ++t) off[t + 1] = off[t] + db.tables[t].slots.size();
        std::unique_ptr<uint64_t[]> slots(new uint64_t[off.back() + 1]);   // (not value-initialised: every slot is copied below)
        parallel_ranges(db.tables.size(), 64, [&](size_t b, size_t e) {
            for (size_t t = b; t < e; ++t) std::copy(db.tables[t].slots.begin(), db.tables[t].slots.end(), slots.get() + off[t]);
        });
        kdbx_tables_view tv{};
        tv.num_tables = db.tables.size(); tv.slot_off = off.data(); tv.slots = slots.get();
        if (kdbx_load_hashtables(c, &tv) != KDBX_OK) throw std::runtime_error(kdbx_last_error(c));
    }
    struct Part {
        kdbx_ctx* ctx = nullptr;
        std::vector<uint32_t> counts;   // "total-kmers" of the part's samples
    };
    const Part& part(int h) const {
        if (h < 0 || (size_t)h >= parts_.size() || !parts_[(size_t)h].ctx) throw std::runtime_error("all2all-parts: no such staged part");
        return parts_[(size_t)h];
    }
    int num_threads_;
    size_t cache_buffer_mb_;
    int device_ = -1;
    kdbx_ctx* ctx_ = nullptr;
    mutable std::vector<Part> parts_;        // all2all-parts: the partial databases staged on this device
    mutable kdbx_stats stats_{};
    mutable uint32_t num_samples_ = 0;
};

}  // namespace kdbx
