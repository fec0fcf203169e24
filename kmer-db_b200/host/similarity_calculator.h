// Host-side C++ mirror of the reference's similarity engine interface for the dense path:
//     class SimilarityCalculator { SimilarityCalculator(int num_threads, size_t cacheBufferMb);
//         void all2all(PrefixKmerDb& db, LowerTriangularMatrix<uint32_t>& matrix) const; ... }
// (src/similarity_calculator.h:4-16).  Same names and argument meaning; the work happens in
// libkdbx.so through the C ABI (include/kdbx.h).  Differences, all deliberate:
//   * `num_threads` / `cacheBufferMb` are accepted for CLI compatibility and ignored
//     ("-buffer stays accepted but is a no-op on GPU");
//   * the database is NOT mutated (the reference adds children's num_kmers into their
//     parents in place, src/similarity_calculator.cpp:64-72, so a second call double counts);
//   * errors surface as std::runtime_error, like every error in the reference
//     (src/main.cpp:56-59); there is no CPU fallback.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/kdbx.h"
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

class SimilarityCalculator {
public:
    SimilarityCalculator(int num_threads, size_t cacheBufferMb, int device = -1)
        : num_threads_(num_threads), cache_buffer_mb_(cacheBufferMb) {
        kdbx_config cfg{};
        cfg.device = device;
        if (kdbx_open(&cfg, &ctx_) != KDBX_OK) throw std::runtime_error(kdbx_last_error(nullptr));
    }
    ~SimilarityCalculator() { kdbx_close(ctx_); }
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
    const kdbx_stats& last_stats() const { return stats_; }

private:
    void check(int rc) const { if (rc != KDBX_OK) throw std::runtime_error(kdbx_last_error(ctx_)); }
    int num_threads_;
    size_t cache_buffer_mb_;
    kdbx_ctx* ctx_ = nullptr;
    mutable kdbx_stats stats_{};
};

}  // namespace kdbx
