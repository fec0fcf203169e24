// Byte-exact emitters of the three CSV tables; see csv_out.cpp.
#pragma once
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/kdbx.h"
#include "metrics.h"
#include "trie.h"

namespace kdbx {

std::string table_header(const Trie& t);
void write_all2all_csv(const std::string& path, const Trie& t, const uint32_t* tri, bool sparse, const OutputFilters* filters = nullptr);
// The dense table with the cells' text formatted on the device (kdbx_csv_dense_rows) from the matrix the context's last
// kdbx_all2all_dense call left in HBM; the host adds names, k-mer counts and newlines.  Same bytes as write_all2all_csv.
void write_all2all_csv_device(const std::string& path, const Trie& t, kdbx_ctx* ctx);
uint64_t write_sparse_csv(const std::string& path, const Trie& t, const kdbx_csr& m, const OutputFilters* filters);

// -sample-rows <criterion>:<count> of all2all-sp and all2all-parts (src/sampler.h, SparseMatrix::add_to_sampler
// src/array.h:451-541, src/console_all2all_sparse.cpp:70-90, src/console_all2all_parts.cpp:137,188-191,272-275,333-345):
// every sample keeps the <count> best of ALL its neighbours — the cells of its own row and of its column, i.e. the
// matrix taken as symmetric — among those that pass the -min/-max filters; "best" = the highest criterion value, ties
// by the lower sample id (what the reference's heap order, score descending then item ascending, comes to, whatever the
// order of insertion); a row is written in ascending sample order.  The criterion of a pair is always evaluated with
// the k-mer count of the pair's ROW sample first (the larger id), also for the mirrored entry.
// The reference's second strategy — random selection when no criterion is given — depends on the iteration order of
// its per-row hash maps and is not offered.
class RowSampler {
public:
    RowSampler(size_t num_samples, uint32_t max_items, metric_fn criterion);
    // one cell of the grid: rows of `m` are samples row_shift + r, its columns samples col_shift + c
    // (all2all-sp: the whole matrix with both shifts 0); row_kmers / col_kmers = k-mer counts of those samples
    void add_cell(const kdbx_csr& m, const OutputFilters* filters, const uint64_t* row_kmers, const uint64_t* col_kmers,
                  uint32_t row_shift, uint32_t col_shift, int k);
    // `<name>,<k-mers>,` + `id+1:value,` pairs per sample; returns the number of pairs written
    uint64_t write_rows(FILE* f, const std::vector<std::string>& names, const std::vector<uint64_t>& kmers);
private:
    struct Item { uint32_t item, value; double score; };
    static bool heap_order(const Item& x, const Item& y) { return x.score != y.score ? x.score > y.score : x.item < y.item; }
    void add(size_t row, uint32_t item, uint32_t value, double score);
    std::vector<std::vector<Item>> rows_;
    size_t max_items_;
    metric_fn criterion_;
};
// all2all-sp table with -sample-rows: headers, then the sampler's rows
uint64_t write_sparse_csv_sampled(const std::string& path, const Trie& t, const kdbx_csr& m, const OutputFilters* filters,
                                  uint32_t max_items, metric_fn criterion);

// one2all table (src/console_one2all.cpp:82-92): the two database header lines, then ONE row `<sample as given>,<its unique
// k-mers>,` + N dense cells — and no newline after it
void write_one2all_csv(const std::string& path, const Trie& db, const std::string& sample, uint64_t kmers, const uint32_t* sims);

class QueryTableWriter {
public:
    QueryTableWriter(const std::string& path, const Trie& db, bool sparse, const OutputFilters* filters);
    ~QueryTableWriter();
    void write_row(const std::string& name, uint64_t kmers, const uint32_t* sims);
    void close();
private:
    const Trie& db_;
    bool sparse_;
    const OutputFilters* filters_;
    FILE* f_ = nullptr;
    std::vector<char> buf_;
};

}  // namespace kdbx
