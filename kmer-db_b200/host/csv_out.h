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
