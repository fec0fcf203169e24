// Byte-exact CSV emitters for the dense / -sparse all2all table.  Layout
// follows All2AllConsole::run (src/console_all2all.cpp:40-78): two header lines, then one
// line per sample with the s cells of packed row s (src/array.h:254-262).  Integer printing
// is plain decimal like NumericConversions::Int2PChar (src/conversion.h:99-165) — written
// independently (two-digit table), rows formatted in parallel and written in order.
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <thread>

#include "csv_out.h"
#include "numfmt.h"

namespace kdbx {
namespace {

size_t format_row(const Trie& t, const uint32_t* tri, size_t s, bool sparse, const OutputFilters* filters, std::string& out) {
    const std::string& name = t.sample_names[s];
    out.resize(name.size() + 32 + s * (sparse ? 22 : 11));
    char* p = out.data();
    std::memcpy(p, name.data(), name.size()); p += name.size();
    *p++ = ',';
    p = put_u64(p, t.sample_kmers[s]);
    *p++ = ',';
    const uint32_t* row = tri + s * (s - 1) / 2;  // src/array.h:140 (row 0 is empty)
    if (!sparse) {
        for (size_t c = 0; c < s; ++c) { p = put_u64(p, row[c]); *p++ = ','; }
    } else {  // <col+1>:<val>, for non-zero cells (src/conversion.h:286-298)
        // cells failing the -min/-max filters are zeroed first (LowerTriangularMatrix::compact,
        // src/array.h:169-181), then zeros are skipped
        const int k = (int)t.hdr.kmer_length;
        for (size_t c = 0; c < s; ++c) if (row[c] != 0) {
            if (filters && !filters->pass(row[c], (uint32_t)t.sample_kmers[s], (uint32_t)t.sample_kmers[c], k)) continue;
            p = put_u64(p, c + 1); *p++ = ':'; p = put_u64(p, row[c]); *p++ = ',';
        }
    }
    *p++ = '\n';
    out.resize((size_t)(p - out.data()));
    return out.size();
}

}  // namespace

std::string table_header(const Trie& t) {
    std::string head = "kmer-length: " + std::to_string(t.hdr.kmer_length) + " fraction: ";
    char num[64];
    std::snprintf(num, sizeof num, "%g", t.hdr.fraction);  // == ostream << double
    head += num;
    head += " ,db-samples ,";
    for (const auto& s : t.sample_names) { head += s; head += ','; }
    head += "\nquery-samples,total-kmers,";
    for (uint64_t c : t.sample_kmers) { head += std::to_string(c); head += ','; }
    head += '\n';
    return head;
}

void write_all2all_csv(const std::string& path, const Trie& t, const uint32_t* tri, bool sparse, const OutputFilters* filters) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("Cannot open output file " + path);
    const std::string head = table_header(t);
    std::fwrite(head.data(), 1, head.size(), f);
    if (filters && filters->trivial()) filters = nullptr;
    const size_t N = t.num_samples();
    const size_t batch = 256;
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::string> rows(batch);
    for (size_t s0 = 0; s0 < N; s0 += batch) {
        const size_t s1 = std::min(N, s0 + batch);
        std::vector<std::thread> th;
        for (unsigned k = 0; k < nt; ++k)
            th.emplace_back([&, k]() {
                for (size_t s = s0 + k; s < s1; s += nt) format_row(t, tri, s, sparse, filters, rows[s - s0]);
            });
        for (auto& x : th) x.join();
        for (size_t s = s0; s < s1; ++s) std::fwrite(rows[s - s0].data(), 1, rows[s - s0].size(), f);
    }
    if (std::fclose(f) != 0) throw std::runtime_error("Cannot write output file " + path);
}

void write_all2all_csv_device(const std::string& path, const Trie& t, kdbx_ctx* ctx) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("Cannot open output file " + path);
    const std::string head = table_header(t);
    std::fwrite(head.data(), 1, head.size(), f);
    const uint32_t N = t.num_samples();
    // row blocks of at most ~256 MB of text; sizes first, then the text into a page-locked buffer
    std::vector<uint64_t> off;
    std::vector<char> line;
    char* text = nullptr;
    uint64_t cap = 0;
    auto fail = [&](const std::string& what) { if (text) kdbx_host_free(text); std::fclose(f); throw std::runtime_error(what); };
    uint32_t r0 = 0;
    while (r0 < N) {
        uint32_t r1 = r0;
        uint64_t est = 0;
        while (r1 < N && (r1 == r0 || est + (uint64_t)r1 * 11 <= ((uint64_t)256 << 20))) { est += (uint64_t)r1 * 11; ++r1; }
        off.assign((size_t)(r1 - r0) + 1, 0);
        uint64_t bytes = 0;
        if (kdbx_csv_dense_rows(ctx, r0, r1, nullptr, 0, off.data(), &bytes) != KDBX_OK) fail(kdbx_last_error(ctx));
        if (bytes > cap) {
            if (text) kdbx_host_free(text);
            text = nullptr;
            void* v = nullptr;
            if (kdbx_host_alloc(&v, bytes + 16) != KDBX_OK) fail("pinned allocation failed");
            text = static_cast<char*>(v); cap = bytes;
        }
        if (kdbx_csv_dense_rows(ctx, r0, r1, text, cap, off.data(), &bytes) != KDBX_OK) fail(kdbx_last_error(ctx));
        for (uint32_t s = r0; s < r1; ++s) {
            const std::string& name = t.sample_names[s];
            const uint64_t b = off[s - r0], e = off[s - r0 + 1];
            line.resize(name.size() + 32 + (size_t)(e - b));
            char* p = line.data();
            std::memcpy(p, name.data(), name.size()); p += name.size();
            *p++ = ',';
            p = put_u64(p, t.sample_kmers[s]);
            *p++ = ',';
            if (e > b) { std::memcpy(p, text + b, (size_t)(e - b)); p += e - b; }
            *p++ = '\n';
            std::fwrite(line.data(), 1, (size_t)(p - line.data()), f);
        }
        r0 = r1;
    }
    if (text) kdbx_host_free(text);
    if (std::fclose(f) != 0) throw std::runtime_error("Cannot write output file " + path);
}

// all2all-sp table (src/console_all2all_sparse.cpp:50-98): same headers, rows of `col+1:val,`
// (SparseMatrix::saveRowSparse, src/array.h:625-637).  `filters` (may be NULL) are applied again
// here for the bounds the device did not evaluate.
uint64_t write_sparse_csv(const std::string& path, const Trie& t, const kdbx_csr& m, const OutputFilters* filters) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("Cannot open output file " + path);
    const std::string head = table_header(t);
    std::fwrite(head.data(), 1, head.size(), f);
    if (filters && filters->trivial()) filters = nullptr;
    const int k = (int)t.hdr.kmer_length;
    std::string row;
    uint64_t saved = 0;
    for (size_t s = 0; s < t.num_samples(); ++s) {
        const uint64_t b = m.row_ptr ? m.row_ptr[s] : 0, e = m.row_ptr ? m.row_ptr[s + 1] : 0;
        row.resize(t.sample_names[s].size() + 32 + (size_t)(e - b) * 22);
        char* p = row.data();
        std::memcpy(p, t.sample_names[s].data(), t.sample_names[s].size()); p += t.sample_names[s].size();
        *p++ = ',';
        p = put_u64(p, t.sample_kmers[s]);
        *p++ = ',';
        for (uint64_t i = b; i < e; ++i) {
            if (filters && !filters->pass(m.val[i], (uint32_t)t.sample_kmers[s], (uint32_t)t.sample_kmers[m.col[i]], k)) continue;
            p = put_u64(p, (uint64_t)m.col[i] + 1); *p++ = ':'; p = put_u64(p, m.val[i]); *p++ = ',';
            ++saved;
        }
        *p++ = '\n';
        std::fwrite(row.data(), 1, (size_t)(p - row.data()), f);
    }
    if (std::fclose(f) != 0) throw std::runtime_error("Cannot write output file " + path);
    return saved;
}

// ---- -sample-rows (csv_out.h) ---------------------------------------------------------------------------------------
RowSampler::RowSampler(size_t num_samples, uint32_t max_items, metric_fn criterion)
    : rows_(num_samples), max_items_(max_items), criterion_(criterion) {
    if (max_items == 0 || !criterion) throw std::runtime_error("Sampling parameters error - a criterion and a positive count are needed");
}

// Bounded like the reference's (src/sampler.h:98-122): a row holds at most max_items; once full it is a heap whose top
// is the worst kept item, and a new item replaces the top iff it is better.
void RowSampler::add(size_t row, uint32_t item, uint32_t value, double score) {
    std::vector<Item>& v = rows_[row];
    const Item it{item, value, score};
    if (v.size() < max_items_) {
        v.push_back(it);
        if (v.size() == max_items_) std::make_heap(v.begin(), v.end(), heap_order);
        return;
    }
    if (!heap_order(it, v.front())) return;   // not better than the worst kept
    std::pop_heap(v.begin(), v.end(), heap_order);
    v.back() = it;
    std::push_heap(v.begin(), v.end(), heap_order);
}

void RowSampler::add_cell(const kdbx_csr& m, const OutputFilters* filters, const uint64_t* row_kmers, const uint64_t* col_kmers,
                          uint32_t row_shift, uint32_t col_shift, int k) {
    if (filters && filters->trivial()) filters = nullptr;
    if (!m.row_ptr) return;
    for (uint32_t r = 0; r < m.num_rows; ++r) {
        const uint32_t rc = (uint32_t)row_kmers[r];
        for (uint64_t i = m.row_ptr[r]; i < m.row_ptr[r + 1]; ++i) {
            const uint32_t c = m.col[i], v = m.val[i];
            const size_t a = (size_t)row_shift + r, b = (size_t)col_shift + c;
            if (a >= rows_.size() || b >= rows_.size()) throw std::runtime_error("sampler: sample id out of range");
            const uint32_t cc = (uint32_t)col_kmers[c];
            if (v == 0 || (filters && !filters->pass(v, rc, cc, k))) continue;
            const double score = criterion_(v, rc, cc, k);
            add(a, (uint32_t)b, v, score);
            add(b, (uint32_t)a, v, score);
        }
    }
}

uint64_t RowSampler::write_rows(FILE* f, const std::vector<std::string>& names, const std::vector<uint64_t>& kmers) {
    uint64_t saved = 0;
    std::string row;
    for (size_t s = 0; s < rows_.size(); ++s) {
        std::vector<Item>& v = rows_[s];
        std::sort(v.begin(), v.end(), [](const Item& x, const Item& y) { return x.item < y.item; });
        row.resize(names[s].size() + 32 + v.size() * 22);
        char* p = row.data();
        std::memcpy(p, names[s].data(), names[s].size()); p += names[s].size();
        *p++ = ',';
        p = put_u64(p, kmers[s]);
        *p++ = ',';
        for (const Item& x : v) { p = put_u64(p, (uint64_t)x.item + 1); *p++ = ':'; p = put_u64(p, x.value); *p++ = ','; }
        *p++ = '\n';
        std::fwrite(row.data(), 1, (size_t)(p - row.data()), f);
        saved += v.size();
    }
    return saved;
}

uint64_t write_sparse_csv_sampled(const std::string& path, const Trie& t, const kdbx_csr& m, const OutputFilters* filters,
                                  uint32_t max_items, metric_fn criterion) {
    RowSampler sampler(t.num_samples(), max_items, criterion);
    sampler.add_cell(m, filters, t.sample_kmers.data(), t.sample_kmers.data(), 0, 0, (int)t.hdr.kmer_length);
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("Cannot open output file " + path);
    const std::string head = table_header(t);
    std::fwrite(head.data(), 1, head.size(), f);
    const uint64_t saved = sampler.write_rows(f, t.sample_names, t.sample_kmers);
    if (std::fclose(f) != 0) throw std::runtime_error("Cannot write output file " + path);
    return saved;
}

void write_one2all_csv(const std::string& path, const Trie& db, const std::string& sample, uint64_t kmers, const uint32_t* sims) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("Cannot open output file " + path);
    const std::string head = table_header(db);
    std::fwrite(head.data(), 1, head.size(), f);
    const size_t N = db.num_samples();
    std::string row(sample.size() + 32 + N * 11, '\0');
    char* p = row.data();
    std::memcpy(p, sample.data(), sample.size()); p += sample.size();
    *p++ = ',';
    p = put_u64(p, kmers);
    *p++ = ',';
    for (size_t c = 0; c < N; ++c) { p = put_u64(p, sims[c]); *p++ = ','; }
    std::fwrite(row.data(), 1, (size_t)(p - row.data()), f);
    if (std::fclose(f) != 0) throw std::runtime_error("Cannot write output file " + path);
}

// new2all table (src/console_new2all.cpp:98-161): database headers, then one row per query in
// input order: `<name>,<unique k-mers>,` + N dense cells, or `col+1:val,` pairs for non-zero cells
// passing the filters (evaluated with the QUERY's k-mer count as the row count).
QueryTableWriter::QueryTableWriter(const std::string& path, const Trie& db, bool sparse, const OutputFilters* filters)
    : db_(db), sparse_(sparse), filters_(filters && !filters->trivial() ? filters : nullptr) {
    f_ = std::fopen(path.c_str(), "wb");
    if (!f_) throw std::runtime_error("Cannot open output file " + path);
    const std::string head = table_header(db);
    std::fwrite(head.data(), 1, head.size(), f_);
}
QueryTableWriter::~QueryTableWriter() { if (f_) std::fclose(f_); }
void QueryTableWriter::write_row(const std::string& name, uint64_t kmers, const uint32_t* sims) {
    const size_t N = db_.num_samples();
    buf_.resize(name.size() + 32 + N * 22);
    char* p = buf_.data();
    std::memcpy(p, name.data(), name.size()); p += name.size();
    *p++ = ',';
    p = put_u64(p, kmers);
    *p++ = ',';
    if (!sparse_) {
        for (size_t c = 0; c < N; ++c) { p = put_u64(p, sims[c]); *p++ = ','; }
    } else {
        const int k = (int)db_.hdr.kmer_length;
        for (size_t c = 0; c < N; ++c) if (sims[c] != 0) {
            if (filters_ && !filters_->pass(sims[c], (uint32_t)kmers, (uint32_t)db_.sample_kmers[c], k)) continue;
            p = put_u64(p, c + 1); *p++ = ':'; p = put_u64(p, sims[c]); *p++ = ',';
        }
    }
    *p++ = '\n';
    std::fwrite(buf_.data(), 1, (size_t)(p - buf_.data()), f_);
}
void QueryTableWriter::close() {
    if (f_ && std::fclose(f_) != 0) { f_ = nullptr; throw std::runtime_error("Cannot write output file"); }
    f_ = nullptr;
}

}  // namespace kdbx
