// Byte-exact CSV emitters for the dense / -sparse all2all table.  Layout
// follows All2AllConsole::run (src/console_all2all.cpp:40-78): two header lines, then one
// line per sample with the s cells of packed row s (src/array.h:254-262).  Integer printing
// is plain decimal like NumericConversions::Int2PChar (src/conversion.h:99-165) — written
// independently (two-digit table), rows formatted in parallel and written in order.
#include <cstdio>
#include <cstring>
#include <thread>

#include "trie.h"

namespace kdbx {
namespace {

const char kDigitPairs[] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839"
    "40414243444546474849505152535455565758596061626364656667686970717273747576777879"
    "8081828384858687888990919293949596979899";

inline char* put_u64(char* p, uint64_t v) {
    char tmp[24];
    int n = 0;
    while (v >= 100) {
        const unsigned r = (unsigned)(v % 100);
        v /= 100;
        tmp[n++] = kDigitPairs[2 * r + 1];
        tmp[n++] = kDigitPairs[2 * r];
    }
    if (v >= 10) { tmp[n++] = kDigitPairs[2 * v + 1]; tmp[n++] = kDigitPairs[2 * v]; }
    else tmp[n++] = (char)('0' + v);
    while (n) *p++ = tmp[--n];
    return p;
}

size_t format_row(const Trie& t, const uint32_t* tri, size_t s, bool sparse, std::string& out) {
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
        for (size_t c = 0; c < s; ++c) if (row[c] != 0) {
            p = put_u64(p, c + 1); *p++ = ':'; p = put_u64(p, row[c]); *p++ = ',';
        }
    }
    *p++ = '\n';
    out.resize((size_t)(p - out.data()));
    return out.size();
}

}  // namespace

void write_all2all_csv(const std::string& path, const Trie& t, const uint32_t* tri, bool sparse) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("Cannot open output file " + path);
    {
        std::string head = "kmer-length: " + std::to_string(t.hdr.kmer_length) + " fraction: ";
        char num[64];
        std::snprintf(num, sizeof num, "%g", t.hdr.fraction);  // == ostream << double
        head += num;
        head += " ,db-samples ,";
        for (const auto& s : t.sample_names) { head += s; head += ','; }
        head += "\nquery-samples,total-kmers,";
        for (uint64_t c : t.sample_kmers) { head += std::to_string(c); head += ','; }
        head += '\n';
        std::fwrite(head.data(), 1, head.size(), f);
    }
    const size_t N = t.num_samples();
    const size_t batch = 256;
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::string> rows(batch);
    for (size_t s0 = 0; s0 < N; s0 += batch) {
        const size_t s1 = std::min(N, s0 + batch);
        std::vector<std::thread> th;
        for (unsigned k = 0; k < nt; ++k)
            th.emplace_back([&, k]() {
                for (size_t s = s0 + k; s < s1; s += nt) format_row(t, tri, s, sparse, rows[s - s0]);
            });
        for (auto& x : th) x.join();
        for (size_t s = s0; s < s1; ++s) std::fwrite(rows[s - s0].data(), 1, rows[s - s0].size(), f);
    }
    if (std::fclose(f) != 0) throw std::runtime_error("Cannot write output file " + path);
}

}  // namespace kdbx
