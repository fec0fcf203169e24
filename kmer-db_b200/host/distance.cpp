// `distance` mode: common-k-mer table (dense or col:val sparse CSV) -> similarity/distance
// table.  Behaviour follows DistanceConsole::run (src/console_distance.cpp:7-213): the first
// header line is re-emitted without its ",db-samples" token, the counts line is consumed, rows
// are transformed cell by cell with the chosen measure; sparse input or -sparse gives
// `col:val,` output restricted to non-zero intersections that pass the -min/-max filters; a
// table whose first row is named like the first database sample and has no value is treated as
// triangular (row r keeps r cells); -phylip-out writes a dense space-separated matrix.
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>

#include "../../include/kdbx.h"

#include "cli.h"
#include "numfmt.h"

namespace kdbx {

namespace {
inline long parse_long(const char* s, const char** end) {
    long v = 0;
    bool neg = false;
    if (*s == '-') { neg = true; ++s; }
    while (*s >= '0' && *s <= '9') v = v * 10 + (*s++ - '0');
    *end = s;
    return neg ? -v : v;
}
}  // namespace

// `distance -device`: the table is parsed here, the measure and its six-decimal text are computed on the GPU
// (kdbx_stage_matrix + kdbx_distance_dense_rows).  For what the device offers: a dense TRIANGULAR table as all2all writes
// it (row r holds r cells), dense output, one of jaccard / min / max / cosine / num-kmers, no sample without k-mers.
// Anything else is an error here — the switch is explicit, nothing falls back silently.
static void run_distance_device(const Params& params) {
    static const std::map<std::string, int> device_metrics = {{"jaccard", KDBX_METRIC_JACCARD}, {"min", KDBX_METRIC_MIN}, {"max", KDBX_METRIC_MAX},
                                                              {"cosine", KDBX_METRIC_COSINE}, {"num-kmers", KDBX_METRIC_NUM_KMERS}};
    const auto mt = device_metrics.find(params.metric_name);
    if (mt == device_metrics.end()) throw std::runtime_error("distance -device offers jaccard, min, max, cosine and num-kmers (the logarithm-based measures run on the host)");
    if (params.phylip_out || params.sparse_out) throw std::runtime_error("distance -device writes the dense table only");
    std::ifstream in(params.files[0]);
    if (!in) throw std::runtime_error("Cannot open common k-mers table: " + params.files[0]);
    std::string tok, rest;
    uint32_t k = 0;
    double fraction = 0;
    in >> tok >> k >> tok >> fraction >> tok;
    std::getline(in, rest);
    const std::string head_rest = rest;
    std::vector<uint32_t> counts;
    {
        std::getline(in, rest);
        std::replace(rest.begin(), rest.end(), ',', ' ');
        std::istringstream iss(rest);
        iss >> tok >> tok;
        size_t v;
        while (iss >> v) counts.push_back((uint32_t)v);
    }
    const size_t N = counts.size();
    std::vector<std::string> names;
    std::vector<uint32_t> tri(N ? N * (N - 1) / 2 : 0);
    std::string line;
    for (size_t r = 0; std::getline(in, line); ++r) {
        if (r >= N) throw std::runtime_error("distance -device: more rows than database samples (not an all2all table)");
        const char* begin = line.data();
        const char* end = begin + line.size();
        const char* p = std::find(begin, end, ',');
        names.emplace_back(begin, p);
        begin = p < end ? p + 1 : end;
        parse_long(begin, &p);   // the row's total-kmers: equals counts[r] in an all2all table
        begin = p < end ? p + 1 : end;
        size_t c = 0;
        uint32_t* row = tri.data() + (r ? r * (r - 1) / 2 : 0);
        for (; end - begin > 1; ++c) {
            const long v = parse_long(begin, &p);
            if (*p == ':' || c >= r) throw std::runtime_error("distance -device needs the dense triangular table all2all writes");
            row[c] = (uint32_t)v;
            begin = p < end ? p + 1 : end;
        }
        if (c != r) throw std::runtime_error("distance -device needs the dense triangular table all2all writes");
    }
    if (names.size() != N) throw std::runtime_error("distance -device: fewer rows than database samples (not an all2all table)");
    kdbx_ctx* ctx = nullptr;
    kdbx_config cfg{};
    cfg.device = params.gpu;
    if (kdbx_open(&cfg, &ctx) != KDBX_OK) throw std::runtime_error(kdbx_last_error(nullptr));
    auto fail = [&](const std::string& what) { kdbx_close(ctx); throw std::runtime_error(what); };
    if (kdbx_stage_matrix(ctx, tri.data(), (uint32_t)N) != KDBX_OK) fail(kdbx_last_error(ctx));
    std::ofstream out(params.files[1]);
    out << "kmer-length: " << k << " fraction: " << fraction << head_rest << std::endl;
    std::vector<uint64_t> off;
    std::vector<char> text;
    uint32_t r0 = 0;
    while (r0 < N) {   // row blocks of at most ~256 MB of text
        uint32_t r1 = r0;
        uint64_t est = 0;
        while (r1 < N && (r1 == r0 || est + (uint64_t)r1 * 10 <= ((uint64_t)256 << 20))) { est += (uint64_t)r1 * 10; ++r1; }
        off.assign((size_t)(r1 - r0) + 1, 0);
        uint64_t bytes = 0;
        if (kdbx_distance_dense_rows(ctx, mt->second, counts.data(), r0, r1, nullptr, 0, off.data(), &bytes) != KDBX_OK) fail(kdbx_last_error(ctx));
        text.resize(bytes + 1);
        if (kdbx_distance_dense_rows(ctx, mt->second, counts.data(), r0, r1, text.data(), bytes, off.data(), &bytes) != KDBX_OK) fail(kdbx_last_error(ctx));
        for (uint32_t s = r0; s < r1; ++s) {
            out << names[s] << ',';
            out.write(text.data() + off[s - r0], (std::streamsize)(off[s - r0 + 1] - off[s - r0]));
            out << std::endl;
        }
        r0 = r1;
    }
    kdbx_close(ctx);
}

void run_distance(const Params& params) {
    if (params.files.size() < 2) throw usage_error(params.mode);
    if (params.device_distance) { run_distance_device(params); return; }
    std::ifstream in(params.files[0]);
    if (!in) throw std::runtime_error("Cannot open common k-mers table: " + params.files[0]);
    std::ofstream out(params.files[1]);
    const metric_fn fn = find_metric(params.metric_name);
    if (!fn) throw std::runtime_error("Unknown distance/similarity metric: " + params.metric_name);

    std::string tok, rest;
    uint32_t k = 0;
    double fraction = 0;
    in >> tok >> k >> tok >> fraction >> tok;  // "kmer-length:" k "fraction:" f ",db-samples"
    std::getline(in, rest);                   // " ,name,name,...,"
    if (!params.phylip_out) out << "kmer-length: " << k << " fraction: " << fraction << rest << std::endl;
    std::vector<std::string> db_names;
    {
        std::string names = rest;
        std::replace(names.begin(), names.end(), ',', ' ');
        std::istringstream iss(names);
        while (iss >> tok) db_names.push_back(tok);
    }
    std::vector<uint32_t> db_counts;
    {
        std::getline(in, rest);
        std::replace(rest.begin(), rest.end(), ',', ' ');
        std::istringstream iss(rest);
        iss >> tok >> tok;
        size_t v;
        while (iss >> v) db_counts.push_back((uint32_t)v);
    }
    if (params.phylip_out) out << db_counts.size() << std::endl;

    const size_t N = db_counts.size();
    std::vector<uint32_t> dense(N, 0);
    std::vector<std::pair<size_t, uint32_t>> sparse;
    std::vector<char> buf;
    bool triangle = false;
    bool sparse_out = params.sparse_out && !params.phylip_out;
    std::string line;
    for (int row_id = 0; std::getline(in, line); ++row_id) {
        const char* begin = line.data();
        const char* end = begin + line.size();
        const char* p = std::find(begin, end, ',');
        const std::string name(begin, p);
        begin = p < end ? p + 1 : end;
        const uint32_t query_count = (uint32_t)parse_long(begin, &p);
        begin = p < end ? p + 1 : end;

        int num_read = 0;
        for (; end - begin > 1; ++num_read) {
            const long v = parse_long(begin, &p);
            if (*p == ':') {
                const uint32_t common = (uint32_t)parse_long(p + 1, &p);
                if (params.phylip_out) { if ((size_t)(v - 1) < N) dense[(size_t)(v - 1)] = common; }
                else {
                    sparse_out = true;  // sparse input always gives sparse output
                    if (common > 0 && (size_t)(v - 1) < N && params.filters.pass(common, query_count, db_counts[(size_t)(v - 1)], (int)k))
                        sparse.emplace_back((size_t)(v - 1), common);
                }
            } else {
                const uint32_t common = (uint32_t)v;
                if (sparse_out) {
                    if (common > 0 && (size_t)num_read < N && params.filters.pass(common, query_count, db_counts[(size_t)num_read], (int)k))
                        sparse.emplace_back((size_t)num_read, common);
                } else if ((size_t)num_read < N) dense[(size_t)num_read] = common;
            }
            begin = p < end ? p + 1 : end;
        }
        const bool empty_diagonal = sparse_out ? sparse.empty() : (N == 0 || dense[0] == 0);
        if (row_id == 0 && !db_names.empty() && name == db_names[0] && empty_diagonal) triangle = true;
        const size_t to_process = !sparse_out ? (triangle ? (size_t)row_id : N) : sparse.size();

        buf.resize(name.size() + 64 + (std::max(to_process, (size_t)num_read) + 1) * 48);
        char* w = buf.data();
        std::memcpy(w, name.data(), name.size());
        w += name.size();
        if (params.phylip_out) {
            *w++ = ' ';
            for (int c = 0; c < num_read && (size_t)c < N; ++c) {
                w = put_f6(w, fn(dense[(size_t)c], query_count, db_counts[(size_t)c], (int)k));
                *w++ = ' ';
            }
        } else {
            *w++ = ',';
            if (sparse_out) {
                for (const auto& e : sparse) {
                    w = put_u64(w, e.first + 1);
                    *w++ = ':';
                    w = put_f6(w, fn(e.second, query_count, db_counts[e.first], (int)k));
                    *w++ = ',';
                }
            } else {
                for (size_t c = 0; c < to_process && c < N; ++c) {
                    w = put_f6(w, fn(dense[c], query_count, db_counts[c], (int)k));
                    *w++ = ',';
                }
            }
        }
        out.write(buf.data(), w - buf.data());
        out << std::endl;
        if (params.phylip_out || !sparse_out) std::fill(dense.begin(), dense.end(), 0u);
        else sparse.clear();
    }
}

}  // namespace kdbx
