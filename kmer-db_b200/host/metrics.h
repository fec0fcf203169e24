// Similarity / distance measures and the -min/-max output filters.  Arithmetic follows the
// reference's lambdas (src/params.cpp:14-42) to the letter where it matters for the bytes:
// counts are num_kmers_t = uint32 (src/types.h:19), so sums, differences and the product in
// `cosine` wrap modulo 2^32 before the conversion to double.  Filters: src/sparse_filters.h:12-61,
// parsing: src/params.cpp:418-455.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace kdbx {

using metric_fn = double (*)(uint32_t common, uint32_t cnt1, uint32_t cnt2, int k);

namespace metric {
inline double mash_of(double j, int k) { return (j == 0) ? 1.0 : (-1.0 / k) * std::log((2 * j) / (j + 1)); }
inline double jaccard(uint32_t c, uint32_t a, uint32_t b, int) { return (double)c / (uint32_t)(a + b - c); }
inline double min(uint32_t c, uint32_t a, uint32_t b, int) { return (double)c / std::min(a, b); }
inline double max(uint32_t c, uint32_t a, uint32_t b, int) { return (double)c / std::max(a, b); }
inline double cosine(uint32_t c, uint32_t a, uint32_t b, int) { return (double)c / std::sqrt((double)(uint32_t)(a * b)); }
inline double mash(uint32_t c, uint32_t a, uint32_t b, int k) { return mash_of(jaccard(c, a, b, k), k); }
inline double ani(uint32_t c, uint32_t a, uint32_t b, int k) { return 1.0 - mash(c, a, b, k); }
inline double ani_shorter(uint32_t c, uint32_t a, uint32_t b, int k) { return 1.0 - mash_of((double)c / std::min(a, b), k); }
inline double mash_query(uint32_t c, uint32_t a, uint32_t, int k) { return mash_of((double)c / a, k); }
inline double num_kmers(uint32_t c, uint32_t, uint32_t, int) { return (double)c; }
}  // namespace metric

inline metric_fn find_metric(const std::string& name) {
    static const std::map<std::string, metric_fn> table = {
        {"jaccard", metric::jaccard}, {"min", metric::min}, {"max", metric::max}, {"cosine", metric::cosine},
        {"mash", metric::mash}, {"ani", metric::ani}, {"ani-shorter", metric::ani_shorter},
        {"mash-query", metric::mash_query}, {"num-kmers", metric::num_kmers}};
    auto it = table.find(name);
    return it == table.end() ? nullptr : it->second;
}

// "-sample-rows [criterion:]count" (src/params.cpp:533-556).  criterion stays null when none is named (the reference's
// random selection, which kmer-db-b200 does not offer: csv_out.h).
inline void parse_sample_rows(const std::string& text, metric_fn& criterion, int& count) {
    std::string num = text;
    criterion = nullptr;
    const size_t sep = text.rfind(':');
    if (sep != std::string::npos) {
        const std::string name = text.substr(0, sep);
        criterion = find_metric(name);
        if (!criterion) throw std::runtime_error("Sampling parameters error - unknown measure: " + name);
        num = text.substr(sep + 1);
    }
    std::istringstream iss(num);
    if (!(iss >> count)) throw std::runtime_error("Sampling parameters error - unable to parse numerical value: " + text);
}

struct MetricBound {
    double lo = std::numeric_limits<double>::lowest(), hi = std::numeric_limits<double>::max();
    metric_fn fn = nullptr;
};

struct OutputFilters {
    std::map<std::string, MetricBound> metrics;  // ordered by name like the reference's std::map
    uint32_t kmers_lo = 0, kmers_hi = std::numeric_limits<uint32_t>::max();

    bool trivial() const { return metrics.empty() && kmers_lo == 0 && kmers_hi == std::numeric_limits<uint32_t>::max(); }
    bool pass(uint32_t common, uint32_t row_cnt, uint32_t col_cnt, int k) const {
        for (const auto& m : metrics) {
            const double v = m.second.fn(common, row_cnt, col_cnt, k);
            if (!(v >= m.second.lo && v <= m.second.hi)) return false;
        }
        return common >= kmers_lo && common <= kmers_hi;
    }
    // one "-min"/"-max" value: "[metric:]number".  `deferred` = name used when no metric is given
    // ("num-kmers" for all2all/new2all, "?" for distance, src/params.cpp:433,640).
    void add(int which /*0=min,1=max*/, const std::string& text, const std::string& deferred) {
        std::string name = deferred, num = text;
        const size_t sep = text.rfind(':');
        if (sep != std::string::npos) { name = text.substr(0, sep); num = text.substr(sep + 1); }
        std::istringstream iss(num);
        double v;
        if (!(iss >> v)) throw std::runtime_error("Filtering error - unable to parse numerical value: " + text);
        if (name == "num-kmers") (which == 0 ? kmers_lo : kmers_hi) = (uint32_t)std::lrint(v);
        else if (name == "?" || find_metric(name)) {
            MetricBound& b = metrics[name];
            b.fn = find_metric(name);
            (which == 0 ? b.lo : b.hi) = v;
        } else throw std::runtime_error("Filtering error - unknown metric: " + name);
    }
};

}  // namespace kdbx
