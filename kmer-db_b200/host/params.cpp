// See cli.h.
#include <algorithm>
#include <iostream>
#include <sstream>
#include <thread>

#include "cli.h"

namespace kdbx {
namespace {

bool find_switch(std::vector<std::string>& a, const std::string& name) {
    auto it = std::find(a.begin(), a.end(), name);
    if (it == a.end()) return false;
    a.erase(it);
    return true;
}

// "<name> <value>": consumed only when the value parses, and never matches the last token
template <class T>
bool find_option(std::vector<std::string>& a, const std::string& name, T& v) {
    if (a.empty()) return false;
    auto stop = std::prev(a.end());
    auto it = std::find(a.begin(), stop, name);
    if (it == stop) return false;
    std::istringstream iss(*std::next(it));
    T tmp;
    if (!(iss >> tmp)) return false;
    v = tmp;
    a.erase(it, it + 2);
    return true;
}

void parse_filters(std::vector<std::string>& a, OutputFilters& f, const std::string& deferred) {
    const char* names[2] = {"-min", "-max"};
    for (int i = 0; i < 2; ++i) {
        std::string text;
        while (find_option(a, names[i], text)) f.add(i, text, deferred);
    }
}

const char* kModes[] = {"build", "minhash", "all2all", "all2all-sp", "all2all-parts", "new2all", "one2all", "distance"};

}  // namespace

void print_usage(const std::string& mode) {
    std::cerr << "kmer-db-b200: B200-native common k-mer counting (kmer-db compatible command line)\n";
    if (mode == "build")
        std::cerr << "  build [-k <len>] [-f <fraction>] [-multisample-fasta] [-extend] [-alphabet <name>] [-preserve-strand] [-t <n>] <sample_list> <database>\n"
                     "  build -from-minhash [-extend] [-t <n>] <sample_list> <database>\n";
    else if (mode == "minhash")
        std::cerr << "  minhash [-f <fraction>] [-k <len>] [-alphabet <name>] [-preserve-strand] [-t <n>] <sample_list>     (writes <sample>.minhash)\n";
    else if (mode == "all2all")
        std::cerr << "  all2all [-sparse [-min [<crit>:]<v>]* [-max [<crit>:]<v>]*] [-gpus <n>] [-gpu <id>] [-t <n>] [-buffer <mb>] <database> <common_table>\n";
    else if (mode == "all2all-sp")
        std::cerr << "  all2all-sp [-min [<crit>:]<v>]* [-max [<crit>:]<v>]* [-sample-rows <crit>:<count>] [-gpu <id>] [-t <n>] [-buffer <mb>] [-bubble-size <n>] <database> <common_table>\n";
    else if (mode == "all2all-parts")
        std::cerr << "  all2all-parts [-min [<crit>:]<v>]* [-max [<crit>:]<v>]* [-sample-rows <crit>:<count>] [-gpus <n>] [-gpu <id>] [-t <n>] [-buffer <mb>] [-bubble-size <n>] <db_list> <common_table>\n";
    else if (mode == "new2all")
        std::cerr << "  new2all [-multisample-fasta | -from-minhash] [-sparse [-min ...]* [-max ...]*] [-gpu <id>] [-t <n>] <database> <sample_list> <common_table>\n";
    else if (mode == "one2all")
        std::cerr << "  one2all [-from-minhash] [-gpu <id>] [-t <n>] <database> <sample> <common_table>\n";
    else if (mode == "distance")
        std::cerr << "  distance <measure> [-sparse [-min [<crit>:]<v>]* [-max [<crit>:]<v>]*] [-phylip-out] <common_table> <output_table>\n"
                     "    measures: jaccard, min, max, cosine, mash, ani, ani-shorter, mash-query, num-kmers\n";
    else
        std::cerr << "  modes: build, minhash, all2all, all2all-sp, all2all-parts, new2all, one2all, distance   (kmer-db-b200 <mode> -help)\n"
                     "  extras: synth (generate a synthetic database), info <database>\n";
}

bool parse_params(int argc, char** argv, Params& p) {
    std::vector<std::string> a(argv + 1, argv + argc);
    if (find_switch(a, "-version")) { std::cout << "kmer-db-b200 0.1 (kmer-db 2.3.1 compatible)" << std::endl; return false; }
    const bool help = find_switch(a, "-help");
    if (a.empty()) { print_usage(""); return false; }
    p.mode = a.front();
    a.erase(a.begin());
    const bool known = std::find(std::begin(kModes), std::end(kModes), p.mode) != std::end(kModes);
    if (help || a.empty() || !known) { print_usage(known ? p.mode : ""); return false; }

    find_switch(a, "-v");
    find_switch(a, "-vv");
    find_option(a, "-t", p.num_threads);
    if (p.num_threads <= 0) p.num_threads = std::max((int)std::thread::hardware_concurrency(), 1);
    find_option(a, "-rt", p.num_reader_threads);
    if (p.num_reader_threads <= 0) p.num_reader_threads = std::max(1, p.num_threads / 2);
    find_option(a, "-gpu", p.gpu);
    find_option(a, "-gpus", p.num_gpus);
    p.host_csv = find_switch(a, "-host-csv");
    p.device_distance = find_switch(a, "-device");
    if (p.num_gpus < 1) p.num_gpus = 1;

    if (p.mode == "build" || p.mode == "minhash") {
        if (find_switch(a, "-from-kmers")) throw std::runtime_error("-from-kmers (KMC databases) is not supported by kmer-db-b200");
        p.from_minhash = find_switch(a, "-from-minhash");
        if (p.from_minhash && p.mode == "minhash") throw std::runtime_error("minhash -from-minhash: the samples are minhashed already");
        p.fraction_given = find_option(a, "-f", p.fraction);
        if (p.mode == "minhash" && !p.fraction_given) p.fraction = 0.01;   // src/params.cpp:130-133
        find_option(a, "-f-start", p.fraction_start);
        p.multisample_fasta = find_switch(a, "-multisample-fasta");
        std::string name;
        if (find_option(a, "-alphabet", name)) p.alphabet = Alphabet::by_name(name);
        if (find_switch(a, "-preserve-strand")) {
            if (p.alphabet.id != kNt) throw std::runtime_error("Switch -preserve-strand applies only to nt alphabet");
            p.alphabet = Alphabet::make(kNtPreserve);
        }
        find_option(a, "-k", p.kmer_length);
        if ((int)p.kmer_length > p.alphabet.max_kmer_len)
            throw std::runtime_error("K-mer length for the given alphabet cannot exceed " + std::to_string(p.alphabet.max_kmer_len));
        p.extend_db = find_switch(a, "-extend");
        p.host_build = find_switch(a, "-host-build");
        if (p.mode == "minhash" && p.multisample_fasta)
            throw std::runtime_error("minhash -multisample-fasta is not supported by kmer-db-b200 (one .minhash file per sample file)");
    } else if (p.mode == "all2all" || p.mode == "all2all-sp" || p.mode == "all2all-parts") {
        find_option(a, "-buffer", p.cache_buffer_mb);
        if (p.cache_buffer_mb <= 0) p.cache_buffer_mb = 8;
        find_option(a, "-bubble-size", p.bubble_size);
        p.sparse_out = find_switch(a, "-sparse");
        if (p.sparse_out || p.mode != "all2all") parse_filters(a, p.filters, "num-kmers");
        std::string rows;
        if (p.mode != "all2all" && find_option(a, "-sample-rows", rows)) {
            parse_sample_rows(rows, p.sampling_criterion, p.sampling_size);
            if (p.sampling_size < 0) throw std::runtime_error("Sampling parameters error - unable to parse numerical value: " + rows);
            if (p.sampling_size > 0 && !p.sampling_criterion)
                throw std::runtime_error("-sample-rows without a criterion (random selection) is not supported by kmer-db-b200: "
                                         "name one, e.g. -sample-rows jaccard:" + std::to_string(p.sampling_size));
        }
    } else if (p.mode == "new2all") {
        if (find_switch(a, "-from-kmers")) throw std::runtime_error("-from-kmers (KMC databases) is not supported by kmer-db-b200");
        p.from_minhash = find_switch(a, "-from-minhash");
        p.multisample_fasta = find_switch(a, "-multisample-fasta");
        p.sparse_out = find_switch(a, "-sparse");
        if (p.sparse_out) parse_filters(a, p.filters, "num-kmers");
    } else if (p.mode == "one2all") {
        if (find_switch(a, "-from-kmers")) throw std::runtime_error("-from-kmers (KMC databases) is not supported by kmer-db-b200");
        p.from_minhash = find_switch(a, "-from-minhash");
    } else if (p.mode == "distance") {
        p.sparse_out = find_switch(a, "-sparse");
        p.phylip_out = find_switch(a, "-phylip-out");
        if (p.phylip_out) p.sparse_out = false;
        parse_filters(a, p.filters, "?");
        if (a.empty()) throw std::runtime_error("No distance/similarity metric specified");
        p.metric_name = a.front();
        a.erase(a.begin());
        auto it = p.filters.metrics.find("?");  // bounds given without a name apply to the chosen measure
        if (it != p.filters.metrics.end()) {
            MetricBound b = it->second;
            b.fn = find_metric(p.metric_name);
            if (!b.fn) throw std::runtime_error("Filtering error - unknown metric: " + p.metric_name);
            p.filters.metrics.erase(it);
            p.filters.metrics[p.metric_name] = b;
        }
    }
    p.files = a;
    return true;
}

}  // namespace kdbx
