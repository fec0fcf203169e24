// FASTA ingest for `build` and `new2all` (host side; the GPU path starts at sorted k-mers).
// Behaviour follows the reference's loader: the sample list is a file of
// whitespace-separated names unless the argument itself ends in a FASTA extension
// (src/loader_ex.cpp:89-116); each entry is opened as given or with one of the known
// extensions appended, gzip detected by content (src/genome_input_file.h:82-95); records are
// split at '>', line ends removed, headers cut at the first space (:287-337); the sample name
// is the last path component of the list entry (src/loader_ex.cpp:165-169) or, in
// -multisample-fasta mode, each record's header (:241-282).
#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <filesystem>
#include <algorithm>
#include <fstream>

#include "ingest.h"

namespace kdbx {

std::vector<std::string> read_sample_list(const std::string& arg) {
    static const char* exts[] = {".fa", ".fna", ".fasta", ".fastq", ".gz", ".fa.gz", ".fna.gz", ".fasta.gz", ".fastq.gz"};
    for (const char* e : exts) {
        const size_t n = std::strlen(e);
        if (arg.size() >= n && arg.compare(arg.size() - n, n, e) == 0) return {arg};
    }
    std::ifstream ifs(arg);
    if (!ifs) throw std::runtime_error("Unable to open input file " + arg);
    std::vector<std::string> names;
    std::string s;
    while (ifs >> s) names.push_back(s);
    return names;
}

bool load_sequence_file(const std::string& entry, std::string& data) {
    static const char* exts[] = {"", ".fa", ".fna", ".fasta", ".gz", ".fa.gz", ".fna.gz", ".fasta.gz"};
    std::string path;
    for (const char* e : exts) {
        std::error_code ec;
        if (std::filesystem::exists(entry + e, ec)) { path = entry + e; break; }
    }
    if (path.empty()) return false;
    gzFile f = gzopen(path.c_str(), "rb");  // transparent for plain files
    if (!f) return false;
    gzbuffer(f, 1 << 20);
    data.clear();
    std::vector<char> buf(1 << 22);
    for (;;) {
        const int n = gzread(f, buf.data(), (unsigned)buf.size());
        if (n < 0) { gzclose(f); return false; }
        if (n == 0) break;
        data.append(buf.data(), (size_t)n);
    }
    gzclose(f);
    return true;
}

void split_fasta(std::string& data, std::vector<FastaRecord>& out) {
    out.clear();
    size_t pos = data.find('>');
    while (pos != std::string::npos) {
        const size_t head = pos + 1;
        size_t eol = data.find('\n', head);
        if (eol == std::string::npos) eol = data.size();
        size_t head_end = eol;
        if (head_end > head && data[head_end - 1] == '\r') --head_end;
        const size_t space = data.find(' ', head);
        if (space != std::string::npos && space < head_end) head_end = space;
        FastaRecord r;
        r.header.assign(data, head, head_end - head);
        const size_t seq_begin = eol < data.size() ? eol + 1 : data.size();
        size_t next = data.find('>', seq_begin);
        const size_t seq_end = next == std::string::npos ? data.size() : next;
        // squeeze line ends out in place
        size_t w = seq_begin;
        for (size_t i = seq_begin; i < seq_end; ++i) {
            const char c = data[i];
            if (c != '\n' && c != '\r') data[w++] = c;
        }
        r.seq = data.data() + seq_begin;
        r.len = w - seq_begin;
        out.push_back(r);
        pos = next;
    }
}

std::string sample_name_of(const std::string& entry) { return std::filesystem::path(entry).filename().string(); }

}  // namespace kdbx

namespace kdbx {

SampleStream::SampleStream(const std::string& list_arg, const Alphabet& alphabet, const MinHash& filter, uint32_t k,
                           bool multisample, int threads)
    : files_(read_sample_list(list_arg)), alphabet_(alphabet), filter_(filter), k_(k), multisample_(multisample),
      ahead_((size_t)std::max(1, threads)) {}

SampleStream::SampleStream(const std::string& list_arg, FromMinhash, int threads)
    : files_(read_sample_list(list_arg)), alphabet_(Alphabet::make(kNt)), filter_(1.0, 0.0, 18), k_(0), multisample_(false),
      from_minhash_(true), ahead_((size_t)std::max(1, threads)) {}
SampleStream::SampleStream(std::vector<std::string> entries, FromMinhash, int threads)
    : files_(std::move(entries)), alphabet_(Alphabet::make(kNt)), filter_(1.0, 0.0, 18), k_(0), multisample_(false),
      from_minhash_(true), ahead_((size_t)std::max(1, threads)) {}

namespace {
constexpr uint32_t kMinhashSignature = 0xfedcba98u;
}

void store_minhash(const std::string& entry, const uint64_t* kmers, size_t count, uint32_t k, double fraction) {
    const std::string path = entry + ".minhash";
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("Cannot write " + path);
    const uint64_t n = count;
    bool ok = std::fwrite(&kMinhashSignature, 4, 1, f) == 1 && std::fwrite(&n, 8, 1, f) == 1;
    if (ok && count) ok = std::fwrite(kmers, 8, count, f) == count;
    ok = ok && std::fwrite(&k, 4, 1, f) == 1 && std::fwrite(&fraction, 8, 1, f) == 1;
    if (std::fclose(f) != 0 || !ok) throw std::runtime_error("Cannot write " + path);
}

bool load_minhash(const std::string& entry, SampleKmers& out) {
    const std::string path = entry + ".minhash";
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    uint32_t sig = 0;
    uint64_t n = 0;
    bool ok = std::fread(&sig, 4, 1, f) == 1 && sig == kMinhashSignature && std::fread(&n, 8, 1, f) == 1;
    if (ok) {   // the count must fit the file
        const long at = std::ftell(f);
        ok = std::fseek(f, 0, SEEK_END) == 0;
        const long end = ok ? std::ftell(f) : 0;
        ok = ok && at >= 0 && end >= at && (uint64_t)(end - at) >= 12 && n == ((uint64_t)(end - at) - 12) / 8 && std::fseek(f, at, SEEK_SET) == 0;
    }
    if (ok) {
        out.kmers.resize(n);
        ok = (n == 0 || std::fread(out.kmers.data(), 8, n, f) == n) && std::fread(&out.k, 4, 1, f) == 1 && std::fread(&out.fraction, 8, 1, f) == 1;
    }
    std::fclose(f);
    if (!ok) { out.kmers.clear(); return false; }
    out.name = sample_name_of(entry);
    out.entry = entry;
    return true;
}

std::vector<SampleKmers> SampleStream::load_file(size_t idx) const {
    std::vector<SampleKmers> out;
    if (from_minhash_) {
        SampleKmers s;
        if (!load_minhash(files_[idx], s)) std::fprintf(stderr, "failed:%s\n", files_[idx].c_str());
        else out.push_back(std::move(s));
        return out;
    }
    std::string data;
    if (!load_sequence_file(files_[idx], data)) {
        std::fprintf(stderr, "failed:%s\n", files_[idx].c_str());
        return out;
    }
    std::vector<FastaRecord> recs;
    split_fasta(data, recs);
    auto finish = [](SampleKmers& s) {
        std::sort(s.kmers.begin(), s.kmers.end());
        s.kmers.erase(std::unique(s.kmers.begin(), s.kmers.end()), s.kmers.end());
    };
    if (multisample_) {
        for (const FastaRecord& r : recs) {
            SampleKmers s;
            s.name = r.header;
            extract_kmers(r.seq, r.len, k_, alphabet_, filter_, s.kmers);
            finish(s);
            out.push_back(std::move(s));
        }
    } else {
        SampleKmers s;
        s.name = sample_name_of(files_[idx]);
        s.entry = files_[idx];
        size_t total = 0;
        for (const FastaRecord& r : recs) total += r.len;
        s.kmers.reserve(total);
        for (const FastaRecord& r : recs) extract_kmers(r.seq, r.len, k_, alphabet_, filter_, s.kmers);
        finish(s);
        out.push_back(std::move(s));
    }
    return out;
}

void SampleStream::refill() {
    while (inflight_.size() < ahead_ && next_file_ < files_.size()) {
        const size_t idx = next_file_++;
        inflight_.push_back(std::async(std::launch::async, [this, idx] { return load_file(idx); }));
    }
}

bool SampleStream::next(SampleKmers& out) {
    for (;;) {
        if (!ready_.empty()) {
            out = std::move(ready_.front());
            ready_.pop_front();
            return true;
        }
        refill();
        if (inflight_.empty()) return false;
        std::vector<SampleKmers> got = inflight_.front().get();
        inflight_.pop_front();
        for (auto& s : got) ready_.push_back(std::move(s));
    }
}

}  // namespace kdbx

namespace kdbx {

SequenceStream::SequenceStream(const std::string& list_arg, bool multisample, int threads)
    : files_(read_sample_list(list_arg)), multisample_(multisample), ahead_((size_t)std::max(1, threads)) {}
SequenceStream::SequenceStream(std::vector<std::string> entries, bool multisample, int threads)
    : files_(std::move(entries)), multisample_(multisample), ahead_((size_t)std::max(1, threads)) {}

std::vector<SampleSeq> SequenceStream::load_file(size_t idx) const {
    std::vector<SampleSeq> out;
    std::string data;
    if (!load_sequence_file(files_[idx], data)) {
        std::fprintf(stderr, "failed:%s\n", files_[idx].c_str());
        return out;
    }
    std::vector<FastaRecord> recs;
    split_fasta(data, recs);
    if (multisample_) {
        for (const FastaRecord& r : recs) {
            SampleSeq s;
            s.name = r.header;
            s.symbols.assign(r.seq, r.len);
            out.push_back(std::move(s));
        }
    } else {
        SampleSeq s;
        s.name = sample_name_of(files_[idx]);
        size_t total = 0;
        for (const FastaRecord& r : recs) total += r.len + 1;
        s.symbols.reserve(total);
        for (const FastaRecord& r : recs) { s.symbols.append(r.seq, r.len); s.symbols.push_back('\0'); }
        out.push_back(std::move(s));
    }
    return out;
}

void SequenceStream::refill() {
    while (inflight_.size() < ahead_ && next_file_ < files_.size()) {
        const size_t idx = next_file_++;
        inflight_.push_back(std::async(std::launch::async, [this, idx] { return load_file(idx); }));
    }
}

bool SequenceStream::next(SampleSeq& out) {
    for (;;) {
        if (!ready_.empty()) {
            out = std::move(ready_.front());
            ready_.pop_front();
            return true;
        }
        refill();
        if (inflight_.empty()) return false;
        std::vector<SampleSeq> got = inflight_.front().get();
        inflight_.pop_front();
        for (auto& s : got) ready_.push_back(std::move(s));
    }
}

}  // namespace kdbx
