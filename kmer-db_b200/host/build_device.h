// Host-side driver of the device builder (include/kdbx.h, kdbx_builder_*): the `build` mode's
// equivalent of PrefixKmerDb as BuildConsole::run uses it (src/console_build.cpp:33-157) —
// addKmers per sample, then serialize.  The samples' sequences go to the GPU, which extracts,
// filters, sorts and de-duplicates the k-mers and maintains the k-mer table and the pattern trie;
// finish() brings the database back in the layout host/db_io.cpp writes.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/kdbx.h"
#include "kmers.h"
#include "trie.h"

namespace kdbx {

class DeviceDbBuilder {
public:
    DeviceDbBuilder(int device, const Alphabet& alphabet, uint32_t k, double fraction, double fraction_start)
        : alphabet_(alphabet) {
        hdr_.kmer_length = k; hdr_.fraction = fraction; hdr_.start_fraction = fraction_start;
        hdr_.alphabet_type = alphabet.id; hdr_.is_initialized = 1; hdr_.format_word = 1;
        hdr_.num_hashtables = num_prefix_tables(k, alphabet.bits_per_symbol);
        kdbx_config cfg{};
        cfg.device = device;
        if (kdbx_open(&cfg, &ctx_) != KDBX_OK) throw std::runtime_error(kdbx_last_error(nullptr));
        kdbx_build_params bp{};
        bp.kmer_length = k; bp.bits_per_symbol = (uint32_t)alphabet.bits_per_symbol; bp.alphabet_size = (uint32_t)alphabet.size;
        bp.preserve_strand = alphabet.preserve_strand ? 1u : 0u;
        bp.fraction = fraction; bp.fraction_start = fraction_start;
        for (int i = 0; i < 256; ++i) bp.symbol_map[i] = alphabet.map[i];
        if (kdbx_builder_open(ctx_, &bp, &b_) != KDBX_OK) {
            const std::string e = kdbx_last_error(ctx_);
            kdbx_close(ctx_);
            throw std::runtime_error(e);
        }
    }
    ~DeviceDbBuilder() { kdbx_builder_close(b_); kdbx_close(ctx_); }
    DeviceDbBuilder(const DeviceDbBuilder&) = delete;
    DeviceDbBuilder& operator=(const DeviceDbBuilder&) = delete;

    // build -extend (src/console_build.cpp:48-57): stage the old database and continue it
    void adopt(const Trie& db) {
        if (db.tables.empty()) throw std::runtime_error("build -extend needs a database with k-mer tables");
        hdr_ = db.hdr;
        names_ = db.sample_names;
        sample_kmers_ = db.sample_kmers;
        const kdbx_trie_view v = db.view();
        check(kdbx_load_patterns(ctx_, &v));
        std::vector<uint64_t> off(db.tables.size() + 1, 0);
        for (size_t t = 0; t < db.tables.size(); ++t) off[t + 1] = off[t] + db.tables[t].slots.size();
        std::vector<uint64_t> slots(off.back());
        for (size_t t = 0; t < db.tables.size(); ++t)
            std::copy(db.tables[t].slots.begin(), db.tables[t].slots.end(), slots.begin() + (std::ptrdiff_t)off[t]);
        kdbx_tables_view tv{};
        tv.num_tables = db.tables.size(); tv.slot_off = off.data(); tv.slots = slots.data();
        check(kdbx_load_hashtables(ctx_, &tv));
        check(kdbx_builder_adopt(b_));
    }

    // symbols: the sample's records, each followed by a byte outside the alphabet (ingest.h: SampleSeq).
    // Returns the sample's number of distinct k-mers ("total-kmers").
    uint64_t add_sample(const std::string& name, const char* symbols, size_t len) {
        uint64_t unique = 0;
        check(kdbx_builder_add_sequence(b_, symbols, len, &unique));
        names_.push_back(name);
        sample_kmers_.push_back((uint32_t)unique);  // the reference keeps uint32 counts (src/kmer_db.h:38)
        return unique;
    }
    uint64_t add_sample_kmers(const std::string& name, const uint64_t* kmers, size_t count) {
        check(kdbx_builder_add_kmers(b_, kmers, count));
        names_.push_back(name);
        sample_kmers_.push_back((uint32_t)count);
        return count;
    }

    // moves the database into `out` (SoA trie + raw k-mer tables); the builder is spent afterwards
    void finish(Trie& out) {
        kdbx_build_result r{};
        check(kdbx_builder_finish(b_, &r));
        result_ = r;
        const uint64_t P = r.num_patterns, T = r.num_tables;
        out.hdr = hdr_;
        out.hdr.kmers_count = r.kmers_count;
        out.hdr.num_hashtables = T;
        out.sample_names = std::move(names_);
        out.sample_kmers = std::move(sample_kmers_);
        out.num_kmers.clear(); out.num_kmers.resize(P); out.parent_id.clear(); out.parent_id.resize(P);
        out.n.clear(); out.n.resize(P); out.l.clear(); out.l.resize(P); out.last.clear(); out.last.resize(P);
        out.bits.clear(); out.bits.resize(P); out.payload_off.clear(); out.payload_off.resize(P);
        out.payload.clear(); out.payload.resize(r.payload_words, 0);
        std::vector<uint64_t> slot_off(T + 1, 0), slots(r.total_slots), filled(T, 0);
        kdbx_build_arrays a{};
        a.num_kmers = out.num_kmers.data(); a.parent_id = out.parent_id.data(); a.num_samples_full = out.n.data();
        a.num_local_samples = out.l.data(); a.last_sample_id = out.last.data(); a.num_bits = out.bits.data();
        a.payload_off = out.payload_off.data(); a.payload = out.payload.data();
        a.slot_off = slot_off.data(); a.slots = slots.data(); a.table_filled = filled.data();
        check(kdbx_builder_export(b_, &a));
        out.tables.assign(T, HashTable());
        for (uint64_t t = 0; t < T; ++t) {
            HashTable& ht = out.tables[t];
            ht.slots.assign(slots.begin() + (std::ptrdiff_t)slot_off[t], slots.begin() + (std::ptrdiff_t)slot_off[t + 1]);
            ht.filled = filled[t];
        }
        out.build_compact();
    }
    const kdbx_build_result& result() const { return result_; }
    const DbHeader& header() const { return hdr_; }

private:
    void check(int rc) const { if (rc != KDBX_OK) throw std::runtime_error(kdbx_last_error(ctx_)); }
    Alphabet alphabet_;
    DbHeader hdr_;
    std::vector<std::string> names_;
    std::vector<uint64_t> sample_kmers_;
    kdbx_ctx* ctx_ = nullptr;
    kdbx_builder* b_ = nullptr;
    kdbx_build_result result_{};
};

}  // namespace kdbx
