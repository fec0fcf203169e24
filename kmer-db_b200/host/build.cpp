// See kmer_db.h.
#include "kmer_db.h"

#include <algorithm>
#include <cstring>
#include <functional>
#include <chrono>
#include <cstdio>
#include <thread>

#include "gamma.h"
#include "kmers.h"

namespace kdbx {

DbBuilder::DbBuilder(int threads) : threads_(std::max(1, threads)) {
    hdr_.is_initialized = 0;
    hdr_.kmers_count = 0;
    pats_.emplace_back();  // pattern 0: the empty sentinel (src/prefix_kmer_db.cpp:24)
}

DbBuilder::~DbBuilder() {
    for (Pattern& p : pats_) std::free(p.data);
}

void DbBuilder::adopt(Trie&& db) {
    if (db.tables.empty()) throw std::runtime_error("build -extend needs a database with k-mer tables");
    db.drop_compact();   // the patterns change from here on
    for (Pattern& p : pats_) std::free(p.data);
    hdr_ = db.hdr;
    names_ = std::move(db.sample_names);
    sample_kmers_ = std::move(db.sample_kmers);
    tables_ = std::move(db.tables);
    const uint64_t P = db.num_patterns();
    pats_.assign(P, Pattern());
    for (uint64_t i = 0; i < P; ++i) {
        Pattern& p = pats_[i];
        p.num_kmers = db.num_kmers[i]; p.parent = db.parent_id[i];
        p.n = db.n[i]; p.l = db.l[i]; p.last = db.last[i]; p.bits = db.bits[i];
        const uint64_t w = Trie::payload_words_for_bits(p.bits);
        if (w) {
            p.cap_words = (uint32_t)w;
            p.data = static_cast<uint64_t*>(std::malloc(w * 8));
            if (!p.data) throw std::bad_alloc();
            std::memcpy(p.data, db.payload.data() + db.payload_off[i], w * 8);
            // gamma_put ORs new codes in: everything past num_bits must be zero
            const uint32_t used = p.bits >> 6, rem = p.bits & 63;
            if (used < w) {
                if (rem) p.data[used] &= ~0ull << (64 - rem); else p.data[used] = 0;
                for (uint64_t x = used + 1; x < w; ++x) p.data[x] = 0;
            }
        }
        if (p.parent >= 0) pats_[(uint64_t)p.parent].is_parent = true;
    }
}

// pattern_t::expand (src/pattern.h:195-203): the new sample goes in as the Elias-gamma code of
// its distance to the previous last sample.
void DbBuilder::append_sample(Pattern& p, uint32_t sample) {
    const uint32_t delta = sample - p.last;
    const uint32_t need_bits = p.bits + gamma_code_len(delta);
    const uint32_t need_words = (need_bits + 127) / 128 * 2;
    if (need_words > p.cap_words) {
        const uint32_t cap = std::max(need_words, p.cap_words * 2);
        uint64_t* d = static_cast<uint64_t*>(std::realloc(p.data, (size_t)cap * 8));
        if (!d) throw std::bad_alloc();
        std::memset(d + p.cap_words, 0, (size_t)(cap - p.cap_words) * 8);
        p.data = d;
        p.cap_words = cap;
    }
    gamma_put(p.data, p.bits, delta);
    p.last = sample;
    ++p.n;
    ++p.l;
}

uint32_t DbBuilder::add_sample(const std::string& name, const uint64_t* kmers, size_t count, uint32_t k, double fraction,
                               int32_t alphabet_id, int bits_per_symbol) {
    if (!hdr_.is_initialized || tables_.empty()) {
        hdr_.kmer_length = k; hdr_.fraction = fraction; hdr_.start_fraction = 0.0;  // (src/kmer_db.h:42-47)
        hdr_.alphabet_type = alphabet_id; hdr_.is_initialized = 1; hdr_.format_word = 1;
        hdr_.num_hashtables = num_prefix_tables(k, bits_per_symbol);
        tables_.assign(hdr_.num_hashtables, HashTable());
    }
    if (hdr_.kmer_length != k) throw std::runtime_error("Error in AbstractKmerDb::addKmers(): adding kmers of different length");
    if (hdr_.fraction != fraction) throw std::runtime_error("Error in AbstractKmerDb::addKmers(): adding kmers of different minhash fraction");
    if (hdr_.alphabet_type != alphabet_id) throw std::runtime_error("Error in AbstractKmerDb::addKmers(): adding samples from different alphabet");
    const uint32_t sample = (uint32_t)names_.size();
    names_.push_back(name);
    sample_kmers_.push_back((uint32_t)count);  // the reference keeps uint32 counts (src/kmer_db.h:38)
    if (count == 0) return sample;

    static const bool trace = std::getenv("KDBX_TRACE") != nullptr;
    static double t_tab = 0, t_sort = 0, t_pat = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    // ---- k-mer tables: find-or-insert, parallel over prefix-aligned blocks -----------------------
    sample_patterns_.resize(count);
    const int T = (int)std::min<size_t>((size_t)threads_, std::max<size_t>(1, count / 4096));
    std::vector<size_t> cut(T + 1, count);
    cut[0] = 0;
    for (int t = 1; t < T; ++t) {
        size_t c = std::max(cut[t - 1], count * (size_t)t / (size_t)T);
        while (c > cut[t - 1] && c < count && (kmers[c] >> 32) == (kmers[c - 1] >> 32)) ++c;  // keep a prefix in one block
        cut[t] = std::min(c, count);
    }
    std::vector<uint64_t> added(T, 0);
    auto work = [&](int t) {
        size_t i = cut[t];
        const size_t end = cut[t + 1];
        while (i < end) {
            const uint64_t prefix = kmers[i] >> 32;
            size_t j = i + 1;
            while (j < end && (kmers[j] >> 32) == prefix) ++j;
            if (prefix >= tables_.size()) throw std::runtime_error("k-mer prefix outside the database's table range");
            HashTable& ht = tables_[prefix];
            const uint64_t before = ht.filled;
            ht.reserve_additional(j - i);
            for (size_t x = i; x < j; ++x) {
                uint64_t* slot = ht.find_or_insert((uint32_t)kmers[x]);
                sample_patterns_[x] = {(int32_t)(*slot >> 32), slot};
            }
            added[t] += ht.filled - before;
            i = j;
        }
    };
    if (T == 1) work(0);
    else {
        std::vector<std::thread> th;
        std::exception_ptr err;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] { try { work(t); } catch (...) { err = std::current_exception(); } });
        for (auto& x : th) x.join();
        if (err) std::rethrow_exception(err);
    }
    for (uint64_t a : added) hdr_.kmers_count += a;

    const double t1 = now();
    // ---- group by pattern, then extend or split (src/prefix_kmer_db.cpp:181-240) ----------------------
    auto run_parallel = [&](int n, const std::function<void(int)>& fn) {
        if (n <= 1) { fn(0); return; }
        std::vector<std::thread> th;
        std::exception_ptr err;
        for (int t = 0; t < n; ++t) th.emplace_back([&, t] { try { fn(t); } catch (...) { err = std::current_exception(); } });
        for (auto& x : th) x.join();
        if (err) std::rethrow_exception(err);
    };
    const int TS = (int)std::min<size_t>((size_t)threads_, std::max<size_t>(1, count / 65536));
    auto slice = [&](int t, int n) { return std::make_pair(count * (size_t)t / (size_t)n, count * (size_t)(t + 1) / (size_t)n); };
    {   // LSD radix sort of (pattern id << 32 | k-mer index): 11-bit passes over the id bits in use,
        // every thread histograms and scatters its own slice
        std::vector<uint64_t>& a = sort_a_;
        std::vector<uint64_t>& b = sort_b_;
        a.resize(count); b.resize(count);
        std::vector<uint32_t> tmax((size_t)TS, 0);
        run_parallel(TS, [&](int t) {
            auto [lo, hi] = slice(t, TS);
            uint32_t m = 0;
            for (size_t i = lo; i < hi; ++i) {
                const uint32_t pid = (uint32_t)sample_patterns_[i].first;
                a[i] = ((uint64_t)pid << 32) | (uint32_t)i;
                m = std::max(m, pid);
            }
            tmax[(size_t)t] = m;
        });
        uint32_t max_pid = 0;
        for (uint32_t m : tmax) max_pid = std::max(max_pid, m);
        std::vector<std::vector<size_t>> hist((size_t)TS, std::vector<size_t>(2048));
        for (int shift = 32; shift < 64 && (shift == 32 || (max_pid >> (shift - 32))); shift += 11) {
            run_parallel(TS, [&](int t) {
                auto [lo, hi] = slice(t, TS);
                std::vector<size_t>& h = hist[(size_t)t];
                std::fill(h.begin(), h.end(), 0);
                for (size_t i = lo; i < hi; ++i) ++h[(a[i] >> shift) & 2047];
            });
            size_t run = 0;  // digit-major, thread-minor exclusive offsets keep the pass stable
            for (int d = 0; d < 2048; ++d)
                for (int t = 0; t < TS; ++t) { const size_t c = hist[(size_t)t][(size_t)d]; hist[(size_t)t][(size_t)d] = run; run += c; }
            run_parallel(TS, [&](int t) {
                auto [lo, hi] = slice(t, TS);
                std::vector<size_t>& h = hist[(size_t)t];
                for (size_t i = lo; i < hi; ++i) b[h[(a[i] >> shift) & 2047]++] = a[i];
            });
            a.swap(b);
        }
        sorted_.resize(count);
        run_parallel(TS, [&](int t) {
            auto [lo, hi] = slice(t, TS);
            for (size_t i = lo; i < hi; ++i) sorted_[i] = sample_patterns_[(uint32_t)a[i]];
        });
        sorted_.swap(sample_patterns_);
    }
    const double t2 = now();
    {   // groups of equal pattern id are independent: threads take group-aligned slices, first decide
        // extend-or-split and count the new patterns, then (after one resize) create them
        std::vector<size_t> gcut((size_t)TS + 1, count);
        gcut[0] = 0;
        for (int t = 1; t < TS; ++t) {
            size_t c = std::max(gcut[(size_t)t - 1], slice(t, TS).first);
            while (c < count && c > 0 && sample_patterns_[c].first == sample_patterns_[c - 1].first) ++c;
            gcut[(size_t)t] = c;
        }
        std::vector<uint64_t> new_count((size_t)TS, 0);
        auto for_groups = [&](int t, auto&& fn) {
            for (size_t i = gcut[(size_t)t]; i < gcut[(size_t)t + 1];) {
                const int32_t pid = sample_patterns_[i].first;
                size_t j = i + 1;
                while (j < count && sample_patterns_[j].first == pid) ++j;
                fn(pid, i, j);
                i = j;
            }
        };
        run_parallel(TS, [&](int t) {
            uint64_t n = 0;
            for_groups(t, [&](int32_t pid, size_t i, size_t j) {
                const Pattern& q = pats_[(size_t)pid];
                if (!(q.num_kmers == (int64_t)(j - i) && !q.is_parent)) ++n;
            });
            new_count[(size_t)t] = n;
        });
        uint64_t next = pats_.size();
        std::vector<uint64_t> first_new((size_t)TS, 0);
        for (int t = 0; t < TS; ++t) { first_new[(size_t)t] = next; next += new_count[(size_t)t]; }
        if (next >= 0x7FFFFFFFull) throw std::runtime_error("too many patterns");
        pats_.resize(next);
        run_parallel(TS, [&](int t) {
            uint64_t new_pid = first_new[(size_t)t];
            for_groups(t, [&](int32_t pid, size_t i, size_t j) {
                const int64_t c = (int64_t)(j - i);
                Pattern& q = pats_[(size_t)pid];
                if (q.num_kmers == c && !q.is_parent) { append_sample(q, sample); return; }
                Pattern& child = pats_[new_pid];
                child.num_kmers = c;
                child.n = q.n + 1; child.l = 1; child.last = sample;
                if (q.n > 0) { q.is_parent = true; child.parent = pid; }  // children of pattern 0 are roots
                if (pid) q.num_kmers -= c;
                for (size_t x = i; x < j; ++x) {
                    uint64_t* slot = sample_patterns_[x].second;
                    *slot = (*slot & 0xFFFFFFFFull) | (new_pid << 32);
                }
                ++new_pid;
            });
        });
    }
    if (trace) {
        t_tab += t1 - t0; t_sort += t2 - t1; t_pat += now() - t2;
        std::fprintf(stderr, "[build] sample %u: tables %.3f s, sort %.3f s, patterns %.3f s (cumulative)\n", sample, t_tab, t_sort, t_pat);
    }
    return sample;
}

void DbBuilder::finish(Trie& out) {
    out.hdr = hdr_;
    out.sample_names = std::move(names_);
    out.sample_kmers = std::move(sample_kmers_);
    out.tables = std::move(tables_);
    const uint64_t P = pats_.size();
    out.num_kmers.resize(P); out.parent_id.resize(P); out.n.resize(P); out.l.resize(P);
    out.last.resize(P); out.bits.resize(P); out.payload_off.resize(P);
    uint64_t words = 0;
    for (const Pattern& p : pats_) words += Trie::payload_words_for_bits(p.bits);
    out.payload.clear();
    out.payload.resize(words, 0);
    uint64_t at = 0;
    for (uint64_t i = 0; i < P; ++i) {
        Pattern& p = pats_[i];
        out.num_kmers[i] = p.num_kmers; out.parent_id[i] = p.parent; out.n[i] = p.n; out.l[i] = p.l;
        out.last[i] = p.last; out.bits[i] = p.bits; out.payload_off[i] = at;
        const uint64_t w = Trie::payload_words_for_bits(p.bits);
        if (w) { std::memcpy(out.payload.data() + at, p.data, w * 8); at += w; }
        std::free(p.data);
        p.data = nullptr;
    }
    pats_.clear();
    pats_.emplace_back();
    hdr_ = DbHeader();
    hdr_.is_initialized = 0;
    out.build_compact();
}

}  // namespace kdbx
