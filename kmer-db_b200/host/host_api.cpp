// extern "C" surface of libkdbx_host.so (include/kdbx_host.h): thin wrappers that translate
// exceptions into error codes.
#include <cstring>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/kdbx_host.h"
#include "gamma.h"
#include "ingest.h"
#include "kmer_db.h"
#include "synth.h"
#include "csv_out.h"
#include "trie.h"

struct kdbxh_trie {
    kdbx::Trie t;
    std::vector<uint64_t> flat_off;   // flattened k-mer tables (kdbxh_tables_view)
    kdbx::Buf<uint64_t> flat_slots;
    explicit kdbxh_trie(bool pinned) : t(pinned) { flat_slots.set_pinned(pinned); }
};
struct kdbxh_builder { kdbx::DbBuilder b; explicit kdbxh_builder(int threads) : b(threads) {} };
struct kdbxh_samples { std::vector<kdbx::SampleKmers> items; };

namespace {
thread_local std::string g_err;
template <class F> int guarded(F&& f) {
    try { f(); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
    catch (...) { g_err = "unknown error"; return -1; }
}
}  // namespace

namespace kdbx {
// Structural validity of a trie (see kdbx_host.h).  Throws with the first violation.
void validate_trie(const Trie& t) {
    const uint64_t P = t.num_patterns();
    const uint32_t N = t.num_samples();
    if (P == 0) throw std::runtime_error("trie has no sentinel pattern");
    std::vector<uint32_t> ids;
    for (uint64_t p = 0; p < P; ++p) {
        const int64_t par = t.parent_id[p];
        const uint32_t n = t.n[p], l = t.l[p];
        auto bad = [&](const char* what) {
            throw std::runtime_error("invalid trie at pattern " + std::to_string(p) + ": " + what);
        };
        if (par >= (int64_t)p || par < -1) bad("parent_id must be -1 or < own id");
        if (l > n) bad("num_local_samples > num_samples");
        if (n > N) bad("num_samples > sample count");
        const uint32_t n_par = par >= 0 ? t.n[par] : 0;
        if (n != n_par + l) bad("num_samples != parent's num_samples + num_local_samples");
        if (l == 0) { if (t.bits[p]) bad("payload without local samples"); continue; }
        const uint64_t words = Trie::payload_words_for_bits(t.bits[p]);
        if (t.payload_off[p] + words > t.payload.size()) bad("payload out of bounds");
        if (t.last[p] >= N) bad("last_sample_id >= sample count");
        ids.resize(l);
        // decode with explicit bounds so that a corrupt stream cannot run away
        uint32_t pos = 0; uint64_t sum = 0;
        for (uint32_t i = 0; i + 1 < l; ++i) {
            if (pos >= t.bits[p]) bad("Elias-gamma stream shorter than num_local_samples-1 codes");
            const uint32_t d = gamma_get(t.payload.data() + t.payload_off[p], pos);
            sum += d;
        }
        if (pos != t.bits[p]) bad("Elias-gamma stream length != num_bits");
        if (sum > t.last[p]) bad("deltas exceed last_sample_id");
        const uint32_t first = t.last[p] - (uint32_t)sum;
        if (par >= 0 && first <= t.last[par]) bad("local list does not continue parent's list");
    }
}

// Sub-database of the first k samples (see kdbx_host.h for the precondition).
void prefix_trie(const Trie& src, uint32_t k, Trie& dst) {
    if (k > src.num_samples()) throw std::runtime_error("prefix: more samples requested than the database has");
    const uint64_t P = src.num_patterns();
    uint64_t keep = 1;
    bool tail = false;
    for (uint64_t p = 1; p < P; ++p) {
        // first id of the full list = first local id of the root of p's chain; last id = last[p]
        const bool low = src.last[p] < k;
        if (low) {
            if (tail) throw std::runtime_error("prefix: patterns of the first samples do not form a prefix");
            keep = p + 1;
        } else {
            tail = true;
            int64_t q = (int64_t)p;
            while (src.parent_id[q] >= 0) q = src.parent_id[q];
            uint32_t pos = 0, sum = 0;
            for (uint32_t i = 0; i + 1 < src.l[q]; ++i) sum += gamma_get(src.payload.data() + src.payload_off[q], pos);
            if (src.last[q] - sum < k) throw std::runtime_error("prefix: a pattern spans the cut");
        }
    }
    dst.hdr = src.hdr;
    dst.sample_names.assign(src.sample_names.begin(), src.sample_names.begin() + k);
    dst.sample_kmers.assign(src.sample_kmers.begin(), src.sample_kmers.begin() + k);
    dst.num_kmers.resize(keep); dst.parent_id.resize(keep); dst.n.resize(keep); dst.l.resize(keep);
    dst.last.resize(keep); dst.bits.resize(keep); dst.payload_off.resize(keep);
    uint64_t words = 0;
    for (uint64_t p = 0; p < keep; ++p) {
        dst.num_kmers[p] = src.num_kmers[p]; dst.parent_id[p] = src.parent_id[p]; dst.n[p] = src.n[p];
        dst.l[p] = src.l[p]; dst.last[p] = src.last[p]; dst.bits[p] = src.bits[p];
        dst.payload_off[p] = words;
        words += Trie::payload_words_for_bits(src.bits[p]);
    }
    dst.payload.resize(words, 0);
    for (uint64_t p = 0; p < keep; ++p) {
        const uint64_t w = Trie::payload_words_for_bits(src.bits[p]);
        if (w) std::memcpy(dst.payload.data() + dst.payload_off[p], src.payload.data() + src.payload_off[p], w * 8);
    }
}
}  // namespace kdbx

extern "C" {

const char* kdbxh_last_error(void) { return g_err.c_str(); }

kdbxh_trie* kdbxh_trie_new(int pinned) {
    try { return new kdbxh_trie(pinned != 0); } catch (...) { g_err = "allocation failed"; return nullptr; }
}
void kdbxh_trie_free(kdbxh_trie* t) { delete t; }

int kdbxh_read_db(kdbxh_trie* t, const char* path) {
    if (!t || !path) { g_err = "null argument"; return -1; }
    return guarded([&] { kdbx::read_db(path, t->t); });
}
int kdbxh_read_db_full(kdbxh_trie* t, const char* path) {
    if (!t || !path) { g_err = "null argument"; return -1; }
    return guarded([&] { t->flat_off.clear(); t->flat_slots.clear(); kdbx::read_db(path, t->t, true); });
}
int kdbxh_tables_view(kdbxh_trie* t, kdbx_tables_view* out) {
    if (!t || !out) { g_err = "null argument"; return -1; }
    return guarded([&] {
        const auto& tabs = t->t.tables;
        if (tabs.empty()) throw std::runtime_error("database was loaded without its k-mer tables");
        if (t->flat_off.size() != tabs.size() + 1) {
            t->flat_off.assign(tabs.size() + 1, 0);
            for (size_t i = 0; i < tabs.size(); ++i) t->flat_off[i + 1] = t->flat_off[i] + tabs[i].slots.size();
            t->flat_slots.clear();
            t->flat_slots.resize(t->flat_off.back());
            kdbx::parallel_ranges(tabs.size(), 64, [&](size_t b, size_t e) {
                for (size_t i = b; i < e; ++i)
                    std::memcpy(t->flat_slots.data() + t->flat_off[i], tabs[i].slots.data(), tabs[i].slots.size() * 8);
            });
        }
        out->num_tables = tabs.size(); out->slot_off = t->flat_off.data(); out->slots = t->flat_slots.data();
    });
}
kdbxh_builder* kdbxh_builder_new(int threads) {
    try { return new kdbxh_builder(threads > 0 ? threads : 1); } catch (...) { g_err = "allocation failed"; return nullptr; }
}
void kdbxh_builder_free(kdbxh_builder* b) { delete b; }
int kdbxh_builder_add_sample(kdbxh_builder* b, const char* name, const uint64_t* kmers, uint64_t count, uint32_t k, double fraction) {
    if (!b || !name || (count && !kmers)) { g_err = "null argument"; return -1; }
    return guarded([&] { b->b.add_sample(name, kmers, (size_t)count, k, fraction, kdbx::kNt, 2); });
}
int kdbxh_builder_finish(kdbxh_builder* b, kdbxh_trie* out) {
    if (!b || !out) { g_err = "null argument"; return -1; }
    return guarded([&] { out->flat_off.clear(); out->flat_slots.clear(); b->b.finish(out->t); });
}
kdbxh_samples* kdbxh_samples_load(const char* list_arg, uint32_t k, double fraction, double fraction_start, int32_t alphabet_id,
                                  int multisample, int threads) {
    if (!list_arg) { g_err = "null argument"; return nullptr; }
    kdbxh_samples* out = nullptr;
    const int rc = guarded([&] {
        const kdbx::Alphabet al = kdbx::Alphabet::make(alphabet_id);
        kdbx::SampleStream stream(list_arg, al, kdbx::MinHash(fraction, fraction_start, k), k, multisample != 0, threads > 0 ? threads : 1);
        out = new kdbxh_samples();
        kdbx::SampleKmers s;
        while (stream.next(s)) { out->items.push_back(std::move(s)); s = kdbx::SampleKmers(); }
    });
    if (rc != 0) { delete out; return nullptr; }
    return out;
}
void kdbxh_samples_free(kdbxh_samples* s) { delete s; }
uint32_t kdbxh_samples_count(const kdbxh_samples* s) { return s ? (uint32_t)s->items.size() : 0; }
const char* kdbxh_samples_name(const kdbxh_samples* s, uint32_t i) { return (s && i < s->items.size()) ? s->items[i].name.c_str() : nullptr; }
const uint64_t* kdbxh_samples_kmers(const kdbxh_samples* s, uint32_t i, uint64_t* count) {
    if (!s || i >= s->items.size()) { if (count) *count = 0; return nullptr; }
    if (count) *count = s->items[i].kmers.size();
    return s->items[i].kmers.data();
}
int kdbxh_write_db(const kdbxh_trie* t, const char* path) {
    if (!t || !path) { g_err = "null argument"; return -1; }
    return guarded([&] { kdbx::write_db(path, t->t); });
}
int kdbxh_synth(kdbxh_trie* t, const kdbxh_synth_params* p) {
    if (!t || !p) { g_err = "null argument"; return -1; }
    return guarded([&] {
        kdbx::SynthParams sp;
        sp.num_samples = p->num_samples; sp.num_clusters = p->num_clusters;
        sp.genome_kmers = p->genome_kmers; sp.k = p->k; sp.mutation_rate = p->mutation_rate;
        sp.seed = p->seed; sp.interleaved = p->interleaved; sp.threads = p->threads; sp.cluster_skew = p->cluster_skew;
        kdbx::synth_generate(sp, t->t);
    });
}
int kdbxh_validate(const kdbxh_trie* t) {
    if (!t) { g_err = "null argument"; return -1; }
    return guarded([&] { kdbx::validate_trie(t->t); });
}
int kdbxh_prefix(const kdbxh_trie* src, uint32_t num_samples, kdbxh_trie* dst) {
    if (!src || !dst || src == dst) { g_err = "bad argument"; return -1; }
    return guarded([&] { kdbx::prefix_trie(src->t, num_samples, dst->t); });
}
int kdbxh_partition(const kdbxh_trie* src, uint32_t num_parts, uint32_t part, kdbxh_trie* dst, uint64_t* owned_updates) {
    if (!src || !dst || src == dst) { g_err = "bad argument"; return -1; }
    return guarded([&] { dst->flat_off.clear(); dst->flat_slots.clear(); kdbx::partition_trie(src->t, num_parts, part, dst->t, owned_updates); });
}
struct kdbxh_partitioner { kdbx::TriePartitioner p; kdbxh_partitioner(const kdbx::Trie& t, uint32_t n) : p(t, n) {} };
kdbxh_partitioner* kdbxh_partitioner_new(const kdbxh_trie* src, uint32_t num_parts) {
    if (!src) { g_err = "null argument"; return nullptr; }
    kdbxh_partitioner* out = nullptr;
    guarded([&] { out = new kdbxh_partitioner(src->t, num_parts); });
    return out;
}
void kdbxh_partitioner_free(kdbxh_partitioner* p) { delete p; }
int kdbxh_partitioner_part(const kdbxh_partitioner* p, uint32_t part, kdbxh_trie* dst, uint64_t* owned_updates, uint32_t* window) {
    if (!p || !dst) { g_err = "null argument"; return -1; }
    return guarded([&] { dst->flat_off.clear(); dst->flat_slots.clear(); p->p.extract(part, dst->t, owned_updates, window); });
}
// All parts of one cut, extracted and written to <prefix><part>of<num_parts>.db by one host thread each.
int kdbxh_partition_write_all(const kdbxh_trie* src, uint32_t num_parts, const char* prefix, uint64_t* owned_updates, uint32_t* windows,
                              uint64_t* part_updates, uint64_t* part_patterns) {
    if (!src || !prefix || num_parts == 0) { g_err = "bad argument"; return -1; }
    return guarded([&] {
        const kdbx::TriePartitioner cut(src->t, num_parts);
        std::vector<std::string> errors(num_parts);
        std::vector<std::thread> th;
        for (uint32_t g = 0; g < num_parts; ++g)
            th.emplace_back([&, g] {
                try {
                    kdbx::Trie part(false);
                    uint64_t owned = 0;
                    uint32_t win[2] = {0, 0};
                    cut.extract(g, part, &owned, win);
                    kdbx::write_db(std::string(prefix) + std::to_string(g) + "of" + std::to_string(num_parts) + ".db", part);
                    const auto tt = part.totals();
                    if (owned_updates) owned_updates[g] = owned;
                    if (windows) { windows[2 * g] = win[0]; windows[2 * g + 1] = win[1]; }
                    if (part_updates) part_updates[g] = tt.U;
                    if (part_patterns) part_patterns[g] = part.num_patterns();
                } catch (const std::exception& e) { errors[g] = e.what(); }
            });
        for (auto& t : th) t.join();
        for (const auto& e : errors) if (!e.empty()) throw std::runtime_error(e);
    });
}
int kdbxh_relabel(kdbxh_trie* t, uint32_t offset, uint32_t new_total) {
    if (!t) { g_err = "null argument"; return -1; }
    return guarded([&] { t->flat_off.clear(); t->flat_slots.clear(); kdbx::relabel_samples(t->t, offset, new_total); });
}
int kdbxh_view(const kdbxh_trie* t, kdbx_trie_view* out) {
    if (!t || !out) { g_err = "null argument"; return -1; }
    *out = t->t.view();
    return 0;
}
int kdbxh_totals_of(const kdbxh_trie* t, kdbxh_totals* out) {
    if (!t || !out) { g_err = "null argument"; return -1; }
    const auto tt = t->t.totals();
    std::memset(out, 0, sizeof *out);
    out->num_patterns = t->t.num_patterns(); out->num_samples = t->t.num_samples();
    out->updates = tt.U; out->sum_n = tt.sum_n; out->sum_l = tt.sum_l;
    out->payload_bytes = tt.payload_bytes; out->kmers_count = t->t.hdr.kmers_count;
    out->kmer_length = t->t.hdr.kmer_length; out->fraction = t->t.hdr.fraction;
    return 0;
}
const char* kdbxh_sample_name(const kdbxh_trie* t, uint32_t i) {
    return (t && i < t->t.num_samples()) ? t->t.sample_names[i].c_str() : nullptr;
}
uint64_t kdbxh_sample_kmers(const kdbxh_trie* t, uint32_t i) {
    return (t && i < t->t.num_samples()) ? t->t.sample_kmers[i] : 0;
}
int kdbxh_write_all2all_csv(const kdbxh_trie* t, const uint32_t* tri, const char* path, int sparse) {
    if (!t || !path || (!tri && t->t.num_samples() > 1)) { g_err = "null argument"; return -1; }
    return guarded([&] { kdbx::write_all2all_csv(path, t->t, tri, sparse != 0); });
}

int kdbxh_write_one2all_csv(const kdbxh_trie* t, const char* sample, uint64_t kmers, const uint32_t* sims, const char* path) {
    if (!t || !sample || !path || (!sims && t->t.num_samples())) { g_err = "null argument"; return -1; }
    return guarded([&] { kdbx::write_one2all_csv(path, t->t, sample, kmers, sims); });
}

int kdbxh_write_sparse_csv(const kdbxh_trie* t, const kdbx_csr* cells, const uint32_t* row_shifts, const uint32_t* col_shifts,
                           uint32_t num_cells, const char* filters, const char* sample_rows, const char* path, uint64_t* saved) {
    if (!t || !path || (num_cells && !cells)) { g_err = "null argument"; return -1; }
    return guarded([&] {
        kdbx::OutputFilters f;
        if (filters) {
            std::istringstream words(filters);
            std::string w, v;
            while (words >> w) {
                if ((w != "-min" && w != "-max") || !(words >> v)) throw std::runtime_error("filters: expected -min <v> / -max <v>");
                f.add(w == "-min" ? 0 : 1, v, "num-kmers");
            }
        }
        const kdbx::Trie& db = t->t;
        const uint32_t N = db.num_samples();
        for (uint32_t i = 0; i < num_cells; ++i) {
            const uint32_t rs = row_shifts ? row_shifts[i] : 0, cs = col_shifts ? col_shifts[i] : 0;
            if ((uint64_t)rs + cells[i].num_rows > N || cs > N) throw std::runtime_error("cell outside the sample table");
            if (cells[i].num_rows && (!cells[i].row_ptr || (cells[i].row_ptr[cells[i].num_rows] && (!cells[i].col || !cells[i].val))))
                throw std::runtime_error("cell with rows but no arrays");
        }
        uint64_t n = 0;
        if (!sample_rows) {
            if (num_cells != 1 || (row_shifts && row_shifts[0]) || (col_shifts && col_shifts[0]))
                throw std::runtime_error("a grid of cells needs sample_rows");
            const kdbx_csr& m = cells[0];
            if (m.num_rows != N || !m.row_ptr || (m.row_ptr[N] && (!m.col || !m.val))) throw std::runtime_error("the matrix must have one row per sample");
            for (uint64_t i = 0; i < m.row_ptr[N]; ++i)
                if (m.col[i] >= N) throw std::runtime_error("column id outside the sample table");
            n = kdbx::write_sparse_csv(path, db, m, &f);
        } else {
            kdbx::metric_fn crit = nullptr;
            int count = 0;
            kdbx::parse_sample_rows(sample_rows, crit, count);
            if (count <= 0 || !crit) throw std::runtime_error("Sampling parameters error - a criterion and a positive count are needed");
            kdbx::RowSampler sampler(N, (uint32_t)count, crit);
            for (uint32_t i = 0; i < num_cells; ++i) {
                const uint32_t rs = row_shifts ? row_shifts[i] : 0, cs = col_shifts ? col_shifts[i] : 0;
                sampler.add_cell(cells[i], &f, db.sample_kmers.data() + rs, db.sample_kmers.data() + cs, rs, cs, (int)db.hdr.kmer_length);
            }
            FILE* out = std::fopen(path, "wb");
            if (!out) throw std::runtime_error(std::string("Cannot open output file ") + path);
            const std::string head = kdbx::table_header(db);
            std::fwrite(head.data(), 1, head.size(), out);
            n = sampler.write_rows(out, db.sample_names, db.sample_kmers);
            if (std::fclose(out) != 0) throw std::runtime_error(std::string("Cannot write output file ") + path);
        }
        if (saved) *saved = n;
    });
}

}  // extern "C"
