// Host-side model of a kmer-db database as the all2all path needs it: database header,
// sample table and the pattern trie in structure-of-arrays form (the layout kdbx_trie_view
// borrows).  Mirrors the *fields* of the reference's pattern_t (src/pattern.h:42-55) and
// AbstractKmerDb (src/kmer_db.h:27-45); storage is ours: one contiguous payload blob instead
// of one heap block per pattern (src/pattern.cpp:84-90).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <algorithm>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/kdbx.h"

namespace kdbx {

// Minimal growable array that can live in page-locked memory (kdbx_host_alloc) so that the
// trie goes to HBM at link rate.  Falls back to malloc only when pinning is not requested.
template <class T>
class Buf {
public:
    Buf() = default;
    Buf(const Buf&) = delete;
    Buf& operator=(const Buf&) = delete;
    Buf(Buf&& o) noexcept { swap(o); }
    Buf& operator=(Buf&& o) noexcept { if (this != &o) { release(); swap(o); } return *this; }
    ~Buf() { release(); }

    void set_pinned(bool p) { if (!ptr_) pinned_ = p; }
    size_t size() const { return size_; }
    T* data() { return ptr_; }
    const T* data() const { return ptr_; }
    T& operator[](size_t i) { return ptr_[i]; }
    const T& operator[](size_t i) const { return ptr_[i]; }
    T& back() { return ptr_[size_ - 1]; }

    void reserve(size_t n) {
        if (n <= cap_) return;
        size_t ncap = cap_ ? cap_ : (n < 1024 ? 1024 : n);   // a first request is met exactly (the readers size every array once;
        while (ncap < n) ncap += ncap / 2 + 1;                //  page-locked memory is too dear to over-allocate by half)
        T* np = nullptr;
        if (pinned_) {
            void* v = nullptr;
            if (kdbx_host_alloc(&v, ncap * sizeof(T)) != KDBX_OK) throw std::bad_alloc();
            np = static_cast<T*>(v);
        } else {
            np = static_cast<T*>(std::malloc(ncap * sizeof(T)));
            if (!np) throw std::bad_alloc();
        }
        for (size_t i = 0; i < size_; ++i) np[i] = ptr_[i];
        free_ptr();
        ptr_ = np;
        cap_ = ncap;
    }
    void resize(size_t n, T fill = T()) {
        reserve(n);
        for (size_t i = size_; i < n; ++i) ptr_[i] = fill;
        size_ = n;
    }
    // n elements whose values the caller is about to write (every one of them): no fill pass over the new storage
    void resize_uninitialized(size_t n) {
        reserve(n);
        size_ = n;
    }
    void push_back(const T& v) {
        if (size_ == cap_) reserve(size_ + 1);
        ptr_[size_++] = v;
    }
    void clear() { size_ = 0; }

private:
    void free_ptr() {
        if (!ptr_) return;
        if (pinned_) kdbx_host_free(ptr_); else std::free(ptr_);
        ptr_ = nullptr;
    }
    void release() { free_ptr(); size_ = cap_ = 0; }
    void swap(Buf& o) {
        std::swap(ptr_, o.ptr_); std::swap(size_, o.size_); std::swap(cap_, o.cap_);
        std::swap(pinned_, o.pinned_);
    }
    T* ptr_ = nullptr;
    size_t size_ = 0, cap_ = 0;
    bool pinned_ = false;
};

// f(begin, end) over [0, n) in contiguous ranges on up to 16 host threads (bulk copies of table slots, headers, ...)
template <class F>
inline void parallel_ranges(size_t n, size_t min_range, F&& f) {
    const size_t hw = std::max<size_t>(1, std::thread::hardware_concurrency());
    const size_t nt = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(hw, 16), n / std::max<size_t>(1, min_range)));
    if (nt <= 1) { if (n) f((size_t)0, n); return; }
    std::vector<std::thread> th;
    for (size_t k = 0; k < nt; ++k) th.emplace_back([&, k] { f(n * k / nt, n * (k + 1) / nt); });
    for (auto& x : th) x.join();
}

// Database header fields in file order (src/prefix_kmer_db.cpp:449-455).
struct DbHeader {
    uint64_t format_word = 1;  // bit0: raw hashtables (console_build.cpp:149 always sets it)
    uint32_t kmer_length = 18;
    double fraction = 1.0;
    double start_fraction = 0.0;
    int32_t alphabet_type = 0;  // enum AlphabetType (src/alphabet.h:10-18), 0 = nt
    uint8_t is_initialized = 1;
    uint64_t kmers_count = 0;
    uint64_t num_hashtables = 256;
};

// One prefix bucket: the raw form of the reference's hash_map_lp<uint32 suffix, int32 pattern id>
// (src/hashmap_lp.h:69-99): open addressing, linear probing from fmix32(suffix) & mask, power-of-two
// capacity, empty slots marked by val == INT32_MAX.  A slot is {u32 key; i32 val}, kept here as one
// little-endian u64 (key in the low half) so that the array can go to HBM as it is.
struct HashTable {
    static constexpr uint64_t kEmptySlot = (uint64_t)0x7FFFFFFFu << 32;  // key 0, val INT32_MAX
    double max_fill = 0.8;
    uint64_t filled = 0;
    uint64_t ht_total = 0, ht_match = 0;  // statistics words of the wire format, carried through
    std::vector<uint64_t> slots = std::vector<uint64_t>(16, kEmptySlot);

    uint64_t allocated() const { return slots.size(); }
    uint64_t mask() const { return slots.size() - 1; }
    static uint32_t hash(uint32_t h) {  // murmur3 fmix32 (src/hashmap_lp.h:52-64)
        h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
        return h;
    }
    static bool is_empty(uint64_t slot) { return (uint32_t)(slot >> 32) == 0x7FFFFFFFu; }
    // pattern id of `key`, or -1
    int32_t find(uint32_t key) const {
        const uint64_t m = mask();
        for (uint64_t h = hash(key) & m;; h = (h + 1) & m) {
            const uint64_t s = slots[h];
            if (is_empty(s)) return -1;
            if ((uint32_t)s == key) return (int32_t)(s >> 32);
        }
    }
    // slot of `key`; a new key is inserted with pattern id 0 (src/prefix_kmer_db.cpp:159-162)
    uint64_t* find_or_insert(uint32_t key) {
        const uint64_t m = mask();
        for (uint64_t h = hash(key) & m;; h = (h + 1) & m) {
            uint64_t& s = slots[h];
            if (is_empty(s)) { s = (uint64_t)key; ++filled; return &s; }
            if ((uint32_t)s == key) return &s;
        }
    }
    // capacity doubles until filled + incoming <= max_fill * capacity (src/hashmap_lp.h:427-464)
    void reserve_additional(uint64_t incoming) {
        uint64_t cap = slots.size();
        while ((double)(filled + incoming) > (double)cap * max_fill) cap *= 2;
        if (cap == slots.size()) return;
        std::vector<uint64_t> old(cap, kEmptySlot);
        old.swap(slots);
        const uint64_t m = mask();
        for (uint64_t s : old) {
            if (is_empty(s)) continue;
            uint64_t h = hash((uint32_t)s) & m;
            while (!is_empty(slots[h])) h = (h + 1) & m;
            slots[h] = s;
        }
    }
};

struct Trie {
    DbHeader hdr;
    // prefix-bucketed k-mer -> pattern id tables (src/prefix_kmer_db.h:198); empty unless the
    // database was read with hashtables or built here.  all2all never needs them.
    std::vector<HashTable> tables;
    std::vector<std::string> sample_names;
    std::vector<uint64_t> sample_kmers;  // per-sample distinct k-mer count ("total-kmers")

    // SoA over patterns, index = pattern id
    Buf<int64_t> num_kmers;
    Buf<int64_t> parent_id;
    Buf<uint32_t> n;     // num_samples (node + ancestors)
    Buf<uint32_t> l;     // num_local_samples
    Buf<uint32_t> last;  // last_sample_id
    Buf<uint32_t> bits;  // num_bits
    Buf<uint64_t> payload_off;  // word offset into payload
    Buf<uint64_t> payload;      // Elias-gamma words, ceil(bits/128)*2 words per pattern
    // 32-bit mirrors of parent_id / num_kmers for the upload (kdbx_trie_view::parent_id32 / num_kmers32): 24 instead of 40
    // bytes of header per pattern over PCIe.  Built by build_compact() where a trie is finished (reader, partitioner);
    // handed out by view() only while they still match the 64-bit arrays in length.
    Buf<int32_t> parent32;
    Buf<uint32_t> num_kmers32;

    explicit Trie(bool pinned = false) {
        num_kmers.set_pinned(pinned); parent_id.set_pinned(pinned); n.set_pinned(pinned);
        l.set_pinned(pinned); last.set_pinned(pinned); bits.set_pinned(pinned);
        payload_off.set_pinned(pinned); payload.set_pinned(pinned);
        parent32.set_pinned(pinned); num_kmers32.set_pinned(pinned);
    }
    // Is the payload densely packed in pattern order (pattern p's words start where pattern p-1's end)?  Every trie this
    // code writes is; view() then hands out payload_off = NULL and the 8 bytes of offset per pattern stay off the link
    // (kdbx_trie_view: "NULL = densely packed in pattern order"; the library derives the offsets from num_bits by a scan).
    bool payload_dense = false;

    // Called where a trie is finished (reader, partitioner, generator, builders): (re)builds the 32-bit mirrors — left
    // empty when a value does not fit — and records whether the payload is densely packed.
    void build_compact() {
        const uint64_t P = num_patterns();
        parent32.clear(); num_kmers32.clear();
        payload_dense = payload_off.size() == P;
        uint64_t at = 0;
        for (uint64_t p = 0; p < P && payload_dense; ++p) {
            if (payload_off[p] != at) payload_dense = false;
            at += payload_words_for_bits(bits[p]);
        }
        if (payload_dense && at != payload.size()) payload_dense = false;
        for (uint64_t p = 0; p < P; ++p)
            if (num_kmers[p] < 0 || num_kmers[p] > 0xFFFFFFFFll || parent_id[p] < -1 || parent_id[p] > 0x7FFFFFFFll) return;
        parent32.resize(P); num_kmers32.resize(P);
        for (uint64_t p = 0; p < P; ++p) { parent32[p] = (int32_t)parent_id[p]; num_kmers32[p] = (uint32_t)num_kmers[p]; }
    }
    void drop_compact() { parent32.clear(); num_kmers32.clear(); payload_dense = false; }

    uint64_t num_patterns() const { return n.size(); }
    uint32_t num_samples() const { return (uint32_t)sample_names.size(); }

    static uint64_t payload_words_for_bits(uint32_t nbits) {
        return nbits == 0 ? 0 : (uint64_t)((nbits + 127) / 128) * 2;  // src/pattern.h:80-82
    }

    kdbx_trie_view view() const {
        kdbx_trie_view v{};
        v.num_patterns = num_patterns();
        v.num_samples = num_samples();
        v.num_kmers = num_kmers.data();
        v.parent_id = parent_id.data();
        v.num_samples_full = n.data();
        v.num_local_samples = l.data();
        v.last_sample_id = last.data();
        v.num_bits = bits.data();
        v.payload_off = payload_dense ? nullptr : payload_off.data();
        v.payload = payload.data();
        v.payload_words = payload.size();
        if (parent32.size() == num_patterns() && num_kmers32.size() == num_patterns() && num_patterns()) {
            v.parent_id32 = parent32.data(); v.num_kmers32 = num_kmers32.data();
        }
        return v;
    }

    // U = sum_p l(2n - l - 1)/2 and friends; pure host arithmetic.
    struct Totals { uint64_t U = 0, sum_n = 0, sum_l = 0, payload_bytes = 0; };
    Totals totals() const {
        Totals t;
        for (uint64_t p = 0; p < num_patterns(); ++p) {
            uint64_t nn = n[p], ll = l[p];
            t.sum_n += nn; t.sum_l += ll;
            t.U += ll * (2 * nn - ll - 1) / 2;
        }
        t.payload_bytes = payload.size() * 8;
        return t;
    }
};

// db_io.cpp — .db wire format (src/prefix_kmer_db.cpp:438-574,578-748)
// with_tables = false is the reference's DeserializationMode::SkipHashtables (all2all needs none)
void read_db(const std::string& path, Trie& out, bool with_tables = false);
// writes t.tables when present, otherwise hdr.num_hashtables EMPTY raw hashtables
void write_db(const std::string& path, const Trie& t);

// partition.cpp — sharding the trie across GPUs: part `part` of `num_parts` as a valid trie of its own
// (owned patterns = a piece of the depth-first preorder, plus their ancestors with num_kmers = 0); the
// parts' all2all matrices add up to the whole database's.
std::vector<uint32_t> trie_preorder(const Trie& t);
void partition_trie(const Trie& src, uint32_t num_parts, uint32_t part, Trie& dst, uint64_t* owned_updates = nullptr);
// the same for all parts of one cut: the preorder and the cut points are computed once
class TriePartitioner {
public:
    TriePartitioner(const Trie& src, uint32_t num_parts);
    // window (optional): [lo, hi) band of sample ids the part's lists lie in
    void extract(uint32_t part, Trie& dst, uint64_t* owned_updates = nullptr, uint32_t* window = nullptr) const;
    uint32_t num_parts() const { return num_parts_; }
private:
    const Trie& src_;
    uint32_t num_parts_;
    std::vector<uint32_t> pre_;
    std::vector<uint64_t> cut_;
};
// shifts all sample ids by `offset` inside a sample table of `new_total` entries
void relabel_samples(Trie& t, uint32_t offset, uint32_t new_total);


}  // namespace kdbx
