// .db wire format reader / writer (little-endian, no magic).  Layout follows SURVEY.md §A.1,
// i.e. what PrefixKmerDb::serialize emits (src/prefix_kmer_db.cpp:438-574) and
// PrefixKmerDb::deserialize(SkipHashtables) consumes (src/prefix_kmer_db.cpp:578-748), with
// packed patterns as in pattern_t::pack (src/pattern.cpp:15-47) and raw hashtables as in
// hash_map_lp::serialize (src/hashmap_lp.h:481-528).  all2all never needs the hashtables
// (src/console_all2all.cpp:26), so by default the reader seeks over them; new2all and build -extend
// read them (DeserializationMode::Everything).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>

#include "trie.h"

namespace kdbx {
namespace {

struct File {
    FILE* f = nullptr;
    std::string path;
    File(const std::string& p, const char* mode) : f(std::fopen(p.c_str(), mode)), path(p) {
        if (!f) throw std::runtime_error("Cannot open k-mer database " + p);
    }
    ~File() { if (f) std::fclose(f); }
    void read(void* dst, size_t bytes) {
        if (bytes && std::fread(dst, 1, bytes, f) != bytes)
            throw std::runtime_error("Cannot open k-mer database " + path + " (truncated)");
    }
    template <class T> T get() { T v; read(&v, sizeof(T)); return v; }
    void skip(uint64_t bytes) {
        if (fseeko(f, (off_t)bytes, SEEK_CUR) != 0)
            throw std::runtime_error("Cannot open k-mer database " + path + " (seek)");
    }
    void write(const void* src, size_t bytes) {
        if (bytes && std::fwrite(src, 1, bytes, f) != bytes)
            throw std::runtime_error("Cannot write k-mer database " + path);
    }
    template <class T> void put(const T& v) { write(&v, sizeof(T)); }
};

constexpr size_t kPatternHeaderBytes = 40;          // src/pattern.cpp:15-37
constexpr size_t kIoBlockBytes = (size_t)64 << 20;  // reader's buffer (src/prefix_kmer_db.h:179)
constexpr size_t kRefPatternStructBytes = 48;       // sizeof(pattern_t), used by the block cut

// ---- pattern blocks, read in parallel ---------------------------------------------------------------------------------
// The pattern section is a chain of blocks {u64 bytes; packed patterns} of at most 64 MB (src/prefix_kmer_db.cpp:540-574).
// A block can only be walked front to back (a pattern's length follows from its num_bits), but blocks are independent once
// the number of patterns and payload words before each of them is known: the file is mapped, every block is walked once to
// count (in parallel), a prefix sum places the blocks, and a second parallel walk fills the structure-of-arrays, the
// payload blob and the 32-bit mirrors.  The reference reads and unpacks the blocks one after the other on one thread
// (src/prefix_kmer_db.cpp:703-745: 4.2 s for the 2.1 GB database of BASELINE.json configs[1] on the build container's 8
// cores, 2.2 s for the sequential reader below, 0.5 s for this one).
struct BlockRef {
    const char* data = nullptr;
    uint64_t bytes = 0, patterns = 0, words = 0;
    bool bad = false;        // a header or a payload runs over the end of the block
    bool fits32 = true;      // num_kmers / parent_id of all its patterns fit the 32-bit mirrors
};

void count_block(BlockRef& b) {
    const char* p = b.data;
    const char* end = p + b.bytes;
    while (p < end) {
        if ((size_t)(end - p) < kPatternHeaderBytes) { b.bad = true; return; }
        uint32_t nb;
        std::memcpy(&nb, p + 28, 4);
        const uint64_t words = Trie::payload_words_for_bits(nb);
        p += kPatternHeaderBytes;
        if ((uint64_t)(end - p) < words * 8) { b.bad = true; return; }
        p += words * 8;
        ++b.patterns;
        b.words += words;
    }
}

// patterns [pid, pid + b.patterns) and payload words [at, at + b.words)
void fill_block(BlockRef& b, uint64_t pid, uint64_t at, Trie& t, int32_t* parent32, uint32_t* num_kmers32) {
    const char* p = b.data;
    for (uint64_t i = 0; i < b.patterns; ++i, ++pid) {
        int64_t nk, par; uint32_t ns, nl, ls, nb;
        std::memcpy(&nk, p, 8); std::memcpy(&par, p + 8, 8);
        std::memcpy(&ns, p + 16, 4); std::memcpy(&nl, p + 20, 4);
        std::memcpy(&ls, p + 24, 4); std::memcpy(&nb, p + 28, 4);
        p += kPatternHeaderBytes;  // bytes 32..39: is_parent (4 valid + 4 undefined bytes)
        const uint64_t words = Trie::payload_words_for_bits(nb);
        t.num_kmers[pid] = nk; t.parent_id[pid] = par; t.n[pid] = ns; t.l[pid] = nl;
        t.last[pid] = ls; t.bits[pid] = nb;
        t.payload_off[pid] = at;
        if (nk < 0 || nk > 0xFFFFFFFFll || par < -1 || par > 0x7FFFFFFFll) b.fits32 = false;
        parent32[pid] = (int32_t)par; num_kmers32[pid] = (uint32_t)nk;
        if (words) {
            std::memcpy(t.payload.data() + at, p, words * 8);
            p += words * 8;
            at += words;
        }
    }
}

// a read-only mapping of the file from byte `from` on (data() = the byte at `from`)
class MappedTail {
public:
    MappedTail(const std::string& path, uint64_t from) {
        const int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return;
        struct stat st;
        if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && (uint64_t)st.st_size > from) {
            const uint64_t page = (uint64_t)sysconf(_SC_PAGESIZE);
            const uint64_t map_off = from / page * page;
            len_ = (size_t)((uint64_t)st.st_size - map_off);
            void* m = mmap(nullptr, len_, PROT_READ, MAP_PRIVATE, fd, (off_t)map_off);
            if (m != MAP_FAILED) {
                map_ = m;
                madvise(m, len_, MADV_WILLNEED);
                data_ = static_cast<const char*>(m) + (from - map_off);
                avail_ = (uint64_t)st.st_size - from;
            }
        }
        ::close(fd);
    }
    ~MappedTail() { if (map_) munmap(map_, len_); }
    MappedTail(const MappedTail&) = delete;
    MappedTail& operator=(const MappedTail&) = delete;
    bool ok() const { return map_ != nullptr; }
    const char* data() const { return data_; }
    uint64_t size() const { return avail_; }
private:
    void* map_ = nullptr;
    size_t len_ = 0;
    const char* data_ = nullptr;
    uint64_t avail_ = 0;
};

template <class F>
void for_each_block(size_t count, F&& f) {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned nt = (unsigned)std::min<size_t>(std::min(hw, 32u), count);
    if (nt <= 1) { for (size_t i = 0; i < count; ++i) f(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (unsigned k = 0; k < nt; ++k)
        th.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < count;) f(i); });
    for (auto& x : th) x.join();
}

// The raw k-mer tables (hash_map_lp::serialize, src/hashmap_lp.h:481-528: 64 bytes of fields, the occupancy bit vector,
// the filled items in slot order), expanded to slot arrays in parallel: the headers are walked once for the tables'
// places in the file, then the tables are dealt to the threads.  Returns false when the file cannot be mapped;
// `end` = file offset right after the last table.
bool read_tables_mapped(const std::string& path, uint64_t here, uint64_t num_tables, std::vector<HashTable>& tables, uint64_t& end) {
    MappedTail map(path, here);
    if (!map.ok()) return false;
    const char* base = map.data();
    const uint64_t avail = map.size();
    const auto corrupt = [&] { return std::runtime_error("Corrupt k-mer database " + path); };
    const auto truncated = [&] { return std::runtime_error("Cannot open k-mer database " + path + " (truncated)"); };
    std::vector<uint64_t> at(num_tables);
    uint64_t o = 0;
    for (uint64_t i = 0; i < num_tables; ++i) {
        if (avail - o < 64) throw truncated();
        uint64_t filled, allocated;
        std::memcpy(&filled, base + o + 8, 8);
        std::memcpy(&allocated, base + o + 16, 8);
        if (allocated == 0 || (allocated & (allocated - 1)) || filled > allocated) throw corrupt();
        const uint64_t body = (allocated + 63) / 64 * 8 + filled * 8;
        if (avail - o - 64 < body) throw truncated();
        at[i] = o;
        o += 64 + body;
    }
    end = here + o;
    tables.clear();
    tables.resize(num_tables);
    std::atomic<bool> bad{false};
    const size_t group = 64;   // tables per grab (k = 25 means 2^18 tables of 16 slots)
    for_each_block((size_t)((num_tables + group - 1) / group), [&](size_t g) {
        for (uint64_t i = g * group; i < std::min<uint64_t>(num_tables, (g + 1) * group); ++i) {
            const char* p = base + at[i];
            HashTable& ht = tables[i];
            uint64_t allocated;
            std::memcpy(&ht.max_fill, p, 8); std::memcpy(&ht.filled, p + 8, 8); std::memcpy(&allocated, p + 16, 8);
            std::memcpy(&ht.ht_total, p + 48, 8); std::memcpy(&ht.ht_match, p + 56, 8);
            const uint64_t bv_words = (allocated + 63) / 64;
            const char* bv = p + 64;
            const char* items = bv + bv_words * 8;
            ht.slots.assign(allocated, HashTable::kEmptySlot);
            uint64_t next = 0;
            for (uint64_t w = 0; w < bv_words; ++w) {
                uint64_t bits;
                std::memcpy(&bits, bv + w * 8, 8);
                while (bits) {
                    const uint64_t slot = w * 64 + (uint64_t)__builtin_ctzll(bits);
                    bits &= bits - 1;
                    if (slot >= allocated || next >= ht.filled) { bad = true; return; }
                    std::memcpy(&ht.slots[slot], items + next * 8, 8);   // {u32 key; i32 val}
                    ++next;
                }
            }
            if (next != ht.filled) { bad = true; return; }
        }
    });
    if (bad) throw corrupt();
    return true;
}

// Reads the P patterns that start at file offset `here` (right after the pattern count) through a mapping of the file.
// Returns false when the file cannot be mapped (the caller then streams it); throws on a corrupt pattern section.
bool read_patterns_mapped(const std::string& path, uint64_t here, uint64_t P, Trie& t) {
    MappedTail map(path, here);
    if (!map.ok()) return false;
    const char* base = map.data();
    const uint64_t avail = map.size();
    const auto corrupt = [&] { return std::runtime_error("Corrupt k-mer database " + path); };
    const auto truncated = [&] { return std::runtime_error("Cannot open k-mer database " + path + " (truncated)"); };

    // the chain of blocks (whatever follows the P-th pattern's block is not looked at, as in the streaming reader)
    std::vector<BlockRef> blocks;
    uint64_t o = 0;
    bool chain_truncated = false, chain_corrupt = false;
    while (o < avail) {
        if (avail - o < 8) { chain_truncated = true; break; }
        uint64_t bytes;
        std::memcpy(&bytes, base + o, 8);
        if (bytes > kIoBlockBytes) { chain_corrupt = true; break; }
        if (avail - o - 8 < bytes) { chain_truncated = true; break; }
        BlockRef b;
        b.data = base + o + 8; b.bytes = bytes;
        blocks.push_back(b);
        o += 8 + bytes;
    }
    for_each_block(blocks.size(), [&](size_t i) { count_block(blocks[i]); });
    // place the blocks; the P-th pattern must be the last one of its block
    std::vector<uint64_t> pid0(blocks.size() + 1, 0), at0(blocks.size() + 1, 0);
    size_t used = 0;
    while (pid0[used] < P) {
        if (used == blocks.size()) { if (chain_corrupt) throw corrupt(); throw truncated(); }
        const BlockRef& b = blocks[used];
        if (b.bad || pid0[used] + b.patterns > P) throw corrupt();
        pid0[used + 1] = pid0[used] + b.patterns;
        at0[used + 1] = at0[used] + b.words;
        ++used;
    }
    (void)chain_truncated;
    t.num_kmers.resize_uninitialized(P); t.parent_id.resize_uninitialized(P); t.n.resize_uninitialized(P);
    t.l.resize_uninitialized(P); t.last.resize_uninitialized(P); t.bits.resize_uninitialized(P);
    t.payload_off.resize_uninitialized(P);
    t.payload.clear();
    t.payload.resize_uninitialized(at0[used]);
    t.parent32.clear(); t.num_kmers32.clear();
    t.parent32.resize_uninitialized(P); t.num_kmers32.resize_uninitialized(P);
    for_each_block(used, [&](size_t i) { fill_block(blocks[i], pid0[i], at0[i], t, t.parent32.data(), t.num_kmers32.data()); });
    // what Trie::build_compact() would find: the payload is dense by construction; the mirrors stay iff every value fits
    t.payload_dense = true;
    for (size_t i = 0; i < used; ++i)
        if (!blocks[i].fits32) { t.parent32.clear(); t.num_kmers32.clear(); break; }
    return true;
}

}  // namespace

void read_db(const std::string& path, Trie& t, bool with_tables) {
    File in(path, "rb");
    DbHeader& h = t.hdr;
    h.format_word = in.get<uint64_t>();
    h.kmer_length = in.get<uint32_t>();
    h.fraction = in.get<double>();
    h.start_fraction = in.get<double>();
    h.alphabet_type = in.get<int32_t>();
    h.is_initialized = in.get<uint8_t>();
    h.kmers_count = in.get<uint64_t>();

    const uint64_t n_samples = in.get<uint64_t>();
    t.sample_names.resize(n_samples);
    t.sample_kmers.resize(n_samples);
    for (uint64_t i = 0; i < n_samples; ++i) {
        t.sample_kmers[i] = in.get<uint64_t>();
        const uint64_t len = in.get<uint64_t>();
        t.sample_names[i].resize(len);
        in.read(t.sample_names[i].data(), len);
    }

    h.num_hashtables = in.get<uint64_t>();
    const bool raw = (h.format_word & 1) != 0;
    t.tables.clear();
    if (with_tables) {
        if (!raw) throw std::runtime_error("Cannot open k-mer database " + path + " (non-raw hashtables are not supported)");
        t.tables.resize(h.num_hashtables);
    }
    const char* force_reader = std::getenv("KDBX_DB_READER");   // "stream": the sequential readers (tests compare the two)
    const bool stream_only = force_reader && std::strcmp(force_reader, "stream") == 0;
    uint64_t tables_done = 0;
    if (with_tables && !stream_only && h.num_hashtables) {
        const off_t here = ftello(in.f);
        uint64_t end = 0;
        if (here >= 0 && read_tables_mapped(path, (uint64_t)here, h.num_hashtables, t.tables, end)) {
            if (fseeko(in.f, (off_t)end, SEEK_SET) != 0) throw std::runtime_error("Cannot open k-mer database " + path + " (seek)");
            tables_done = h.num_hashtables;
        }
    }
    std::vector<uint64_t> bv, items;
    for (uint64_t i = tables_done; i < h.num_hashtables; ++i) {
        if (raw) {
            const double max_fill = in.get<double>();
            const uint64_t filled = in.get<uint64_t>();
            const uint64_t allocated = in.get<uint64_t>();
            in.skip(2 * sizeof(uint64_t));  // size_when_restruct, allocated_mask (both derived)
            in.skip(sizeof(uint64_t));      // ht_memory (derived)
            const uint64_t ht_total = in.get<uint64_t>(), ht_match = in.get<uint64_t>();
            const uint64_t bv_words = (allocated + 63) / 64;
            if (!with_tables) { in.skip(bv_words * sizeof(uint64_t) + filled * 8); continue; }
            if (allocated == 0 || (allocated & (allocated - 1)) || filled > allocated)
                throw std::runtime_error("Corrupt k-mer database " + path);
            HashTable& ht = t.tables[i];
            ht.max_fill = max_fill; ht.filled = filled; ht.ht_total = ht_total; ht.ht_match = ht_match;
            bv.resize(bv_words); items.resize(filled);
            in.read(bv.data(), bv_words * 8);
            in.read(items.data(), filled * 8);   // {u32 key; i32 val} in slot order
            ht.slots.assign(allocated, HashTable::kEmptySlot);
            uint64_t next = 0;
            for (uint64_t w = 0; w < bv_words; ++w) {
                uint64_t bits = bv[w];
                while (bits) {
                    const uint64_t slot = w * 64 + (uint64_t)__builtin_ctzll(bits);
                    bits &= bits - 1;
                    if (slot >= allocated || next >= filled) throw std::runtime_error("Corrupt k-mer database " + path);
                    ht.slots[slot] = items[next++];
                }
            }
            if (next != filled) throw std::runtime_error("Corrupt k-mer database " + path);
        } else {  // portioned form (src/prefix_kmer_db.cpp:657-697)
            const uint64_t total = in.get<uint64_t>();
            uint64_t seen = 0;
            while (seen < total) {
                const uint64_t portion = in.get<uint64_t>();
                // NB: the reference seeks `portion` BYTES here (src/prefix_kmer_db.cpp:676)
                // although the portion holds 8-byte items; a correct reader skips the items.
                in.skip(portion * 8);
                seen += portion;
            }
        }
    }

    const uint64_t P = in.get<uint64_t>();
    {
        const off_t here = ftello(in.f);
        if (here >= 0 && !stream_only && read_patterns_mapped(path, (uint64_t)here, P, t)) return;
    }
    t.num_kmers.resize(P); t.parent_id.resize(P); t.n.resize(P); t.l.resize(P);
    t.last.resize(P); t.bits.resize(P); t.payload_off.resize(P);
    t.payload.clear();
    {   // everything left in the file is pattern blocks: an upper bound for the payload, so that
        // (possibly page-locked) storage is allocated once
        const off_t here = ftello(in.f);
        if (fseeko(in.f, 0, SEEK_END) == 0) {
            const off_t end = ftello(in.f);
            const uint64_t rest = end > here ? (uint64_t)(end - here) : 0;
            const uint64_t headers = P * kPatternHeaderBytes;
            if (rest > headers) t.payload.reserve((rest - headers) / 8 + 2);
        }
        fseeko(in.f, here, SEEK_SET);
    }

    std::unique_ptr<char[]> block(new char[kIoBlockBytes]);
    uint64_t pid = 0;
    while (pid < P) {
        const uint64_t block_bytes = in.get<uint64_t>();
        if (block_bytes > kIoBlockBytes) throw std::runtime_error("Corrupt k-mer database " + path);
        in.read(block.get(), block_bytes);
        const char* p = block.get();
        const char* end = p + block_bytes;
        while (p < end) {
            if (pid >= P || p + kPatternHeaderBytes > end)
                throw std::runtime_error("Corrupt k-mer database " + path);
            int64_t nk, par; uint32_t ns, nl, ls, nb;
            std::memcpy(&nk, p, 8); std::memcpy(&par, p + 8, 8);
            std::memcpy(&ns, p + 16, 4); std::memcpy(&nl, p + 20, 4);
            std::memcpy(&ls, p + 24, 4); std::memcpy(&nb, p + 28, 4);
            p += kPatternHeaderBytes;  // bytes 32..39: is_parent (4 valid + 4 undefined bytes)
            const uint64_t words = Trie::payload_words_for_bits(nb);
            if (p + words * 8 > end) throw std::runtime_error("Corrupt k-mer database " + path);
            t.num_kmers[pid] = nk; t.parent_id[pid] = par; t.n[pid] = ns; t.l[pid] = nl;
            t.last[pid] = ls; t.bits[pid] = nb;
            t.payload_off[pid] = t.payload.size();
            if (words) {
                const size_t at = t.payload.size();
                t.payload.resize(at + words);
                std::memcpy(t.payload.data() + at, p, words * 8);
                p += words * 8;
            }
            ++pid;
        }
    }
    t.build_compact();
}

void write_db(const std::string& path, const Trie& t) {
    File out(path, "wb");
    const DbHeader& h = t.hdr;
    out.put<uint64_t>(1);  // raw hashtables
    out.put(h.kmer_length); out.put(h.fraction); out.put(h.start_fraction);
    out.put(h.alphabet_type); out.put(h.is_initialized); out.put(h.kmers_count);
    out.put<uint64_t>(t.sample_names.size());
    for (size_t i = 0; i < t.sample_names.size(); ++i) {
        out.put<uint64_t>(t.sample_kmers[i]);
        out.put<uint64_t>(t.sample_names[i].size());
        out.write(t.sample_names[i].data(), t.sample_names[i].size());
    }
    if (!t.tables.empty()) {   // raw form of every table (src/hashmap_lp.h:481-528)
        out.put<uint64_t>(t.tables.size());
        std::vector<uint64_t> bv, items;
        for (const HashTable& ht : t.tables) {
            const uint64_t allocated = ht.allocated();
            out.put<double>(ht.max_fill);
            out.put<uint64_t>(ht.filled);
            out.put<uint64_t>(allocated);
            out.put<uint64_t>((uint64_t)((double)allocated * ht.max_fill));  // size_when_restruct
            out.put<uint64_t>(allocated - 1);                                // allocated_mask
            out.put<uint64_t>(allocated * 8);                                // ht_memory
            out.put<uint64_t>(ht.ht_total); out.put<uint64_t>(ht.ht_match);
            bv.assign((allocated + 63) / 64, 0);
            items.clear(); items.reserve(ht.filled);
            for (uint64_t i = 0; i < allocated; ++i)
                if (!HashTable::is_empty(ht.slots[i])) { bv[i >> 6] |= 1ull << (i & 63); items.push_back(ht.slots[i]); }
            out.write(bv.data(), bv.size() * 8);
            out.write(items.data(), items.size() * 8);
        }
    } else {
        // empty raw hashtables: capacity 16, nothing filled (SURVEY.md §8d cfg3 note)
        out.put<uint64_t>(h.num_hashtables);
        for (uint64_t i = 0; i < h.num_hashtables; ++i) {
            out.put<double>(0.8);
            out.put<uint64_t>(0);    // filled
            out.put<uint64_t>(16);   // allocated
            out.put<uint64_t>(12);   // size_when_restruct = allocated * max_fill
            out.put<uint64_t>(15);   // allocated_mask
            out.put<uint64_t>(0); out.put<uint64_t>(0); out.put<uint64_t>(0);
            out.put<uint64_t>(0);    // one bit-vector word, no slot used
        }
    }
    const uint64_t P = t.num_patterns();
    out.put<uint64_t>(P);
    std::vector<uint8_t> has_child(P, 0);  // pattern_t::is_parent (consulted by build -extend)
    for (uint64_t p = 0; p < P; ++p)
        if (t.parent_id[p] >= 0) has_child[(uint64_t)t.parent_id[p]] = 1;
    std::unique_ptr<char[]> block(new char[kIoBlockBytes]);
    char* cur = block.get();
    auto flush = [&]() {
        const uint64_t bytes = (uint64_t)(cur - block.get());
        out.put<uint64_t>(bytes);
        out.write(block.get(), bytes);
        cur = block.get();
    };
    for (uint64_t p = 0; p < P; ++p) {
        const uint64_t words = Trie::payload_words_for_bits(t.bits[p]);
        if (cur + kRefPatternStructBytes + words * 8 > block.get() + kIoBlockBytes) flush();
        std::memcpy(cur, &t.num_kmers[p], 8); std::memcpy(cur + 8, &t.parent_id[p], 8);
        std::memcpy(cur + 16, &t.n[p], 4); std::memcpy(cur + 20, &t.l[p], 4);
        std::memcpy(cur + 24, &t.last[p], 4); std::memcpy(cur + 28, &t.bits[p], 4);
        const uint64_t is_parent = has_child[p];
        std::memcpy(cur + 32, &is_parent, 8);
        cur += kPatternHeaderBytes;
        if (words) { std::memcpy(cur, t.payload.data() + t.payload_off[p], words * 8); cur += words * 8; }
    }
    flush();
}

}  // namespace kdbx
