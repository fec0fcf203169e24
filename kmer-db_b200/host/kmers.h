// k-mer extraction for the host side of `build` and `new2all`: alphabets, rolling canonical
// k-mers with the >= 8-bit prefix shift, and the minhash filter.  Behaviour follows the
// reference: KmerHelper::extract (src/kmer_extract.h:13-97), the alphabet
// table (src/alphabet.h:79-86) and MinHashFilter (src/filter.h:40-115).  Written from that
// description; the rolling update below keeps forward and reverse-complement words in the
// same way any 2-bit k-mer scanner does.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

namespace kdbx {

// enum AlphabetType of the reference (src/alphabet.h:10-18); the value is stored in the .db.
enum AlphabetId : int32_t { kNt = 0, kNtPreserve = 1, kAa = 2, kAa11Diamond = 3, kAa12Mmseqs = 4, kAa6Dayhoff = 5 };

struct Alphabet {
    int32_t id = kNt;
    std::string name;
    bool preserve_strand = false;
    int size = 4;
    int bits_per_symbol = 2;
    int max_kmer_len = 31;
    int8_t map[256];

    static Alphabet make(int32_t id) {
        struct Desc { int32_t id; const char* name; const char* groups; bool preserve; };
        static const Desc table[] = {
            {kNt, "nt", "A,C,G,TU", false},
            {kNtPreserve, "nt-preserve", "A,C,G,TU", true},
            {kAa, "aa", "K,R,E,D,Q,N,C,G,H,I,L,V,M,F,Y,W,P,S,T,A", true},
            {kAa11Diamond, "aa11_diamond", "KREDQN,C,G,H,ILV,M,F,Y,W,P,STA", true},
            {kAa12Mmseqs, "aa12_mmseqs", "AST,C,DN,EQ,FY,G,H,IV,KR,LM,P,W", true},
            {kAa6Dayhoff, "aa6_dayhoff", "STPAG,NDEQ,HRK,MILV,FYW,C", true},
        };
        for (const Desc& d : table) {
            if (d.id != id) continue;
            Alphabet a;
            a.id = id; a.name = d.name; a.preserve_strand = d.preserve;
            for (int i = 0; i < 256; ++i) a.map[i] = -1;
            int group = 0;
            for (const char* c = d.groups; *c; ++c) {
                if (*c == ',') { ++group; continue; }
                const unsigned char u = (unsigned char)*c;
                a.map[u] = (int8_t)group;                    // groups are given in upper case
                a.map[u - 'A' + 'a'] = (int8_t)group;
            }
            a.size = group + 1;
            a.bits_per_symbol = 0;
            while ((1 << a.bits_per_symbol) < a.size) ++a.bits_per_symbol;
            a.max_kmer_len = 64 / a.bits_per_symbol - 1;  // top bit is reserved (src/alphabet.h:41)
            return a;
        }
        throw std::runtime_error("Invalid alphabet type");
    }
    static Alphabet by_name(const std::string& name) {
        for (int32_t id = kNt; id <= kAa6Dayhoff; ++id) {
            Alphabet a = make(id);
            if (a.name == name) return a;
        }
        throw std::runtime_error("Invalid alphabet type");
    }
};

// Keep k-mer x iff lo <= h(x) < hi (src/filter.h:48-51).  f >= 1 accepts everything
// (NullFilter, src/filter.h:120-145).
struct MinHash {
    bool accept_all = true;
    uint64_t lo = 0, hi = 0, k_div_4 = 0, seed_mix = 0;

    MinHash() = default;
    MinHash(double fraction, double start, uint32_t k) {
        accept_all = !(fraction < 1.0);
        const double top = (double)std::numeric_limits<uint64_t>::max();
        lo = (uint64_t)(top * start);
        hi = (uint64_t)(top * (start + fraction));
        k_div_4 = (uint64_t)std::ceil((double)k / 4);
        seed_mix = 42 ^ k_div_4;
    }
    static uint64_t fmix64(uint64_t x) {
        x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
        x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
        x ^= x >> 33;
        return x;
    }
    // one MurmurHash3-x64-128 block round on the k-mer word, seeds 42 (src/filter.h:96-115)
    uint64_t hash(uint64_t x) const {
        uint64_t h = x * 0x87c37b91114253d5ull;
        h = (h << 31) | (h >> 33);
        h *= 0x4cf5ad432745937full;
        uint64_t h1 = (42 ^ h) ^ k_div_4;
        uint64_t h2 = seed_mix;
        h1 += h2; h2 += h1;
        h1 = fmix64(h1); h2 = fmix64(h2);
        h1 += h2; h2 += h1;
        return h1 ^ h2;
    }
    bool operator()(uint64_t x) const {
        if (accept_all) return true;
        const uint64_t h = hash(x);
        return h >= lo && h < hi;
    }
};

// Shift applied to every stored k-mer so that prefix = kmer >> 32 has at least 8 bits; the
// shifted-out low bits are duplicated (src/kmer_extract.h:36-45,87-88).
inline uint32_t prefix_shift(uint32_t k, int bits_per_symbol) {
    const int prefix_bits = (int)k * bits_per_symbol - 32;
    return prefix_bits < 8 ? (uint32_t)(8 - prefix_bits) : 0u;
}
// number of prefix buckets (hashtables) of a database (src/prefix_kmer_db.cpp:54-62)
inline uint64_t num_prefix_tables(uint32_t k, int bits_per_symbol) {
    int prefix_bits = (int)k * bits_per_symbol - 32;
    if (prefix_bits < 8) prefix_bits = 8;
    return 1ull << prefix_bits;
}

// Appends the (filtered) k-mers of one sequence to `out`.  Symbols outside the alphabet
// invalidate every window that contains them.
inline void extract_kmers(const char* seq, size_t len, uint32_t k, const Alphabet& al, const MinHash& filter,
                          std::vector<uint64_t>& out) {
    if (len < k || k == 0) return;
    const int b = al.bits_per_symbol;
    const uint64_t mask = (b * k >= 64) ? ~0ull : ((1ull << (b * k)) - 1);
    const uint32_t top_shift = (k - 1) * b;
    const uint32_t shift = prefix_shift(k, b);
    const uint64_t tail_mask = shift ? ((1ull << shift) - 1) : 0;
    uint64_t fwd = 0, rev = 0;
    uint32_t valid = 0;  // number of consecutive in-alphabet symbols ending here, capped at k
    for (size_t i = 0; i < len; ++i) {
        int s = al.map[(unsigned char)seq[i]];
        if (s < 0) { s = 0; valid = 0; }
        else if (valid < k) ++valid;
        fwd = ((fwd << b) | (uint64_t)s) & mask;
        rev = (rev >> b) | ((uint64_t)(al.size - 1 - s) << top_shift);
        if (valid < k) continue;
        uint64_t can = (al.preserve_strand || fwd < rev) ? fwd : rev;
        can = (can << shift) | (can & tail_mask);
        if (filter(can)) out.push_back(can);
    }
}

}  // namespace kdbx
