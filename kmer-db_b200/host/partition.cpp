// Sharding a database's pattern trie across GPUs (ours; the reference is single-process).
//
// The shared-k-mer matrix is LINEAR in the num_kmers vector: every pattern p adds num_kmers_p to
// each pair of its full sample list (the flat form the reference's sparse path executes,
// src/similarity_calculator.cpp:596-638; the tree form of the dense path, :64-72 + :206-241,
// regroups the same sum).  So any split of the patterns into disjoint OWNED sets gives partial
// matrices that add up to the whole one, provided each part is itself a valid trie: it carries
// its owned patterns plus their ancestors, the latter with num_kmers = 0.
//
// To keep the replicated ancestors few, the owned sets are contiguous pieces of the trie's
// depth-first preorder: the ancestor closure of such a piece is the piece plus ONE root-to-node
// chain (the ancestors of its first pattern).  Pieces are balanced on a per-pattern cost model of
// the GPU pipeline (updates of the scatter-add + list entries expanded + local ids decoded + a constant per node).
// A part is renumbered in preorder (parents stay before children, as kdbx_load_patterns requires),
// keeps the whole sample table, and its Elias-gamma payload is gathered into a compact blob, so a
// GPU uploads and decodes only its own share.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>

#include "gamma.h"
#include "trie.h"

namespace kdbx {

namespace {
// cost of a pattern in "updates" of the scatter kernel (0.32 ps each at config 2, profiles/r02_bench_cfg2_v3.json):
// l(2n-l-1)/2 updates; n list entries expanded at ~3.7 ps each (12 updates); l local ids decoded and walked by the fill
// pass at ~18 ps each (57 updates); ~220 ps per node for packing, sorting, scans and the level passes (680 updates)
inline uint64_t pattern_cost(uint32_t n, uint32_t l) {
    const uint64_t nn = n, ll = l;
    return ll * (2 * nn - ll - 1) / 2 + 12 * nn + 57 * ll + 680;
}
}  // namespace

// preorder[i] = pattern visited i-th by a depth-first walk (children in ascending id order)
std::vector<uint32_t> trie_preorder(const Trie& t) {
    const uint64_t P = t.num_patterns();
    if (P >= ((uint64_t)1 << 31)) throw std::runtime_error("partition: too many patterns");
    // children lists by counting sort on the parent id; slot P collects the roots (parent = -1)
    std::vector<uint32_t> first(P + 2, 0), child(P);
    for (uint64_t p = 0; p < P; ++p) {
        const int64_t q = t.parent_id[p];
        if (q < -1 || q >= (int64_t)p) throw std::runtime_error("partition: parent_id must be -1 or < own id");
        ++first[(q < 0 ? P : (uint64_t)q) + 1];
    }
    for (uint64_t i = 0; i <= P; ++i) first[i + 1] += first[i];
    {
        std::vector<uint32_t> at(first.begin(), first.end() - 1);
        for (uint64_t p = 0; p < P; ++p) {
            const int64_t q = t.parent_id[p];
            child[at[q < 0 ? P : (uint64_t)q]++] = (uint32_t)p;
        }
    }
    std::vector<uint32_t> order;
    order.reserve(P);
    std::vector<uint32_t> stack;
    for (uint32_t r = first[P + 1]; r-- > first[P];) stack.push_back(child[r]);
    while (!stack.empty()) {
        const uint32_t p = stack.back();
        stack.pop_back();
        order.push_back(p);
        for (uint32_t c = first[p + 1]; c-- > first[p];) stack.push_back(child[c]);
    }
    if (order.size() != P) throw std::runtime_error("partition: trie is not a forest");
    return order;
}

// Cuts of the preorder at equal shares of the cost: part g owns pre[cut[g] .. cut[g+1]).
TriePartitioner::TriePartitioner(const Trie& src, uint32_t num_parts) : src_(src), num_parts_(num_parts) {
    if (num_parts == 0) throw std::runtime_error("partition: bad number of parts");
    const uint64_t P = src.num_patterns();
    pre_ = trie_preorder(src);
    long double total = 0;
    for (uint64_t p = 0; p < P; ++p) total += (long double)pattern_cost(src.n[p], src.l[p]);
    cut_.assign((size_t)num_parts + 1, P);
    cut_[0] = 0;
    long double run = 0;
    uint32_t g = 1;
    for (uint64_t i = 0; i < P && g < num_parts; ++i) {
        while (g < num_parts && run >= total * g / num_parts) cut_[g++] = i;
        run += (long double)pattern_cost(src.n[pre_[i]], src.l[pre_[i]]);
    }
}

// dst := part `part` (see the header of this file).  The parts' all2all matrices sum to the all2all matrix
// of src.  owned_updates (optional) receives U of the owned patterns; window (optional) the band of sample ids
// [lo, hi) the part's lists lie in (what kdbx_set_sample_window wants to hear).
void TriePartitioner::extract(uint32_t part, Trie& dst, uint64_t* owned_updates, uint32_t* window) const {
    if (part >= num_parts_) throw std::runtime_error("partition: bad part index");
    const Trie& src = src_;
    const uint64_t P = src.num_patterns();
    const std::vector<uint32_t>& pre = pre_;
    const uint64_t a = cut_[part], b = std::max(cut_[part], cut_[(size_t)part + 1]);
    // nodes of the part: the ancestor chain of pre[a] (root first), then the owned piece
    std::vector<uint32_t> nodes;
    if (a < b) {
        for (int64_t q = src.parent_id[pre[a]]; q >= 0; q = src.parent_id[q]) nodes.push_back((uint32_t)q);
        std::reverse(nodes.begin(), nodes.end());
    }
    const size_t chain = nodes.size();
    nodes.insert(nodes.end(), pre.begin() + a, pre.begin() + b);
    // the sentinel pattern 0 (no samples) opens every kmer-db trie; keep that convention
    const bool add_sentinel = nodes.empty() || src.n[nodes[0]] != 0;
    const size_t Q = nodes.size() + (add_sentinel ? 1 : 0);
    std::vector<int32_t> new_id(P, -1);
    for (size_t i = 0; i < nodes.size(); ++i) new_id[nodes[i]] = (int32_t)(i + (add_sentinel ? 1 : 0));

    dst.hdr = src.hdr;
    dst.tables.clear();
    dst.sample_names = src.sample_names;
    dst.sample_kmers = src.sample_kmers;
    dst.num_kmers.resize(Q); dst.parent_id.resize(Q); dst.n.resize(Q); dst.l.resize(Q);
    dst.last.resize(Q); dst.bits.resize(Q); dst.payload_off.resize(Q);
    uint64_t words = 0, U = 0;
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    std::vector<uint32_t> ids;
    size_t o = 0;
    if (add_sentinel) {
        dst.num_kmers[0] = 0; dst.parent_id[0] = -1; dst.n[0] = 0; dst.l[0] = 0; dst.last[0] = 0; dst.bits[0] = 0;
        dst.payload_off[0] = 0;
        o = 1;
    }
    for (size_t i = 0; i < nodes.size(); ++i, ++o) {
        const uint32_t p = nodes[i];
        const bool owned = i >= chain;
        dst.num_kmers[o] = owned ? src.num_kmers[p] : 0;
        const int64_t q = src.parent_id[p];
        dst.parent_id[o] = q < 0 ? -1 : (int64_t)new_id[q];
        dst.n[o] = src.n[p]; dst.l[o] = src.l[p]; dst.last[o] = src.last[p]; dst.bits[o] = src.bits[p];
        dst.payload_off[o] = words;
        words += Trie::payload_words_for_bits(src.bits[p]);
        if (owned) { const uint64_t nn = src.n[p], ll = src.l[p]; U += ll * (2 * nn - ll - 1) / 2; }
        if (src.l[p]) {
            hi = std::max(hi, src.last[p] + 1);
            if (q < 0) {   // a root's first id is the smallest id of every list below it
                ids.resize(src.l[p]);
                decode_local_ids(src.payload.data() + src.payload_off[p], src.l[p], src.last[p], ids.data());
                lo = std::min(lo, ids[0]);
            }
        }
    }
    dst.payload.resize(words, 0);
    o = add_sentinel ? 1 : 0;
    for (size_t i = 0; i < nodes.size(); ++i, ++o) {
        const uint32_t p = nodes[i];
        const uint64_t w = Trie::payload_words_for_bits(src.bits[p]);
        if (w) std::memcpy(dst.payload.data() + dst.payload_off[o], src.payload.data() + src.payload_off[p], w * 8);
    }
    dst.build_compact();
    if (owned_updates) *owned_updates = U;
    if (window) { window[0] = hi ? lo : 0; window[1] = hi; }
}

void partition_trie(const Trie& src, uint32_t num_parts, uint32_t part, Trie& dst, uint64_t* owned_updates) {
    if (num_parts == 0 || part >= num_parts) throw std::runtime_error("partition: bad part index");
    TriePartitioner(src, num_parts).extract(part, dst, owned_updates, nullptr);
}

// Moves every sample of the database `offset` places up inside a table of `new_total` samples
// (the new places before and after are empty samples).  Only last_sample_id is absolute in a
// pattern (the Elias-gamma payload holds differences, src/pattern.cpp:99-109), so the trie keeps
// its shape.  Used to lay shards of a cluster-structured workload side by side (bench.py, weak scaling).
void relabel_samples(Trie& t, uint32_t offset, uint32_t new_total) {
    const uint32_t N = t.num_samples();
    if ((uint64_t)offset + N > new_total) throw std::runtime_error("relabel: samples do not fit the new table");
    const uint64_t P = t.num_patterns();
    for (uint64_t p = 0; p < P; ++p)
        if (t.l[p]) t.last[p] += offset;
    std::vector<std::string> names(new_total);
    std::vector<uint64_t> kmers(new_total, 0);
    for (uint32_t s = 0; s < new_total; ++s) {
        if (s >= offset && s < offset + N) { names[s] = t.sample_names[s - offset]; kmers[s] = t.sample_kmers[s - offset]; }
        else { char buf[32]; std::snprintf(buf, sizeof buf, "e%06u", s); names[s] = buf; }
    }
    t.sample_names.swap(names);
    t.sample_kmers.swap(kmers);
    t.tables.clear();
}

}  // namespace kdbx
