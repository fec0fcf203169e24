// libkdbx.so — B200 (sm_100a) implementation of kmer-db's dense all2all common-k-mer counting.
//
// Replaces SimilarityCalculator::all2all (src/similarity_calculator.cpp:42-438) and what it
// calls: the W accumulation up the trie (:64-72), pattern_t::decodeSamples + CEliasGamma
// (src/pattern.cpp:99-109, src/elias_gamma.h:133-256), the counting sort of row-add jobs by
// sample id (:166-203, :244-280) and the row_add inner loop (src/simd/row_add_avx2.cpp:31-124).
// Nothing here is a translation: the CPU code streams 8 MB cache blocks through four thread
// pools; this file keeps the whole trie in HBM, expands full sample-id lists in L2-sized
// chunks and runs the scatter-add on CTA-shared shared-memory accumulator tiles.
//
// Pipeline of one kdbx_all2all_dense* call (all on one stream, two host syncs up front):
//   prepare : scans of l / n / chunk cost (CUB), node packing, W accumulation, gamma decode
//   per chunk of patterns [p0,p1):
//     K_expand     full list of p = locals of its ancestors (root first) ++ locals of p
//     K_job_hist   one job per (pattern, local position i, column tile t): row = full[i],
//                  cols = full[0..i) restricted to tile t; histogram by key = row*T + t
//     scan         bucket offsets
//     K_job_fill   counting-sort scatter of 16-byte job records
//     K_units      split every bucket into work units of ~unit_updates updates
//     K_scatter    THE hot kernel: a CTA owns a (row, tile) accumulator in shared memory, its
//                  warps stream the jobs' id runs with coalesced loads and red.shared.add.u32,
//                  then the tile is flushed into the packed triangle with red.global.add.u32
// Integer adds commute, so any schedule gives the reference's bits (uint32 wrap included).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "../../include/kdbx.h"

namespace {

// ------------------------------------------------------------------------------------------
// device-side records
// ------------------------------------------------------------------------------------------
struct __align__(32) Node {  // one 32-byte sector per chain step
    int32_t parent;
    uint32_t n;       // samples in node + ancestors
    uint32_t l;       // local samples
    uint32_t last;    // last local sample id
    uint64_t loff;    // offset of the node's decoded local ids in d_loc
    int32_t up2;      // grandparent and great-grandparent (-1: none): skip pointers that let the
    int32_t up3;      // expansion keep three chain steps in flight per round trip
};

struct __align__(16) Unit {  // jobs [job_begin, job_end) all belong to one key = row_block*T + tile
    uint32_t key;
    uint32_t job_begin;
    uint32_t job_end;
    uint32_t pad;
};


__device__ __forceinline__ uint64_t tri_offset(uint64_t row) { return row * (row - 1) / 2; }

// ------------------------------------------------------------------------------------------
// prepare kernels
// ------------------------------------------------------------------------------------------
struct U32AsU64 {
    const uint32_t* p; uint64_t n;
    __host__ __device__ uint64_t operator()(uint64_t i) const { return i < n ? (uint64_t)p[i] : 0ull; }
};
// slots of a pattern's full list in `flat`: lists start on 16-byte boundaries so that the level-order
// expansion can copy a parent's list with 128-bit loads and stores
struct ListSlots {
    const uint32_t* p; uint64_t n;
    __host__ __device__ uint64_t operator()(uint64_t i) const { return i < n ? (uint64_t)((p[i] + 3u) & ~3u) : 0ull; }
};
struct PayloadWords {
    const uint32_t* bits; uint64_t n;
    __host__ __device__ uint64_t operator()(uint64_t i) const {
        if (i >= n) return 0ull;
        const uint32_t b = bits[i];
        return b == 0 ? 0ull : (uint64_t)((b + 127u) / 128u) * 2ull;
    }
};
// chunking cost of a pattern: it needs n ids of flat space and at most l*(tiles its ids can
// span) job slots; both buffers hold `chunk_ids` entries.
struct ChunkCost {
    const uint32_t* n; const uint32_t* l; const uint32_t* last; uint64_t cnt; uint32_t tile_cols;
    __host__ __device__ uint64_t operator()(uint64_t i) const {
        if (i >= cnt) return 0ull;
        const uint64_t jobs = (uint64_t)l[i] * (uint64_t)(last[i] / tile_cols + 2u);  // windows a row of p can reach
        const uint64_t ids = (n[i] + 3u) & ~3u;
        return jobs > ids ? jobs : ids;
    }
};

__global__ void k_build_nodes(uint64_t P, const int64_t* __restrict__ parent, const int64_t* __restrict__ num_kmers,
                              const uint32_t* __restrict__ n, const uint32_t* __restrict__ l,
                              const uint32_t* __restrict__ last, const uint64_t* __restrict__ loff,
                              Node* __restrict__ nodes, uint32_t* __restrict__ W, uint32_t N, int* __restrict__ err) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    {   // structure every later kernel relies on: parent before child, n = n_parent + l <= N
        const int64_t q = parent[p];
        // (a pattern with a parent holds at least one sample of its own: the level-order passes rely on
        //  num_samples growing strictly along every parent chain)
        const bool ok = q >= -1 && q < (int64_t)p && l[p] <= n[p] && n[p] <= N &&
                        n[p] == (q >= 0 ? n[q] : 0u) + l[p] && (l[p] == 0 || last[p] < N) && (l[p] != 0 || q < 0);
        if (!ok) atomicExch(err, 4);
    }
    Node nd;
    nd.parent = (int32_t)parent[p];
    nd.n = n[p]; nd.l = l[p]; nd.last = last[p]; nd.loff = loff[p];
    const int64_t q1 = parent[p];
    const int64_t q2 = (q1 >= 0 && q1 < (int64_t)p) ? parent[q1] : -1;
    const int64_t q3 = (q2 >= 0 && q2 < q1) ? parent[q2] : -1;
    nd.up2 = (int32_t)q2; nd.up3 = (q3 < q2) ? (int32_t)q3 : -1;
    nodes[p] = nd;
    W[p] = (uint32_t)num_kmers[p];  // the reference adds (uint32_t)num_kmers (similarity_calculator.cpp:222)
}

// W_p = sum of num_kmers over p's subtree, mod 2^32 (reference: serial reverse sweep over the
// pattern array, similarity_calculator.cpp:64-72).  num_samples is a topological key: a child
// has strictly more samples than its parent (n_child = n_parent + l_child, l_child >= 1), so
// patterns are radix-sorted by n (descending) and every level pushes its finished sums to the
// parents in one launch.  (A per-node walk to the root with atomics was 80 % of the prepare
// stage: the ancestors near the roots are hit by millions of serialized L2 atomics.)
// the 32-bit mirrors of parent_id / num_kmers (kdbx_trie_view) back to the 64-bit arrays the kernels read
__global__ void k_widen_headers(uint64_t P, const int32_t* __restrict__ parent32, const uint32_t* __restrict__ num_kmers32,
                                int64_t* __restrict__ parent, int64_t* __restrict__ num_kmers) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    parent[p] = (int64_t)parent32[p];
    num_kmers[p] = (int64_t)num_kmers32[p];
}
__global__ void k_iota(uint64_t P, uint32_t* __restrict__ out) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) out[p] = (uint32_t)p;
}
// level_start[v] = first index of key v in the descending-sorted key array
// (keys come straight from the caller's trie: a num_samples beyond N is reported by k_build_nodes, and must
// not be used as an index here before the host has seen that flag)
__global__ void k_level_starts(uint64_t P, uint32_t N, const uint32_t* __restrict__ keys, uint32_t* __restrict__ level_start) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint32_t v = keys[i];
    if (v <= N && (i == 0 || v != keys[i - 1])) level_start[v] = (uint32_t)i;
}
__global__ void k_push_level(uint32_t count, const uint32_t* __restrict__ order, const int64_t* __restrict__ parent,
                             uint32_t* __restrict__ W) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t p = order[i];
    const int64_t q = parent[p];
    if (q < 0) return;
    const uint32_t v = W[p];
    if (v) atomicAdd(&W[q], v);
}

#include "gamma_tokens.cuh"   // the Elias-gamma parser: one walk of the bits, one token per run of zero bits

// Decodes the LOCAL ids of every pattern into d_loc (ascending), one thread per pattern.  The
// stream holds the deltas in append order and only the LAST id is stored (src/pattern.cpp:99-109):
// a thread walks its bits once, parking one token per delta >= 2 and per run of deltas of 1, then turns the tokens into
// ids front to back (gamma_tokens.cuh; measured equal to parking one delta per id, 24.6 ms of prepare either way at
// config 2 — runs average 3.4 ids there — and kept because the same code is compiled for the host and unit-tested
// without a GPU, tests/test_host.py).
// The local lists of the 128 consecutive patterns of a block are contiguous in d_loc, so they are
// staged in shared memory and written out with coalesced stores (a thread writing its own list
// straight to HBM costs one 32-byte sector per 4-byte id); blocks whose lists do not fit the stage
// write directly.
//
// With one column window (N <= tile_cols) a job is fully described by a run of local ids inside one
// block of matrix rows, so the decoder — which has every id in hand — also counts the jobs per key
// and their updates (DecodeHist): the first of the two job-enumeration passes of the bucketing
// stage is then not needed at all.  The rule must equal the fill pass's (walk_runs) exactly: a run
// ends where the row block changes; a job counts iff its weight and its number of updates
// k*i + k(k-1)/2 are non-zero.
// What a job costs the scatter kernel besides its updates (its record, two dependent loads, the row table, the dispatch),
// in updates: without it a row block with ten million tiny jobs (the first rows of every cluster: short lists, one or two
// updates each) weighed next to nothing, got ONE work unit, and that unit kept a single CTA busy for as long as the rest of
// the pass took (sm__cycles_active: 55 % of elapsed at 2000 genomes, 84 % at config 2; profiles/r02_ab_notes.txt).
constexpr unsigned long long kJobCost = 4096;
constexpr int kDecodeThreads = 128;
constexpr uint32_t kDecodeStage = 10240;  // ids (40 KB)
constexpr uint32_t kSmallL = 8;           // see enumerate_jobs
constexpr uint32_t kDecodeHistKeys = 512;   // row blocks the decoder can count for: 16384 samples in 32-row blocks
struct DecodeHist {
    uint32_t enabled, rb_shift, nkeys;
    uint64_t per;                       // patterns per block of the fill pass (a multiple of kDecodeThreads)
    uint32_t* blockhist;                // [fill block][key]
    unsigned long long* work;           // [key]
    unsigned long long* total_updates;
    const uint32_t* W;
    // boundary form of the local lists (diff.cuh): per pattern (entries appended to the parent's boundary list) << 1 |
    // (the list continues the parent's last run), and the sum of the appended entries over all patterns
    uint32_t* ownb;
    unsigned long long* sum_app;
};
__global__ void __launch_bounds__(kDecodeThreads)
k_decode_locals(uint64_t P, const Node* __restrict__ nodes, const uint64_t* __restrict__ loff,
                const uint32_t* __restrict__ bits, const uint64_t* __restrict__ poff, const uint64_t* __restrict__ payload,
                uint64_t payload_words, uint32_t* __restrict__ loc, uint32_t win_lo, uint32_t win_n, int* __restrict__ err, DecodeHist dh,
                uint64_t w_lo, uint64_t w_hi) {
    // Chunked launches while the payload is still arriving (prepare()): this launch takes the blocks whose payload —
    // the words of its 128 patterns plus the two guard words the bit reader may touch — ends inside (w_lo, w_hi] words
    // of the densely packed payload, i.e. inside the chunk that has just arrived.  (0, ~0] = every block.
    if (w_lo != 0 || w_hi != ~0ull) {
        const uint64_t pe = (uint64_t)blockIdx.x * kDecodeThreads + kDecodeThreads;
        const uint64_t endw = poff[pe < P ? pe : P] + 2;
        if (endw <= w_lo || endw > w_hi) return;
    }
    __shared__ uint32_t s_ids[kDecodeStage];
    // per key: jobs in the top 12 bits, their updates in units of 1024 in the low 20 (a block's 128 patterns
    // hold at most 128 runs per key and 128 * 32 * N / 1024 < 2^20 such units with N <= 3072) — ONE
    // 32-bit shared-memory reduction per run.  The updates only size the scatter kernel's work units, so the
    // rounding is harmless; the job counts are exact.
    __shared__ uint32_t s_pack[kDecodeHistKeys];
    if (dh.enabled) {
        for (uint32_t k = threadIdx.x; k < dh.nkeys; k += kDecodeThreads) s_pack[k] = 0;
        __syncthreads();
    }
    unsigned long long my_updates = 0;
    uint32_t my_app = 0;
    // closes the run of k rows that starts at list position i (all in row block rb)
    auto close_run = [&](uint32_t rb, uint32_t i, uint32_t k, uint32_t w) {
        const uint32_t upd = k * i + k * (k - 1u) / 2u;   // < 32 * 3072 + 496 (k <= 32: the rows of one block)
        if (w != 0 && upd != 0 && rb < dh.nkeys) atomicAdd(&s_pack[rb], (1u << 20) | ((upd + 512u) >> 10));   // (rb >= nkeys: id outside the window, flagged)
    };
    const uint64_t p0 = (uint64_t)blockIdx.x * kDecodeThreads;
    const uint64_t p = p0 + threadIdx.x;
    const uint64_t p_end = p0 + kDecodeThreads < P ? p0 + kDecodeThreads : P;
    const uint64_t base = loff[p0];
    const uint64_t total = loff[p_end] - base;
    const bool staged = total <= kDecodeStage;
    if (p < P) {
        const Node nd = nodes[p];
        uint32_t own = 0;
        if (nd.l) {
            uint32_t* out = staged ? s_ids + (nd.loff - base) : loc + nd.loff;
            // the parent's list must end before this one starts, or full lists would not ascend
            const bool has_par = nd.parent >= 0 && nodes[nd.parent].l;
            const uint32_t floor_id = has_par ? nodes[nd.parent].last + 1u : 0u;
            const uint32_t first = nd.n - nd.l;
            if (dh.enabled) my_updates = (unsigned long long)nd.l * first + (unsigned long long)nd.l * (nd.l - 1u) / 2u;
            // every id must lie in the sample window; the kernels downstream see ids relative to win_lo
            if (nd.last - win_lo >= win_n) atomicExch(err, 7);
            if (nd.l == 1) {
                out[0] = nd.last - win_lo;
                if (nd.last < floor_id) atomicExch(err, 3);
                if (nd.last < win_lo) atomicExch(err, 7);
                if (dh.enabled) close_run((nd.last - win_lo) >> dh.rb_shift, first, 1u, dh.W[p]);
                const uint32_t joined = has_par && nd.last == floor_id;
                own = ((2u - joined) << 1) | joined;
            } else {
                const uint32_t nb = bits[p];
                const uint64_t po = poff[p];
                bool ok = po + ((uint64_t)(nb + 127u) / 128u) * 2u <= payload_words;
                if (!ok) atomicExch(err, 5);
                uint32_t runs = 1;
                uint64_t sum = 0;
                if (ok) {
                    const int rc = gamma_parse_tokens(payload + po, nb, nd.l, out, sum, runs);
                    if (rc == 1) { atomicExch(err, 1); ok = false; }
                    else if (rc == 2 || sum > nd.last) { atomicExch(err, 2); ok = false; }
                }
                if (ok) {
                    uint32_t cur = nd.last - (uint32_t)sum;
                    if (cur < floor_id) atomicExch(err, 3);
                    if (cur < win_lo) { atomicExch(err, 7); ok = false; }
                    const uint32_t joined = has_par && cur == floor_id;
                    own = ((2u * runs - joined) << 1) | joined;
                    cur -= win_lo;
                    out[0] = cur;
                    if (!ok) {
                        for (uint32_t i = 1; i < nd.l; ++i) out[i] = 0;
                    } else if (!dh.enabled) {
                        tokens_to_ids(out, nd.l, cur);
                    } else {
                        const uint32_t w = dh.W[p];
                        tokens_to_ids_blocks(out, nd.l, cur, dh.rb_shift, [&](uint32_t rb, uint32_t j, uint32_t k) { close_run(rb, first + j, k, w); });
                    }
                } else {
                    for (uint32_t i = 0; i < nd.l; ++i) out[i] = 0;  // never chased: the call fails on the error flag
                }
            }
        }
        if (dh.ownb) { dh.ownb[p] = own; my_app = own >> 1; }
    }
    if (staged) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < (uint32_t)total; i += kDecodeThreads) loc[base + i] = s_ids[i];
    }
    if (dh.ownb) {
        for (int o = 16; o; o >>= 1) my_app += __shfl_xor_sync(0xffffffffu, my_app, o);
        if ((threadIdx.x & 31) == 0 && my_app) atomicAdd(dh.sum_app, (unsigned long long)my_app);
    }
    if (dh.enabled) {
        for (int o = 16; o; o >>= 1) my_updates += __shfl_xor_sync(0xffffffffu, my_updates, o);
        if ((threadIdx.x & 31) == 0 && my_updates) atomicAdd(dh.total_updates, my_updates);
        __syncthreads();
        // this block's 128 patterns lie inside one block of the fill pass (dh.per is a multiple of 128)
        for (uint32_t k = threadIdx.x; k < dh.nkeys; k += kDecodeThreads) {
            const uint32_t pk = s_pack[k];
            if (!pk) continue;
            atomicAdd(&dh.blockhist[(p0 / dh.per) * dh.nkeys + k], pk >> 20);
            atomicAdd(&dh.work[k], ((unsigned long long)(pk & 0xFFFFFu) << 10) + (unsigned long long)(pk >> 20) * kJobCost);  // never 0 for a used key
        }
    }
}

// smallest p with off[p] >= c * chunk  (off is the exclusive scan of the chunk cost, P+1 long)
__global__ void k_chunk_bounds(uint64_t P, const uint64_t* __restrict__ coff, uint64_t chunk, uint32_t nchunks,
                               uint64_t* __restrict__ bounds) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nchunks) return;
    if (c == nchunks) { bounds[c] = P; return; }
    const uint64_t target = (uint64_t)c * chunk;
    uint64_t lo = 0, hi = P;  // first index with coff[idx] >= target
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (coff[mid] >= target) hi = mid; else lo = mid + 1;
    }
    bounds[c] = lo;
}

// per-row update counts (sum over jobs of the position i), for work-balanced row sharding;
// block-private shared-memory accumulators when the rows fit (global atomics otherwise).
constexpr uint32_t kRowUpdSmemRows = 4096;
__global__ void k_row_updates(uint64_t P, uint32_t N, const Node* __restrict__ nodes, const uint32_t* __restrict__ loc,
                              unsigned long long* __restrict__ upd) {
    __shared__ unsigned long long s_upd[kRowUpdSmemRows];
    const bool priv = N <= kRowUpdSmemRows;
    if (priv) {
        for (uint32_t r = threadIdx.x; r < N; r += blockDim.x) s_upd[r] = 0;
        __syncthreads();
    }
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t p = gw; p < P; p += nw) {
        const Node nd = nodes[p];
        const uint32_t first = nd.n - nd.l;
        for (uint32_t j = lane; j < nd.l; j += 32) {
            const uint32_t i = first + j;
            if (!i) continue;
            const uint32_t row = loc[nd.loff + j];
            if (priv) atomicAdd(&s_upd[row], (unsigned long long)i);
            else atomicAdd(&upd[row], (unsigned long long)i);
        }
    }
    if (priv) {
        __syncthreads();
        for (uint32_t r = threadIdx.x; r < N; r += blockDim.x)
            if (s_upd[r]) atomicAdd(&upd[r], s_upd[r]);
    }
}

// ------------------------------------------------------------------------------------------
// per-chunk kernels
// ------------------------------------------------------------------------------------------

// Full list of p = local lists of its ancestors (root first) followed by its own; node q's
// locals land at positions [n_q - l_q, n_q).  Eight lanes per pattern walk the parent chain
// (most chain steps copy only a few ids, so a full warp per pattern would idle and — worse —
// keep only one pointer chase in flight); the parent's node is requested before the copy.
constexpr uint32_t kExpandLanes = 8;
__global__ void k_expand(uint64_t p0, uint64_t p1, const Node* __restrict__ nodes, const uint64_t* __restrict__ noff,
                         const uint32_t* __restrict__ loc, uint32_t* __restrict__ flat) {
    const uint64_t gid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kExpandLanes;
    const uint64_t ng = ((uint64_t)gridDim.x * blockDim.x) / kExpandLanes;
    const uint32_t sub = threadIdx.x & (kExpandLanes - 1);
    const uint64_t base0 = noff[p0];
    for (uint64_t p = p0 + gid; p < p1; p += ng) {
        const Node nd = nodes[p];
        if (nd.n == 0) continue;
        uint32_t* dst = flat + (noff[p] - base0);
        auto copy_locals = [&](const Node& q) {
            const uint32_t* src = loc + q.loff;
            const uint32_t at = q.n - q.l;
            for (uint32_t j = sub; j < q.l; j += kExpandLanes) dst[at + j] = src[j];
        };
        copy_locals(nd);
        // three ancestors per round trip: parent, grandparent and great-grandparent are known from
        // the node (skip pointers), so their records are requested together
        int32_t a1 = nd.parent, a2 = nd.up2, a3 = nd.up3;
        while (a1 >= 0) {
            const Node n1 = nodes[a1];
            Node n2 = n1, n3 = n1;
            if (a2 >= 0) n2 = nodes[a2];
            if (a3 >= 0) n3 = nodes[a3];
            copy_locals(n1);
            if (a2 < 0) break;
            copy_locals(n2);
            if (a3 < 0) break;
            copy_locals(n3);
            a1 = n3.parent; a2 = n3.up2; a3 = n3.up3;
        }
    }
}

// Resident expansion (used when the full lists of ALL patterns fit in HBM, 4 * sum n bytes): a
// pattern's full list is its parent's full list followed by its own local ids, so the patterns of
// one num_samples level — whose parents all sit on earlier levels, n_parent < n — are expanded by
// two contiguous copies each.  One launch per level, ascending; no pointer chasing at all.
// (8 lanes per pattern measured best: 17.1 ms per step at config 2 against 18.4 with 16 and 24.2 with 32,
// profiles/r01_small_ab.txt)
constexpr uint32_t kLevelLanes = 8;
template <uint32_t kLanes>
__global__ void k_expand_level(uint32_t count, const uint32_t* __restrict__ order, const Node* __restrict__ nodes,
                               const uint64_t* __restrict__ noff, const uint32_t* __restrict__ loc, uint32_t* flat, uint32_t* first_id) {
    const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) / kLanes;
    const uint32_t sub = threadIdx.x & (kLanes - 1);
    const bool have = gid < count;
    Node nd; nd.parent = -1; nd.n = 0; nd.l = 0; nd.last = 0; nd.loff = 0; nd.up2 = nd.up3 = -1;
    uint32_t* dst = flat;
    if (have) {
        const uint32_t p = order[gid];
        nd = nodes[p];
        dst = flat + noff[p];
        // first id of the full list (the job enumeration places a pattern's columns with it)
        if (sub == 0 && nd.n) first_id[p] = nd.parent >= 0 ? first_id[nd.parent] : loc[nd.loff];
    }
    const uint32_t npar = nd.n - nd.l;
    if (have && nd.parent >= 0) {
        // both lists start 16-byte aligned (ListSlots); the copy may run up to 3 ids past the parent's
        // list — still inside this pattern's slots, overwritten by its own ids below
        const uint4* src = reinterpret_cast<const uint4*>(flat + noff[nd.parent]);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        const uint32_t n4 = (npar + 3u) >> 2;
        for (uint32_t j = sub; j < n4; j += kLanes) d4[j] = src[j];
    }
    __syncwarp();  // the tail of the 128-bit copy must land before the own ids that overwrite it
    if (have) {
        const uint32_t* own = loc + nd.loff;
        for (uint32_t j = sub; j < nd.l; j += kLanes) dst[npar + j] = own[j];
    }
}

// A job = one pattern x one block of matrix rows x one column tile: rows full[A0 .. A0+k) (k local
// samples of the pattern that fall into the same block of `tile_rows` consecutive matrix rows) and,
// for row j, the columns full[a .. min(b, A0+j)) — [a,b) being the part of the (ascending) list that
// lies inside the column tile.  All k rows share the ids full[a .. min(b,A0)): one pass over them
// feeds k accumulator rows, which is what makes the scatter kernel shared-memory-bound instead
// of id-stream-bound (the reference re-reads the list once per row, src/similarity_calculator.cpp:
// 214-231; so did the first three versions of this file).
struct __align__(32) Job {
    uint32_t off;  // start of the pattern's full list in the chunk's flat id array (low 32 bits)
    uint32_t a, b;
    uint32_t A0;
    uint32_t k;
    uint32_t w;    // (uint32_t) W_p
    uint32_t off_hi;
    uint32_t pad;
};

// sum over rows j < k of max(0, min(b, A0 + j) - a): the number of row[col] += w updates of a job
__device__ __forceinline__ unsigned long long job_updates(uint32_t a, uint32_t b, uint32_t A0, uint32_t k) {
    if (b <= a) return 0;
    long long lo = (long long)a + 1 - (long long)A0;  // first row that sees any id of [a,b)
    if (lo < 0) lo = 0;
    if (lo >= (long long)k) return 0;
    long long sat = (long long)b - (long long)A0;     // first row that sees all of [a,b)
    if (sat < lo) sat = lo;
    if (sat > (long long)k) sat = k;
    const long long cnt = sat - lo;
    long long s = cnt * ((long long)A0 - (long long)a) + (lo + sat - 1) * cnt / 2;
    s += ((long long)k - sat) * (long long)(b - a);
    return (unsigned long long)s;
}

__device__ __forceinline__ uint32_t lower_bound_ids(const uint32_t* __restrict__ ids, uint32_t n, uint32_t target) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (ids[mid] >= target) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// Column windows.  The accumulator tile of row block rb does not sit on a fixed column grid: its
// window 0 ends right above the block's rows (at win_hi(rb), the 32-aligned end of the block) and
// window t covers the tile_cols columns [win_hi - (t+1) tile_cols, win_hi - t tile_cols) below it,
// clipped at column 0.  Samples that entered the database close together (a clade, a cluster, a
// shard of a larger database) have their ids within a few hundred places below the row, so all of
// a pattern's columns fall into window 0 whatever N is; with N <= tile_cols there is one window and
// it starts at column 0.  key = rb * T + t.
__device__ __forceinline__ uint32_t win_hi(uint32_t rb, uint32_t rb_shift) { return (((rb + 1u) << rb_shift) + 31u) & ~31u; }
__device__ __forceinline__ uint32_t win_col0(uint32_t hi, uint32_t t, uint32_t tile_cols) {
    const uint64_t span = (uint64_t)(t + 1u) * tile_cols;
    return span >= hi ? 0u : hi - (uint32_t)span;
}

// One run of k rows of pattern `nd` (list positions i .. i+k, all in row block rb): one job per
// column window the run's ids reach.  emit(key, job, updates).
// The ids a run sees are list[0 .. reach): they ascend from `first_id` = list[0] and stay below
// `row_last`, the id of the run's last row.  When both ends lie in one window (always when T == 1)
// the job is known without touching the list; otherwise the list is cut at the window boundaries by
// binary search, far window first.
template <class Emit>
__device__ __forceinline__ void emit_run(const Node& nd, uint64_t base, const uint32_t* __restrict__ list, uint32_t T,
                                         uint32_t tile_cols, uint32_t rb_shift, uint32_t rb, uint32_t i, uint32_t k, uint32_t first_id,
                                         uint32_t row_last, Emit& emit) {
    const uint32_t reach = i + k - 1;  // the last row of the run sees the ids [0, reach)
    if (reach == 0) return;
    Job jb;
    jb.off = (uint32_t)base; jb.off_hi = (uint32_t)(base >> 32); jb.A0 = i; jb.k = k; jb.w = 0; jb.pad = 0;
    uint32_t t_near = 0, t_far = 0, hi = 0;
    if (T > 1) {
        hi = win_hi(rb, rb_shift);               // row_last < hi, and first_id <= list[reach-1] < row_last
        t_near = (hi - row_last) / tile_cols;     // window of the id row_last - 1
        t_far = (hi - 1u - first_id) / tile_cols;
    }
    if (t_near == t_far) {  // one window: every row j of the run sees the ids [0, i + j)
        jb.a = 0; jb.b = nd.n;
        emit(rb * T + t_near, jb, (unsigned long long)k * i + (unsigned long long)(k * (k - 1u) / 2u));
        return;
    }
    uint32_t a = 0;
    for (uint32_t t = t_far; a < reach; --t) {
        // ids below the upper end of window t belong to it (lower windows were cut off before)
        const uint32_t b = (t == t_near) ? nd.n : a + lower_bound_ids(list + a, reach - a, hi - t * tile_cols);
        if (b > a) { jb.a = a; jb.b = b; emit(rb * T + t, jb, job_updates(a, b, i, k)); }
        a = b;
        if (t == t_near) break;
    }
}

// Warp-cooperative enumeration of one (long) pattern: lanes over the pattern's local positions.
template <class Emit>
__device__ __forceinline__ void jobs_of_pattern_warp(const Node& nd, uint64_t base, const uint32_t* __restrict__ flat,
                                                     const uint32_t* __restrict__ loc, uint32_t T, uint32_t tile_cols, uint32_t rb_shift, uint32_t row_begin,
                                                     uint32_t row_end, uint32_t lane, uint32_t first_id, unsigned long long& updates, Emit& emit) {
    const uint32_t first = nd.n - nd.l;
    const uint32_t rounds = (nd.l + 31) / 32;
    const uint32_t* list = flat + base;
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t j = r * 32 + lane;
        const bool have = j < nd.l;
        const uint32_t i = first + j;
        uint32_t row = 0xFFFFFFFFu;
        bool active = false;
        if (have) {
            row = loc[nd.loff + j];  // the rows are the node's own local ids: contiguous in d_loc
            active = row >= row_begin && row < row_end;
        }
        if (active) updates += i;
        const uint32_t rb = row >> rb_shift;
        const uint32_t prev_rb = __shfl_up_sync(0xffffffffu, rb, 1);
        const uint32_t amask = __ballot_sync(0xffffffffu, active);
        const bool start = active && (lane == 0 || !((amask >> (lane - 1)) & 1u) || prev_rb != rb);
        const uint32_t smask = __ballot_sync(0xffffffffu, start);
        // the run ends at the next start or after the last active lane (active lanes are contiguous:
        // rows ascend and [row_begin,row_end) is an interval)
        const uint32_t above = lane == 31 ? 0u : ((smask >> (lane + 1)) << (lane + 1));
        const uint32_t next_start = above ? (uint32_t)__ffs((int)above) - 1u : 32u;
        const uint32_t last_active = amask ? 32u - (uint32_t)__clz((int)amask) : 0u;
        const uint32_t run_end = min(next_start, last_active);           // one past the run's last lane
        const uint32_t row_last = __shfl_sync(0xffffffffu, row, (run_end > lane ? run_end : lane + 1u) - 1u);
        if (!start) continue;
        emit_run(nd, base, list, T, tile_cols, rb_shift, rb, i, run_end - lane, first_id, row_last, emit);
    }
}

// Enumerates the jobs of the patterns [lo, hi) handled by this warp (`warp` of `nwarps`, 32 patterns
// per step).  Most patterns hold a handful of local samples (a split creates a node with one), so a
// lane walks the rows of its own pattern when there are at most kSmallL of them — 32 patterns in
// flight per warp instead of one; longer local lists are then taken one by one by the whole warp.
// emit(key, job, updates, w) may be called by any lane; `updates` accumulates U per lane.
template <class Emit>
__device__ __forceinline__ void enumerate_jobs(uint64_t lo, uint64_t hi, uint32_t warp, uint32_t nwarps,
                                               const Node* __restrict__ nodes, const uint64_t* __restrict__ noff, uint64_t base0,
                                               const uint32_t* __restrict__ W, const uint32_t* __restrict__ flat,
                                               const uint32_t* __restrict__ loc, const uint32_t* __restrict__ first_ids, uint32_t T,
                                               uint32_t tile_cols, uint32_t rb_shift, uint32_t row_begin, uint32_t row_end, uint32_t lane,
                                               unsigned long long& updates, Emit emit) {
    for (uint64_t b = lo + (uint64_t)warp * 32; b < hi; b += (uint64_t)nwarps * 32) {
        const uint64_t p = b + lane;
        Node nd; nd.parent = -1; nd.n = 0; nd.l = 0; nd.last = 0; nd.loff = 0; nd.up2 = nd.up3 = -1;
        uint32_t w = 0, fid = 0;
        uint64_t base = 0;
        if (p < hi) {
            nd = nodes[p];
            if (nd.l) {
                w = W[p]; base = noff[p] - base0;
                // first id of the full list: kept per pattern by the level-order expansion, else read from the list
                if (T > 1) fid = first_ids ? first_ids[p] : flat[base];
            }
        }
        auto emit_w = [&](uint32_t key, const Job& jb, unsigned long long upd) { emit(key, jb, upd, w); };
        if (nd.l && nd.l <= kSmallL) {
            const uint32_t first = nd.n - nd.l;
            const uint32_t* list = flat + base;
            const uint32_t* rows = loc + nd.loff;
            const uint32_t first_id = fid;
            uint32_t run_i = 0, run_k = 0, run_rb = 0, run_last = 0;
            for (uint32_t j = 0; j < nd.l; ++j) {
                const uint32_t row = rows[j];
                const bool active = row >= row_begin && row < row_end;
                const uint32_t rb = row >> rb_shift;
                if (active) updates += first + j;
                if (run_k && (!active || rb != run_rb)) { emit_run(nd, base, list, T, tile_cols, rb_shift, run_rb, run_i, run_k, first_id, run_last, emit_w); run_k = 0; }
                if (active) {
                    if (!run_k) { run_i = first + j; run_rb = rb; }
                    ++run_k;
                    run_last = row;
                }
            }
            if (run_k) emit_run(nd, base, list, T, tile_cols, rb_shift, run_rb, run_i, run_k, first_id, run_last, emit_w);
        }
        uint32_t big = __ballot_sync(0xffffffffu, nd.l > kSmallL);
        while (big) {
            const int src = __ffs((int)big) - 1;
            big &= big - 1;
            Node bn;
            bn.parent = __shfl_sync(0xffffffffu, nd.parent, src);
            bn.n = __shfl_sync(0xffffffffu, nd.n, src);
            bn.l = __shfl_sync(0xffffffffu, nd.l, src);
            bn.last = 0; bn.loff = __shfl_sync(0xffffffffu, nd.loff, src); bn.up2 = bn.up3 = -1;
            const uint32_t bw = __shfl_sync(0xffffffffu, w, src);
            const uint64_t bbase = __shfl_sync(0xffffffffu, base, src);
            const uint32_t bfid = __shfl_sync(0xffffffffu, fid, src);
            auto emit_b = [&](uint32_t key, const Job& jb, unsigned long long upd) { emit(key, jb, upd, bw); };
            jobs_of_pattern_warp(bn, bbase, flat, loc, T, tile_cols, rb_shift, row_begin, row_end, lane, bfid, updates, emit_b);
        }
    }
}

// ---- job bucketing, small key spaces: block-private histograms in shared memory ------------
// With nkeys <= kSmemKeys every block counts its contiguous slice of patterns into shared memory
// and publishes one row of blockhist[block][key]; a per-key column scan then gives every
// (block, key) pair its exact slot range, so the fill pass needs no global atomics at all.
constexpr uint32_t kSmemKeys = 2048;
constexpr int kBucketThreads = 256;
#define KDBX_BUCKET_BOUNDS __launch_bounds__(kBucketThreads)

__device__ __forceinline__ void block_slice(uint64_t p0, uint64_t p1, uint64_t& lo, uint64_t& hi, uint64_t per_given = 0) {
    const uint64_t per = per_given ? per_given : (p1 - p0 + gridDim.x - 1) / gridDim.x;
    lo = p0 + (uint64_t)blockIdx.x * per;
    if (lo > p1) lo = p1;
    hi = lo + per < p1 ? lo + per : p1;
}

__global__ void KDBX_BUCKET_BOUNDS
k_job_hist_smem(uint64_t p0, uint64_t p1, const Node* __restrict__ nodes, const uint64_t* __restrict__ noff,
                const uint32_t* __restrict__ W, const uint32_t* __restrict__ flat_all, const uint32_t* __restrict__ loc, const uint32_t* __restrict__ first_ids, uint32_t resident,
                uint32_t T, uint32_t tile_cols, uint32_t rb_shift, uint32_t row_begin, uint32_t row_end, uint32_t nkeys, uint32_t* __restrict__ blockhist,
                unsigned long long* __restrict__ work, unsigned long long* __restrict__ total_updates, uint64_t per) {
    __shared__ uint32_t s_hist[kSmemKeys];
    __shared__ unsigned long long s_work[kSmemKeys];
    for (uint32_t k = threadIdx.x; k < nkeys; k += blockDim.x) { s_hist[k] = 0; s_work[k] = 0; }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint64_t lo, hi;
    block_slice(p0, p1, lo, hi, per);
    unsigned long long updates = 0;
    const uint32_t* flat = flat_all + (resident ? noff[p0] : 0ull);  // the chunk's lists start at its first pattern
    enumerate_jobs(lo, hi, warp, nwarps, nodes, noff, noff[p0], W, flat, loc, first_ids, T, tile_cols, rb_shift, row_begin, row_end, lane, updates,
                   [&](uint32_t key, const Job&, unsigned long long upd, uint32_t w) {
                       if (w == 0 || upd == 0) return;  // adds of 0 are skipped, but still counted in U
                       atomicAdd(&s_hist[key], 1u);
                       atomicAdd(&s_work[key], upd + kJobCost);
                   });
    for (int o = 16; o; o >>= 1) updates += __shfl_xor_sync(0xffffffffu, updates, o);
    if (lane == 0 && updates) atomicAdd(total_updates, updates);
    __syncthreads();
    uint32_t* out = blockhist + (size_t)blockIdx.x * nkeys;
    for (uint32_t k = threadIdx.x; k < nkeys; k += blockDim.x) {
        out[k] = s_hist[k];
        if (s_work[k]) atomicAdd(&work[k], s_work[k]);
    }
}

// hist[key] = sum over blocks (column sums of blockhist)
__global__ void k_key_totals(uint32_t nkeys, uint32_t nblocks, const uint32_t* __restrict__ blockhist, uint32_t* __restrict__ hist) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > nkeys) return;
    uint32_t sum = 0;
    if (k < nkeys)
        for (uint32_t b = 0; b < nblocks; ++b) sum += blockhist[(size_t)b * nkeys + k];
    hist[k] = sum;
}

// blockhist[b][key] := bucket_off[key] + sum_{b' < b} blockhist[b'][key]   (first slot of the pair)
// One warp per key: a running exclusive scan over the blocks, 32 at a time.
__global__ void k_block_offsets(uint32_t nkeys, uint32_t nblocks, const uint32_t* __restrict__ bucket_off,
                                uint32_t* __restrict__ blockhist) {
    const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (k >= nkeys) return;
    uint32_t run = bucket_off[k];
    for (uint32_t b0 = 0; b0 < nblocks; b0 += 32) {
        const uint32_t b = b0 + lane;
        const uint32_t c = b < nblocks ? blockhist[(size_t)b * nkeys + k] : 0u;
        uint32_t incl = c;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += up;
        }
        if (b < nblocks) blockhist[(size_t)b * nkeys + k] = run + incl - c;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

__global__ void KDBX_BUCKET_BOUNDS
k_job_fill_smem(uint64_t p0, uint64_t p1, const Node* __restrict__ nodes, const uint64_t* __restrict__ noff,
                const uint32_t* __restrict__ W, const uint32_t* __restrict__ flat_all, const uint32_t* __restrict__ loc, const uint32_t* __restrict__ first_ids, uint32_t resident,
                uint32_t T, uint32_t tile_cols, uint32_t rb_shift, uint32_t row_begin, uint32_t row_end, uint32_t nkeys, const uint32_t* __restrict__ blockbase,
                Job* __restrict__ jobs, uint64_t per) {
    __shared__ uint32_t s_next[kSmemKeys];
    const uint32_t* mine = blockbase + (size_t)blockIdx.x * nkeys;
    for (uint32_t k = threadIdx.x; k < nkeys; k += blockDim.x) s_next[k] = mine[k];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint64_t lo, hi;
    block_slice(p0, p1, lo, hi, per);
    unsigned long long updates = 0;
    const uint32_t* flat = flat_all + (resident ? noff[p0] : 0ull);  // the chunk's lists start at its first pattern
    enumerate_jobs(lo, hi, warp, nwarps, nodes, noff, noff[p0], W, flat, loc, first_ids, T, tile_cols, rb_shift, row_begin, row_end, lane, updates,
                   [&](uint32_t key, Job jb, unsigned long long upd, uint32_t w) {
                       if (w == 0 || upd == 0) return;
                       const uint32_t slot = atomicAdd(&s_next[key], 1u);
                       jb.w = w;
                       jobs[slot] = jb;
                   });
}

// ---- job bucketing, large key spaces: global atomics (contention is low when keys are many) --
__global__ void KDBX_BUCKET_BOUNDS
k_job_hist(uint64_t p0, uint64_t p1, const Node* __restrict__ nodes, const uint64_t* __restrict__ noff,
           const uint32_t* __restrict__ W, const uint32_t* __restrict__ flat_all, const uint32_t* __restrict__ loc, const uint32_t* __restrict__ first_ids, uint32_t resident,
           uint32_t T, uint32_t tile_cols, uint32_t rb_shift,
           uint32_t row_begin, uint32_t row_end, uint32_t* __restrict__ hist, unsigned long long* __restrict__ work,
           unsigned long long* __restrict__ total_updates) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint64_t lo, hi;
    block_slice(p0, p1, lo, hi);
    unsigned long long updates = 0;
    const uint32_t* flat = flat_all + (resident ? noff[p0] : 0ull);  // the chunk's lists start at its first pattern
    enumerate_jobs(lo, hi, warp, nwarps, nodes, noff, noff[p0], W, flat, loc, first_ids, T, tile_cols, rb_shift, row_begin, row_end, lane, updates,
                   [&](uint32_t key, const Job&, unsigned long long upd, uint32_t w) {
                       if (w == 0 || upd == 0) return;
                       atomicAdd(&hist[key], 1u);
                       atomicAdd(&work[key], upd + kJobCost);
                   });
    for (int o = 16; o; o >>= 1) updates += __shfl_xor_sync(0xffffffffu, updates, o);
    if (lane == 0 && updates) atomicAdd(total_updates, updates);
}

__global__ void KDBX_BUCKET_BOUNDS
k_job_fill(uint64_t p0, uint64_t p1, const Node* __restrict__ nodes, const uint64_t* __restrict__ noff,
           const uint32_t* __restrict__ W, const uint32_t* __restrict__ flat_all, const uint32_t* __restrict__ loc, const uint32_t* __restrict__ first_ids, uint32_t resident,
           uint32_t T, uint32_t tile_cols, uint32_t rb_shift,
           uint32_t row_begin, uint32_t row_end, uint32_t* __restrict__ cursor, Job* __restrict__ jobs) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint64_t lo, hi;
    block_slice(p0, p1, lo, hi);
    unsigned long long updates = 0;
    const uint32_t* flat = flat_all + (resident ? noff[p0] : 0ull);  // the chunk's lists start at its first pattern
    enumerate_jobs(lo, hi, warp, nwarps, nodes, noff, noff[p0], W, flat, loc, first_ids, T, tile_cols, rb_shift, row_begin, row_end, lane, updates,
                   [&](uint32_t key, Job jb, unsigned long long upd, uint32_t w) {
                       if (w == 0 || upd == 0) return;
                       const uint32_t slot = atomicAdd(&cursor[key], 1u);
                       jb.w = w;
                       jobs[slot] = jb;
                   });
}

// work[nkeys] = total work of the pass (one block)
__global__ void k_work_total(uint32_t nkeys, unsigned long long* __restrict__ work) {
    __shared__ unsigned long long s_part[32];
    unsigned long long sum = 0;
    for (uint32_t k = threadIdx.x; k < nkeys; k += blockDim.x) sum += work[k];
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) t += s_part[w];
        work[nkeys] = t;
    }
}

// units per key = ceil(work / unit) (>= 1 when the bucket is non-empty).  unit = unit_updates, or — when the plan
// leaves it to the library (target_units > 0) — the larger of that and total work / target_units: a CTA of the
// scatter kernel meets a barrier and re-reads its schedule once per unit, so a pass is cut into a few units per
// CTA (enough to balance the tail) rather than into units of a fixed size.
__global__ void k_unit_count(uint32_t nkeys, const uint32_t* __restrict__ hist, const unsigned long long* __restrict__ work,
                             uint32_t unit_updates_min, uint32_t target_units, uint32_t* __restrict__ ucount) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > nkeys) return;
    unsigned long long unit_updates = unit_updates_min;
    if (target_units) unit_updates = max(unit_updates, work[nkeys] / target_units);
    uint32_t c = 0;
    if (k < nkeys && hist[k]) {
        const unsigned long long u = (work[k] + unit_updates - 1) / unit_updates;
        c = (uint32_t)(u < 1 ? 1 : (u > hist[k] ? hist[k] : u));
    }
    ucount[k] = c;  // ucount[nkeys] = 0 so that the exclusive scan yields the total there
}

__global__ void k_unit_fill(uint32_t nkeys, const uint32_t* __restrict__ hist, const uint32_t* __restrict__ bucket_off,
                            const uint32_t* __restrict__ ucount, const uint32_t* __restrict__ uoff,
                            Unit* __restrict__ units) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkeys) return;
    const uint32_t c = ucount[k];
    if (c == 0) return;
    const uint32_t jobs = hist[k], b0 = bucket_off[k];
    const uint32_t per = (jobs + c - 1) / c;
    for (uint32_t u = 0; u < c; ++u) {
        Unit un;
        un.key = k;
        un.job_begin = b0 + min(jobs, u * per);
        un.job_end = b0 + min(jobs, (u + 1) * per);
        un.pad = 0;
        units[uoff[k] + u] = un;
    }
}

// THE hot kernel.  Persistent CTAs pull work units (one (row block, column tile) key and a range
// of its jobs) from a global counter.  The CTA owns a tile of tile_rows x tile_cols uint32
// accumulators in shared memory.  A warp takes a small batch of jobs; for each job it loads the
// job's k row ids once, then streams the ids shared by all k rows in 128-id slices (coalesced
// 128-byte loads) and, for every row, issues red.shared.add.u32 tile[row][id] += w — one id load
// feeds k reductions.  The short triangular tail (row j also receives the rows before it) follows.
// Measured on B200 (profiles/r01_microbench_atomics.txt): shared-memory reductions run at 2.3-5.1e12
// updates/s when ids come from registers, i.e. the shared-memory pipe — not HBM — is the bound.
// When a unit ends the tile is added into the packed lower-triangular matrix (src/array.h:140)
// with red.global.add.u32, skipping zero cells; consecutive units of one key keep the tile.
__device__ __forceinline__ void red_shared_add(uint32_t saddr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ldg_nc_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldg_nc_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

constexpr uint32_t kUnitsPerCta = 24;  // work units per CTA of a scatter pass when the plan does not fix their size (k_unit_count)
constexpr uint32_t kJobBatch = 8;  // jobs a warp claims at once (16 lanes load them as uint4 halves)

// Last pass of a job over its k rows (see k_scatter_add): F complete 32-id groups (x0..x2), the partial
// group xt under `tail`, and the triangular part (lanes below j hold the ids of the rows before row j).
template <uint32_t F>
__device__ __forceinline__ void last_rows(uint32_t k, uint32_t rowoff, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t xt, bool tail,
                                          bool has_tail, uint32_t wt, uint32_t my_id4, uint32_t lane, uint32_t w) {
#pragma unroll 2
    for (uint32_t j = 0; j < k; ++j) {
        const uint32_t ro = __shfl_sync(0xffffffffu, rowoff, j);
        if (F > 0) red_shared_add(ro + x0, w);
        if (F > 1) red_shared_add(ro + x1, w);
        if (F > 2) red_shared_add(ro + x2, w);
        if (tail) red_shared_add(ro + xt, w);
        if (lane < j) red_shared_add(ro + my_id4, wt);   // (wt = 0 for lanes without a row: their my_id4 is the padding word)
    }
}
constexpr uint32_t kRowPad = 32;   // padding words after every accumulator row (see k_scatter_add)
// Tried and dropped (profiles/r01_scatter_rows_as_lanes.txt): taking the shared ids one at a time with the
// job's ROWS on the lanes (row stride odd modulo 32, so every instruction is one conflict-free wavefront with
// k of 32 lanes busy).  The wavefront count falls by a third on config 2, the time rises (84 -> 112 ms at
// k >= 18): a shared-memory reduction costs ~1 clock per INSTRUCTION plus ~0.7 per wavefront, and this
// form issues 32/k times more instructions.

__global__ void __launch_bounds__(1024)
k_scatter_add(const Unit* __restrict__ units, const uint32_t* __restrict__ n_units_ptr, const Job* __restrict__ jobs,
              const uint32_t* __restrict__ flat_all, const uint64_t* __restrict__ flat_shift, uint32_t* __restrict__ tri,
              uint64_t tri_base, uint32_t id_lo, uint32_t T, uint32_t tile_cols, uint32_t rb_shift, uint32_t* __restrict__ unit_counter) {
    extern __shared__ uint4 tile4[];
    const uint32_t* flat = flat_all + (flat_shift ? *flat_shift : 0ull);
    uint32_t* tile = reinterpret_cast<uint32_t*>(tile4);
    __shared__ uint32_t s_unit, s_next_job;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n_units = *n_units_ptr;
    const uint32_t R = 1u << rb_shift;
    // every accumulator row is followed by 32 padding words: lanes beyond the end of an id slice
    // reduce into them (one bank each), which keeps the inner loop free of branches
    const uint32_t stride = tile_cols + kRowPad;
    const uint32_t tile_saddr = (uint32_t)__cvta_generic_to_shared(tile);
    constexpr uint32_t kNone = 0xFFFFFFFFu;
    uint32_t cur_key = kNone;  // key whose partial sums the tile holds
    for (;;) {
        if (threadIdx.x == 0) s_unit = atomicAdd(unit_counter, 1u);
        __syncthreads();
        const uint32_t u = s_unit;
        const bool done = u >= n_units;
        Unit un; un.key = kNone; un.job_begin = un.job_end = 0; un.pad = 0;
        if (!done) un = units[u];
        if (un.key != cur_key) {
            if (cur_key != kNone) {  // flush the finished tile
                const uint32_t rb = cur_key / T, t = cur_key - rb * T;
                const uint32_t row0 = rb << rb_shift, col0 = win_col0(win_hi(rb, rb_shift), t, tile_cols);
                for (uint32_t r = 0; r < R; ++r) {
                    const uint32_t row = row0 + r;
                    if (row <= col0) continue;
                    const uint32_t nc = min(tile_cols, row - col0);  // only columns < row exist
                    const uint64_t out0 = tri_offset((uint64_t)row + id_lo) - tri_base + col0 + id_lo;   // rows and columns are relative to the sample window
                    const uint32_t* src = tile + r * stride;
                    for (uint32_t c = threadIdx.x; c < nc; c += blockDim.x) {
                        const uint32_t v = src[c];
                        if (v) atomicAdd(&tri[out0 + c], v);
                    }
                }
                __syncthreads();
            }
            if (!done) {
                const uint32_t n4 = (R * stride + 3u) >> 2;
                for (uint32_t c = threadIdx.x; c < n4; c += blockDim.x) tile4[c] = make_uint4(0, 0, 0, 0);
            }
            cur_key = un.key;
        }
        if (done) break;
        if (threadIdx.x == 0) s_next_job = un.job_begin;
        __syncthreads();
        const uint32_t rb = un.key / T, t = un.key - rb * T;
        const uint32_t row0 = rb << rb_shift, col0 = win_col0(win_hi(rb, rb_shift), t, tile_cols);
        // shared address of (row0, column 0); may lie below the tile when col0 > 0 — only in-tile ids occur
        const uint32_t base = tile_saddr - col0 * 4u;
        const uint32_t pad4 = (col0 + tile_cols + lane) * 4u;  // this lane's padding word, relative to `base`
        for (;;) {
            uint32_t jb = 0;
            if (lane == 0) jb = atomicAdd(&s_next_job, kJobBatch);
            jb = __shfl_sync(0xffffffffu, jb, 0);
            if (jb >= un.job_end) break;
            const uint32_t cnt = min(kJobBatch, un.job_end - jb);
            uint4 part = make_uint4(0, 0, 0, 0);  // lane 2q: (off, a, b, A0) of job q; lane 2q+1: (k, w, off_hi, -)
            if (lane < 2 * cnt) part = ldg_nc_v4(reinterpret_cast<const uint4*>(jobs + jb) + lane);
            for (uint32_t q = 0; q < cnt; ++q) {
                const uint32_t off = __shfl_sync(0xffffffffu, part.x, 2 * q), a = __shfl_sync(0xffffffffu, part.y, 2 * q);
                const uint32_t b = __shfl_sync(0xffffffffu, part.z, 2 * q), A0 = __shfl_sync(0xffffffffu, part.w, 2 * q);
                const uint32_t k = __shfl_sync(0xffffffffu, part.x, 2 * q + 1), w = __shfl_sync(0xffffffffu, part.y, 2 * q + 1);
                const uint32_t off_hi = __shfl_sync(0xffffffffu, part.z, 2 * q + 1);
                const uint32_t* list = flat + (((uint64_t)off_hi << 32) | off);
                uint32_t my_id4 = 0, rowoff = 0;  // lane j < k: 4 * id of row j, shared address of its accumulator row
                if (lane < k) {
                    const uint32_t my_row = ldg_nc_u32(list + A0 + lane);
                    my_id4 = my_row * 4u;
                    rowoff = base + (my_row - row0) * stride * 4u;
                }
                // ids seen by every row of the job: [a, min(b, A0)).  Whole 128-id slices first ...
                const uint32_t bc = min(b, A0);
                const uint32_t c_last = a < bc ? a + ((bc - a) & ~127u) : a;
                for (uint32_t c = a; c < c_last; c += 128) {  // full slice: no lane is idle, no predicates
                    const uint32_t* p = list + c + lane;
                    const uint32_t x0 = ldg_nc_u32(p) * 4u, x1 = ldg_nc_u32(p + 32) * 4u;
                    const uint32_t x2 = ldg_nc_u32(p + 64) * 4u, x3 = ldg_nc_u32(p + 96) * 4u;
                    for (uint32_t j = 0; j < k; ++j) {
                        const uint32_t ro = __shfl_sync(0xffffffffu, rowoff, j);
                        red_shared_add(ro + x0, w); red_shared_add(ro + x1, w);
                        red_shared_add(ro + x2, w); red_shared_add(ro + x3, w);
                    }
                }
                // ... then ONE pass over the rows for what is left: the ragged last slice (0..127 ids: complete
                // 32-id groups reduce unconditionally, the partial group under a lane predicate — an idle lane
                // parked on a padding word would still occupy a bank) together with the triangular part (row j
                // also receives the rows before it that lie in [a, b)).  Most jobs are nothing but this pass
                // (a parent's list is ~70 ids), so it is specialised on the number of complete groups.
                {
                    const uint32_t rem = a < bc ? bc - c_last : 0u;
                    const uint32_t* p = list + c_last + lane;
                    const uint32_t full = rem >> 5;           // complete 32-id groups (0..3)
                    const bool tail = lane < (rem & 31u);      // this lane has an id in the partial group
                    uint32_t x0 = pad4, x1 = pad4, x2 = pad4, x3 = pad4;
                    if (lane < rem) x0 = ldg_nc_u32(p) * 4u;
                    if (lane + 32 < rem) x1 = ldg_nc_u32(p + 32) * 4u;
                    if (lane + 64 < rem) x2 = ldg_nc_u32(p + 64) * 4u;
                    if (lane + 96 < rem) x3 = ldg_nc_u32(p + 96) * 4u;
                    const bool tv = lane < k && A0 + lane >= a && A0 + lane < b;
                    const uint32_t wt = tv ? w : 0u;              // weight of this lane's row id as a column of later rows
                    const uint32_t tcol = tv ? my_id4 : pad4;     // (lanes without one add 0 to their padding word)
                    const bool has_tail = (rem & 31u) != 0;
                    switch (full) {
                        case 0: last_rows<0>(k, rowoff, x0, x1, x2, x0, tail, has_tail, wt, tcol, lane, w); break;
                        case 1: last_rows<1>(k, rowoff, x0, x1, x2, x1, tail, has_tail, wt, tcol, lane, w); break;
                        case 2: last_rows<2>(k, rowoff, x0, x1, x2, x2, tail, has_tail, wt, tcol, lane, w); break;
                        default: last_rows<3>(k, rowoff, x0, x1, x2, x3, tail, has_tail, wt, tcol, lane, w); break;
                    }
                }
            }
        }
        __syncthreads();
    }
}

#include "diff.cuh"

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; bytes = 0; }
        const size_t want = need + need / 16 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

thread_local std::string g_open_error;   // kdbx_open may run on several host threads at once (one per device)

}  // namespace

struct kdbx_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t up_stream = nullptr;     // second H2D stream: the payload travels while the scans run
    cudaEvent_t ev_up_begin = nullptr, ev_up_hdr = nullptr, ev_up_payload = nullptr;
    bool upload_pending = false;          // KDBX_FLAG_ASYNC_UPLOAD: copies may still be in flight
    // the payload travels in chunks (asynchronous uploads of a densely packed payload): chunk k = words
    // [up_bounds[k], up_bounds[k+1]), complete when ev_up_chunk[k] is; the decoder starts on a chunk as soon as it is there
    std::vector<uint64_t> up_bounds;
    std::vector<cudaEvent_t> ev_up_chunk;
    kdbx_config cfg{};
    std::string err;

    // staged trie (raw, as uploaded)
    uint64_t P = 0;
    uint32_t N = 0;
    uint64_t payload_words = 0;
    bool loaded = false;
    bool dense_payload = false;
    DevBuf num_kmers, parent, n, l, last, bits, poff, payload, hdr32;
    float ms_upload = 0.f;

    // prepared
    DevBuf nodes, W, loc, loff, noff, coff, bounds, err_flag, cub_tmp, order_in, order, keys_sorted, level_start;
    std::vector<uint32_t> h_level_start;
    std::vector<std::pair<uint32_t, uint32_t>> levels;  // ranges of `order` per distinct num_samples, ascending
    // per chunk
    DevBuf flat, jobs, hist, work, bucket_off, cursor, ucount, uoff, units, counters, blockhist;
    DevBuf tri, rowupd, first_id;
    uint64_t sum_l = 0, sum_n = 0, sum_cost = 0;
    bool prepared = false;  // nodes / W / loc are valid for the loaded trie

    // Sample window [win_lo, win_hi): every sample id of the staged trie lies inside it (kdbx_set_sample_window;
    // default [0, N)).  The dense kernels then work on ids relative to win_lo, so the shard of a large database
    // whose samples sit close together is planned like a database of win_hi - win_lo samples.
    uint32_t win_lo = 0, win_hi = 0;

    // What the host has to know about the staged trie to enqueue a step — sizes, the level table, the plan
    // of the last call — is read back ONCE per staged trie (load_gen) and call signature.  Later calls with
    // the same signature recompute everything on the device but never wait for it: one synchronisation at the
    // end of the call, none inside.
    uint64_t load_gen = 0;
    uint64_t no_slide_gen = ~0ull;   // load_gen of a staged trie whose lists do not fit the sliding column window (make_plan)
    struct StepMeta {
        bool prep_valid = false;   // sum_l / sum_n / sum_cost / levels belong to load_gen
        uint64_t prep_gen = 0;
        bool valid = false;
        uint64_t gen = 0;
        uint32_t row_begin = 0, row_end = 0, part = 0, num_parts = 0, win_lo = 0, win_hi = 0;
        uint32_t tile_cols = 0, tile_rows = 0, threads = 0, unit_updates = 0, flags = 0;
        uint64_t chunk = 0;
        bool resident = false, diff = false, out_aligned = true;
        uint32_t nchunks = 0;
        unsigned scatter_grid = 0;
        std::vector<uint64_t> bounds;     // chunk boundaries (chunked mode)
        std::vector<uint64_t> jobs_in_pass; // exact job counts of the passes executed
        uint64_t diff_slots = 0;          // entries of the boundary lists (diff mode)
    } meta;
    // boundary lists (diff.cuh)
    DevBuf ownb, nb, boff;
    // multi-GPU (comm.cuh): the NCCL communicator this context is a rank of, and its block of the result
    void* comm = nullptr;
    int comm_nranks = 1, comm_rank = 0;
    DevBuf rs_block;
    // dense table text (csvfmt.cuh): which rows of the matrix ctx->tri holds after a host-output all2all call
    bool tri_rows_valid = false;
    uint32_t tri_row_begin = 0, tri_row_end = 0;
    DevBuf csv_text;
    float ms_csv = 0.f;
    uint64_t* h_pinned = nullptr;          // small pinned read-back area

    // sparse delivery (sparse.cuh)
    DevBuf sp_cnt, sp_counts, sp_rowptr, sp_col, sp_val;
    // k-mer tables + query buffers (query.cuh)
    bool tables_loaded = false;
    uint64_t num_tables = 0, total_slots = 0;
    float ms_upload_tables = 0.f;
    DevBuf slot_off, slots, q_off, q_kmers, q_keys, q_keys2, q_runkeys, q_runcnt, q_out;
    DevBuf qx_alpha, qx_seq, qx_raw, qx_sorted, qx_count;   // queries from sequences (build.cuh)

    std::vector<cudaEvent_t> events;
    size_t ev_used = 0;

    // The W accumulation and the level-order expansion are one tiny launch per num_samples level
    // (hundreds per step): they are captured once per staged trie into CUDA graphs and replayed, so
    // the step does not depend on how fast the host can issue launches.
    struct LevelGraph {
        cudaGraphExec_t exec = nullptr;
        const void* ptr[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        uint64_t P = 0, levels_hash = 0;
        size_t nlevels = 0;
    } g_push, g_expand, g_pull, g_expand_diff;

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        err = buf;
        return code;
    }
    cudaEvent_t event() {
        if (ev_used == events.size()) { cudaEvent_t e; cudaEventCreate(&e); events.push_back(e); }
        cudaEvent_t e = events[ev_used++];
        cudaEventRecord(e, stream);
        return e;
    }
};

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return ctx->fail(e__ == cudaErrorMemoryAllocation ? KDBX_ERR_NOMEM : KDBX_ERR_CUDA, \
                             "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

namespace {

inline unsigned blocks_for(uint64_t items, unsigned threads) { return (unsigned)((items + threads - 1) / threads); }

template <class InIt>
int scan_exclusive(kdbx_ctx* ctx, InIt in, uint64_t* out, uint64_t count) {
    size_t tmp = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, count, ctx->stream));
    CK(ctx->cub_tmp.ensure(tmp));
    CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, in, out, count, ctx->stream));
    return KDBX_OK;
}
int scan_exclusive_u32(kdbx_ctx* ctx, const uint32_t* in, uint32_t* out, uint64_t count) {
    size_t tmp = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, count, ctx->stream));
    CK(ctx->cub_tmp.ensure(tmp));
    CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, in, out, count, ctx->stream));
    return KDBX_OK;
}

struct Plan {
    uint32_t tile_cols = 0, tile_rows = 0, rb_shift = 0, T = 0, RB = 0, unit_updates = 0, threads = 0;
    uint32_t lo = 0, Nw = 0;     // sample window: ids lo .. lo + Nw, the kernels see ids relative to lo
    bool slide = false;          // one SLIDING column window per row block (see make_plan)
    uint64_t chunk = 0;
    size_t smem = 0, smem_diff = 0;
};

constexpr size_t kMaxTileBytes = 220 * 1024;  // of the 227 KB a CTA may use (the boundary-form kernel adds 4 KB of static shared memory): 32 rows x 1728 columns
constexpr uint32_t kMaxOneWindowCols = 1728;  // the widest tile of 32 rows (one column window holding whole rows)

// allow_slide: beyond kMaxOneWindowCols samples, try ONE column window per row block that ends right above the block's rows
// and covers the kMaxOneWindowCols columns below them (col0 = block end - tile_cols).  It holds every list entry a row of
// the block can receive iff no pattern's full list reaches further down than that — true for databases whose related
// samples sit close together in sample order (clusters, clades, the parts of a sharded database); k_pull_level checks it
// per pattern, and a database that fails is re-planned with 1024-column windows and the id form.  With it the headline
// path (decoder job counts, boundary lists, no second enumeration) is not limited to 1728 samples but to
// 32 * kDecodeHistKeys.
int make_plan(kdbx_ctx* ctx, Plan& pl, bool allow_slide = false) {
    pl.lo = ctx->win_lo;
    const uint32_t N = ctx->win_hi - ctx->win_lo;
    pl.Nw = N;
    uint32_t tc = ctx->cfg.tile_cols;
    // default: whole rows (one column window) up to 1728 samples; beyond, a sliding window of that width if allowed,
    // else 1024-column windows
    pl.slide = false;
    if (tc == 0) {
        if (N <= kMaxOneWindowCols) tc = std::max<uint32_t>(32u, (N + 31u) & ~31u);
        else if (allow_slide && N <= 32u * kDecodeHistKeys && ctx->cfg.tile_rows == 0) { tc = kMaxOneWindowCols; pl.slide = true; }
        else tc = 1024u;
    }
    if (tc < 32 || (tc & 31)) return ctx->fail(KDBX_ERR_ARG, "tile_cols must be a multiple of 32");
    uint32_t tr = ctx->cfg.tile_rows;
    if (tr == 0) {
        tr = 32;
        while (tr > 1 && (size_t)tr * (tc + kRowPad) * 4 > kMaxTileBytes) tr >>= 1;
    }
    if (tr == 0 || tr > 32 || (tr & (tr - 1))) return ctx->fail(KDBX_ERR_ARG, "tile_rows must be a power of two <= 32");
    if ((size_t)tr * (tc + kRowPad) * 4 > kMaxTileBytes) return ctx->fail(KDBX_ERR_ARG, "tile_rows x tile_cols too large for shared memory");
    pl.tile_cols = tc; pl.tile_rows = tr;
    pl.rb_shift = 0;
    while ((1u << pl.rb_shift) < tr) ++pl.rb_shift;
    pl.RB = N == 0 ? 1 : (N + tr - 1) / tr;
    {   // windows per row block: enough to reach column 0 from the end of the last block (see win_hi)
        const uint64_t hi_max = ((uint64_t)pl.RB * tr + 31u) & ~(uint64_t)31u;
        pl.T = (uint32_t)std::max<uint64_t>(1, (hi_max + tc - 1) / tc);
        if (pl.slide) pl.T = 1;
    }
    if ((uint64_t)pl.T * pl.RB >= ((uint64_t)1 << 31)) return ctx->fail(KDBX_ERR_ARG, "too many (row block, column tile) keys");
    pl.smem = ((size_t)tr * (tc + kRowPad) * 4 + 15) & ~(size_t)15;  // padding words per row, see k_scatter_add
    pl.smem_diff = ((size_t)tr * tc * 4 + 15) & ~(size_t)15;         // no padding in the boundary form
    pl.threads = ctx->cfg.scatter_threads ? ctx->cfg.scatter_threads : (pl.smem > 100 * 1024 ? 1024u : pl.smem > 48 * 1024 ? 512u : 256u);
    if (pl.threads < 32 || pl.threads > 1024 || (pl.threads & 31)) return ctx->fail(KDBX_ERR_ARG, "scatter_threads must be a multiple of 32 in [32, 1024]");
    // a unit flushes at most tile_rows x tile_cols cells with global reductions: keep >= 64 updates per cell
    pl.unit_updates = ctx->cfg.unit_updates ? ctx->cfg.unit_updates : std::max<uint32_t>(131072u, 64u * tr * tc);
    pl.chunk = ctx->cfg.chunk_ids ? ctx->cfg.chunk_ids : ((uint64_t)64 << 20);
    if (pl.chunk < 4096) pl.chunk = 4096;
    if (pl.chunk > ((uint64_t)1 << 31)) pl.chunk = (uint64_t)1 << 31;
    return KDBX_OK;
}

int error_from_flag(kdbx_ctx* ctx, int flag) {
    if (flag) { ctx->prepared = false; ctx->meta.valid = false; ctx->meta.prep_valid = false; }   // nothing decoded is usable: the next call starts over
    if (flag == 1) return ctx->fail(KDBX_ERR_ARG, "malformed trie: Elias-gamma stream does not match num_bits");
    if (flag == 2) return ctx->fail(KDBX_ERR_ARG, "malformed trie: sample id out of range");
    if (flag == 3) return ctx->fail(KDBX_ERR_ARG, "malformed trie: a local list does not continue its parent's list");
    if (flag == 4) return ctx->fail(KDBX_ERR_ARG, "malformed trie: parent_id / num_samples / num_local_samples inconsistent");
    if (flag == 5) return ctx->fail(KDBX_ERR_ARG, "malformed trie: payload offset out of bounds");
    if (flag == 6) return ctx->fail(KDBX_ERR_ARG, "k-mer table points at a pattern that does not exist");
    if (flag == 7) return ctx->fail(KDBX_ERR_ARG, "a sample id lies outside the declared sample window");
    if (flag == 8) return ctx->fail(KDBX_ERR_STATE, "internal: boundary lists exceed their buffer");
    if (flag == 9) return ctx->fail(KDBX_ERR_STATE, "internal: a sample list reaches below its sliding column window");
    return KDBX_OK;
}
int check_device_error(kdbx_ctx* ctx) {
    int flag = 0;
    CK(cudaMemcpy(&flag, ctx->err_flag.p, sizeof flag, cudaMemcpyDeviceToHost));
    return error_from_flag(ctx, flag);
}

// entry points that hand decoded sample ids to the caller (or index per-sample arrays with them) need ids that
// are not shifted by a sample window
int require_full_window(kdbx_ctx* ctx, const char* what) {
    if (ctx->win_lo != 0 || ctx->win_hi != ctx->N)
        return ctx->fail(KDBX_ERR_STATE, "%s: a sample window is set (kdbx_set_sample_window); only the dense and sparse all2all honour it", what);
    return KDBX_OK;
}

uint64_t levels_hash_of(const std::vector<std::pair<uint32_t, uint32_t>>& levels) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (const auto& lv : levels) { h ^= ((uint64_t)lv.first << 32) | lv.second; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 29; }
    return h;
}

// Replays `record` (kernel launches on ctx->stream only) from a graph.  The first call for a given
// staged trie launches directly (a one-shot CLI run should not pay for capture and instantiation); the
// second call with the same trie, levels and buffers captures the graph, later calls replay it.
template <class Record>
int launch_level_graph(kdbx_ctx* ctx, kdbx_ctx::LevelGraph& g, std::initializer_list<const void*> ptrs, Record&& record) {
    kdbx_ctx::LevelGraph want;
    size_t i = 0;
    for (const void* p : ptrs) want.ptr[i++] = p;
    want.P = ctx->P; want.nlevels = ctx->levels.size(); want.levels_hash = levels_hash_of(ctx->levels);
    const bool same = g.P == want.P && g.nlevels == want.nlevels && g.levels_hash == want.levels_hash &&
                      std::memcmp(g.ptr, want.ptr, sizeof want.ptr) == 0;
    if (!same) {
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        g = want;          // remember what was launched; exec stays empty
        record();
        CK(cudaGetLastError());
        return KDBX_OK;
    }
    if (!g.exec) {
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        record();
        CK(cudaStreamEndCapture(ctx->stream, &graph));
        cudaGraphExec_t exec = nullptr;
        const cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return ctx->fail(KDBX_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
        g.exec = exec;
    }
    CK(cudaGraphLaunch(g.exec, ctx->stream));
    return KDBX_OK;
}

// scans + node packing + W + gamma decode.  The sizes and the level table reach the host once per staged trie
// (one synchronisation); later calls re-run every kernel but take those from the context.
int prepare(kdbx_ctx* ctx, const Plan& pl, uint32_t& launches, const DecodeHist* decode_hist = nullptr) {
    const uint64_t P = ctx->P;
    cudaStream_t st = ctx->stream;
    const bool cached = ctx->meta.prep_valid && ctx->meta.prep_gen == ctx->load_gen;
    CK(ctx->loff.ensure((P + 1) * 8)); CK(ctx->noff.ensure((P + 1) * 8)); CK(ctx->coff.ensure((P + 1) * 8));
    CK(ctx->nodes.ensure(P * sizeof(Node))); CK(ctx->W.ensure(P * 4)); CK(ctx->err_flag.ensure(16));
    cub::CountingInputIterator<uint64_t> idx(0);
    {
        cub::TransformInputIterator<uint64_t, U32AsU64, cub::CountingInputIterator<uint64_t>> it(idx, U32AsU64{ctx->l.as<uint32_t>(), P});
        if (int rc = scan_exclusive(ctx, it, ctx->loff.as<uint64_t>(), P + 1)) return rc;
    }
    {
        cub::TransformInputIterator<uint64_t, ListSlots, cub::CountingInputIterator<uint64_t>> it(idx, ListSlots{ctx->n.as<uint32_t>(), P});
        if (int rc = scan_exclusive(ctx, it, ctx->noff.as<uint64_t>(), P + 1)) return rc;
    }
    {
        ChunkCost cc{ctx->n.as<uint32_t>(), ctx->l.as<uint32_t>(), ctx->last.as<uint32_t>(), P, pl.tile_cols};
        cub::TransformInputIterator<uint64_t, ChunkCost, cub::CountingInputIterator<uint64_t>> it(idx, cc);
        if (int rc = scan_exclusive(ctx, it, ctx->coff.as<uint64_t>(), P + 1)) return rc;
    }
    launches += 3;
    if (ctx->dense_payload) {
        CK(ctx->poff.ensure((P + 1) * 8));
        cub::TransformInputIterator<uint64_t, PayloadWords, cub::CountingInputIterator<uint64_t>> it(idx, PayloadWords{ctx->bits.as<uint32_t>(), P});
        if (int rc = scan_exclusive(ctx, it, ctx->poff.as<uint64_t>(), P + 1)) return rc;
        launches += 1;
    }
    CK(cudaMemsetAsync(ctx->err_flag.p, 0, 16, st));
    k_build_nodes<<<blocks_for(P, 256), 256, 0, st>>>(P, ctx->parent.as<int64_t>(), ctx->num_kmers.as<int64_t>(),
                                                       ctx->n.as<uint32_t>(), ctx->l.as<uint32_t>(), ctx->last.as<uint32_t>(),
                                                       ctx->loff.as<uint64_t>(), ctx->nodes.as<Node>(), ctx->W.as<uint32_t>(), ctx->N,
                                                       ctx->err_flag.as<int>());
    launches += 1;
    // order patterns by num_samples, descending (levels of the W accumulation)
    const uint32_t N = ctx->N;
    CK(ctx->order_in.ensure(P * 4)); CK(ctx->order.ensure(P * 4)); CK(ctx->keys_sorted.ensure(P * 4));
    CK(ctx->level_start.ensure(((size_t)N + 2) * 4));
    {
        int end_bit = 1;
        while (end_bit < 32 && (N >> end_bit)) ++end_bit;
        k_iota<<<blocks_for(P, 256), 256, 0, st>>>(P, ctx->order_in.as<uint32_t>());
        size_t tmp = 0;
        CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, ctx->n.as<uint32_t>(), ctx->keys_sorted.as<uint32_t>(),
                                                     ctx->order_in.as<uint32_t>(), ctx->order.as<uint32_t>(), P, 0, end_bit, st));
        CK(ctx->cub_tmp.ensure(tmp));
        CK(cub::DeviceRadixSort::SortPairsDescending(ctx->cub_tmp.p, tmp, ctx->n.as<uint32_t>(), ctx->keys_sorted.as<uint32_t>(),
                                                     ctx->order_in.as<uint32_t>(), ctx->order.as<uint32_t>(), P, 0, end_bit, st));
        launches += 2;
    }
    if (!cached) {
        ctx->h_level_start.resize((size_t)N + 2);
        CK(cudaMemsetAsync(ctx->level_start.p, 0xFF, ((size_t)N + 2) * 4, st));
        k_level_starts<<<blocks_for(P, 256), 256, 0, st>>>(P, N, ctx->keys_sorted.as<uint32_t>(), ctx->level_start.as<uint32_t>());
        launches += 2;
        uint64_t sums[3] = {0, 0, 0};
        CK(cudaMemcpyAsync(&sums[0], ctx->loff.as<uint64_t>() + P, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&sums[1], ctx->noff.as<uint64_t>() + P, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&sums[2], ctx->coff.as<uint64_t>() + P, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ctx->h_level_start.data(), ctx->level_start.p, ((size_t)N + 2) * 4, cudaMemcpyDeviceToHost, st));
        int h_flag = 0;
        CK(cudaMemcpyAsync(&h_flag, ctx->err_flag.p, sizeof h_flag, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (int rc = error_from_flag(ctx, h_flag)) return rc;  // structural errors stop here, before any id is chased
        ctx->sum_l = sums[0]; ctx->sum_n = sums[1]; ctx->sum_cost = sums[2];
        // deepest level first; level 0 (the sentinel, n = 0) has no parent to feed
        uint32_t end = (uint32_t)P;
        std::vector<std::pair<uint32_t, uint32_t>>& levels = ctx->levels;  // ascending n; ranges in `order`
        levels.clear();
        for (uint32_t v = 0; v <= N; ++v) {
            const uint32_t b = ctx->h_level_start[v];
            if (b == 0xFFFFFFFFu) continue;
            levels.emplace_back(b, end);   // keys are sorted descending: smaller n sits further back
            end = b;
        }
        ctx->meta.prep_valid = true; ctx->meta.prep_gen = ctx->load_gen;
    }
    {
        const std::vector<std::pair<uint32_t, uint32_t>>& levels = ctx->levels;
        // levels[k] = (begin of key v_k, end); built for ascending v, so begin decreases
        uint32_t n_launch = 0;
        for (size_t k = levels.size(); k-- > 0;) if (levels[k].second > levels[k].first) ++n_launch;
        if (n_launch) {
            if (int rc = launch_level_graph(ctx, ctx->g_push, {ctx->order.p, ctx->parent.p, ctx->W.p}, [&] {
                    for (size_t k = levels.size(); k-- > 0;) {
                        const uint32_t b = levels[k].first, e = levels[k].second;
                        if (e <= b) continue;
                        k_push_level<<<blocks_for(e - b, 256), 256, 0, st>>>(e - b, ctx->order.as<uint32_t>() + b, ctx->parent.as<int64_t>(),
                                                                              ctx->W.as<uint32_t>());
                    }
                })) return rc;
        }
        launches += n_launch;
    }
    CK(ctx->loc.ensure((ctx->sum_l + 32) * 4));
    DecodeHist dh{};
    if (decode_hist) { dh = *decode_hist; dh.W = ctx->W.as<uint32_t>(); }
    auto decode = [&](uint64_t w_lo, uint64_t w_hi) {
        k_decode_locals<<<blocks_for(P, kDecodeThreads), kDecodeThreads, 0, st>>>(P, ctx->nodes.as<Node>(), ctx->loff.as<uint64_t>(), ctx->bits.as<uint32_t>(), ctx->poff.as<uint64_t>(),
                                                             ctx->payload.as<uint64_t>(), ctx->payload_words, ctx->loc.as<uint32_t>(), ctx->win_lo,
                                                             ctx->win_hi - ctx->win_lo, ctx->err_flag.as<int>(), dh, w_lo, w_hi);
        launches += 1;
    };
    // the payload may still be on its way (second H2D stream).  When it travels in chunks, one launch per chunk takes
    // the blocks whose payload that chunk completes: decoding overlaps the rest of the transfer
    const size_t K = ctx->up_bounds.size() > 1 ? ctx->up_bounds.size() - 1 : 1;
    bool chunked = ctx->upload_pending && ctx->dense_payload && K > 1 && ctx->ev_up_chunk.size() >= K;
    if (chunked && cudaEventQuery(ctx->ev_up_payload) == cudaSuccess) chunked = false;   // (everything has arrived already)
    cudaGetLastError();   // (cudaErrorNotReady of the query is not an error)
    if (!chunked) {
        CK(cudaStreamWaitEvent(st, ctx->ev_up_payload, 0));
        decode(0, ~0ull);
    } else {
        for (size_t k = 0; k < K; ++k) {
            CK(cudaStreamWaitEvent(st, k + 1 == K ? ctx->ev_up_payload : ctx->ev_up_chunk[k], 0));
            decode(k == 0 ? 0 : ctx->up_bounds[k], k + 1 == K ? ~0ull : ctx->up_bounds[k + 1]);
        }
    }
    CK(cudaGetLastError());
    // chunk boundaries
    const uint64_t total_cost = ctx->sum_cost;
    const uint32_t nchunks = (uint32_t)std::max<uint64_t>(1, (total_cost + pl.chunk - 1) / pl.chunk);
    CK(ctx->bounds.ensure(((size_t)nchunks + 1) * 8));
    k_chunk_bounds<<<blocks_for(nchunks + 1, 128), 128, 0, st>>>(P, ctx->coff.as<uint64_t>(), pl.chunk, nchunks, ctx->bounds.as<uint64_t>());
    launches += 1;
    CK(cudaGetLastError());
    ctx->prepared = true;
    return (int)nchunks;
}

float elapsed(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }

// Waits for the copies of kdbx_load_patterns and records their duration.
int finish_upload(kdbx_ctx* ctx) {
    if (!ctx->upload_pending) return KDBX_OK;
    CK(cudaEventSynchronize(ctx->ev_up_hdr));
    CK(cudaEventSynchronize(ctx->ev_up_payload));
    float hdr = 0.f, all = 0.f;
    cudaEventElapsedTime(&hdr, ctx->ev_up_begin, ctx->ev_up_hdr);
    cudaEventElapsedTime(&all, ctx->ev_up_begin, ctx->ev_up_payload);
    ctx->ms_upload = std::max(hdr, all);
    ctx->upload_pending = false;
    return KDBX_OK;
}

// One dense all2all over the staged trie: rows [row_begin, row_end) of the packed triangle into d_out.
// part / num_parts: only the chunks c with c % num_parts == part are executed (pattern sharding: the partial
// matrices of all parts sum to the full one).
//
// Host round trips.  The FIRST call for a staged trie and call signature reads back what the host needs to size
// buffers and launches (sums and the level table in prepare(), then the job / list totals and the choice of the
// list form) — three synchronisations.  It leaves them in ctx->meta; every later call with the same signature
// enqueues the whole step without waiting for the device and synchronises once, at the end, for U and the error
// flag.  Nothing of the RESULT is cached: every call decodes, expands, buckets and scatters again.
constexpr int kRetryWithoutSlide = 1;   // internal: the sliding column window does not hold this database's lists

int all2all_rows_device_impl(kdbx_ctx* ctx, uint32_t row_begin, uint32_t row_end, uint32_t* d_out, kdbx_stats* stats,
                             uint32_t part, uint32_t num_parts, bool allow_slide) {
    if (!ctx->loaded) return ctx->fail(KDBX_ERR_STATE, "no patterns loaded (call kdbx_load_patterns first)");
    if (row_begin > row_end || row_end > ctx->N) return ctx->fail(KDBX_ERR_ARG, "bad row range [%u,%u) for %u samples", row_begin, row_end, ctx->N);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Plan pl;
    {   // the sliding window serves the headline path only (all rows of the window, one part, boundary lists, aligned output)
        const bool all_rows0 = row_begin <= ctx->win_lo && row_end >= ctx->win_hi;
        const bool slide_ok = allow_slide && all_rows0 && num_parts == 1 && !(ctx->cfg.flags & (KDBX_FLAG_CHUNKED_LISTS | KDBX_FLAG_ID_LISTS)) &&
                              (reinterpret_cast<uintptr_t>(d_out) & 15u) == 0;
        if (int rc = make_plan(ctx, pl, slide_ok)) return rc;
    }
    kdbx_ctx::StepMeta& M = ctx->meta;
    const bool cached = M.valid && M.gen == ctx->load_gen && M.row_begin == row_begin && M.row_end == row_end && M.part == part &&
                        M.num_parts == num_parts && M.win_lo == ctx->win_lo && M.win_hi == ctx->win_hi && M.tile_cols == pl.tile_cols &&
                        M.tile_rows == pl.tile_rows && M.threads == pl.threads && M.unit_updates == pl.unit_updates &&
                        M.flags == ctx->cfg.flags && M.chunk == pl.chunk &&
                        M.out_aligned == ((reinterpret_cast<uintptr_t>(d_out) & 15u) == 0);   // (the boundary form needs an aligned output)
    kdbx_stats s{};
    s.ms_upload = ctx->ms_upload;
    ctx->ev_used = 0;
    uint32_t launches = 0;
    auto tri_off = [](uint64_t r) { return r == 0 ? 0ull : r * (r - 1) / 2; };
    const uint64_t tri_base = tri_off(row_begin);
    const uint64_t cells = tri_off(row_end) - tri_base;
    // the rows that can hold anything lie inside the sample window; the kernels see them relative to it
    const uint32_t wb = std::min(std::max(row_begin, ctx->win_lo), ctx->win_hi) - ctx->win_lo;
    const uint32_t we = std::max(std::min(row_end, ctx->win_hi), ctx->win_lo) - ctx->win_lo;
    const bool all_rows = row_begin <= ctx->win_lo && row_end >= ctx->win_hi;

    cudaEvent_t ev_start = ctx->event();
    if (cells) CK(cudaMemsetAsync(d_out, 0, cells * 4, st));
    const uint32_t nkeys = pl.RB * pl.T;
    const unsigned wide_grid = (unsigned)(ctx->sm_count * 8);
    const bool smem_buckets = nkeys <= kSmemKeys;
    // patterns per block of the bucketing passes: a multiple of the decoder's block, so that the decoder can
    // count jobs per (fill block, key) itself (DecodeHist)
    const uint64_t per_block = (((ctx->P + wide_grid - 1) / wide_grid) + kDecodeThreads - 1) / kDecodeThreads * kDecodeThreads;
    CK(ctx->counters.ensure(64));
    if (!ctx->h_pinned) CK(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_pinned), 64, cudaHostAllocDefault));
    unsigned long long* d_total_updates = ctx->counters.as<unsigned long long>();
    uint32_t* d_unit_counter = ctx->counters.as<uint32_t>() + 4;
    unsigned long long* d_sum_app = ctx->counters.as<unsigned long long>() + 3;
    unsigned long long* d_physical = ctx->counters.as<unsigned long long>() + 4;
    CK(cudaMemsetAsync(ctx->counters.p, 0, 64, st));
    // one column window, all rows, one part: the decoder counts the jobs and the first enumeration pass is skipped
    bool hist_by_decoder = pl.T == 1 && nkeys <= kDecodeHistKeys && nkeys > 0 && cells > 0 && all_rows &&
                           num_parts == 1 && !(ctx->cfg.flags & KDBX_FLAG_CHUNKED_LISTS);
    if (cached && !M.resident) hist_by_decoder = false;
    // (the boundary form flushes its tiles with 16-byte bulk reductions: the output must be aligned accordingly)
    const bool diff_possible = hist_by_decoder && !(ctx->cfg.flags & KDBX_FLAG_ID_LISTS) && (reinterpret_cast<uintptr_t>(d_out) & 15u) == 0;
    DecodeHist dh{};
    if (hist_by_decoder) {
        CK(ctx->blockhist.ensure((size_t)wide_grid * nkeys * 4 + 16));
        CK(ctx->work.ensure(((size_t)nkeys + 1) * 8));
        CK(cudaMemsetAsync(ctx->blockhist.p, 0, (size_t)wide_grid * nkeys * 4, st));
        CK(cudaMemsetAsync(ctx->work.p, 0, ((size_t)nkeys + 1) * 8, st));
        dh.enabled = 1; dh.rb_shift = pl.rb_shift; dh.nkeys = nkeys; dh.per = per_block;
        dh.blockhist = ctx->blockhist.as<uint32_t>(); dh.work = ctx->work.as<unsigned long long>(); dh.total_updates = d_total_updates;
        if (diff_possible && (!cached || M.diff)) {
            CK(ctx->ownb.ensure(ctx->P * 4));
            dh.ownb = ctx->ownb.as<uint32_t>(); dh.sum_app = d_sum_app;
        }
    }
    const int nchunks_or_err = prepare(ctx, pl, launches, hist_by_decoder ? &dh : nullptr);
    if (nchunks_or_err < 0) return nchunks_or_err;
    const uint32_t nchunks = (uint32_t)nchunks_or_err;
    std::vector<uint64_t> bounds(nchunks + 1);
    cudaEvent_t ev_prepared = ctx->event();

    if (cells == 0 || nkeys == 0 || we <= wb) {  // no cell of the requested rows can be non-zero
        CK(cudaStreamSynchronize(st));
        s.ms_prepare = elapsed(ev_start, ev_prepared);
        s.ms_total = s.ms_prepare;
        s.flat_ids = ctx->sum_n; s.local_ids = ctx->sum_l; s.kernel_launches = launches;
        if (int rc = check_device_error(ctx)) return rc;
        if (stats) *stats = s;
        return KDBX_OK;
    }
    const uint64_t cap = pl.chunk + (uint64_t)pl.Nw * (pl.T + 1) + 64;  // a chunk may overshoot by one pattern
    // full lists of all patterns resident (level-order expansion) when they take at most 40 % of the
    // free HBM; otherwise they are expanded chunk by chunk by walking parent chains
    // (a pattern-sharded part needs only its own chunks' lists: walking their chains is cheaper than
    // expanding everything on every GPU)
    bool resident = !(ctx->cfg.flags & KDBX_FLAG_CHUNKED_LISTS) && num_parts == 1;
    if (cached) resident = M.resident;
    else if (resident) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t have = (uint64_t)free_b + ctx->flat.bytes;
        resident = (ctx->sum_n + 64) * 4 <= have / 5 * 2;
    }
    CK(ctx->hist.ensure(((size_t)nkeys + 1) * 4)); CK(ctx->work.ensure(((size_t)nkeys + 1) * 8));
    CK(ctx->bucket_off.ensure(((size_t)nkeys + 1) * 4)); CK(ctx->cursor.ensure(((size_t)nkeys + 1) * 4));
    CK(ctx->ucount.ensure(((size_t)nkeys + 1) * 4)); CK(ctx->uoff.ensure(((size_t)nkeys + 1) * 4));
    if (hist_by_decoder && !resident) {   // the lists are streamed after all: count per chunk as usual
        hist_by_decoder = false;
        CK(cudaMemsetAsync(ctx->counters.p, 0, 64, st));
    }

    // ---- the decoder counted the jobs: bucket offsets now; first call: totals and the list form to the host ----
    bool diff = false;
    uint64_t diff_slots = 0, jobs_total = 0;
    cudaEvent_t ev_bucket0 = nullptr, ev_bucket0_end = nullptr;
    if (hist_by_decoder) {
        ev_bucket0 = ctx->event();
        k_key_totals<<<blocks_for((uint64_t)nkeys + 1, 128), 128, 0, st>>>(nkeys, wide_grid, ctx->blockhist.as<uint32_t>(), ctx->hist.as<uint32_t>());
        if (int rc = scan_exclusive_u32(ctx, ctx->hist.as<uint32_t>(), ctx->bucket_off.as<uint32_t>(), (uint64_t)nkeys + 1)) return rc;
        k_block_offsets<<<blocks_for((uint64_t)nkeys * 32, 256), 256, 0, st>>>(nkeys, wide_grid, ctx->bucket_off.as<uint32_t>(), ctx->blockhist.as<uint32_t>());
        launches += 3;
        const bool want_nb = dh.ownb != nullptr;
        if (want_nb) {   // entries per boundary list (parents first) and their offsets
            CK(ctx->nb.ensure(ctx->P * 4)); CK(ctx->boff.ensure((ctx->P + 1) * 8)); CK(ctx->first_id.ensure(ctx->P * 4));
            uint32_t n_launch = 0;
            for (const auto& lv : ctx->levels) if (lv.second > lv.first) ++n_launch;
            if (n_launch) {
                if (int rc = launch_level_graph(ctx, ctx->g_pull, {ctx->order.p, ctx->nodes.p, ctx->ownb.p, ctx->nb.p, ctx->loc.p, ctx->first_id.p,
                                                                  reinterpret_cast<const void*>((uintptr_t)(pl.slide ? pl.tile_cols : 0u) | ((uintptr_t)ctx->win_lo << 32))}, [&] {
                        for (size_t k = 0; k < ctx->levels.size(); ++k) {
                            const uint32_t b = ctx->levels[k].first, e = ctx->levels[k].second;
                            if (e <= b) continue;
                            k_pull_level<<<blocks_for(e - b, 256), 256, 0, st>>>(e - b, ctx->order.as<uint32_t>() + b, ctx->nodes.as<Node>(),
                                                                                  ctx->ownb.as<uint32_t>(), ctx->nb.as<uint32_t>(), ctx->loc.as<uint32_t>(),
                                                                                  ctx->first_id.as<uint32_t>(), pl.slide ? pl.tile_cols : 0u, ctx->win_lo,
                                                                                  pl.rb_shift, ctx->err_flag.as<int>());
                        }
                    })) return rc;
            }
            launches += n_launch;
            cub::CountingInputIterator<uint64_t> idx(0);
            cub::TransformInputIterator<uint64_t, BoundSlots, cub::CountingInputIterator<uint64_t>> it(idx, BoundSlots{ctx->nb.as<uint32_t>(), ctx->P});
            if (int rc = scan_exclusive(ctx, it, ctx->boff.as<uint64_t>(), ctx->P + 1)) return rc;
            launches += 1;
        }
        ev_bucket0_end = ctx->event();
        if (cached) { diff = M.diff; diff_slots = M.diff_slots; jobs_total = M.jobs_in_pass.empty() ? 0 : M.jobs_in_pass[0]; }
        else {
            uint64_t* h = ctx->h_pinned;   // [0] sum of appended entries, [1] boundary slots, [2] jobs, [3] error flag
            h[0] = h[1] = h[2] = h[3] = 0;
            if (want_nb) {
                CK(cudaMemcpyAsync(&h[0], d_sum_app, 8, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(&h[1], ctx->boff.as<uint64_t>() + ctx->P, 8, cudaMemcpyDeviceToHost, st));
            }
            CK(cudaMemcpyAsync(&h[2], ctx->bucket_off.as<uint32_t>() + nkeys, 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(&h[3], ctx->err_flag.p, 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (pl.slide && (uint32_t)h[3] == 9u) {   // a list reaches below its window: plan again without it
                CK(cudaMemsetAsync(ctx->err_flag.p, 0, 16, st));
                return kRetryWithoutSlide;
            }
            if (int rc = error_from_flag(ctx, (int)(uint32_t)h[3])) return rc;   // decode errors stop the call before any list is built
            jobs_total = (uint32_t)h[2];
            diff_slots = h[1];
            // boundary form when it holds at most 3/4 of the entries of the id form (a run of consecutive ids costs 2
            // entries whatever its length; a lone id costs 2 instead of 1), or when the caller insists
            diff = want_nb && (((ctx->cfg.flags & KDBX_FLAG_BOUNDARY_LISTS) != 0) || h[0] * 4 <= ctx->sum_l * 3);
            if (pl.slide && !diff) return kRetryWithoutSlide;   // (the sliding window is built for the boundary form only)
        }
    }
    if (pl.slide && !hist_by_decoder) return kRetryWithoutSlide;   // (lists not resident after all)
    CK(ctx->flat.ensure(diff ? (diff_slots + 64) * 4 : resident ? (ctx->sum_n + 64) * 4 : cap * 4));
    if (resident && !diff) CK(ctx->first_id.ensure(ctx->P * 4));
    if (!resident) { CK(ctx->jobs.ensure(cap * sizeof(Job))); CK(ctx->units.ensure(cap * sizeof(Unit))); }
    if (!resident) {   // chunk boundaries come from the device; decode errors stop the call before any chain is walked
        if (cached) bounds = M.bounds;
        else {
            CK(cudaMemcpyAsync(bounds.data(), ctx->bounds.p, (nchunks + 1) * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (int rc = check_device_error(ctx)) return rc;
        }
    }   // (resident mode: the expansion only copies by the validated sizes; the flag is read before the jobs are filled in)

    cudaEvent_t ev_expand_a = nullptr, ev_expand_b = nullptr;
    if (resident) {
        cudaEvent_t a = ctx->event();
        uint32_t n_launch = 0;
        for (const auto& lv : ctx->levels) if (lv.second > lv.first) ++n_launch;
        if (n_launch && diff) {
            const uint64_t capacity = ctx->flat.bytes / 4;
            if (int rc = launch_level_graph(ctx, ctx->g_expand_diff, {ctx->order.p, ctx->nodes.p, ctx->boff.p, ctx->nb.p, ctx->ownb.p, ctx->loc.p, ctx->flat.p, ctx->err_flag.p}, [&] {
                    for (size_t k = 0; k < ctx->levels.size(); ++k) {  // ascending num_samples: parents first
                        const uint32_t b = ctx->levels[k].first, e = ctx->levels[k].second;
                        if (e <= b) continue;
                        k_expand_level_diff<kLevelLanes><<<blocks_for((uint64_t)(e - b) * kLevelLanes, 256), 256, 0, st>>>(
                            e - b, ctx->order.as<uint32_t>() + b, ctx->nodes.as<Node>(), ctx->boff.as<uint64_t>(), ctx->nb.as<uint32_t>(),
                            ctx->ownb.as<uint32_t>(), ctx->loc.as<uint32_t>(), ctx->flat.as<uint32_t>(), capacity, ctx->err_flag.as<int>());
                    }
                })) return rc;
        } else if (n_launch) {
            if (int rc = launch_level_graph(ctx, ctx->g_expand, {ctx->order.p, ctx->nodes.p, ctx->noff.p, ctx->loc.p, ctx->flat.p, ctx->first_id.p}, [&] {
                    for (size_t k = 0; k < ctx->levels.size(); ++k) {  // ascending num_samples: parents first
                        const uint32_t b = ctx->levels[k].first, e = ctx->levels[k].second;
                        if (e <= b) continue;
                        k_expand_level<kLevelLanes><<<blocks_for((uint64_t)(e - b) * kLevelLanes, 256), 256, 0, st>>>(
                            e - b, ctx->order.as<uint32_t>() + b, ctx->nodes.as<Node>(), ctx->noff.as<uint64_t>(), ctx->loc.as<uint32_t>(),
                            ctx->flat.as<uint32_t>(), ctx->first_id.as<uint32_t>());
                    }
                })) return rc;
        }
        launches += n_launch;
        ev_expand_a = a;
        ev_expand_b = ctx->event();
    }
    unsigned scatter_grid = M.scatter_grid;
    const size_t smem = pl.smem;   // both forms pad every accumulator row with kRowPad words
    if (!cached || M.diff != diff) {
        int blocks_per_sm = 0;
        if (diff) {
            CK(cudaFuncSetAttribute(k_scatter_diff, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_scatter_diff, (int)pl.threads, smem));
        } else {
            CK(cudaFuncSetAttribute(k_scatter_add, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_scatter_add, (int)pl.threads, smem));
        }
        if (blocks_per_sm < 1) return ctx->fail(KDBX_ERR_CUDA, "scatter kernel does not fit on an SM");
        scatter_grid = (unsigned)(ctx->sm_count * blocks_per_sm);
    }
    if (smem_buckets) CK(ctx->blockhist.ensure((size_t)wide_grid * nkeys * 4 + 16));

    struct ChunkEv { cudaEvent_t a, b, c, d; };
    std::vector<ChunkEv> cev;
    cev.reserve(nchunks);
    if (resident && num_parts == 1) {  // nothing is streamed: one pass over all patterns
        bounds.assign(2, 0);
        bounds[1] = ctx->P;
    }
    const uint32_t nloops = (uint32_t)bounds.size() - 1;
    uint64_t jobs_cap = resident ? std::min<uint64_t>(ctx->jobs.bytes / sizeof(Job), ctx->units.bytes / sizeof(Unit)) : cap;  // slots allocated
    std::vector<uint64_t> jobs_in_pass;
    uint32_t pass = 0;
    // grows the job / unit arrays of a resident pass to its exact job count: known from the last call, or read now
    auto ensure_job_slots = [&](uint64_t known) -> int {
        if (!resident) return KDBX_OK;
        uint64_t total = known;
        if (!cached) {
            uint32_t h_total = 0;
            int h_flag = 0;   // decode errors surface here: before any job is filled in or executed
            CK(cudaMemcpyAsync(&h_total, ctx->bucket_off.as<uint32_t>() + nkeys, 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(&h_flag, ctx->err_flag.p, sizeof h_flag, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (int rc = error_from_flag(ctx, h_flag)) return rc;
            total = h_total;
        }
        jobs_in_pass.push_back(total);
        if (total + 64 > jobs_cap) {
            jobs_cap = total + 64;
            CK(ctx->jobs.ensure(jobs_cap * sizeof(Job)));
            CK(ctx->units.ensure(jobs_cap * sizeof(Unit)));
        }
        return KDBX_OK;
    };
    for (uint32_t c = 0; c < nloops; ++c) {
        const uint64_t p0 = bounds[c], p1 = bounds[c + 1];
        if (p1 <= p0 || c % num_parts != part) continue;
        ChunkEv e;
        e.a = ctx->event();
        if (!resident)
            k_expand<<<wide_grid, 256, 0, st>>>(p0, p1, ctx->nodes.as<Node>(), ctx->noff.as<uint64_t>(), ctx->loc.as<uint32_t>(), ctx->flat.as<uint32_t>());
        e.b = ctx->event();
        if (!hist_by_decoder) CK(cudaMemsetAsync(ctx->work.p, 0, ((size_t)nkeys + 1) * 8, st));
        const uint64_t known = cached && pass < M.jobs_in_pass.size() ? M.jobs_in_pass[pass] : 0;
        if (hist_by_decoder) {
            // counted by the decoder, offsets already in place (above): the job total is known
            jobs_in_pass.push_back(jobs_total);
            if (jobs_total + 64 > jobs_cap) {
                jobs_cap = jobs_total + 64;
                CK(ctx->jobs.ensure(jobs_cap * sizeof(Job)));
                CK(ctx->units.ensure(jobs_cap * sizeof(Unit)));
            }
            if (diff)
                k_job_fill_runs<true><<<wide_grid, kBucketThreads, 0, st>>>(ctx->P, ctx->nodes.as<Node>(), ctx->boff.as<uint64_t>(), ctx->nb.as<uint32_t>(),
                                                                            ctx->ownb.as<uint32_t>(), ctx->W.as<uint32_t>(), ctx->loc.as<uint32_t>(), pl.rb_shift, nkeys,
                                                                            ctx->blockhist.as<uint32_t>(), ctx->jobs.as<Job>(), per_block, d_physical);
            else
                k_job_fill_runs<false><<<wide_grid, kBucketThreads, 0, st>>>(ctx->P, ctx->nodes.as<Node>(), ctx->noff.as<uint64_t>(), nullptr, nullptr,
                                                                             ctx->W.as<uint32_t>(), ctx->loc.as<uint32_t>(), pl.rb_shift, nkeys,
                                                                             ctx->blockhist.as<uint32_t>(), ctx->jobs.as<Job>(), per_block, d_physical);
            launches += 1;
        } else if (smem_buckets) {
            // per-block slices of whole decoder blocks when the pass covers all patterns; chunks are sliced evenly
            const uint64_t per = (p0 == 0 && p1 == ctx->P) ? per_block : 0;
            k_job_hist_smem<<<wide_grid, kBucketThreads, 0, st>>>(p0, p1, ctx->nodes.as<Node>(), ctx->noff.as<uint64_t>(), ctx->W.as<uint32_t>(),
                                                                  ctx->flat.as<uint32_t>(), ctx->loc.as<uint32_t>(), resident ? ctx->first_id.as<uint32_t>() : nullptr, resident ? 1u : 0u, pl.T, pl.tile_cols, pl.rb_shift, wb, we, nkeys,
                                                                  ctx->blockhist.as<uint32_t>(), ctx->work.as<unsigned long long>(), d_total_updates, per);
            k_key_totals<<<blocks_for((uint64_t)nkeys + 1, 128), 128, 0, st>>>(nkeys, wide_grid, ctx->blockhist.as<uint32_t>(), ctx->hist.as<uint32_t>());
            if (int rc = scan_exclusive_u32(ctx, ctx->hist.as<uint32_t>(), ctx->bucket_off.as<uint32_t>(), (uint64_t)nkeys + 1)) return rc;
            if (int rc = ensure_job_slots(known)) return rc;
            k_block_offsets<<<blocks_for((uint64_t)nkeys * 32, 256), 256, 0, st>>>(nkeys, wide_grid, ctx->bucket_off.as<uint32_t>(), ctx->blockhist.as<uint32_t>());
            k_job_fill_smem<<<wide_grid, kBucketThreads, 0, st>>>(p0, p1, ctx->nodes.as<Node>(), ctx->noff.as<uint64_t>(), ctx->W.as<uint32_t>(),
                                                                  ctx->flat.as<uint32_t>(), ctx->loc.as<uint32_t>(), resident ? ctx->first_id.as<uint32_t>() : nullptr, resident ? 1u : 0u, pl.T, pl.tile_cols, pl.rb_shift, wb, we, nkeys,
                                                                  ctx->blockhist.as<uint32_t>(), ctx->jobs.as<Job>(), per);
            launches += 5;
        } else {
            CK(cudaMemsetAsync(ctx->hist.p, 0, ((size_t)nkeys + 1) * 4, st));
            k_job_hist<<<wide_grid, kBucketThreads, 0, st>>>(p0, p1, ctx->nodes.as<Node>(), ctx->noff.as<uint64_t>(), ctx->W.as<uint32_t>(),
                                                   ctx->flat.as<uint32_t>(), ctx->loc.as<uint32_t>(), resident ? ctx->first_id.as<uint32_t>() : nullptr, resident ? 1u : 0u, pl.T, pl.tile_cols, pl.rb_shift, wb, we,
                                                   ctx->hist.as<uint32_t>(), ctx->work.as<unsigned long long>(), d_total_updates);
            if (int rc = scan_exclusive_u32(ctx, ctx->hist.as<uint32_t>(), ctx->bucket_off.as<uint32_t>(), (uint64_t)nkeys + 1)) return rc;
            if (int rc = ensure_job_slots(known)) return rc;
            CK(cudaMemcpyAsync(ctx->cursor.p, ctx->bucket_off.p, ((size_t)nkeys + 1) * 4, cudaMemcpyDeviceToDevice, st));
            k_job_fill<<<wide_grid, kBucketThreads, 0, st>>>(p0, p1, ctx->nodes.as<Node>(), ctx->noff.as<uint64_t>(), ctx->W.as<uint32_t>(),
                                                   ctx->flat.as<uint32_t>(), ctx->loc.as<uint32_t>(), resident ? ctx->first_id.as<uint32_t>() : nullptr, resident ? 1u : 0u, pl.T, pl.tile_cols, pl.rb_shift, wb, we,
                                                   ctx->cursor.as<uint32_t>(), ctx->jobs.as<Job>());
            launches += 4;
        }
        ++pass;
        k_work_total<<<1, 256, 0, st>>>(nkeys, ctx->work.as<unsigned long long>());
        k_unit_count<<<blocks_for((uint64_t)nkeys + 1, 256), 256, 0, st>>>(nkeys, ctx->hist.as<uint32_t>(), ctx->work.as<unsigned long long>(),
                                                                            pl.unit_updates, ctx->cfg.unit_updates ? 0u : (uint32_t)ctx->sm_count * kUnitsPerCta,
                                                                            ctx->ucount.as<uint32_t>());
        if (int rc = scan_exclusive_u32(ctx, ctx->ucount.as<uint32_t>(), ctx->uoff.as<uint32_t>(), (uint64_t)nkeys + 1)) return rc;
        k_unit_fill<<<blocks_for(nkeys, 256), 256, 0, st>>>(nkeys, ctx->hist.as<uint32_t>(), ctx->bucket_off.as<uint32_t>(),
                                                             ctx->ucount.as<uint32_t>(), ctx->uoff.as<uint32_t>(), ctx->units.as<Unit>());
        CK(cudaMemsetAsync(d_unit_counter, 0, 4, st));
        e.c = ctx->event();
        if (diff)
            k_scatter_diff<<<scatter_grid, pl.threads, smem, st>>>(ctx->units.as<Unit>(), ctx->uoff.as<uint32_t>() + nkeys, ctx->jobs.as<Job>(),
                                                                   ctx->flat.as<uint32_t>(), ctx->loc.as<uint32_t>(), d_out, pl.lo, pl.Nw, pl.tile_cols, pl.rb_shift,
                                                                   d_unit_counter);
        else
            k_scatter_add<<<scatter_grid, pl.threads, smem, st>>>(ctx->units.as<Unit>(), ctx->uoff.as<uint32_t>() + nkeys, ctx->jobs.as<Job>(),
                                                                  ctx->flat.as<uint32_t>(), resident ? ctx->noff.as<uint64_t>() + p0 : nullptr, d_out, tri_base, pl.lo,
                                                                  pl.T, pl.tile_cols, pl.rb_shift, d_unit_counter);
        e.d = ctx->event();
        launches += 5;
        s.scatter_launches += 1;
        cev.push_back(e);
    }
    cudaEvent_t ev_end = ctx->event();
    CK(cudaGetLastError());
    uint64_t* h = ctx->h_pinned;
    h[4] = h[5] = h[6] = 0;
    CK(cudaMemcpyAsync(&h[4], d_total_updates, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h[5], d_physical, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h[6], ctx->err_flag.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (int rc = error_from_flag(ctx, (int)(uint32_t)h[6])) return rc;
    s.ms_prepare = elapsed(ev_start, ev_prepared);
    for (const ChunkEv& e : cev) {
        s.ms_expand += elapsed(e.a, e.b);
        s.ms_bucket += elapsed(e.b, e.c);
        s.ms_scatter += elapsed(e.c, e.d);
    }
    if (ev_bucket0) s.ms_bucket += elapsed(ev_bucket0, ev_bucket0_end);
    if (int rc = finish_upload(ctx)) return rc;
    s.ms_upload = ctx->ms_upload;
    if (ev_expand_a) s.ms_expand += elapsed(ev_expand_a, ev_expand_b);
    s.ms_total = elapsed(ev_start, ev_end);
    s.updates = h[4];
    s.physical_updates = diff ? h[5] : h[4];
    s.list_form = diff ? 1u : 0u;
    s.flat_ids = diff ? diff_slots : ctx->sum_n; s.local_ids = ctx->sum_l;
    s.chunks = nloops; s.kernel_launches = launches;
    if (stats) *stats = s;
    // what the next call with the same signature may take for granted
    M.valid = true; M.gen = ctx->load_gen; M.row_begin = row_begin; M.row_end = row_end; M.part = part; M.num_parts = num_parts;
    M.win_lo = ctx->win_lo; M.win_hi = ctx->win_hi; M.tile_cols = pl.tile_cols; M.tile_rows = pl.tile_rows; M.threads = pl.threads;
    M.unit_updates = pl.unit_updates; M.flags = ctx->cfg.flags; M.chunk = pl.chunk;
    M.out_aligned = (reinterpret_cast<uintptr_t>(d_out) & 15u) == 0;
    M.resident = resident; M.diff = diff; M.nchunks = nchunks; M.scatter_grid = scatter_grid; M.diff_slots = diff_slots;
    if (!resident) M.bounds = bounds;
    M.jobs_in_pass = jobs_in_pass;
    return KDBX_OK;
}

int all2all_rows_device(kdbx_ctx* ctx, uint32_t row_begin, uint32_t row_end, uint32_t* d_out, kdbx_stats* stats,
                        uint32_t part = 0, uint32_t num_parts = 1) {
    int rc = all2all_rows_device_impl(ctx, row_begin, row_end, d_out, stats, part, num_parts, ctx->no_slide_gen != ctx->load_gen);
    if (rc == kRetryWithoutSlide) {
        ctx->no_slide_gen = ctx->load_gen;   // remembered for this staged trie
        rc = all2all_rows_device_impl(ctx, row_begin, row_end, d_out, stats, part, num_parts, false);
    }
    return rc;
}

#include "sparse.cuh"
#include "comm.cuh"
#include "csvfmt.cuh"
#include "query.cuh"
#include "db2db.cuh"
#include "build.cuh"

}  // namespace

// Device-side database builder (build.cuh / build_host.cuh); lives on a context's device and stream.
struct kdbx_builder {
    kdbx_ctx* ctx = nullptr;
    BuildAlphabet alpha{};          // host copy of the extraction parameters
    DevBuf d_alpha;
    uint32_t sentinel_bit = 0;
    uint64_t num_tables = 0;
    // k-mer -> pattern id: open addressing over the whole 64-bit k-mer
    DevBuf keys, vals;
    uint64_t cap = 0, filled = 0;
    uint32_t table_growths = 0;
    // patterns (SoA, index = pattern id) and the log of in-place extensions
    DevBuf num_kmers, parent, n, l, last, born, is_parent;
    uint64_t P = 0, pat_cap = 0;
    DevBuf ev_pat, ev_sample;
    uint64_t ev_count = 0, ev_cap = 0;
    // per-sample scratch
    DevBuf seq, raw, sorted, uniq, nsel, slot_of, pid, pid2, idx, idx2, head, head_incl, run_start, split, split_incl, run_tag, ps;
    BuildPerSample* h_ps = nullptr;   // pinned
    uint32_t num_samples = 0;
    uint64_t total_kmers = 0;
    uint32_t launches = 0;
    // finish
    bool finished = false;
    DevBuf bits, eoff, poff, payload, ev_pat2, ev_sample2, tfilled, slot_off, slots;
    uint64_t payload_words = 0, total_slots = 0;
    float ms_finish = 0.f;
};

namespace {
#include "build_host.cuh"
}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int kdbx_abi_version(void) { return KDBX_ABI_VERSION; }

int kdbx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return -1; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp pr;
        if (cudaGetDeviceProperties(&pr, d) == cudaSuccess && pr.major == 10) ++ok;
    }
    return ok;
}

const char* kdbx_last_error(const kdbx_ctx* ctx) { return ctx ? ctx->err.c_str() : g_open_error.c_str(); }

int kdbx_open(const kdbx_config* cfg, kdbx_ctx** out) {
    if (!out) { g_open_error = "kdbx_open: out is NULL"; return KDBX_ERR_ARG; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        g_open_error = std::string("kdbx_open: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
        return KDBX_ERR_CUDA;
    }
    int dev = cfg ? cfg->device : -1;
    if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
    if (dev >= ndev) { g_open_error = "kdbx_open: device ordinal out of range"; return KDBX_ERR_ARG; }
    cudaDeviceProp pr;
    if ((e = cudaGetDeviceProperties(&pr, dev)) != cudaSuccess) { g_open_error = cudaGetErrorString(e); return KDBX_ERR_CUDA; }
    if (pr.major != 10) {
        g_open_error = "kdbx_open: device is sm_" + std::to_string(pr.major) + std::to_string(pr.minor) +
                       "; this library carries sm_100a code only";
        return KDBX_ERR_CUDA;
    }
    kdbx_ctx* ctx = new kdbx_ctx();
    ctx->device = dev;
    ctx->sm_count = pr.multiProcessorCount;
    if (cfg) ctx->cfg = *cfg;
    if ((e = cudaSetDevice(dev)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev_up_begin)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev_up_hdr)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev_up_payload)) != cudaSuccess) {
        g_open_error = std::string("kdbx_open: ") + cudaGetErrorString(e);
        delete ctx;
        return KDBX_ERR_CUDA;
    }
    *out = ctx;
    return KDBX_OK;
}

void kdbx_close(kdbx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (DevBuf* b : {&ctx->num_kmers, &ctx->parent, &ctx->n, &ctx->l, &ctx->last, &ctx->bits, &ctx->poff, &ctx->payload,
                      &ctx->nodes, &ctx->W, &ctx->loc, &ctx->loff, &ctx->noff, &ctx->coff, &ctx->bounds, &ctx->err_flag,
                      &ctx->cub_tmp, &ctx->order_in, &ctx->order, &ctx->keys_sorted, &ctx->level_start, &ctx->flat, &ctx->jobs, &ctx->hist, &ctx->work, &ctx->bucket_off, &ctx->cursor,
                      &ctx->ucount, &ctx->uoff, &ctx->units, &ctx->counters, &ctx->blockhist, &ctx->tri, &ctx->rowupd, &ctx->first_id,
                      &ctx->sp_cnt, &ctx->sp_counts, &ctx->sp_rowptr, &ctx->sp_col, &ctx->sp_val, &ctx->slot_off, &ctx->slots,
                      &ctx->q_off, &ctx->q_kmers, &ctx->q_keys, &ctx->q_keys2, &ctx->q_runkeys, &ctx->q_runcnt, &ctx->q_out,
                      &ctx->qx_alpha, &ctx->qx_seq, &ctx->qx_raw, &ctx->qx_sorted, &ctx->qx_count, &ctx->ownb, &ctx->nb, &ctx->boff, &ctx->rs_block, &ctx->csv_text, &ctx->hdr32})
        b->release();
    if (ctx->comm) { nccl_api()->CommDestroy(static_cast<ncclComm_t>(ctx->comm)); ctx->comm = nullptr; }
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    if (ctx->g_push.exec) cudaGraphExecDestroy(ctx->g_push.exec);
    if (ctx->g_expand.exec) cudaGraphExecDestroy(ctx->g_expand.exec);
    if (ctx->g_pull.exec) cudaGraphExecDestroy(ctx->g_pull.exec);
    if (ctx->g_expand_diff.exec) cudaGraphExecDestroy(ctx->g_expand_diff.exec);
    if (ctx->up_stream) { cudaStreamSynchronize(ctx->up_stream); cudaStreamDestroy(ctx->up_stream); }
    for (cudaEvent_t e : {ctx->ev_up_begin, ctx->ev_up_hdr, ctx->ev_up_payload}) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_up_chunk) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int kdbx_host_alloc(void** out, size_t bytes) {
    if (!out) return KDBX_ERR_ARG;
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return KDBX_ERR_NOMEM; }
    return KDBX_OK;
}
void kdbx_host_free(void* p) { if (p) cudaFreeHost(p); }

int kdbx_load_patterns(kdbx_ctx* ctx, const kdbx_trie_view* v) {
    if (!ctx) return KDBX_ERR_ARG;
    if (!v) return ctx->fail(KDBX_ERR_ARG, "kdbx_load_patterns: view is NULL");
    const uint64_t P = v->num_patterns;
    if (P == 0 || P >= ((uint64_t)1 << 31)) return ctx->fail(KDBX_ERR_ARG, "kdbx_load_patterns: num_patterns must be in [1, 2^31)");
    if (!v->num_kmers || !v->parent_id || !v->num_samples_full || !v->num_local_samples || !v->last_sample_id || !v->num_bits)
        return ctx->fail(KDBX_ERR_ARG, "kdbx_load_patterns: NULL array in view");
    if (v->payload_words && !v->payload) return ctx->fail(KDBX_ERR_ARG, "kdbx_load_patterns: payload is NULL");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->loaded = false;
    ctx->prepared = false;
    CK(ctx->num_kmers.ensure(P * 8)); CK(ctx->parent.ensure(P * 8)); CK(ctx->n.ensure(P * 4)); CK(ctx->l.ensure(P * 4));
    CK(ctx->last.ensure(P * 4)); CK(ctx->bits.ensure(P * 4)); CK(ctx->payload.ensure((v->payload_words + 2) * 8));
    // headers on the compute stream (the scans need them first), the Elias-gamma payload on a second
    // stream: only the decode kernel waits for it (prepare())
    CK(cudaStreamSynchronize(ctx->up_stream));
    CK(cudaEventRecord(ctx->ev_up_begin, st));
    CK(cudaMemcpyAsync(ctx->n.p, v->num_samples_full, P * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->l.p, v->num_local_samples, P * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->last.p, v->last_sample_id, P * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->bits.p, v->num_bits, P * 4, cudaMemcpyHostToDevice, st));
    if (v->parent_id32 && v->num_kmers32) {   // 8 instead of 16 bytes per pattern over PCIe, widened on the device
        CK(ctx->hdr32.ensure(P * 8));
        CK(cudaMemcpyAsync(ctx->hdr32.p, v->parent_id32, P * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->hdr32.as<uint32_t>() + P, v->num_kmers32, P * 4, cudaMemcpyHostToDevice, st));
        k_widen_headers<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(P, ctx->hdr32.as<int32_t>(), ctx->hdr32.as<uint32_t>() + P,
                                                                      ctx->parent.as<int64_t>(), ctx->num_kmers.as<int64_t>());
        CK(cudaGetLastError());
    } else {
        CK(cudaMemcpyAsync(ctx->parent.p, v->parent_id, P * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->num_kmers.p, v->num_kmers, P * 8, cudaMemcpyHostToDevice, st));
    }
    ctx->dense_payload = (v->payload_off == nullptr);
    if (!ctx->dense_payload) {
        CK(ctx->poff.ensure((P + 1) * 8));
        CK(cudaMemcpyAsync(ctx->poff.p, v->payload_off, P * 8, cudaMemcpyHostToDevice, st));
    }
    CK(cudaEventRecord(ctx->ev_up_hdr, st));
    // the payload follows the headers on the link (one copy at a time runs at link rate; sharing it would only delay the
    // headers, which everything before the decoder waits for), in chunks when the upload is asynchronous, so that the
    // decoder can start on the first chunk while the others travel (prepare())
    CK(cudaStreamWaitEvent(ctx->up_stream, ctx->ev_up_hdr, 0));   // (also: not before earlier work on the buffers is done)
    {
        uint64_t K = 1;
        if ((ctx->cfg.flags & KDBX_FLAG_ASYNC_UPLOAD) && ctx->dense_payload)
            K = std::min<uint64_t>(8, std::max<uint64_t>(1, v->payload_words * 8 / (ctx->cfg.upload_chunk_bytes ? ctx->cfg.upload_chunk_bytes : ((uint64_t)96 << 20))));
        ctx->up_bounds.assign(K + 1, 0);
        for (uint64_t k = 0; k <= K; ++k) ctx->up_bounds[k] = k == K ? v->payload_words : (v->payload_words / K * k) & ~(uint64_t)1;
        while (ctx->ev_up_chunk.size() < K) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->ev_up_chunk.push_back(e); }
        for (uint64_t k = 0; k < K; ++k) {
            const uint64_t b = ctx->up_bounds[k], e = ctx->up_bounds[k + 1];
            if (e > b) CK(cudaMemcpyAsync(ctx->payload.as<uint64_t>() + b, v->payload + b, (e - b) * 8, cudaMemcpyHostToDevice, ctx->up_stream));
            CK(cudaEventRecord(ctx->ev_up_chunk[k], ctx->up_stream));
        }
    }
    // two zero guard words: the decoder may touch word i+1 of a run that ends at a word edge
    CK(cudaMemsetAsync(ctx->payload.as<uint64_t>() + v->payload_words, 0, 16, ctx->up_stream));
    CK(cudaEventRecord(ctx->ev_up_payload, ctx->up_stream));
    ctx->upload_pending = true;
    if (!(ctx->cfg.flags & KDBX_FLAG_ASYNC_UPLOAD)) {
        if (int rc = finish_upload(ctx)) return rc;
    }
    ctx->P = P; ctx->N = v->num_samples; ctx->payload_words = v->payload_words;
    ctx->win_lo = 0; ctx->win_hi = ctx->N;
    ++ctx->load_gen;
    ctx->loaded = true;
    return KDBX_OK;
}

int kdbx_set_sample_window(kdbx_ctx* ctx, uint32_t lo, uint32_t hi) {
    if (!ctx) return KDBX_ERR_ARG;
    if (!ctx->loaded) return ctx->fail(KDBX_ERR_STATE, "no patterns loaded (call kdbx_load_patterns first)");
    if (lo > hi || hi > ctx->N) return ctx->fail(KDBX_ERR_ARG, "bad sample window [%u,%u) for %u samples", lo, hi, ctx->N);
    lo &= ~31u;   // row blocks and column windows are laid on a 32-id grid
    if (lo != ctx->win_lo || hi != ctx->win_hi) { ctx->win_lo = lo; ctx->win_hi = hi; ctx->prepared = false; ++ctx->load_gen; }
    return KDBX_OK;
}

int kdbx_all2all_dense_rows_device(kdbx_ctx* ctx, uint32_t row_begin, uint32_t row_end, void* d_out_rows, kdbx_stats* stats) {
    if (!ctx) return KDBX_ERR_ARG;
    if (!d_out_rows && row_end > row_begin && row_end > 1) return ctx->fail(KDBX_ERR_ARG, "output pointer is NULL");
    return all2all_rows_device(ctx, row_begin, row_end, static_cast<uint32_t*>(d_out_rows), stats);
}

int kdbx_all2all_dense_part_device(kdbx_ctx* ctx, uint32_t part, uint32_t num_parts, void* d_out_tri, kdbx_stats* stats) {
    if (!ctx) return KDBX_ERR_ARG;
    if (num_parts == 0 || part >= num_parts) return ctx->fail(KDBX_ERR_ARG, "bad part %u of %u", part, num_parts);
    if (!d_out_tri && ctx->N > 1) return ctx->fail(KDBX_ERR_ARG, "output pointer is NULL");
    return all2all_rows_device(ctx, 0, ctx->N, static_cast<uint32_t*>(d_out_tri), stats, part, num_parts);
}

int kdbx_all2all_dense_rows(kdbx_ctx* ctx, uint32_t row_begin, uint32_t row_end, uint32_t* out_rows, kdbx_stats* stats) {
    if (!ctx) return KDBX_ERR_ARG;
    if (!ctx->loaded) return ctx->fail(KDBX_ERR_STATE, "no patterns loaded (call kdbx_load_patterns first)");
    if (row_begin > row_end || row_end > ctx->N) return ctx->fail(KDBX_ERR_ARG, "bad row range [%u,%u) for %u samples", row_begin, row_end, ctx->N);
    auto tri_off = [](uint64_t r) { return r == 0 ? 0ull : r * (r - 1) / 2; };
    const uint64_t cells = tri_off(row_end) - tri_off(row_begin);
    if (cells && !out_rows) return ctx->fail(KDBX_ERR_ARG, "output pointer is NULL");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->tri.ensure((cells + 1) * 4));
    kdbx_stats s{};
    ctx->tri_rows_valid = false;
    if (int rc = all2all_rows_device(ctx, row_begin, row_end, ctx->tri.as<uint32_t>(), &s)) return rc;
    ctx->tri_rows_valid = true; ctx->tri_row_begin = row_begin; ctx->tri_row_end = row_end;   // (kdbx_csv_dense_rows formats from here)
    cudaEvent_t a = ctx->event();
    if (cells) CK(cudaMemcpyAsync(out_rows, ctx->tri.p, cells * 4, cudaMemcpyDeviceToHost, ctx->stream));
    cudaEvent_t b = ctx->event();
    CK(cudaStreamSynchronize(ctx->stream));
    s.ms_download = elapsed(a, b);
    if (stats) *stats = s;
    return KDBX_OK;
}

int kdbx_all2all_dense(kdbx_ctx* ctx, uint32_t* out_tri, kdbx_stats* stats) {
    if (!ctx) return KDBX_ERR_ARG;
    return kdbx_all2all_dense_rows(ctx, 0, ctx->N, out_tri, stats);
}

int kdbx_comm_unique_id(void* id) {
    if (!id) return KDBX_ERR_ARG;
    NcclApi* api = nccl_api();
    if (!api->error.empty()) { g_open_error = api->error; return KDBX_ERR_STATE; }
    ncclUniqueId u;
    const ncclResult_t r = api->GetUniqueId(&u);
    if (r != ncclSuccess) { g_open_error = std::string("ncclGetUniqueId failed: ") + api->GetErrorString(r); return KDBX_ERR_CUDA; }
    std::memcpy(id, &u, KDBX_COMM_ID_BYTES);
    return KDBX_OK;
}

int kdbx_comm_init_rank(kdbx_ctx* ctx, int nranks, int rank, const void* id) {
    if (!ctx) return KDBX_ERR_ARG;
    if (!id || nranks < 1 || rank < 0 || rank >= nranks) return ctx->fail(KDBX_ERR_ARG, "kdbx_comm_init_rank: bad rank %d of %d", rank, nranks);
    NcclApi* api = nccl_api();
    if (!api->error.empty()) return ctx->fail(KDBX_ERR_STATE, "%s", api->error.c_str());
    CK(cudaSetDevice(ctx->device));
    if (ctx->comm) { api->CommDestroy(static_cast<ncclComm_t>(ctx->comm)); ctx->comm = nullptr; }
    ncclUniqueId u;
    std::memcpy(&u, id, KDBX_COMM_ID_BYTES);
    ncclComm_t c = nullptr;
    NCK(api->CommInitRank(&c, nranks, u, rank));
    ctx->comm = c; ctx->comm_nranks = nranks; ctx->comm_rank = rank;
    return KDBX_OK;
}

int kdbx_comm_init_all(kdbx_ctx* const* ctxs, int n) {
    if (!ctxs || n < 1) return KDBX_ERR_ARG;
    kdbx_ctx* ctx = ctxs[0];
    if (!ctx) return KDBX_ERR_ARG;
    NcclApi* api = nccl_api();
    if (!api->error.empty()) return ctx->fail(KDBX_ERR_STATE, "%s", api->error.c_str());
    std::vector<int> devs((size_t)n);
    std::vector<ncclComm_t> comms((size_t)n, nullptr);
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i]) return ctx->fail(KDBX_ERR_ARG, "kdbx_comm_init_all: context %d is NULL", i);
        devs[(size_t)i] = ctxs[i]->device;
        if (ctxs[i]->comm) { api->CommDestroy(static_cast<ncclComm_t>(ctxs[i]->comm)); ctxs[i]->comm = nullptr; }
    }
    NCK(api->CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) { ctxs[i]->comm = comms[(size_t)i]; ctxs[i]->comm_nranks = n; ctxs[i]->comm_rank = i; }
    return KDBX_OK;
}

void kdbx_comm_destroy(kdbx_ctx* ctx) {
    if (!ctx || !ctx->comm) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    nccl_api()->CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nullptr; ctx->comm_nranks = 1; ctx->comm_rank = 0;
}

int kdbx_all2all_dense_reduce_scatter_device(kdbx_ctx* ctx, void* d_block, uint64_t* first_cell, uint64_t* num_cells, kdbx_stats* stats) {
    if (!ctx) return KDBX_ERR_ARG;
    return all2all_reduce_scatter_device(ctx, static_cast<uint32_t*>(d_block), first_cell, num_cells, stats);
}

int kdbx_all2all_dense_reduce_scatter(kdbx_ctx* ctx, uint32_t* out_block, uint64_t* first_cell, uint64_t* num_cells, kdbx_stats* stats) {
    if (!ctx) return KDBX_ERR_ARG;
    if (!ctx->loaded) return ctx->fail(KDBX_ERR_STATE, "no patterns loaded (call kdbx_load_patterns first)");
    const uint64_t N = ctx->N, cells = N ? N * (N - 1) / 2 : 0;
    const uint64_t B = comm_block_cells(cells, ctx->comm ? ctx->comm_nranks : 1);
    CK(cudaSetDevice(ctx->device));
    CK(ctx->rs_block.ensure((B + 4) * 4));
    kdbx_stats s{};
    uint64_t first = 0, count = 0;
    if (int rc = all2all_reduce_scatter_device(ctx, ctx->rs_block.as<uint32_t>(), &first, &count, &s)) return rc;
    if (count && !out_block) return ctx->fail(KDBX_ERR_ARG, "output pointer is NULL");
    cudaEvent_t a = ctx->event();
    if (count) CK(cudaMemcpyAsync(out_block, ctx->rs_block.p, count * 4, cudaMemcpyDeviceToHost, ctx->stream));
    cudaEvent_t b = ctx->event();
    CK(cudaStreamSynchronize(ctx->stream));
    s.ms_download = elapsed(a, b);
    if (first_cell) *first_cell = first;
    if (num_cells) *num_cells = count;
    if (stats) *stats = s;
    return KDBX_OK;
}

int kdbx_csv_dense_rows(kdbx_ctx* ctx, uint32_t row_begin, uint32_t row_end, char* text, uint64_t capacity, uint64_t* row_off, uint64_t* bytes) {
    if (!ctx) return KDBX_ERR_ARG;
    return csv_dense_rows(ctx, row_begin, row_end, text, capacity, row_off, bytes);
}

int kdbx_stage_matrix(kdbx_ctx* ctx, const uint32_t* tri, uint32_t num_samples) {
    if (!ctx) return KDBX_ERR_ARG;
    return stage_matrix(ctx, tri, num_samples);
}

int kdbx_distance_dense_rows(kdbx_ctx* ctx, int metric, const uint32_t* sample_kmers, uint32_t row_begin, uint32_t row_end, char* text,
                             uint64_t capacity, uint64_t* row_off, uint64_t* bytes) {
    if (!ctx) return KDBX_ERR_ARG;
    return distance_dense_rows(ctx, metric, sample_kmers, row_begin, row_end, text, capacity, row_off, bytes);
}

int kdbx_row_updates(kdbx_ctx* ctx, uint64_t* out) {
    if (!ctx) return KDBX_ERR_ARG;
    if (!ctx->loaded) return ctx->fail(KDBX_ERR_STATE, "no patterns loaded (call kdbx_load_patterns first)");
    if (!out) return ctx->fail(KDBX_ERR_ARG, "output pointer is NULL");
    if (int rc = require_full_window(ctx, "kdbx_row_updates")) return rc;
    CK(cudaSetDevice(ctx->device));
    Plan pl;
    if (int rc = make_plan(ctx, pl)) return rc;
    uint32_t launches = 0;
    const int rc = prepare(ctx, pl, launches);
    if (rc < 0) return rc;
    CK(ctx->rowupd.ensure(((size_t)ctx->N + 1) * 8));
    CK(cudaMemsetAsync(ctx->rowupd.p, 0, ((size_t)ctx->N + 1) * 8, ctx->stream));
    k_row_updates<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->P, ctx->N, ctx->nodes.as<Node>(), ctx->loc.as<uint32_t>(),
                                                              ctx->rowupd.as<unsigned long long>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->rowupd.p, (size_t)ctx->N * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return check_device_error(ctx);
}

int kdbx_all2all_sparse(kdbx_ctx* ctx, const kdbx_filter* filter, kdbx_csr* out, kdbx_stats* stats) {
    if (!ctx) return KDBX_ERR_ARG;
    return all2all_sparse_impl(ctx, filter, out, stats, 0, ctx->N);
}

int kdbx_all2all_sparse_rows(kdbx_ctx* ctx, uint32_t row_begin, uint32_t row_end, const kdbx_filter* filter, kdbx_csr* out, kdbx_stats* stats) {
    if (!ctx) return KDBX_ERR_ARG;
    return all2all_sparse_impl(ctx, filter, out, stats, row_begin, row_end);
}

int kdbx_db2db_sparse(kdbx_ctx* rows_db, kdbx_ctx* cols_db, const kdbx_filter* filter, const uint32_t* cols_sample_kmers, kdbx_csr* out,
                      kdbx_stats* stats) {
    if (!rows_db) return KDBX_ERR_ARG;
    return db2db_sparse_impl(rows_db, cols_db, filter, cols_sample_kmers, out, stats);
}

void kdbx_free_csr(kdbx_csr* csr) {
    if (!csr) return;
    std::free(csr->row_ptr);
    if (csr->_pad == 1) { if (csr->col) cudaFreeHost(csr->col); if (csr->val) cudaFreeHost(csr->val); }
    else { std::free(csr->col); std::free(csr->val); }
    std::memset(csr, 0, sizeof *csr);
}

int kdbx_load_hashtables(kdbx_ctx* ctx, const kdbx_tables_view* view) {
    if (!ctx) return KDBX_ERR_ARG;
    return load_hashtables_impl(ctx, view);
}

int kdbx_new2all_batch(kdbx_ctx* ctx, const uint64_t* kmers, const uint64_t* q_off, uint32_t n_queries, uint32_t* out,
                       kdbx_stats* stats) {
    if (!ctx) return KDBX_ERR_ARG;
    return new2all_impl(ctx, kmers, q_off, n_queries, out, stats);
}

int kdbx_new2all_sequences(kdbx_ctx* ctx, const kdbx_build_params* params, const char* symbols, const uint64_t* q_off,
                           uint32_t n_queries, uint32_t* out, uint64_t* unique_kmers, kdbx_stats* stats) {
    if (!ctx) return KDBX_ERR_ARG;
    return new2all_sequences_impl(ctx, params, symbols, q_off, n_queries, out, unique_kmers, stats);
}

int kdbx_builder_open(kdbx_ctx* ctx, const kdbx_build_params* p, kdbx_builder** out) {
    if (!ctx) return KDBX_ERR_ARG;
    if (!p || !out) return ctx->fail(KDBX_ERR_ARG, "kdbx_builder_open: NULL argument");
    *out = nullptr;
    BuildAlphabet alpha{};
    uint32_t sentinel_bit = 0;
    uint64_t num_tables = 0;
    if (int rc = make_build_alphabet(ctx, p, "kdbx_builder_open", alpha, sentinel_bit, num_tables)) return rc;
    CK(cudaSetDevice(ctx->device));
    kdbx_builder* b = new kdbx_builder();
    b->ctx = ctx;
    b->alpha = alpha;
    b->sentinel_bit = sentinel_bit;
    b->num_tables = num_tables;
    auto fail = [&](int code, const char* what) { kdbx_builder_close(b); return ctx->fail(code, "kdbx_builder_open: %s", what); };
    if (b->d_alpha.ensure(sizeof(BuildAlphabet)) != cudaSuccess || b->ps.ensure(sizeof(BuildPerSample)) != cudaSuccess ||
        b->nsel.ensure(16) != cudaSuccess)
        return fail(KDBX_ERR_NOMEM, "device allocation failed");
    if (cudaHostAlloc(reinterpret_cast<void**>(&b->h_ps), sizeof(BuildPerSample), cudaHostAllocDefault) != cudaSuccess)
        return fail(KDBX_ERR_NOMEM, "pinned allocation failed");
    if (cudaMemcpyAsync(b->d_alpha.p, &b->alpha, sizeof(BuildAlphabet), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
        cudaMemsetAsync(b->ps.p, 0, sizeof(BuildPerSample), ctx->stream) != cudaSuccess)
        return fail(KDBX_ERR_CUDA, "staging the parameters failed");
    if (builder_reserve_patterns(b, 4096) != KDBX_OK || builder_reserve_events(b, 4096) != KDBX_OK) { kdbx_builder_close(b); return KDBX_ERR_NOMEM; }
    // pattern 0: the empty sentinel every database starts with (src/prefix_kmer_db.cpp:24)
    const long long minus1 = -1;
    for (DevBuf* d : {&b->num_kmers, &b->n, &b->l, &b->last, &b->born, &b->is_parent}) cudaMemsetAsync(d->p, 0, 64, ctx->stream);
    cudaMemcpyAsync(b->parent.p, &minus1, 8, cudaMemcpyHostToDevice, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(KDBX_ERR_CUDA, "initialisation failed");
    b->P = 1;
    if (p->table_capacity_hint) {
        if (int rc = builder_reserve_table(b, p->table_capacity_hint / 2)) { kdbx_builder_close(b); return rc; }
    }
    *out = b;
    return KDBX_OK;
}

void kdbx_builder_close(kdbx_builder* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    for (DevBuf* d : {&b->d_alpha, &b->keys, &b->vals, &b->num_kmers, &b->parent, &b->n, &b->l, &b->last, &b->born, &b->is_parent,
                      &b->ev_pat, &b->ev_sample, &b->seq, &b->raw, &b->sorted, &b->uniq, &b->nsel, &b->slot_of, &b->pid, &b->pid2,
                      &b->idx, &b->idx2, &b->head, &b->head_incl, &b->run_start, &b->split, &b->split_incl, &b->run_tag, &b->ps,
                      &b->bits, &b->eoff, &b->poff, &b->payload, &b->ev_pat2, &b->ev_sample2, &b->tfilled, &b->slot_off, &b->slots})
        d->release();
    if (b->h_ps) cudaFreeHost(b->h_ps);
    delete b;
}

int kdbx_builder_adopt(kdbx_builder* b) { return b ? builder_adopt(b) : KDBX_ERR_ARG; }
int kdbx_builder_add_sequence(kdbx_builder* b, const char* symbols, uint64_t len, uint64_t* unique_kmers) {
    return b ? builder_add_sequence(b, symbols, len, unique_kmers) : KDBX_ERR_ARG;
}
int kdbx_builder_add_kmers(kdbx_builder* b, const uint64_t* kmers, uint64_t count) { return b ? builder_add_kmers(b, kmers, count) : KDBX_ERR_ARG; }
int kdbx_builder_finish(kdbx_builder* b, kdbx_build_result* out) { return b ? builder_finish(b, out) : KDBX_ERR_ARG; }
int kdbx_builder_export(kdbx_builder* b, const kdbx_build_arrays* arrays) { return b ? builder_export(b, arrays) : KDBX_ERR_ARG; }

int64_t kdbx_debug_fetch(kdbx_ctx* ctx, int what, void* out, uint64_t max_elems) {
    if (!ctx || !out) return KDBX_ERR_ARG;
    if (!ctx->loaded || !ctx->nodes.p) return ctx->fail(KDBX_ERR_STATE, "nothing prepared yet");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return KDBX_ERR_CUDA;
    cudaStreamSynchronize(ctx->stream);
    const void* src = nullptr; uint64_t n = 0; size_t es = 4;
    switch (what) {
        case 0: src = ctx->W.p; n = ctx->P; es = 4; break;
        case 1: src = ctx->loc.p; n = ctx->sum_l; es = 4; break;
        case 2: src = ctx->loff.p; n = ctx->P + 1; es = 8; break;
        default: return ctx->fail(KDBX_ERR_ARG, "unknown debug tap %d", what);
    }
    n = std::min(n, max_elems);
    if (n && cudaMemcpy(out, src, n * es, cudaMemcpyDeviceToHost) != cudaSuccess)
        return ctx->fail(KDBX_ERR_CUDA, "debug fetch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return (int64_t)n;
}

}  // extern "C"
