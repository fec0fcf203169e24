// build: database construction on the device (included by kdbx.cu; shares its anonymous namespace).
//
// Replaces, for the `build` mode, what the reference runs on host threads per sample
// (src/console_build.cpp:91-118): KmerHelper::extract + MinHashFilter (src/kmer_extract.h:13-97,
// src/filter.h:40-115), the sort + unique of the sample's k-mers (ParallelSort / pdqsort; the dead
// parallel_sorter.cpp), the prefix histogram + hashtable find-or-insert (PrefixKmerDb::hashtableJobATP,
// src/prefix_kmer_db.cpp:67-178), the sort of (pattern id, slot) pairs and the extend-or-split of
// patterns (PrefixKmerDb::patternJob, :181-240).  The semantics are §A.4 of SURVEY.md; nothing of
// the reference's task-queue structure is kept:
//   * the k-mer -> pattern id map is ONE open-addressing table in HBM keyed by the whole 64-bit
//     k-mer (linear probing, fmix64); the reference's prefix-bucketed raw tables are produced once,
//     at the end, by a re-insertion kernel (same slot layout and hash as src/hashmap_lp.h, so the
//     result serialises into a .db the reference loads and kdbx_load_hashtables probes);
//   * a sample's windows are cut by one thread per position, sorted with a radix sort, made unique,
//     probed/inserted by one thread per k-mer; (pattern id, k-mer) pairs are radix-sorted by pattern
//     id; runs are found with a flag scan; one thread per run decides extend-or-split; new pattern
//     ids come from a prefix sum over the split flags (ascending old pattern id: deterministic);
//   * patterns do not carry growing Elias-gamma streams: an in-place extension appends one
//     (pattern, sample) event to a log; at the end the log is stably sorted by pattern and every
//     pattern's deltas are gamma-coded in one pass into the wire format of src/elias_gamma.h:104-128.
#pragma once

struct BuildPerSample {            // read back by the host after every sample
    unsigned long long inserted;   // k-mers new to the table
    unsigned long long num_runs;   // distinct patterns touched
    unsigned long long num_splits; // new patterns created
    unsigned long long unique_kmers;
    int err;
    int pad;
};

struct BuildAlphabet {
    int8_t map[256];
    uint32_t k, bits, size, preserve, shift;
    uint32_t accept_all;
    unsigned long long lo, hi, k_div_4, sentinel;
};

constexpr unsigned long long kEmptyKey = ~0ull;

__device__ __forceinline__ unsigned long long fmix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

// MinHashFilter::operator() (src/filter.h:96-115): one MurmurHash3-x64-128 block round, seeds 42
__device__ __forceinline__ unsigned long long minhash_of(unsigned long long x, unsigned long long k_div_4) {
    unsigned long long h = x * 0x87c37b91114253d5ull;
    h = (h << 31) | (h >> 33);
    h *= 0x4cf5ad432745937full;
    unsigned long long h1 = (42ull ^ h) ^ k_div_4;
    unsigned long long h2 = 42ull ^ k_div_4;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    return h1 ^ h2;
}

// One thread per window start.  A window is a k-mer iff its k symbols are all in the alphabet
// (records are separated by an out-of-alphabet byte, so windows never span records,
// src/genome_input_file.h:195-203).  Value = big-endian packing; canonical = min(forward, reverse
// complement) unless the alphabet preserves the strand (src/kmer_extract.h:80-85); then the
// >= 8-bit-prefix shift (:87-88) and the minhash filter on the shifted value (:90-92).
// Rejected windows get the sentinel (one bit above every k-mer), which the sort puts last.
__global__ void k_extract_kmers(const uint8_t* __restrict__ seq, uint64_t len, const BuildAlphabet* __restrict__ alpha,
                                unsigned long long* __restrict__ out) {
    __shared__ BuildAlphabet a;
    for (uint32_t i = threadIdx.x; i < sizeof(BuildAlphabet) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(&a)[i] = reinterpret_cast<const uint32_t*>(alpha)[i];
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    unsigned long long res = a.sentinel;
    if (i + a.k <= len) {
        unsigned long long fwd = 0, rev = 0;
        bool ok = true;
        for (uint32_t j = 0; j < a.k; ++j) {
            const int s = a.map[seq[i + j]];
            if (s < 0) { ok = false; break; }
            fwd = (fwd << a.bits) | (unsigned long long)s;
            rev |= (unsigned long long)(a.size - 1 - s) << (a.bits * j);
        }
        if (ok) {
            unsigned long long can = (a.preserve || fwd < rev) ? fwd : rev;
            if (a.shift) can = (can << a.shift) | (can & ((1ull << a.shift) - 1));
            bool keep = true;
            if (!a.accept_all) {
                const unsigned long long h = minhash_of(can, a.k_div_4);
                keep = h >= a.lo && h < a.hi;
            }
            if (keep) res = can;
        }
    }
    out[i] = res;
}

// kmers[0..count) must ascend strictly (kdbx_builder_add_kmers hands in caller data)
__global__ void k_check_sorted(const unsigned long long* __restrict__ kmers, uint64_t count, unsigned long long sentinel, int* __restrict__ err) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    if (kmers[i] >= sentinel || (i && kmers[i] <= kmers[i - 1])) atomicExch(err, 1);
}

// find-or-insert (src/prefix_kmer_db.cpp:139-165): a k-mer new to the database enters on pattern 0.
// The sample's k-mers are unique, so no two threads insert the same key; slots never change once
// claimed, so a stale read can only show "empty" for a taken slot, which the CAS then corrects.
__global__ void k_find_or_insert(const unsigned long long* __restrict__ kmers, uint64_t count, unsigned long long* keys,
                                 const uint32_t* __restrict__ vals, unsigned long long mask, unsigned long long* __restrict__ slot_of,
                                 uint32_t* __restrict__ pid_of, uint32_t* __restrict__ idx_of, BuildPerSample* __restrict__ ps) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool inserted = false;
    if (i < count) {
        const unsigned long long kmer = kmers[i];
        unsigned long long h = fmix64(kmer) & mask;
        for (;;) {
            const unsigned long long cur = keys[h];
            if (cur == kmer) break;
            if (cur == kEmptyKey) {
                const unsigned long long prev = atomicCAS(&keys[h], kEmptyKey, kmer);
                if (prev == kEmptyKey) { inserted = true; break; }
                if (prev == kmer) break;
            }
            h = (h + 1) & mask;
        }
        slot_of[i] = h;
        pid_of[i] = inserted ? 0u : vals[h];
        idx_of[i] = (uint32_t)i;
    }
    const unsigned m = __ballot_sync(0xffffffffu, inserted);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&ps->inserted, (unsigned long long)__popc(m));
}

__global__ void k_run_heads(const uint32_t* __restrict__ pid_sorted, uint64_t count, uint32_t* __restrict__ head) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    head[j] = (j == 0 || pid_sorted[j] != pid_sorted[j - 1]) ? 1u : 0u;
}
// run r starts at run_start[r]; run_start[num_runs] = count
__global__ void k_run_starts(const uint32_t* __restrict__ head, const uint32_t* __restrict__ head_incl, uint64_t count,
                             uint32_t* __restrict__ run_start, BuildPerSample* __restrict__ ps) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    if (head[j]) run_start[head_incl[j] - 1] = (uint32_t)j;
    if (j == count - 1) { run_start[head_incl[j]] = (uint32_t)count; ps->num_runs = head_incl[j]; }
}
// split[r] = 1 when the run needs a new pattern: the sample's k-mers on pattern q are not ALL of
// q's k-mers, or q already has children (src/prefix_kmer_db.cpp:210-216); entries past the last
// run are zeroed so that the scan over `count` items is well defined
__global__ void k_decide(const uint32_t* __restrict__ pid_sorted, const uint32_t* __restrict__ run_start, uint64_t count,
                         const BuildPerSample* __restrict__ ps, const long long* __restrict__ num_kmers,
                         const uint32_t* __restrict__ is_parent, uint32_t* __restrict__ split) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= count) return;
    uint32_t s = 0;
    if (r < ps->num_runs) {
        const uint32_t b = run_start[r], e = run_start[r + 1];
        const uint32_t q = pid_sorted[b];
        s = (num_kmers[q] == (long long)(e - b) && !is_parent[q]) ? 0u : 1u;
    }
    split[r] = s;
}
// extend in place (pattern_t::expand, src/pattern.h:195-203) or create the child (:216-226)
__global__ void k_apply_runs(const uint32_t* __restrict__ pid_sorted, const uint32_t* __restrict__ run_start, uint64_t count,
                             BuildPerSample* __restrict__ ps, const uint32_t* __restrict__ split, const uint32_t* __restrict__ split_incl,
                             uint32_t sample, uint64_t P, uint64_t ev_count, long long* __restrict__ num_kmers,
                             long long* __restrict__ parent, uint32_t* __restrict__ n, uint32_t* __restrict__ l,
                             uint32_t* __restrict__ last, uint32_t* __restrict__ born, uint32_t* __restrict__ is_parent,
                             uint32_t* __restrict__ ev_pat, uint32_t* __restrict__ ev_sample, uint32_t* __restrict__ run_tag) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= count) return;
    const uint64_t runs = ps->num_runs;
    if (r >= runs) return;
    const uint32_t b = run_start[r], e = run_start[r + 1];
    const uint32_t q = pid_sorted[b];
    const long long c = (long long)(e - b);
    const uint32_t splits_before = split_incl[r] - split[r];
    if (!split[r]) {
        last[q] = sample; n[q] += 1; l[q] += 1;
        const uint64_t at = ev_count + (r - splits_before);
        ev_pat[at] = q; ev_sample[at] = sample;
        run_tag[r] = 0xFFFFFFFFu;
    } else {
        const uint64_t child = P + splits_before;
        const uint32_t nq = n[q];
        num_kmers[child] = c;
        n[child] = nq + 1; l[child] = 1; last[child] = sample; born[child] = sample; is_parent[child] = 0;
        parent[child] = nq > 0 ? (long long)q : -1ll;   // children of pattern 0 are roots (src/pattern.h:106-114)
        if (nq > 0) is_parent[q] = 1;
        if (q != 0) num_kmers[q] -= c;
        run_tag[r] = (uint32_t)child;
    }
    if (r == runs - 1) ps->num_splits = split_incl[r];
}
// the k-mers of a split run now belong to the child (src/prefix_kmer_db.cpp:228-230)
__global__ void k_repoint(const uint32_t* __restrict__ head_incl, const uint32_t* __restrict__ idx_sorted, uint64_t count,
                          const uint32_t* __restrict__ run_tag, const unsigned long long* __restrict__ slot_of, uint32_t* __restrict__ vals) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const uint32_t t = run_tag[head_incl[j] - 1];
    if (t != 0xFFFFFFFFu) vals[slot_of[idx_sorted[j]]] = t;
}

__global__ void k_rehash(const unsigned long long* __restrict__ old_keys, const uint32_t* __restrict__ old_vals, uint64_t old_cap,
                         unsigned long long* keys, uint32_t* __restrict__ vals, unsigned long long mask) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= old_cap) return;
    const unsigned long long kmer = old_keys[i];
    if (kmer == kEmptyKey) return;
    unsigned long long h = fmix64(kmer) & mask;
    for (;;) {
        if (keys[h] == kEmptyKey && atomicCAS(&keys[h], kEmptyKey, kmer) == kEmptyKey) break;
        h = (h + 1) & mask;
    }
    vals[h] = old_vals[i];
}

// ---- finish: Elias-gamma payloads from the event log ---------------------------------------------
struct EventsOf {
    const uint32_t* l; uint64_t n;
    __host__ __device__ uint64_t operator()(uint64_t i) const { return (i < n && l[i] > 1) ? (uint64_t)(l[i] - 1) : 0ull; }
};
__device__ __forceinline__ uint32_t gamma_len(uint32_t v) { return 2u * (32u - (uint32_t)__clz((int)v)) - 1u; }

__global__ void k_gamma_bits(uint64_t P, const uint32_t* __restrict__ l, const uint32_t* __restrict__ born, const uint32_t* __restrict__ last,
                             const uint64_t* __restrict__ eoff, const uint32_t* __restrict__ ev_pat, const uint32_t* __restrict__ ev_sample,
                             uint32_t* __restrict__ bits, int* __restrict__ err) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    uint32_t nb = 0;
    if (l[p] > 1) {
        uint32_t prev = born[p];
        const uint64_t b = eoff[p], e = eoff[p + 1];
        for (uint64_t x = b; x < e; ++x) {
            const uint32_t id = ev_sample[x];
            if (ev_pat[x] != (uint32_t)p || id <= prev) { atomicExch(err, 7); break; }
            nb += gamma_len(id - prev);
            prev = id;
        }
        if (prev != last[p]) atomicExch(err, 7);
    } else if (l[p] == 1 && born[p] != last[p]) atomicExch(err, 7);
    bits[p] = nb;
}
// the code of v (b = bit length): b-1 ones, a zero, the low b-1 bits; MSB-first (src/elias_gamma.h:104-128).
// Every pattern owns its own words (2-word granules), so plain read-modify-write is safe.
__global__ void k_gamma_encode(uint64_t P, const uint32_t* __restrict__ l, const uint32_t* __restrict__ born,
                               const uint64_t* __restrict__ eoff, const uint32_t* __restrict__ ev_sample,
                               const uint64_t* __restrict__ poff, unsigned long long* __restrict__ payload) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P || l[p] <= 1) return;
    unsigned long long* w = payload + poff[p];
    uint32_t pos = 0, prev = born[p];
    unsigned long long cur = 0;      // word being filled
    uint32_t wi = 0;
    for (uint64_t x = eoff[p]; x < eoff[p + 1]; ++x) {
        const uint32_t id = ev_sample[x], v = id - prev;
        prev = id;
        const uint32_t b = 32u - (uint32_t)__clz((int)v), len = 2 * b - 1;
        const unsigned long long ones = b > 1 ? ((1ull << (b - 1)) - 1) : 0ull;
        const unsigned long long code = (ones << b) | (unsigned long long)(v - (1u << (b - 1)));
        const uint32_t used = pos & 63, room = 64 - used;
        if (len <= room) {
            cur |= code << (room - len);
            if (len == room) { w[wi++] = cur; cur = 0; }
        } else {
            const uint32_t rest = len - room;
            cur |= code >> rest;
            w[wi++] = cur;
            cur = code << (64 - rest);
        }
        pos += len;
    }
    if (pos & 63) w[wi] = cur;
}

// ---- finish: the reference's prefix-bucketed raw tables -------------------------------------------
__global__ void k_prefix_count(const unsigned long long* __restrict__ keys, uint64_t cap, uint64_t num_tables,
                               unsigned long long* __restrict__ filled, int* __restrict__ err) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const unsigned long long kmer = keys[i];
    if (kmer == kEmptyKey) return;
    const unsigned long long prefix = kmer >> 32;
    if (prefix >= num_tables) { atomicExch(err, 8); return; }
    atomicAdd(&filled[prefix], 1ull);
}
// capacity: the smallest power of two >= 16 with filled <= 0.8 * capacity (src/hashmap_lp.h:427-464)
struct TableCapacity {
    const unsigned long long* filled; uint64_t n;
    __host__ __device__ uint64_t operator()(uint64_t t) const {
        if (t >= n) return 0ull;
        uint64_t cap = 16;
        while ((double)filled[t] > (double)cap * 0.8) cap *= 2;
        return cap;
    }
};
__global__ void k_fill_u64(unsigned long long* __restrict__ p, uint64_t count, unsigned long long v) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[i] = v;
}
__global__ void k_export_tables(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t cap,
                                const uint64_t* __restrict__ slot_off, unsigned long long* slots) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const unsigned long long kmer = keys[i];
    if (kmer == kEmptyKey) return;
    const uint64_t t = kmer >> 32;
    const uint32_t suffix = (uint32_t)kmer;
    const uint64_t off = slot_off[t], mask = slot_off[t + 1] - off - 1;
    const unsigned long long empty = 0x7FFFFFFFull << 32;  // {key 0, val INT32_MAX}
    const unsigned long long item = ((unsigned long long)vals[i] << 32) | suffix;
    uint64_t h = fmix32(suffix) & mask;
    for (;;) {
        if (atomicCAS(&slots[off + h], empty, item) == empty) break;
        h = (h + 1) & mask;
    }
}

// ---- adopt: continue an existing database (build -extend) ------------------------------------------
__global__ void k_adopt_patterns(uint64_t P, const Node* __restrict__ nodes, const uint32_t* __restrict__ loc,
                                 const uint64_t* __restrict__ eoff, uint32_t* __restrict__ born, uint32_t* __restrict__ is_parent,
                                 uint32_t* __restrict__ ev_pat, uint32_t* __restrict__ ev_sample) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const Node nd = nodes[p];
    born[p] = nd.l ? loc[nd.loff] : 0u;
    if (nd.parent >= 0) is_parent[nd.parent] = 1u;   // benign race: everyone writes 1
    uint64_t at = eoff[p];
    for (uint32_t j = 1; j < nd.l; ++j, ++at) { ev_pat[at] = (uint32_t)p; ev_sample[at] = loc[nd.loff + j]; }
}
__global__ void k_adopt_tables(uint64_t total_slots, uint64_t num_tables, const uint64_t* __restrict__ slot_off,
                               const unsigned long long* __restrict__ slots, unsigned long long* keys, uint32_t* __restrict__ vals,
                               unsigned long long mask, uint64_t P, BuildPerSample* __restrict__ ps) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool inserted = false;
    if (i < total_slots) {
        const unsigned long long s = slots[i];
        const uint32_t val = (uint32_t)(s >> 32);
        if (val != 0x7FFFFFFFu) {
            uint64_t lo = 0, hi = num_tables;   // slot_off[lo] <= i < slot_off[hi]
            while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (slot_off[mid] <= i) lo = mid; else hi = mid; }
            if ((uint64_t)val >= P) atomicExch(&ps->err, 6);
            else {
                const unsigned long long kmer = ((unsigned long long)lo << 32) | (uint32_t)s;
                unsigned long long h = fmix64(kmer) & mask;
                for (;;) {
                    if (keys[h] == kEmptyKey && atomicCAS(&keys[h], kEmptyKey, kmer) == kEmptyKey) break;
                    h = (h + 1) & mask;
                }
                vals[h] = val;
                inserted = true;
            }
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, inserted);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&ps->inserted, (unsigned long long)__popc(m));
}

// ---- extraction parameters from the caller's description (shared by the builder and the query path) ----
int make_build_alphabet(kdbx_ctx* ctx, const kdbx_build_params* p, const char* who, BuildAlphabet& a, uint32_t& sentinel_bit,
                        uint64_t& num_tables) {
    if (!p) return ctx->fail(KDBX_ERR_ARG, "%s: parameters are NULL", who);
    if (p->kmer_length == 0 || p->bits_per_symbol == 0 || p->bits_per_symbol > 8 || p->alphabet_size < 2 ||
        p->alphabet_size > (1u << p->bits_per_symbol))
        return ctx->fail(KDBX_ERR_ARG, "%s: bad alphabet / k-mer length", who);
    const int kb = (int)p->kmer_length * (int)p->bits_per_symbol;
    const int prefix_bits = kb - 32;
    const uint32_t shift = prefix_bits < 8 ? (uint32_t)(8 - prefix_bits) : 0u;   // src/kmer_extract.h:36-45
    if (kb > 62 || kb + (int)shift > 62) return ctx->fail(KDBX_ERR_ARG, "%s: k-mer does not fit 62 bits", who);
    const int table_bits = prefix_bits < 8 ? 8 : prefix_bits;                     // src/prefix_kmer_db.cpp:54-62
    if (table_bits > 24) return ctx->fail(KDBX_ERR_ARG, "%s: 2^%d prefix tables are not supported on the device", who, table_bits);
    if (!(p->fraction > 0.0)) return ctx->fail(KDBX_ERR_ARG, "%s: fraction must be positive", who);
    std::memcpy(a.map, p->symbol_map, 256);
    a.k = p->kmer_length; a.bits = p->bits_per_symbol; a.size = p->alphabet_size;
    a.preserve = p->preserve_strand ? 1u : 0u; a.shift = shift;
    a.accept_all = !(p->fraction < 1.0) ? 1u : 0u;                                // NullFilter, src/filter.h:120-145
    a.lo = 0; a.hi = ~0ull;
    if (!a.accept_all) {                                                          // src/filter.h:40-51
        const double top = (double)UINT64_MAX;
        a.lo = (unsigned long long)(top * p->fraction_start);
        const double hi = top * (p->fraction_start + p->fraction);
        a.hi = hi >= top ? ~0ull : (unsigned long long)hi;
    }
    a.k_div_4 = (p->kmer_length + 3) / 4;
    sentinel_bit = (uint32_t)kb + shift;
    a.sentinel = 1ull << sentinel_bit;
    num_tables = 1ull << table_bits;
    return KDBX_OK;
}

__global__ void k_unique_count(const unsigned long long* __restrict__ uniq, const unsigned long long* __restrict__ nsel,
                               unsigned long long sentinel, unsigned long long* __restrict__ out) {
    unsigned long long c = *nsel;
    if (c && uniq[c - 1] >= sentinel) --c;
    *out = c;
}

// new2all from the queries' SEQUENCES: what New2AllConsole's loader threads do per query on the host
// (KmerHelper::extract + MinHashFilter + sort + unique, src/console_new2all.cpp:64-94, src/kmer_extract.h:13-119)
// happens here on the device, straight into the k-mer buffer the probe kernel reads: the host uploads
// 1 byte per base instead of 8 bytes per k-mer and does no sorting.  Queries are taken in sub-batches
// bounded by the k-mer budget of the probe pipeline.
int new2all_sequences_impl(kdbx_ctx* ctx, const kdbx_build_params* params, const char* symbols, const uint64_t* q_off,
                           uint32_t n_queries, uint32_t* out, uint64_t* unique_kmers, kdbx_stats* stats) {
    if (!ctx->loaded) return ctx->fail(KDBX_ERR_STATE, "no patterns loaded (call kdbx_load_patterns first)");
    if (!ctx->tables_loaded) return ctx->fail(KDBX_ERR_STATE, "no k-mer tables loaded (call kdbx_load_hashtables first)");
    if (int rcw = require_full_window(ctx, "kdbx_new2all_sequences")) return rcw;
    if (n_queries && (!q_off || !out)) return ctx->fail(KDBX_ERR_ARG, "kdbx_new2all_sequences: NULL argument");
    for (uint32_t q = 0; q < n_queries; ++q)
        if (q_off[q + 1] < q_off[q]) return ctx->fail(KDBX_ERR_ARG, "kdbx_new2all_sequences: q_off must be non-decreasing");
    if (n_queries && q_off[n_queries] > q_off[0] && !symbols) return ctx->fail(KDBX_ERR_ARG, "kdbx_new2all_sequences: symbols is NULL");
    BuildAlphabet alpha{};
    uint32_t sentinel_bit = 0;
    uint64_t num_tables = 0;
    if (int rc = make_build_alphabet(ctx, params, "kdbx_new2all_sequences", alpha, sentinel_bit, num_tables)) return rc;
    if (num_tables != ctx->num_tables)
        return ctx->fail(KDBX_ERR_ARG, "kdbx_new2all_sequences: the database has %llu k-mer tables, these parameters imply %llu",
                         (unsigned long long)ctx->num_tables, (unsigned long long)num_tables);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t N = ctx->N;
    kdbx_stats s{};
    uint32_t launches = 0;
    ctx->ev_used = 0;
    cudaEvent_t ev0 = ctx->event();
    if (!ctx->prepared) {  // decoded local lists + nodes, shared with all2all
        Plan pl;
        if (int rc = make_plan(ctx, pl)) return rc;
        const int rc = prepare(ctx, pl, launches);
        if (rc < 0) return rc;
        if (int rc2 = check_device_error(ctx)) return rc2;
    }
    CK(cudaMemsetAsync(ctx->err_flag.p, 0, 16, st));   // stale probe errors of earlier calls (see new2all_impl)
    cudaEvent_t ev1 = ctx->event();
    if (n_queries == 0 || N == 0) { if (stats) *stats = s; return KDBX_OK; }
    CK(ctx->counters.ensure(64));
    CK(ctx->qx_alpha.ensure(sizeof(BuildAlphabet))); CK(ctx->qx_count.ensure(16));
    CK(cudaMemcpyAsync(ctx->qx_alpha.p, &alpha, sizeof alpha, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));   // `alpha` lives on this stack frame

    const uint64_t max_kmers = ctx->cfg.query_batch_kmers ? ctx->cfg.query_batch_kmers : ((uint64_t)1 << 28);
    const uint64_t max_rows = std::max<uint64_t>(1, ((uint64_t)1 << 30) / std::max<uint32_t>(1, N));
    float ms_probe = 0.f, ms_scatter = 0.f, ms_download = 0.f, ms_extract = 0.f;
    std::vector<uint64_t> koff;   // k-mer offsets of the sub-batch's queries
    uint32_t q0 = 0;
    while (q0 < n_queries) {
        // a sub-batch: whole queries while their WINDOWS (an upper bound of their k-mers) fit the budget
        uint32_t q1 = q0 + 1;
        while (q1 < n_queries && q_off[q1 + 1] - q_off[q0] <= max_kmers && (uint64_t)(q1 + 1 - q0) <= max_rows) ++q1;
        const uint64_t sym0 = q_off[q0], nsym = q_off[q1] - sym0;
        CK(ctx->qx_seq.ensure(nsym + 16)); CK(ctx->q_kmers.ensure((nsym + 1) * 8));
        if (nsym) CK(cudaMemcpyAsync(ctx->qx_seq.p, symbols + sym0, nsym, cudaMemcpyHostToDevice, st));
        koff.assign(1, 0);
        cudaEvent_t xa = ctx->event();
        for (uint32_t q = q0; q < q1; ++q) {
            const uint64_t len = q_off[q + 1] - q_off[q];
            unsigned long long cnt = 0;
            if (len >= alpha.k) {
                CK(ctx->qx_raw.ensure(len * 8)); CK(ctx->qx_sorted.ensure(len * 8));
                k_extract_kmers<<<blocks_for(len, 256), 256, 0, st>>>(ctx->qx_seq.as<uint8_t>() + (q_off[q] - sym0), len,
                                                                      ctx->qx_alpha.as<BuildAlphabet>(), ctx->qx_raw.as<unsigned long long>());
                size_t tmp = 0;
                const int end_bit = (int)sentinel_bit + 1;
                CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp, ctx->qx_raw.as<unsigned long long>(), ctx->qx_sorted.as<unsigned long long>(), len, 0, end_bit, st));
                CK(ctx->cub_tmp.ensure(tmp));
                CK(cub::DeviceRadixSort::SortKeys(ctx->cub_tmp.p, tmp, ctx->qx_raw.as<unsigned long long>(), ctx->qx_sorted.as<unsigned long long>(), len, 0, end_bit, st));
                // the query's distinct k-mers land behind those of the queries before it (<= their windows, so they fit)
                unsigned long long* dst = ctx->q_kmers.as<unsigned long long>() + koff.back();
                unsigned long long* nsel = ctx->qx_count.as<unsigned long long>();
                CK(cub::DeviceSelect::Unique(nullptr, tmp, ctx->qx_sorted.as<unsigned long long>(), dst, nsel, len, st));
                CK(ctx->cub_tmp.ensure(tmp));
                CK(cub::DeviceSelect::Unique(ctx->cub_tmp.p, tmp, ctx->qx_sorted.as<unsigned long long>(), dst, nsel, len, st));
                k_unique_count<<<1, 1, 0, st>>>(dst, nsel, alpha.sentinel, nsel + 1);
                CK(cudaMemcpyAsync(&cnt, nsel + 1, 8, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                launches += 5;
            }
            if (unique_kmers) unique_kmers[q] = cnt;
            koff.push_back(koff.back() + cnt);
        }
        cudaEvent_t xb = ctx->event();
        const uint32_t nq = q1 - q0;
        CK(ctx->q_off.ensure(((size_t)nq + 1) * 8));
        CK(cudaMemcpyAsync(ctx->q_off.p, koff.data(), ((size_t)nq + 1) * 8, cudaMemcpyHostToDevice, st));
        if (int rc = new2all_core(ctx, koff.back(), 0, 0, nq, out + (size_t)q0 * N, s, ms_probe, ms_scatter, ms_download, launches)) return rc;
        ms_extract += elapsed(xa, xb);
        q0 = q1;
    }
    cudaEvent_t ev2 = ctx->event();
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (int rc = finish_upload(ctx)) return rc;
    s.ms_prepare = elapsed(ev0, ev1);
    s.ms_probe = ms_probe; s.ms_scatter = ms_scatter; s.ms_download = ms_download;
    s.ms_expand = ms_extract;   // reported in the "expand" slot: k-mer extraction + sort + unique of the queries
    s.ms_total = elapsed(ev0, ev2);
    s.kernel_launches = launches;
    s.local_ids = ctx->sum_l; s.flat_ids = ctx->sum_n;
    if (stats) *stats = s;
    return KDBX_OK;
}
