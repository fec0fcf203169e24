// new2all: query samples against the database (included by kdbx.cu; shares its anonymous namespace).
//
// Replaces SimilarityCalculator::one2all<false> / one2all_sp as New2AllConsole calls them once per
// query from its worker threads (src/similarity_calculator.cpp:810-925,929-1051;
// src/console_new2all.cpp:64-94).  Per query the reference (1) probes the prefix bucket's
// hash_map_lp for every k-mer (prefetch distance 48, :825-852), (2) counts hits per pattern in an
// unordered_map, (3) decodes each distinct pattern's full sample list and adds the count to
// similarities[id] (:890-919).  Here a whole batch of queries goes through three device stages:
//   k_probe          one thread per k-mer: linear probing in the raw slot arrays staged in HBM
//                    (same hash, same probe sequence); emits key = (query << 32 | pattern id)
//   radix sort + RLE (CUB) the per-(query, pattern) hit counts — the unordered_map
//   k_query_scatter  8 lanes per (query, pattern, count): walk the parent chain over the decoded
//                    local lists (shared with all2all's prepare stage) and red.global.add the
//                    count into out[query][sample]; a query's row (4 N bytes) stays L2 resident.
#pragma once


__device__ __forceinline__ uint32_t fmix32(uint32_t h) {  // murmur3 finalizer, the tables' hash (src/hashmap_lp.h:52-64)
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

constexpr unsigned long long kMissKey = ~0ull;

__global__ void k_probe(uint64_t count, const uint64_t* __restrict__ kmers, uint64_t kmer_base, const uint64_t* __restrict__ q_off,
                        uint32_t q_begin, uint32_t q_end, uint64_t num_tables, const uint64_t* __restrict__ slot_off,
                        const uint64_t* __restrict__ slots, const int64_t* __restrict__ num_kmers, uint64_t P,
                        unsigned long long* __restrict__ keys, unsigned long long* __restrict__ hits, int* __restrict__ err) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long key = kMissKey;
    if (i < count) {
        const uint64_t kmer = kmers[i];
        const uint64_t prefix = kmer >> 32;
        const uint32_t suffix = (uint32_t)kmer;
        if (prefix < num_tables) {
            const uint64_t off = slot_off[prefix];
            const uint64_t mask = slot_off[prefix + 1] - off - 1;
            uint64_t h = fmix32(suffix) & mask;
            for (uint64_t step = 0; step <= mask; ++step) {
                const uint64_t s = slots[off + h];
                const uint32_t val = (uint32_t)(s >> 32);
                if (val == 0x7FFFFFFFu) break;  // empty slot: not in the database
                if ((uint32_t)s == suffix) {
                    if ((uint64_t)val >= P) { atomicExch(err, 6); break; }
                    if (num_kmers[val] != 0) {  // (src/similarity_calculator.cpp:846-847)
                        // query of global k-mer index kmer_base + i: last q with q_off[q] <= index
                        const uint64_t g = kmer_base + i;
                        uint32_t lo = q_begin, hi = q_end;  // q_off[lo] <= g < q_off[hi]
                        while (hi - lo > 1) {
                            const uint32_t mid = (lo + hi) >> 1;
                            if (q_off[mid] <= g) lo = mid; else hi = mid;
                        }
                        key = ((unsigned long long)(lo - q_begin) << 32) | val;
                    }
                    break;
                }
                h = (h + 1) & mask;
            }
        }
    }
    const unsigned found = __ballot_sync(0xffffffffu, key != kMissKey);
    if ((threadIdx.x & 31) == 0 && found) atomicAdd(hits, (unsigned long long)__popc(found));
    if (i < count) keys[i] = key;
}

constexpr uint32_t kQueryLanes = 8;
__global__ void k_query_scatter(const int* __restrict__ num_runs, const unsigned long long* __restrict__ run_keys,
                                const uint32_t* __restrict__ run_counts, const Node* __restrict__ nodes,
                                const uint32_t* __restrict__ loc, uint32_t N, uint32_t* __restrict__ out) {
    const uint64_t gid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kQueryLanes;
    const uint64_t ng = ((uint64_t)gridDim.x * blockDim.x) / kQueryLanes;
    const uint32_t sub = threadIdx.x & (kQueryLanes - 1);
    const uint64_t runs = (uint64_t)*num_runs;
    for (uint64_t r = gid; r < runs; r += ng) {
        const unsigned long long key = run_keys[r];
        if (key == kMissKey) continue;
        const uint32_t cnt = run_counts[r];
        uint32_t* row = out + (size_t)(key >> 32) * N;
        Node nd = nodes[(uint32_t)key];
        for (;;) {
            const bool more = nd.parent >= 0;
            Node up = nd;
            if (more) up = nodes[nd.parent];
            const uint32_t* src = loc + nd.loff;
            for (uint32_t j = sub; j < nd.l; j += kQueryLanes) atomicAdd(&row[src[j]], cnt);
            if (!more) break;
            nd = up;
        }
    }
}

// Probe, count and scatter the k-mers ctx->q_kmers[0 .. count) of the queries [q_begin, q_end): query q owns the
// k-mers with global index in [q_off[q], q_off[q+1]) (ctx->q_off, device), the first of the batch being kmer_base.
// The rows of the batch go to out_rows (HOST memory, (q_end - q_begin) x N).
int new2all_core(kdbx_ctx* ctx, uint64_t count, uint64_t kmer_base, uint32_t q_begin, uint32_t q_end, uint32_t* out_rows,
                 kdbx_stats& s, float& ms_probe, float& ms_scatter, float& ms_download, uint32_t& launches) {
    cudaStream_t st = ctx->stream;
    const uint32_t N = ctx->N;
    const uint32_t nq = q_end - q_begin;
    const uint64_t base = kmer_base;
    const uint32_t q0 = q_begin, q1 = q_end;
    {
        CK(ctx->q_out.ensure((size_t)nq * N * 4 + 16));
        CK(cudaMemsetAsync(ctx->q_out.p, 0, (size_t)nq * N * 4, st));
        if (count) {
            CK(ctx->q_keys.ensure(count * 8)); CK(ctx->q_keys2.ensure(count * 8));
            CK(ctx->q_runkeys.ensure(count * 8)); CK(ctx->q_runcnt.ensure(count * 4));
            CK(cudaMemsetAsync(ctx->counters.p, 0, 64, st));
            unsigned long long* d_hits = ctx->counters.as<unsigned long long>();
            int* d_runs = ctx->counters.as<int>() + 4;
            cudaEvent_t a = ctx->event();
            k_probe<<<blocks_for(count, 256), 256, 0, st>>>(count, ctx->q_kmers.as<uint64_t>(), base, ctx->q_off.as<uint64_t>(), q0, q1,
                                                             ctx->num_tables, ctx->slot_off.as<uint64_t>(), ctx->slots.as<uint64_t>(),
                                                             ctx->num_kmers.as<int64_t>(), ctx->P, ctx->q_keys.as<unsigned long long>(), d_hits,
                                                             ctx->err_flag.as<int>());
            cudaEvent_t b = ctx->event();
            int pid_bits = 1; while (pid_bits < 32 && (ctx->P >> pid_bits)) ++pid_bits;
            int q_bits = 1; while (q_bits < 32 && (nq >> q_bits)) ++q_bits;
            // keys of misses are all ones: sorting the low 32+q_bits bits still puts them last within
            // their (truncated) query, and k_query_scatter skips them by value
            size_t tmp = 0;
            CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp, ctx->q_keys.as<unsigned long long>(), ctx->q_keys2.as<unsigned long long>(), count, 0, 64, st));
            CK(ctx->cub_tmp.ensure(tmp));
            CK(cub::DeviceRadixSort::SortKeys(ctx->cub_tmp.p, tmp, ctx->q_keys.as<unsigned long long>(), ctx->q_keys2.as<unsigned long long>(), count, 0, 64, st));
            (void)pid_bits; (void)q_bits;
            CK(cub::DeviceRunLengthEncode::Encode(nullptr, tmp, ctx->q_keys2.as<unsigned long long>(), ctx->q_runkeys.as<unsigned long long>(),
                                                  ctx->q_runcnt.as<uint32_t>(), d_runs, count, st));
            CK(ctx->cub_tmp.ensure(tmp));
            CK(cub::DeviceRunLengthEncode::Encode(ctx->cub_tmp.p, tmp, ctx->q_keys2.as<unsigned long long>(), ctx->q_runkeys.as<unsigned long long>(),
                                                  ctx->q_runcnt.as<uint32_t>(), d_runs, count, st));
            k_query_scatter<<<ctx->sm_count * 16, 256, 0, st>>>(d_runs, ctx->q_runkeys.as<unsigned long long>(), ctx->q_runcnt.as<uint32_t>(),
                                                                 ctx->nodes.as<Node>(), ctx->loc.as<uint32_t>(), N, ctx->q_out.as<uint32_t>());
            cudaEvent_t c = ctx->event();
            launches += 8;
            unsigned long long h_hits = 0;
            CK(cudaMemcpyAsync(&h_hits, d_hits, 8, cudaMemcpyDeviceToHost, st));
            cudaEvent_t d0 = ctx->event();
            CK(cudaMemcpyAsync(out_rows, ctx->q_out.p, (size_t)nq * N * 4, cudaMemcpyDeviceToHost, st));
            cudaEvent_t d1 = ctx->event();
            CK(cudaStreamSynchronize(st));
            ms_probe += elapsed(a, b); ms_scatter += elapsed(b, c); ms_download += elapsed(d0, d1);
            s.hits += h_hits; s.probes += count;
        } else {
            CK(cudaMemcpyAsync(out_rows, ctx->q_out.p, (size_t)nq * N * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        }
    }
    return check_device_error(ctx);
}

int new2all_impl(kdbx_ctx* ctx, const uint64_t* kmers, const uint64_t* q_off, uint32_t n_queries, uint32_t* out, kdbx_stats* stats) {
    if (!ctx->loaded) return ctx->fail(KDBX_ERR_STATE, "no patterns loaded (call kdbx_load_patterns first)");
    if (!ctx->tables_loaded) return ctx->fail(KDBX_ERR_STATE, "no k-mer tables loaded (call kdbx_load_hashtables first)");
    if (int rc = require_full_window(ctx, "kdbx_new2all_batch")) return rc;
    if (n_queries && (!q_off || !out)) return ctx->fail(KDBX_ERR_ARG, "kdbx_new2all_batch: NULL argument");
    for (uint32_t q = 0; q < n_queries; ++q)
        if (q_off[q + 1] < q_off[q]) return ctx->fail(KDBX_ERR_ARG, "kdbx_new2all_batch: q_off must be non-decreasing");
    if (n_queries && q_off[n_queries] > q_off[0] && !kmers) return ctx->fail(KDBX_ERR_ARG, "kdbx_new2all_batch: kmers is NULL");
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
    // errors of earlier calls (a table that pointed at a missing pattern, since reloaded) must not fail this one
    CK(cudaMemsetAsync(ctx->err_flag.p, 0, 16, st));
    cudaEvent_t ev1 = ctx->event();
    if (n_queries == 0 || N == 0) { if (stats) *stats = s; return KDBX_OK; }

    CK(ctx->q_off.ensure(((size_t)n_queries + 1) * 8));
    CK(cudaMemcpyAsync(ctx->q_off.p, q_off, ((size_t)n_queries + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(ctx->counters.ensure(64));
    // sub-batches: whole queries, bounded by k-mers (sort buffers) and by output rows
    const uint64_t max_kmers = ctx->cfg.query_batch_kmers ? ctx->cfg.query_batch_kmers : ((uint64_t)1 << 28);
    const uint64_t max_rows = std::max<uint64_t>(1, ((uint64_t)1 << 30) / std::max<uint32_t>(1, N));  // <= 4 GB of output rows
    uint32_t q0 = 0;
    float ms_probe = 0.f, ms_scatter = 0.f, ms_download = 0.f;
    while (q0 < n_queries) {
        uint32_t q1 = q0 + 1;
        while (q1 < n_queries && q_off[q1 + 1] - q_off[q0] <= max_kmers && (uint64_t)(q1 + 1 - q0) <= max_rows) ++q1;
        const uint64_t base = q_off[q0], count = q_off[q1] - base;
        if (count) {
            CK(ctx->q_kmers.ensure(count * 8));
            CK(cudaMemcpyAsync(ctx->q_kmers.p, kmers + base, count * 8, cudaMemcpyHostToDevice, st));
        }
        if (int rc = new2all_core(ctx, count, base, q0, q1, out + (size_t)q0 * N, s, ms_probe, ms_scatter, ms_download, launches)) return rc;
        q0 = q1;
    }
    cudaEvent_t ev2 = ctx->event();
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (int rc = finish_upload(ctx)) return rc;
    s.ms_prepare = elapsed(ev0, ev1);
    s.ms_probe = ms_probe; s.ms_scatter = ms_scatter; s.ms_download = ms_download;
    s.ms_total = elapsed(ev0, ev2);
    s.kernel_launches = launches;
    s.local_ids = ctx->sum_l; s.flat_ids = ctx->sum_n;
    if (stats) *stats = s;
    return KDBX_OK;
}

int load_hashtables_impl(kdbx_ctx* ctx, const kdbx_tables_view* v) {
    if (!v || !v->slot_off || v->num_tables == 0) return ctx->fail(KDBX_ERR_ARG, "kdbx_load_hashtables: bad view");
    const uint64_t T = v->num_tables;
    if (v->slot_off[0] != 0) return ctx->fail(KDBX_ERR_ARG, "kdbx_load_hashtables: slot_off[0] must be 0");
    for (uint64_t t = 0; t < T; ++t) {
        if (v->slot_off[t + 1] <= v->slot_off[t]) return ctx->fail(KDBX_ERR_ARG, "kdbx_load_hashtables: table %llu is empty", (unsigned long long)t);
        const uint64_t size = v->slot_off[t + 1] - v->slot_off[t];
        if (size & (size - 1)) return ctx->fail(KDBX_ERR_ARG, "kdbx_load_hashtables: size of table %llu is not a power of two", (unsigned long long)t);
    }
    const uint64_t total = v->slot_off[T];
    if (!v->slots) return ctx->fail(KDBX_ERR_ARG, "kdbx_load_hashtables: slots is NULL");
    CK(cudaSetDevice(ctx->device));
    ctx->tables_loaded = false;
    if (ctx->err_flag.p) CK(cudaMemsetAsync(ctx->err_flag.p, 0, 16, ctx->stream));
    CK(ctx->slot_off.ensure((T + 1) * 8)); CK(ctx->slots.ensure(total * 8));
    ctx->ev_used = 0;
    cudaEvent_t a = ctx->event();
    CK(cudaMemcpyAsync(ctx->slot_off.p, v->slot_off, (T + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->slots.p, v->slots, total * 8, cudaMemcpyHostToDevice, ctx->stream));
    cudaEvent_t b = ctx->event();
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->ms_upload_tables = elapsed(a, b);
    ctx->num_tables = T;
    ctx->total_slots = total;
    ctx->tables_loaded = true;
    return KDBX_OK;
}
