// Database against database (included by kdbx.cu; shares its anonymous namespace): one cell of the grid of
// all2all-parts.
//
// Replaces SimilarityCalculator::db2db_sp + SparseMatrix::compact2 (src/similarity_calculator.cpp:1225-1540,
// src/array.h:391-446) as All2AllPartsConsole calls them for the cells (row part, column part) below the diagonal
// (src/console_all2all_parts.cpp:163-254).  The reference (1) sorts the (suffix, pattern) pairs of every prefix
// bucket of both databases and merges them to find the k-mers both hold, (2) sorts the (pattern of db1, pattern of
// db2) pairs and counts the repeats, (3) decodes the two full sample lists of every distinct pair and adds the
// count to matrix[s1][s2] for every s1 of the first and s2 of the second list (hash map per row, bubbles for the
// largest pairs).  Here:
//   k_match_tables   one thread per slot of the ROW database's raw tables: the k-mer's suffix is looked up in the
//                    same prefix bucket of the COLUMN database (same hash, same probe sequence as k_probe); hits
//                    are appended as key = (pattern1 << 32 | pattern2) with one atomic per warp
//   radix sort + RLE (CUB) the per-pair counts
//   k_pair_scatter   one warp per distinct pair: both full lists are gathered tile by tile into shared memory by
//                    walking the parent chains over the decoded local lists (shared with all2all's prepare
//                    stage), and the lanes go over the tile's (row, column) pairs with red.global.add into a dense
//                    block of rows x N2 cells in HBM
//   k_rect_count / scan / k_rect_fill   filter (CombinedFilter with the row database's and the column database's
//                    k-mer counts) and compaction to ascending (col, val) pairs, like sparse.cuh
// Bubbles have no analogue (dense accumulator).
#pragma once

__global__ void k_match_tables(uint64_t total_slots, uint64_t num_tables, const uint64_t* __restrict__ slot_off1,
                               const uint64_t* __restrict__ slots1, uint64_t P1, const uint64_t* __restrict__ slot_off2,
                               const uint64_t* __restrict__ slots2, uint64_t P2, unsigned long long* __restrict__ keys,
                               unsigned long long* __restrict__ counters /* [0] matched, [1] non-empty slots */, int* __restrict__ err) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long key = kMissKey;
    bool used = false;
    if (i < total_slots) {
        const uint64_t s1 = slots1[i];
        const uint32_t pid1 = (uint32_t)(s1 >> 32);
        if (pid1 != 0x7FFFFFFFu) {
            used = true;
            if ((uint64_t)pid1 >= P1) atomicExch(err, 6);
            else {
                // prefix bucket of slot i: last t with slot_off1[t] <= i
                uint64_t lo = 0, hi = num_tables;
                while (hi - lo > 1) {
                    const uint64_t mid = (lo + hi) >> 1;
                    if (slot_off1[mid] <= i) lo = mid; else hi = mid;
                }
                const uint32_t suffix = (uint32_t)s1;
                const uint64_t off = slot_off2[lo];
                const uint64_t mask = slot_off2[lo + 1] - off - 1;
                uint64_t h = fmix32(suffix) & mask;
                for (uint64_t step = 0; step <= mask; ++step) {
                    const uint64_t s2 = slots2[off + h];
                    const uint32_t pid2 = (uint32_t)(s2 >> 32);
                    if (pid2 == 0x7FFFFFFFu) break;
                    if ((uint32_t)s2 == suffix) {
                        if ((uint64_t)pid2 >= P2) atomicExch(err, 6);
                        else key = ((unsigned long long)pid1 << 32) | pid2;
                        break;
                    }
                    h = (h + 1) & mask;
                }
            }
        }
    }
    const uint32_t lane = threadIdx.x & 31;
    const unsigned found = __ballot_sync(0xffffffffu, key != kMissKey);
    const unsigned nonempty = __ballot_sync(0xffffffffu, used);
    unsigned long long base = 0;
    if (lane == 0) {
        if (found) base = atomicAdd(&counters[0], (unsigned long long)__popc(found));
        if (nonempty) atomicAdd(&counters[1], (unsigned long long)__popc(nonempty));
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (key != kMissKey) keys[base + __popc(found & ((1u << lane) - 1u))] = key;
}

// Gathers into buf (shared memory of the calling warp) the ids of the full list of pattern pid that lie in [lo, hi),
// starting at logical position `pos` of the list taken in chain order (the pattern's own local ids first, then its
// parent's, ...), until buf holds cap ids or the list ends.  Returns the number of ids gathered; `pos` advances to the
// position to continue from (== n when the list is exhausted).
__device__ __forceinline__ uint32_t gather_chain(const Node* __restrict__ nodes, const uint32_t* __restrict__ loc, uint32_t pid,
                                                 uint32_t& pos, uint32_t lo, uint32_t hi, uint32_t* buf, uint32_t cap, uint32_t lane) {
    uint32_t kept = 0, base = 0;
    Node nd = nodes[pid];
    for (;;) {
        const uint32_t seg_end = base + nd.l;
        while (pos < seg_end && kept + 32u <= cap) {   // (a round may keep up to 32 ids)
            const uint32_t j = pos - base + lane;
            uint32_t id = 0;
            bool ok = false;
            if (j < nd.l) { id = loc[nd.loff + j]; ok = id >= lo && id < hi; }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) buf[kept + __popc(m & ((1u << lane) - 1u))] = id;
            kept += __popc(m);
            pos = min(seg_end, pos + 32u);
        }
        if (pos < seg_end || nd.parent < 0) break;   // buffer full, or the root's ids are done
        base = seg_end;
        nd = nodes[nd.parent];
    }
    __syncwarp();
    return kept;
}

constexpr uint32_t kPairWarps = 8;       // warps per block of k_pair_scatter
constexpr uint32_t kPairRowCap = 96;     // ids of the row list per tile
constexpr uint32_t kPairColCap = 256;    // ids of the column list per tile
__global__ void __launch_bounds__(kPairWarps * 32)
k_pair_scatter(const int* __restrict__ num_runs, const unsigned long long* __restrict__ run_keys, const uint32_t* __restrict__ run_counts,
               const Node* __restrict__ nodes1, const uint32_t* __restrict__ loc1, const Node* __restrict__ nodes2,
               const uint32_t* __restrict__ loc2, uint32_t r0, uint32_t r1, uint32_t N2, uint32_t* __restrict__ M,
               unsigned long long* __restrict__ updates) {
    __shared__ uint32_t s_rows[kPairWarps][kPairRowCap];
    __shared__ uint32_t s_cols[kPairWarps][kPairColCap];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t gw = (uint64_t)blockIdx.x * kPairWarps + wib, nw = (uint64_t)gridDim.x * kPairWarps;
    const uint64_t runs = (uint64_t)*num_runs;
    uint32_t* rows = s_rows[wib];
    uint32_t* cols = s_cols[wib];
    unsigned long long my_updates = 0;
    for (uint64_t r = gw; r < runs; r += nw) {
        const unsigned long long key = run_keys[r];
        const uint32_t pid1 = (uint32_t)(key >> 32), pid2 = (uint32_t)key;
        const uint32_t cnt = run_counts[r];
        const uint32_t n1 = nodes1[pid1].n, n2 = nodes2[pid2].n;
        uint32_t pos1 = 0;
        while (pos1 < n1) {
            const uint32_t before1 = pos1;
            const uint32_t nr = gather_chain(nodes1, loc1, pid1, pos1, r0, r1, rows, kPairRowCap, lane);
            if (pos1 == before1) break;   // (cannot happen on a validated trie: n = sum of l along the chain)
            if (nr == 0) continue;        // none of these ids belongs to the row block
            uint32_t pos2 = 0;
            while (pos2 < n2) {
                const uint32_t before2 = pos2;
                const uint32_t nc = gather_chain(nodes2, loc2, pid2, pos2, 0u, N2, cols, kPairColCap, lane);
                if (pos2 == before2) break;
                const uint32_t pairs = nr * nc;
                for (uint32_t t = lane; t < pairs; t += 32) {
                    const uint32_t j = t / nc, k = t - j * nc;
                    atomicAdd(&M[(size_t)(rows[j] - r0) * N2 + cols[k]], cnt);
                }
                __syncwarp();
            }
            if (lane == 0) my_updates += (unsigned long long)nr * n2;
        }
    }
    if (lane == 0 && my_updates) atomicAdd(updates, my_updates);
}

// one warp per row of the dense block (rows [r0, r0 + rows) of the row database, N2 columns)
__global__ void k_rect_count(const uint32_t* __restrict__ M, uint32_t rows, uint32_t N2, uint32_t r0, FilterDev f,
                             unsigned long long* __restrict__ counts) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t row = gw; row < rows; row += nw) {
        const uint32_t* src = M + (size_t)row * N2;
        uint32_t n = 0;
        for (uint32_t c0 = 0; c0 < N2; c0 += 32) {
            const uint32_t c = c0 + lane;
            const bool keep = c < N2 && cell_passes(f, src[c], r0 + row, c);
            n += __popc(__ballot_sync(0xffffffffu, keep));
        }
        if (lane == 0) counts[row] = n;
    }
}

__global__ void k_rect_fill(const uint32_t* __restrict__ M, uint32_t rows, uint32_t N2, uint32_t r0, FilterDev f,
                            const unsigned long long* __restrict__ row_ptr, uint32_t* __restrict__ col, uint32_t* __restrict__ val) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t row = gw; row < rows; row += nw) {
        const uint32_t* src = M + (size_t)row * N2;
        unsigned long long at = row_ptr[row];
        for (uint32_t c0 = 0; c0 < N2; c0 += 32) {
            const uint32_t c = c0 + lane;
            const uint32_t v = c < N2 ? src[c] : 0u;
            const bool keep = c < N2 && cell_passes(f, v, r0 + row, c);
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const unsigned long long o = at + __popc(m & ((1u << lane) - 1u));
                col[o] = c;
                val[o] = v;
            }
            at += __popc(m);
        }
    }
}

// decoded local lists + nodes of a staged database (shared with all2all and new2all)
int ensure_prepared(kdbx_ctx* ctx, uint32_t& launches) {
    if (ctx->prepared) return KDBX_OK;
    Plan pl;
    if (int rc = make_plan(ctx, pl)) return rc;
    const int rc = prepare(ctx, pl, launches);
    if (rc < 0) return rc;
    return check_device_error(ctx);
}

int db2db_sparse_impl(kdbx_ctx* ctx, kdbx_ctx* cdb, const kdbx_filter* filter, const uint32_t* col_sample_kmers, kdbx_csr* out,
                      kdbx_stats* stats) {
    if (!cdb) return ctx->fail(KDBX_ERR_ARG, "kdbx_db2db_sparse: the column database's context is NULL");
    if (cdb == ctx) return ctx->fail(KDBX_ERR_ARG, "kdbx_db2db_sparse: row and column database must be staged on two contexts (use kdbx_all2all_sparse for a database against itself)");
    if (!out) return ctx->fail(KDBX_ERR_ARG, "kdbx_db2db_sparse: out is NULL");
    std::memset(out, 0, sizeof *out);
    if (!ctx->loaded || !cdb->loaded) return ctx->fail(KDBX_ERR_STATE, "no patterns loaded (call kdbx_load_patterns on both contexts first)");
    if (!ctx->tables_loaded || !cdb->tables_loaded) return ctx->fail(KDBX_ERR_STATE, "no k-mer tables loaded (call kdbx_load_hashtables on both contexts first)");
    if (ctx->device != cdb->device) return ctx->fail(KDBX_ERR_ARG, "kdbx_db2db_sparse: both databases must be staged on the same device (%d and %d)", ctx->device, cdb->device);
    if (ctx->num_tables != cdb->num_tables)
        return ctx->fail(KDBX_ERR_ARG, "kdbx_db2db_sparse: the databases have %llu and %llu prefix buckets (different k-mer lengths or alphabets)",
                         (unsigned long long)ctx->num_tables, (unsigned long long)cdb->num_tables);
    if (int rc = require_full_window(ctx, "kdbx_db2db_sparse")) return rc;
    if (cdb->win_lo != 0 || cdb->win_hi != cdb->N) return ctx->fail(KDBX_ERR_STATE, "kdbx_db2db_sparse: a sample window is set on the column database");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->tri_rows_valid = false;   // row blocks pass through ctx->tri
    const uint32_t N1 = ctx->N, N2 = cdb->N;
    kdbx_stats s{};
    uint32_t launches = 0;
    ctx->ev_used = 0;
    cudaEvent_t ev0 = ctx->event();
    {   // the column database is prepared on its own stream; everything after runs on the row database's
        uint32_t l2 = 0;
        if (int rc = ensure_prepared(cdb, l2)) return ctx->fail(rc, "column database: %s", cdb->err.c_str());
        if (int rc = finish_upload(cdb)) return ctx->fail(rc, "column database: %s", cdb->err.c_str());
        CK(cudaStreamSynchronize(cdb->stream));
        launches += l2;
    }
    if (int rc = ensure_prepared(ctx, launches)) return rc;
    CK(cudaMemsetAsync(ctx->err_flag.p, 0, 16, st));   // (stale flags of earlier calls, as in new2all)
    cudaEvent_t ev1 = ctx->event();

    FilterDev f{};
    f.lo = 0; f.hi = 0xFFFFFFFFu;
    if (filter) {
        f.lo = filter->min_common; f.hi = filter->max_common; f.nb = filter->num_metric_bounds;
        if (f.nb > 4) return ctx->fail(KDBX_ERR_ARG, "kdbx_db2db_sparse: at most 4 metric bounds");
        for (uint32_t b = 0; b < f.nb; ++b) {
            f.metric[b] = filter->metric_bounds[b].metric;
            if (f.metric[b] < KDBX_METRIC_JACCARD || f.metric[b] > KDBX_METRIC_COSINE)
                return ctx->fail(KDBX_ERR_ARG, "kdbx_db2db_sparse: unknown metric %d", f.metric[b]);
            f.mlo[b] = filter->metric_bounds[b].lo; f.mhi[b] = filter->metric_bounds[b].hi;
        }
        if (f.nb) {
            if (!filter->sample_kmers || !col_sample_kmers) return ctx->fail(KDBX_ERR_ARG, "kdbx_db2db_sparse: metric bounds need the k-mer counts of both databases");
            CK(ctx->sp_cnt.ensure(((size_t)N1 + N2 + 2) * 4));
            CK(cudaMemcpyAsync(ctx->sp_cnt.p, filter->sample_kmers, (size_t)N1 * 4, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(ctx->sp_cnt.as<uint32_t>() + N1, col_sample_kmers, (size_t)N2 * 4, cudaMemcpyHostToDevice, st));
            f.cnt = ctx->sp_cnt.as<uint32_t>();
            f.cnt_col = ctx->sp_cnt.as<uint32_t>() + N1;
        }
    }
    out->num_rows = N1;
    out->row_ptr = static_cast<uint64_t*>(std::calloc((size_t)N1 + 1, 8));
    if (!out->row_ptr) return ctx->fail(KDBX_ERR_NOMEM, "host allocation failed");
    auto bail = [&](int rc) { kdbx_free_csr(out); return rc; };
    if (N1 == 0 || N2 == 0) { if (stats) *stats = s; return KDBX_OK; }

    // ---- the k-mers both databases hold, as (pattern1, pattern2) pairs; then the count of every distinct pair ----
    const uint64_t slots1 = ctx->total_slots;
    CK(ctx->counters.ensure(64));
    CK(cudaMemsetAsync(ctx->counters.p, 0, 64, st));
    unsigned long long* d_cnt = ctx->counters.as<unsigned long long>();          // [0] matched, [1] non-empty slots
    int* d_runs = ctx->counters.as<int>() + 4;                                    // (bytes 16..19)
    unsigned long long* d_updates = ctx->counters.as<unsigned long long>() + 3;   // (bytes 24..31)
    if (slots1 == 0) { if (stats) *stats = s; return KDBX_OK; }
    {
        const cudaError_t e = ctx->q_keys.ensure(slots1 * 8);
        if (e != cudaSuccess) { kdbx_free_csr(out); CK(e); }
    }
    cudaEvent_t ea = ctx->event();
    k_match_tables<<<blocks_for(slots1, 256), 256, 0, st>>>(slots1, ctx->num_tables, ctx->slot_off.as<uint64_t>(), ctx->slots.as<uint64_t>(), ctx->P,
                                                             cdb->slot_off.as<uint64_t>(), cdb->slots.as<uint64_t>(), cdb->P,
                                                             ctx->q_keys.as<unsigned long long>(), d_cnt, ctx->err_flag.as<int>());
    cudaEvent_t eb = ctx->event();
    launches += 1;
    uint64_t* h = ctx->h_pinned;
    if (!h) { CK(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_pinned), 64, cudaHostAllocDefault)); h = ctx->h_pinned; }
    h[0] = h[1] = h[2] = 0;
    CK(cudaMemcpyAsync(&h[0], d_cnt, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h[2], ctx->err_flag.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (int rc = error_from_flag(ctx, (int)(uint32_t)h[2])) return bail(rc);
    const uint64_t matched = h[0];
    s.probes = h[1]; s.hits = matched;
    s.ms_probe = elapsed(ea, eb);
    if (matched == 0) {
        s.ms_prepare = elapsed(ev0, ev1); s.ms_total = elapsed(ev0, eb); s.kernel_launches = launches;
        if (stats) *stats = s;
        return KDBX_OK;
    }
    if (matched >= ((uint64_t)1 << 31)) return bail(ctx->fail(KDBX_ERR_ARG, "kdbx_db2db_sparse: %llu common k-mers exceed the 2^31 this call sorts at once", (unsigned long long)matched));
    {
        cudaError_t e = ctx->q_keys2.ensure(matched * 8);
        if (e == cudaSuccess) e = ctx->q_runkeys.ensure(matched * 8);
        if (e == cudaSuccess) e = ctx->q_runcnt.ensure(matched * 4);
        if (e != cudaSuccess) { kdbx_free_csr(out); CK(e); }
        int b1 = 1; while (b1 < 32 && (ctx->P >> b1)) ++b1;
        size_t tmp = 0;
        e = cub::DeviceRadixSort::SortKeys(nullptr, tmp, ctx->q_keys.as<unsigned long long>(), ctx->q_keys2.as<unsigned long long>(), matched, 0, 32 + b1, st);
        if (e == cudaSuccess) e = ctx->cub_tmp.ensure(tmp);
        if (e == cudaSuccess) e = cub::DeviceRadixSort::SortKeys(ctx->cub_tmp.p, tmp, ctx->q_keys.as<unsigned long long>(), ctx->q_keys2.as<unsigned long long>(), matched, 0, 32 + b1, st);
        if (e == cudaSuccess) e = cub::DeviceRunLengthEncode::Encode(nullptr, tmp, ctx->q_keys2.as<unsigned long long>(), ctx->q_runkeys.as<unsigned long long>(),
                                                                     ctx->q_runcnt.as<uint32_t>(), d_runs, matched, st);
        if (e == cudaSuccess) e = ctx->cub_tmp.ensure(tmp);
        if (e == cudaSuccess) e = cub::DeviceRunLengthEncode::Encode(ctx->cub_tmp.p, tmp, ctx->q_keys2.as<unsigned long long>(), ctx->q_runkeys.as<unsigned long long>(),
                                                                     ctx->q_runcnt.as<uint32_t>(), d_runs, matched, st);
        if (e != cudaSuccess) { kdbx_free_csr(out); CK(e); }
        launches += 4;
    }

    // ---- blocks of rows: dense accumulation in HBM, then filter + compaction ---------------------------------------
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    uint64_t budget_cells = ctx->cfg.sparse_block_cells ? ctx->cfg.sparse_block_cells : (uint64_t)(free_b / 4 / 4);  // a quarter of free HBM
    const uint32_t rows_per_block = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(N1, budget_cells / N2));
    std::vector<std::pair<uint32_t*, uint64_t>> col_parts, val_parts;  // pinned chunks per block
    auto cleanup = [&]() {
        for (auto& c : col_parts) cudaFreeHost(c.first);
        for (auto& c : val_parts) cudaFreeHost(c.first);
    };
    auto fail_cuda = [&](cudaError_t e, const char* what) {
        cleanup(); kdbx_free_csr(out); cudaGetLastError();
        return ctx->fail(e == cudaErrorMemoryAllocation ? KDBX_ERR_NOMEM : KDBX_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
    };
    uint64_t nnz = 0;
    const unsigned grid = (unsigned)(ctx->sm_count * 8);
    for (uint32_t r0 = 0; r0 < N1; r0 += rows_per_block) {
        const uint32_t r1 = (uint32_t)std::min<uint64_t>(N1, (uint64_t)r0 + rows_per_block);
        const uint32_t rows = r1 - r0;
        const uint64_t cells = (uint64_t)rows * N2;
        cudaError_t e = ctx->tri.ensure((cells + 1) * 4);
        if (e == cudaSuccess) e = ctx->sp_counts.ensure(((size_t)rows + 1) * 8);
        if (e == cudaSuccess) e = ctx->sp_rowptr.ensure(((size_t)rows + 1) * 8);
        if (e == cudaSuccess) e = cudaMemsetAsync(ctx->tri.p, 0, cells * 4, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(ctx->sp_counts.p, 0, ((size_t)rows + 1) * 8, st);
        if (e != cudaSuccess) return fail_cuda(e, "block allocation");
        cudaEvent_t e0 = ctx->event();
        k_pair_scatter<<<grid, kPairWarps * 32, 0, st>>>(d_runs, ctx->q_runkeys.as<unsigned long long>(), ctx->q_runcnt.as<uint32_t>(),
                                                         ctx->nodes.as<Node>(), ctx->loc.as<uint32_t>(), cdb->nodes.as<Node>(), cdb->loc.as<uint32_t>(),
                                                         r0, r1, N2, ctx->tri.as<uint32_t>(), d_updates);
        cudaEvent_t e1 = ctx->event();
        k_rect_count<<<grid, 256, 0, st>>>(ctx->tri.as<uint32_t>(), rows, N2, r0, f, ctx->sp_counts.as<unsigned long long>());
        if (int rc = scan_exclusive(ctx, ctx->sp_counts.as<uint64_t>(), ctx->sp_rowptr.as<uint64_t>(), (uint64_t)rows + 1)) { cleanup(); return bail(rc); }
        std::vector<uint64_t> h_ptr((size_t)rows + 1);
        e = cudaMemcpyAsync(h_ptr.data(), ctx->sp_rowptr.p, ((size_t)rows + 1) * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail_cuda(e, "row pointers");
        const uint64_t block_nnz = h_ptr[rows];
        launches += 3;
        s.scatter_launches += 1;
        if (block_nnz) {
            uint32_t *h_col = nullptr, *h_val = nullptr;
            if ((e = ctx->sp_col.ensure(block_nnz * 4)) != cudaSuccess || (e = ctx->sp_val.ensure(block_nnz * 4)) != cudaSuccess) return fail_cuda(e, "compaction buffers");
            k_rect_fill<<<grid, 256, 0, st>>>(ctx->tri.as<uint32_t>(), rows, N2, r0, f, ctx->sp_rowptr.as<unsigned long long>(), ctx->sp_col.as<uint32_t>(),
                                              ctx->sp_val.as<uint32_t>());
            launches += 1;
            if ((e = cudaHostAlloc((void**)&h_col, block_nnz * 4, cudaHostAllocDefault)) != cudaSuccess) return fail_cuda(e, "pinned allocation");
            col_parts.emplace_back(h_col, block_nnz);
            if ((e = cudaHostAlloc((void**)&h_val, block_nnz * 4, cudaHostAllocDefault)) != cudaSuccess) return fail_cuda(e, "pinned allocation");
            val_parts.emplace_back(h_val, block_nnz);
            cudaEvent_t e2 = ctx->event();
            e = cudaMemcpyAsync(h_col, ctx->sp_col.p, block_nnz * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_val, ctx->sp_val.p, block_nnz * 4, cudaMemcpyDeviceToHost, st);
            cudaEvent_t e3 = ctx->event();
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return fail_cuda(e, "download");
            s.ms_compact += elapsed(e1, e2);
            s.ms_download += elapsed(e2, e3);
        } else {
            cudaEvent_t e2 = ctx->event();
            cudaStreamSynchronize(st);
            s.ms_compact += elapsed(e1, e2);
        }
        s.ms_scatter += elapsed(e0, e1);
        for (uint32_t r = 0; r < rows; ++r) out->row_ptr[(size_t)r0 + r] = nnz + h_ptr[r];
        nnz += block_nnz;
    }
    out->row_ptr[N1] = nnz;
    out->nnz = nnz;
    cudaEvent_t ev2 = ctx->event();
    h[0] = h[2] = 0;
    CK(cudaMemcpyAsync(&h[0], d_updates, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h[2], ctx->err_flag.p, 4, cudaMemcpyDeviceToHost, st));
    {
        const cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail_cuda(e, "cudaStreamSynchronize");
    }
    if (int rc = error_from_flag(ctx, (int)(uint32_t)h[2])) { cleanup(); return bail(rc); }
    if (col_parts.size() == 1) {
        out->col = col_parts[0].first; out->val = val_parts[0].first;
        out->_pad = 1;  // pinned
    } else if (nnz) {
        out->col = static_cast<uint32_t*>(std::malloc(nnz * 4));
        out->val = static_cast<uint32_t*>(std::malloc(nnz * 4));
        if (!out->col || !out->val) { cleanup(); kdbx_free_csr(out); return ctx->fail(KDBX_ERR_NOMEM, "host allocation failed"); }
        uint64_t at = 0;
        for (size_t i = 0; i < col_parts.size(); ++i) {
            std::memcpy(out->col + at, col_parts[i].first, col_parts[i].second * 4);
            std::memcpy(out->val + at, val_parts[i].first, val_parts[i].second * 4);
            at += col_parts[i].second;
        }
        cleanup();
        out->_pad = 0;
    }
    if (int rc = finish_upload(ctx)) return rc;
    s.updates = h[0];
    s.physical_updates = h[0];
    s.ms_upload = ctx->ms_upload;
    s.ms_prepare = elapsed(ev0, ev1);
    s.ms_total = elapsed(ev0, ev2);
    s.kernel_launches = launches;
    s.local_ids = ctx->sum_l; s.flat_ids = ctx->sum_n;
    if (stats) *stats = s;
    return KDBX_OK;
}
