// Sparse delivery of the all2all matrix (included by kdbx.cu; shares its anonymous namespace).
//
// Replaces SimilarityCalculator::all2all_sp + SparseMatrix::compact2 (src/similarity_calculator.cpp:
// 442-657, src/array.h:391-446).  The reference accumulates into one hash map per row because a
// dense N x N accumulator does not fit a CPU's memory budget at large N; on a B200 it does: blocks of
// rows are accumulated densely in HBM by the dense kernels, then every row is filtered
// (CombinedFilter, src/sparse_filters.h:33-61) and compacted to ascending (col, val) pairs — which is
// what compact2's per-row sort produces.  Counting and filling are two passes of the same traversal
// with a CUB scan between them; both are pure HBM streams (4 B read per cell).
#pragma once

struct FilterDev {
    uint32_t lo, hi, nb;
    int32_t metric[4];
    double mlo[4], mhi[4];
    const uint32_t* cnt;       // k-mer counts of the row samples
    const uint32_t* cnt_col;   // ... of the column samples (the same array in all2all; another database's in db2db.cuh)
};

__device__ __forceinline__ bool cell_passes(const FilterDev& f, uint32_t v, uint32_t row, uint32_t col) {
    if (v == 0 || v < f.lo || v > f.hi) return false;
    for (uint32_t b = 0; b < f.nb; ++b) {
        const uint32_t a = f.cnt[row], c = f.cnt_col[col];  // uint32 arithmetic wraps like the reference's num_kmers_t
        double x;
        switch (f.metric[b]) {
            case KDBX_METRIC_JACCARD: x = __ddiv_rn((double)v, (double)(uint32_t)(a + c - v)); break;
            case KDBX_METRIC_MIN: x = __ddiv_rn((double)v, (double)(a < c ? a : c)); break;
            case KDBX_METRIC_MAX: x = __ddiv_rn((double)v, (double)(a > c ? a : c)); break;
            default: x = __ddiv_rn((double)v, __dsqrt_rn((double)(uint32_t)(a * c))); break;
        }
        if (!(x >= f.mlo[b] && x <= f.mhi[b])) return false;
    }
    return true;
}

// one warp per row; rows [row_begin, row_end) of the packed triangle that starts at tri_base
__global__ void k_sparse_count(const uint32_t* __restrict__ tri, uint64_t tri_base, uint32_t row_begin, uint32_t row_end,
                               FilterDev f, unsigned long long* __restrict__ counts) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t row = row_begin + gw; row < row_end; row += nw) {
        const uint32_t* src = tri + (tri_offset(row) - tri_base);
        uint32_t n = 0;
        for (uint32_t c0 = 0; c0 < row; c0 += 32) {
            const uint32_t c = c0 + lane;
            const bool keep = c < row && cell_passes(f, src[c], row, c);
            n += __popc(__ballot_sync(0xffffffffu, keep));
        }
        if (lane == 0) counts[row - row_begin] = n;
    }
}

__global__ void k_sparse_fill(const uint32_t* __restrict__ tri, uint64_t tri_base, uint32_t row_begin, uint32_t row_end,
                              FilterDev f, const unsigned long long* __restrict__ row_ptr, uint32_t* __restrict__ col,
                              uint32_t* __restrict__ val) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t row = row_begin + gw; row < row_end; row += nw) {
        const uint32_t* src = tri + (tri_offset(row) - tri_base);
        unsigned long long at = row_ptr[row - row_begin];
        for (uint32_t c0 = 0; c0 < row; c0 += 32) {
            const uint32_t c = c0 + lane;
            const uint32_t v = c < row ? src[c] : 0u;
            const bool keep = c < row && cell_passes(f, v, row, c);
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

// rows [row_begin, row_end) only (the other rows of the result are empty): the unit of work of a multi-GPU run, where
// every device holds the whole trie and the CSR rows of the devices are concatenated by the caller
int all2all_sparse_impl(kdbx_ctx* ctx, const kdbx_filter* filter, kdbx_csr* out, kdbx_stats* stats, uint32_t row_begin, uint32_t row_end) {
    if (!ctx->loaded) return ctx->fail(KDBX_ERR_STATE, "no patterns loaded (call kdbx_load_patterns first)");
    if (!out) return ctx->fail(KDBX_ERR_ARG, "kdbx_all2all_sparse: out is NULL");
    std::memset(out, 0, sizeof *out);
    ctx->tri_rows_valid = false;   // row blocks pass through ctx->tri
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t N = ctx->N;
    if (row_begin > row_end || row_end > N) return ctx->fail(KDBX_ERR_ARG, "bad row range [%u,%u) for %u samples", row_begin, row_end, N);
    FilterDev f{};
    f.lo = 0; f.hi = 0xFFFFFFFFu;
    if (filter) {
        f.lo = filter->min_common; f.hi = filter->max_common; f.nb = filter->num_metric_bounds;
        if (f.nb > 4) return ctx->fail(KDBX_ERR_ARG, "kdbx_all2all_sparse: at most 4 metric bounds");
        for (uint32_t b = 0; b < f.nb; ++b) {
            f.metric[b] = filter->metric_bounds[b].metric;
            if (f.metric[b] < KDBX_METRIC_JACCARD || f.metric[b] > KDBX_METRIC_COSINE)
                return ctx->fail(KDBX_ERR_ARG, "kdbx_all2all_sparse: unknown metric %d", f.metric[b]);
            f.mlo[b] = filter->metric_bounds[b].lo; f.mhi[b] = filter->metric_bounds[b].hi;
        }
        if (f.nb) {
            if (!filter->sample_kmers) return ctx->fail(KDBX_ERR_ARG, "kdbx_all2all_sparse: metric bounds need sample_kmers");
            CK(ctx->sp_cnt.ensure(((size_t)N + 1) * 4));
            CK(cudaMemcpyAsync(ctx->sp_cnt.p, filter->sample_kmers, (size_t)N * 4, cudaMemcpyHostToDevice, st));
            f.cnt = ctx->sp_cnt.as<uint32_t>();
            f.cnt_col = f.cnt;
        }
    }
    // row blocks: as many rows as fit the accumulator budget (free HBM minus working buffers)
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    uint64_t budget_cells = ctx->cfg.sparse_block_cells ? ctx->cfg.sparse_block_cells : (uint64_t)(free_b / 4 / 4);  // a quarter of free HBM
    if (budget_cells < N) budget_cells = N;
    auto tri_off = [](uint64_t r) { return r == 0 ? 0ull : r * (r - 1) / 2; };

    std::vector<uint64_t> row_ptr((size_t)N + 1, 0);
    std::vector<std::pair<uint32_t*, uint64_t>> col_parts, val_parts;  // pinned chunks per block
    auto cleanup = [&]() {
        for (auto& c : col_parts) cudaFreeHost(c.first);
        for (auto& c : val_parts) cudaFreeHost(c.first);
    };
    kdbx_stats total{};
    uint64_t nnz = 0;
    uint32_t r0 = row_begin;
    while (r0 < row_end) {
        uint32_t r1 = r0 + 1;
        while (r1 < row_end && tri_off(r1 + 1) - tri_off(r0) <= budget_cells) ++r1;
        const uint64_t cells = tri_off(r1) - tri_off(r0);
        const uint32_t rows = r1 - r0;
        kdbx_stats s{};
        cudaError_t e = ctx->tri.ensure((cells + 1) * 4);
        if (e != cudaSuccess) { cleanup(); CK(e); }
        if (int rc = all2all_rows_device(ctx, r0, r1, ctx->tri.as<uint32_t>(), &s)) { cleanup(); return rc; }
        total.updates += s.updates; total.chunks += s.chunks; total.kernel_launches += s.kernel_launches;
        total.ms_prepare += s.ms_prepare; total.ms_expand += s.ms_expand; total.ms_bucket += s.ms_bucket;
        total.ms_scatter += s.ms_scatter; total.ms_total += s.ms_total; total.scatter_launches += s.scatter_launches;
        total.flat_ids = s.flat_ids; total.local_ids = s.local_ids; total.ms_upload = s.ms_upload;
        // count -> scan -> fill
        cudaEvent_t ea = ctx->event();
        if ((e = ctx->sp_counts.ensure(((size_t)rows + 1) * 8)) != cudaSuccess || (e = ctx->sp_rowptr.ensure(((size_t)rows + 1) * 8)) != cudaSuccess) { cleanup(); CK(e); }
        e = cudaMemsetAsync(ctx->sp_counts.p, 0, ((size_t)rows + 1) * 8, st);
        if (e != cudaSuccess) { cleanup(); CK(e); }
        const unsigned grid = (unsigned)(ctx->sm_count * 8);
        k_sparse_count<<<grid, 256, 0, st>>>(ctx->tri.as<uint32_t>(), tri_off(r0), r0, r1, f, ctx->sp_counts.as<unsigned long long>());
        if (int rc = scan_exclusive(ctx, ctx->sp_counts.as<uint64_t>(), ctx->sp_rowptr.as<uint64_t>(), (uint64_t)rows + 1)) { cleanup(); return rc; }
        std::vector<uint64_t> h_ptr((size_t)rows + 1);
        e = cudaMemcpyAsync(h_ptr.data(), ctx->sp_rowptr.p, ((size_t)rows + 1) * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { cleanup(); CK(e); }
        const uint64_t block_nnz = h_ptr[rows];
        uint32_t *h_col = nullptr, *h_val = nullptr;
        if (block_nnz) {
            if ((e = ctx->sp_col.ensure(block_nnz * 4)) != cudaSuccess || (e = ctx->sp_val.ensure(block_nnz * 4)) != cudaSuccess) { cleanup(); CK(e); }
            k_sparse_fill<<<grid, 256, 0, st>>>(ctx->tri.as<uint32_t>(), tri_off(r0), r0, r1, f, ctx->sp_rowptr.as<unsigned long long>(),
                                                ctx->sp_col.as<uint32_t>(), ctx->sp_val.as<uint32_t>());
            if ((e = cudaHostAlloc((void**)&h_col, block_nnz * 4, cudaHostAllocDefault)) != cudaSuccess) { cleanup(); cudaGetLastError(); return ctx->fail(KDBX_ERR_NOMEM, "pinned allocation of %llu bytes failed", (unsigned long long)(block_nnz * 4)); }
            col_parts.emplace_back(h_col, block_nnz);
            if ((e = cudaHostAlloc((void**)&h_val, block_nnz * 4, cudaHostAllocDefault)) != cudaSuccess) { cleanup(); cudaGetLastError(); return ctx->fail(KDBX_ERR_NOMEM, "pinned allocation of %llu bytes failed", (unsigned long long)(block_nnz * 4)); }
            val_parts.emplace_back(h_val, block_nnz);
            cudaEvent_t eb = ctx->event();
            e = cudaMemcpyAsync(h_col, ctx->sp_col.p, block_nnz * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_val, ctx->sp_val.p, block_nnz * 4, cudaMemcpyDeviceToHost, st);
            cudaEvent_t ec = ctx->event();
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { cleanup(); CK(e); }
            total.ms_compact += elapsed(ea, eb);
            total.ms_download += elapsed(eb, ec);
        } else {
            cudaEvent_t eb = ctx->event();
            cudaStreamSynchronize(st);
            total.ms_compact += elapsed(ea, eb);
        }
        total.kernel_launches += block_nnz ? 4 : 3;
        for (uint32_t r = 0; r < rows; ++r) row_ptr[(size_t)r0 + r] = nnz + h_ptr[r];
        nnz += block_nnz;
        r0 = r1;
    }
    for (size_t r = row_end; r <= N; ++r) row_ptr[r] = nnz;   // rows after the range are empty
    // hand over: row_ptr always malloc'ed; col/val are the pinned chunk itself when there was one block
    out->num_rows = N;
    out->nnz = nnz;
    out->row_ptr = static_cast<uint64_t*>(std::malloc(((size_t)N + 1) * 8));
    if (!out->row_ptr) { cleanup(); return ctx->fail(KDBX_ERR_NOMEM, "host allocation failed"); }
    std::memcpy(out->row_ptr, row_ptr.data(), ((size_t)N + 1) * 8);
    if (col_parts.size() == 1) {
        out->col = col_parts[0].first; out->val = val_parts[0].first;
        out->_pad = 1;  // pinned
    } else if (nnz) {
        out->col = static_cast<uint32_t*>(std::malloc(nnz * 4));
        out->val = static_cast<uint32_t*>(std::malloc(nnz * 4));
        if (!out->col || !out->val) { cleanup(); std::free(out->col); std::free(out->val); std::free(out->row_ptr); std::memset(out, 0, sizeof *out); return ctx->fail(KDBX_ERR_NOMEM, "host allocation failed"); }
        uint64_t at = 0;
        for (size_t i = 0; i < col_parts.size(); ++i) {
            std::memcpy(out->col + at, col_parts[i].first, col_parts[i].second * 4);
            std::memcpy(out->val + at, val_parts[i].first, val_parts[i].second * 4);
            at += col_parts[i].second;
        }
        cleanup();
        out->_pad = 0;
    }
    if (stats) *stats = total;
    return KDBX_OK;
}
