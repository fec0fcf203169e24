// Decimal text of dense table rows on the device (included by kdbx.cu; shares its anonymous namespace).
//
// The step after the path: All2AllConsole::run prints row s of the matrix as its s cells, each followed by ','
// (src/console_all2all.cpp:65-78 -> LowerTriangularMatrix::saveRow, src/array.h:254-258 -> num2str of a collection,
// src/conversion.h:275-284, plain decimal by Int2PChar, :99-165).  At N = 10^4 that is 5*10^7 numbers, 0.3 GB of text,
// and costs the reference's single writer thread more than the matrix itself.  Here the cells are
// formatted where the matrix already is: a byte count per row (one warp per row), an exclusive scan, then every warp
// writes its row's digits straight to their final place; the host adds the sample name, the k-mer count and the newline.
#pragma once

__device__ __forceinline__ uint32_t dec_len(uint32_t v) {
    return v < 10u ? 1u : v < 100u ? 2u : v < 1000u ? 3u : v < 10000u ? 4u : v < 100000u ? 5u : v < 1000000u ? 6u :
           v < 10000000u ? 7u : v < 100000000u ? 8u : v < 1000000000u ? 9u : 10u;
}

__global__ void k_csv_row_bytes(const uint32_t* __restrict__ tri, uint64_t tri_base, uint32_t row_begin, uint32_t row_end,
                                unsigned long long* __restrict__ bytes) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t row = row_begin + gw; row < row_end; row += nw) {
        const uint32_t* src = tri + (tri_offset(row) - tri_base);
        unsigned long long n = 0;
        for (uint32_t c = lane; c < row; c += 32) n += dec_len(src[c]) + 1u;
        for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        if (lane == 0) bytes[row - row_begin] = n;
    }
}

__global__ void k_csv_fill(const uint32_t* __restrict__ tri, uint64_t tri_base, uint32_t row_begin, uint32_t row_end,
                           const unsigned long long* __restrict__ row_off, char* __restrict__ text) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t row = row_begin + gw; row < row_end; row += nw) {
        const uint32_t* src = tri + (tri_offset(row) - tri_base);
        unsigned long long at = row_off[row - row_begin];
        for (uint32_t c0 = 0; c0 < row; c0 += 32) {
            const uint32_t c = c0 + lane;
            uint32_t v = c < row ? src[c] : 0u;
            const uint32_t len = c < row ? dec_len(v) + 1u : 0u;
            uint32_t incl = len;
            for (int o = 1; o < 32; o <<= 1) { const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += up; }
            if (len) {
                char* p = text + at + incl;   // one past this cell's comma
                *--p = ',';
                do { *--p = (char)('0' + v % 10u); v /= 10u; } while (v);
            }
            at += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

// The cells' text of rows [row_begin, row_end) of the matrix that the last host-output all2all call left in ctx->tri.
// text == NULL: only row_off (rows + 1 byte offsets) and *bytes are produced.
int csv_dense_rows(kdbx_ctx* ctx, uint32_t row_begin, uint32_t row_end, char* text, uint64_t capacity, uint64_t* row_off, uint64_t* bytes) {
    if (!ctx->tri_rows_valid) return ctx->fail(KDBX_ERR_STATE, "kdbx_csv_dense_rows: no matrix on the device (call kdbx_all2all_dense or kdbx_all2all_dense_rows first)");
    if (row_begin > row_end || row_begin < ctx->tri_row_begin || row_end > ctx->tri_row_end)
        return ctx->fail(KDBX_ERR_ARG, "kdbx_csv_dense_rows: rows [%u,%u) are not in the resident block [%u,%u)", row_begin, row_end, ctx->tri_row_begin, ctx->tri_row_end);
    if (!row_off) return ctx->fail(KDBX_ERR_ARG, "kdbx_csv_dense_rows: row_off is NULL");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t rows = row_end - row_begin;
    auto tri_off = [](uint64_t r) { return r == 0 ? 0ull : r * (r - 1) / 2; };
    const uint64_t tri_base = tri_off(ctx->tri_row_begin);
    CK(ctx->sp_counts.ensure(((size_t)rows + 1) * 8)); CK(ctx->sp_rowptr.ensure(((size_t)rows + 1) * 8));
    CK(cudaMemsetAsync(ctx->sp_counts.p, 0, ((size_t)rows + 1) * 8, st));
    const unsigned grid = (unsigned)(ctx->sm_count * 8);
    ctx->ev_used = 0;
    cudaEvent_t a = ctx->event();
    if (rows) k_csv_row_bytes<<<grid, 256, 0, st>>>(ctx->tri.as<uint32_t>(), tri_base, row_begin, row_end, ctx->sp_counts.as<unsigned long long>());
    if (int rc = scan_exclusive(ctx, ctx->sp_counts.as<uint64_t>(), ctx->sp_rowptr.as<uint64_t>(), (uint64_t)rows + 1)) return rc;
    CK(cudaMemcpyAsync(row_off, ctx->sp_rowptr.p, ((size_t)rows + 1) * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint64_t total = row_off[rows];
    if (bytes) *bytes = total;
    if (!text) return KDBX_OK;
    if (capacity < total) return ctx->fail(KDBX_ERR_ARG, "kdbx_csv_dense_rows: %llu bytes needed, %llu given", (unsigned long long)total, (unsigned long long)capacity);
    if (total) {
        CK(ctx->csv_text.ensure(total));
        k_csv_fill<<<grid, 256, 0, st>>>(ctx->tri.as<uint32_t>(), tri_base, row_begin, row_end, ctx->sp_rowptr.as<unsigned long long>(),
                                         ctx->csv_text.as<char>());
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(text, ctx->csv_text.p, total, cudaMemcpyDeviceToHost, st));
    }
    cudaEvent_t b = ctx->event();
    CK(cudaStreamSynchronize(st));
    ctx->ms_csv = elapsed(a, b);
    return KDBX_OK;
}

// ---- `distance` on the device -----------------------------------------------------------------------------------
// DistanceConsole::run (src/console_distance.cpp:7-213) maps every cell of a common-k-mer table through a measure
// (src/params.cpp:14-42) and prints it with six decimals: num2str(double) -> Double2PChar (src/conversion.h:167-219,
// 254-260): "0" for an exact zero, otherwise (uint64)(v * 10^6 + 0.5) split into integer part, '.', six digits.  The
// measures offered here are the ones whose arithmetic is one correctly rounded IEEE operation or two (a division, a square
// root), written with the _rn intrinsics so that nothing is contracted into a fused multiply-add: their bytes are the
// reference's.  The logarithm-based measures (mash, ani, ...) stay on the host (glibc's log is not reproducible here).
// Counts are uint32 and wrap like the reference's num_kmers_t (src/types.h:19).
__device__ __forceinline__ double measure_value(int metric, uint32_t c, uint32_t a, uint32_t b) {
    switch (metric) {
        case KDBX_METRIC_JACCARD: return __ddiv_rn((double)c, (double)(uint32_t)(a + b - c));
        case KDBX_METRIC_MIN: return __ddiv_rn((double)c, (double)(a < b ? a : b));
        case KDBX_METRIC_MAX: return __ddiv_rn((double)c, (double)(a > b ? a : b));
        case KDBX_METRIC_COSINE: return __ddiv_rn((double)c, __dsqrt_rn((double)(uint32_t)(a * b)));
        default: return (double)c;   // KDBX_METRIC_NUM_KMERS
    }
}
// length of the text of v followed by ','; x receives the scaled value
__device__ __forceinline__ uint32_t f6_len(double v, unsigned long long& x) {
    if (v == 0) { x = 0; return 2u; }
    x = (unsigned long long)__dadd_rn(__dmul_rn(v, 1000000.0), 0.5);
    unsigned long long ip = x / 1000000ull;
    uint32_t d = 1;
    while (ip >= 10) { ip /= 10; ++d; }
    return d + 8u;   // digits '.' six digits ','
}

__global__ void k_dist_row_bytes(const uint32_t* __restrict__ tri, uint64_t tri_base, uint32_t row_begin, uint32_t row_end, int metric,
                                 const uint32_t* __restrict__ cnt, unsigned long long* __restrict__ bytes) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t row = row_begin + gw; row < row_end; row += nw) {
        const uint32_t* src = tri + (tri_offset(row) - tri_base);
        const uint32_t a = cnt[row];
        unsigned long long n = 0, x;
        for (uint32_t c = lane; c < row; c += 32) n += f6_len(measure_value(metric, src[c], a, cnt[c]), x);
        for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        if (lane == 0) bytes[row - row_begin] = n;
    }
}

__global__ void k_dist_fill(const uint32_t* __restrict__ tri, uint64_t tri_base, uint32_t row_begin, uint32_t row_end, int metric,
                            const uint32_t* __restrict__ cnt, const unsigned long long* __restrict__ row_off, char* __restrict__ text) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t row = row_begin + gw; row < row_end; row += nw) {
        const uint32_t* src = tri + (tri_offset(row) - tri_base);
        const uint32_t a = cnt[row];
        unsigned long long at = row_off[row - row_begin];
        for (uint32_t c0 = 0; c0 < row; c0 += 32) {
            const uint32_t c = c0 + lane;
            unsigned long long x = 0;
            double v = 0;
            uint32_t len = 0;
            if (c < row) { v = measure_value(metric, src[c], a, cnt[c]); len = f6_len(v, x); }
            uint32_t incl = len;
            for (int o = 1; o < 32; o <<= 1) { const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += up; }
            if (len) {
                char* p = text + at + incl;   // one past this cell's comma
                *--p = ',';
                if (v == 0) *--p = '0';
                else {
                    uint32_t frac = (uint32_t)(x % 1000000ull);
                    unsigned long long ip = x / 1000000ull;
                    for (int i = 0; i < 6; ++i) { *--p = (char)('0' + frac % 10u); frac /= 10u; }
                    *--p = '.';
                    do { *--p = (char)('0' + (uint32_t)(ip % 10ull)); ip /= 10ull; } while (ip);
                }
            }
            at += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

// A packed triangle from the host (what `distance` parsed from a table) staged as the resident matrix.
int stage_matrix(kdbx_ctx* ctx, const uint32_t* tri_host, uint32_t num_samples) {
    const uint64_t N = num_samples, cells = N ? N * (N - 1) / 2 : 0;
    if (cells && !tri_host) return ctx->fail(KDBX_ERR_ARG, "kdbx_stage_matrix: matrix is NULL");
    CK(cudaSetDevice(ctx->device));
    ctx->tri_rows_valid = false;
    CK(ctx->tri.ensure((cells + 1) * 4));
    if (cells) CK(cudaMemcpyAsync(ctx->tri.p, tri_host, cells * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->tri_rows_valid = true; ctx->tri_row_begin = 0; ctx->tri_row_end = num_samples;
    return KDBX_OK;
}

int distance_dense_rows(kdbx_ctx* ctx, int metric, const uint32_t* sample_kmers, uint32_t row_begin, uint32_t row_end, char* text,
                        uint64_t capacity, uint64_t* row_off, uint64_t* bytes) {
    if (!ctx->tri_rows_valid) return ctx->fail(KDBX_ERR_STATE, "kdbx_distance_dense_rows: no matrix on the device (kdbx_all2all_dense, kdbx_all2all_dense_rows or kdbx_stage_matrix first)");
    if (row_begin > row_end || row_begin < ctx->tri_row_begin || row_end > ctx->tri_row_end)
        return ctx->fail(KDBX_ERR_ARG, "kdbx_distance_dense_rows: rows [%u,%u) are not in the resident block [%u,%u)", row_begin, row_end, ctx->tri_row_begin, ctx->tri_row_end);
    if (metric < KDBX_METRIC_JACCARD || metric > KDBX_METRIC_NUM_KMERS) return ctx->fail(KDBX_ERR_ARG, "kdbx_distance_dense_rows: measure %d is not offered on the device", metric);
    if (!row_off || !sample_kmers) return ctx->fail(KDBX_ERR_ARG, "kdbx_distance_dense_rows: NULL argument");
    for (uint32_t s = 0; s < row_end; ++s)   // a zero count makes 0/0 or x/0: what the reference prints for those is not defined arithmetic
        if (sample_kmers[s] == 0 && metric != KDBX_METRIC_NUM_KMERS) return ctx->fail(KDBX_ERR_ARG, "kdbx_distance_dense_rows: sample %u has no k-mers", s);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t rows = row_end - row_begin;
    auto tri_off = [](uint64_t r) { return r == 0 ? 0ull : r * (r - 1) / 2; };
    const uint64_t tri_base = tri_off(ctx->tri_row_begin);
    CK(ctx->sp_counts.ensure(((size_t)rows + 1) * 8)); CK(ctx->sp_rowptr.ensure(((size_t)rows + 1) * 8));
    CK(ctx->sp_cnt.ensure(((size_t)row_end + 1) * 4));
    CK(cudaMemcpyAsync(ctx->sp_cnt.p, sample_kmers, (size_t)row_end * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->sp_counts.p, 0, ((size_t)rows + 1) * 8, st));
    const unsigned grid = (unsigned)(ctx->sm_count * 8);
    if (rows) k_dist_row_bytes<<<grid, 256, 0, st>>>(ctx->tri.as<uint32_t>(), tri_base, row_begin, row_end, metric, ctx->sp_cnt.as<uint32_t>(),
                                                     ctx->sp_counts.as<unsigned long long>());
    if (int rc = scan_exclusive(ctx, ctx->sp_counts.as<uint64_t>(), ctx->sp_rowptr.as<uint64_t>(), (uint64_t)rows + 1)) return rc;
    CK(cudaMemcpyAsync(row_off, ctx->sp_rowptr.p, ((size_t)rows + 1) * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint64_t total = row_off[rows];
    if (bytes) *bytes = total;
    if (!text) return KDBX_OK;
    if (capacity < total) return ctx->fail(KDBX_ERR_ARG, "kdbx_distance_dense_rows: %llu bytes needed, %llu given", (unsigned long long)total, (unsigned long long)capacity);
    if (total) {
        CK(ctx->csv_text.ensure(total));
        k_dist_fill<<<grid, 256, 0, st>>>(ctx->tri.as<uint32_t>(), tri_base, row_begin, row_end, metric, ctx->sp_cnt.as<uint32_t>(),
                                          ctx->sp_rowptr.as<unsigned long long>(), ctx->csv_text.as<char>());
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(text, ctx->csv_text.p, total, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return KDBX_OK;
}
