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
