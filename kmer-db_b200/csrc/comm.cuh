// Multi-GPU all2all inside the library (included by kdbx.cu; shares its anonymous namespace).
//
// The reference is single-process; its only sharding template is the grid of partial databases of all2all-parts
// (src/console_all2all_parts.cpp:143-331) and, between threads, row ownership (src/similarity_calculator.cpp:371-395).
// Here the DATABASE is sharded: every GPU holds one sub-trie (kdbxh_partition: the matrix is linear in num_kmers, so
// the partial matrices of the parts add up), runs the complete single-GPU pipeline on it, and ONE collective —
// ncclReduceScatter(uint32, sum) over NVLink / NVSwitch — adds the partial matrices and leaves every GPU with its
// own block of the packed triangle: GPU r owns the cells [r B, (r+1) B), B = ceil(cells / ranks).  The host reads
// every block once.  One context per GPU: one process per GPU (kdbx_comm_init_rank, the unique id travels by the
// caller's launcher — bench.py: torch.distributed) or several contexts in one process (kdbx_comm_init_all — the CLI's
// -gpus, one host thread per device).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy a Python process already holds through torch, or the
// system one), so that libkdbx.so loads — and the single-GPU path works — on a machine without NCCL.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) { api.error = std::string("NCCL is not available: ") + dlerror(); return &api; }
    auto sym = [&](const char* n) { void* p = dlsym(api.handle, n); if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + n; return p; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.ReduceScatter = reinterpret_cast<decltype(api.ReduceScatter)>(sym("ncclReduceScatter"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    return &api;
}

#define NCK(call)                                                                                          \
    do {                                                                                                   \
        ncclResult_t r__ = (call);                                                                         \
        if (r__ != ncclSuccess)                                                                            \
            return ctx->fail(KDBX_ERR_CUDA, "%s failed: %s", #call, nccl_api()->GetErrorString(r__));     \
    } while (0)

// cells of the packed triangle one rank owns after the reduce-scatter
inline uint64_t comm_block_cells(uint64_t cells, int nranks) { return nranks > 0 ? (cells + (uint64_t)nranks - 1) / (uint64_t)nranks : cells; }

// The partial matrix of this rank's sub-trie into ctx->tri (nranks * B cells, the tail zero), then the collective.
// d_block receives B cells; *first_cell / *num_cells describe the part of the triangle they are.
int all2all_reduce_scatter_device(kdbx_ctx* ctx, uint32_t* d_block, uint64_t* first_cell, uint64_t* num_cells, kdbx_stats* stats) {
    if (!ctx->loaded) return ctx->fail(KDBX_ERR_STATE, "no patterns loaded (call kdbx_load_patterns first)");
    const int nranks = ctx->comm ? ctx->comm_nranks : 1, rank = ctx->comm ? ctx->comm_rank : 0;
    const uint64_t N = ctx->N;
    const uint64_t cells = N ? N * (N - 1) / 2 : 0;
    const uint64_t B = comm_block_cells(cells, nranks);
    if (first_cell) *first_cell = std::min<uint64_t>(cells, (uint64_t)rank * B);
    if (num_cells) *num_cells = std::min<uint64_t>(cells, (uint64_t)(rank + 1) * B) - std::min<uint64_t>(cells, (uint64_t)rank * B);
    if (B && !d_block) return ctx->fail(KDBX_ERR_ARG, "output pointer is NULL");
    CK(cudaSetDevice(ctx->device));
    ctx->tri_rows_valid = false;   // ctx->tri holds a PARTIAL matrix from here on
    CK(ctx->tri.ensure(((uint64_t)nranks * B + 4) * 4));
    if ((uint64_t)nranks * B > cells)
        CK(cudaMemsetAsync(ctx->tri.as<uint32_t>() + cells, 0, ((uint64_t)nranks * B - cells) * 4, ctx->stream));
    kdbx_stats s{};
    if (int rc = all2all_rows_device(ctx, 0, (uint32_t)N, ctx->tri.as<uint32_t>(), &s)) return rc;
    cudaEvent_t a = ctx->event();
    if (B) {
        if (nranks > 1)
            NCK(nccl_api()->ReduceScatter(ctx->tri.p, d_block, B, ncclUint32, ncclSum, static_cast<ncclComm_t>(ctx->comm), ctx->stream));
        else
            CK(cudaMemcpyAsync(d_block, ctx->tri.p, B * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    cudaEvent_t b = ctx->event();
    CK(cudaStreamSynchronize(ctx->stream));
    s.ms_collective = elapsed(a, b);
    s.ms_total += s.ms_collective;
    if (stats) *stats = s;
    return KDBX_OK;
}
