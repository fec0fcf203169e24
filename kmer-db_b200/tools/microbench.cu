// Microbenchmarks behind the design of k_scatter_add (DESIGN.md "measured rates"): how many
// `cell += w` updates per second a B200 sustains with (a) warp-private shared-memory tiles and
// plain LDS/IADD/STS, (b) red.shared.add.u32 on a CTA-shared tile, (c) red.global.add.u32 into
// an L2-resident matrix.  Ids are pre-generated (sorted runs or random) and streamed from HBM
// exactly like the real kernel streams decoded sample-id runs.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int TILE = 1024;

template <int MODE>  // 0: warp-private LDS/STS, 1: CTA-shared red.shared, 2: red.global
__global__ void __launch_bounds__(256) k_bench(const uint32_t* __restrict__ ids, size_t per_warp, uint32_t* __restrict__ mat,
                                               uint32_t* __restrict__ sink) {
    extern __shared__ uint32_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* tile = (MODE == 0) ? smem + warp * TILE : smem;
    const int tile_words = (MODE == 0) ? 8 * TILE : TILE;
    for (int i = threadIdx.x; i < tile_words; i += blockDim.x) smem[i] = 0;
    __syncthreads();
    const size_t gw = (size_t)blockIdx.x * 8 + warp;
    const uint32_t* p = ids + gw * per_warp;
    uint32_t* row = mat + (gw % 977) * TILE;
    for (size_t k = lane; k + 96 < per_warp; k += 128) {
        const uint32_t i0 = p[k], i1 = p[k + 32], i2 = p[k + 64], i3 = p[k + 96];
        if (MODE == 0) {
            const uint32_t v0 = tile[i0], v1 = tile[i1], v2 = tile[i2], v3 = tile[i3];
            tile[i0] = v0 + 3; tile[i1] = v1 + 3; tile[i2] = v2 + 3; tile[i3] = v3 + 3;
        } else if (MODE == 1) {
            atomicAdd(&tile[i0], 3u); atomicAdd(&tile[i1], 3u); atomicAdd(&tile[i2], 3u); atomicAdd(&tile[i3], 3u);
        } else {
            atomicAdd(&row[i0], 3u); atomicAdd(&row[i1], 3u); atomicAdd(&row[i2], 3u); atomicAdd(&row[i3], 3u);
        }
    }
    __syncthreads();
    uint32_t acc = 0;
    for (int i = threadIdx.x; i < tile_words; i += blockDim.x) acc += smem[i];
    if (acc == 0xdeadbeef) sink[0] = acc;
}

// ids resident in registers (8 per lane = a 256-id list), reductions into many different rows
// of a CTA-shared tile: the pure red.shared issue rate, no global traffic in the loop.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_regs(const uint32_t* __restrict__ ids, int rows, int reps, uint32_t* __restrict__ sink) {
    extern __shared__ uint32_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cells = rows * 256;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) smem[i] = 0;
    __syncthreads();
    const uint32_t* p = ids + ((size_t)blockIdx.x * WARPS + warp) * 256;
    uint32_t c[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) c[m] = (p[lane + 32 * m] & 255u) * 4u;
    const uint32_t base0 = (uint32_t)__cvta_generic_to_shared(smem);
    uint32_t r = warp * 7 + blockIdx.x;
    for (int it = 0; it < reps; ++it) {
        r = (r * 13 + 5) % (uint32_t)rows;
        const uint32_t base = base0 + r * 1024u;
        const uint32_t lim = 200 + (it & 31);  // most of the 256 columns are live
#pragma unroll
        for (int m = 0; m < 8; ++m)
            if (lane + 32 * m < lim) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(base + c[m]), "r"(3u) : "memory");
    }
    __syncthreads();
    uint32_t acc = 0;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) acc += smem[i];
    if (acc == 0xdeadbeef) sink[0] = acc;
}

template <int WARPS>
void run_regs(const char* name, const uint32_t* d_ids, int blocks, int rows, uint32_t* d_sink) {
    const size_t smem = (size_t)rows * 1024;
    CK(cudaFuncSetAttribute(k_regs<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int reps = 20000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_regs<WARPS><<<blocks, WARPS * 32, smem>>>(d_ids, rows, 100, d_sink);
    CK(cudaEventRecord(a));
    k_regs<WARPS><<<blocks, WARPS * 32, smem>>>(d_ids, rows, reps, d_sink);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    double upd = 0;
    for (int it = 0; it < reps; ++it) upd += 200 + (it & 31);
    upd *= (double)blocks * WARPS;
    printf("%-60s %8.3f ms  %.3e updates/s\n", name, ms, upd / (ms / 1e3));
}

template <int MODE>
void run(const char* name, const uint32_t* d_ids, size_t per_warp, int blocks, uint32_t* d_mat, uint32_t* d_sink) {
    const size_t smem = (MODE == 0) ? 8 * TILE * 4 : TILE * 4;
    CK(cudaFuncSetAttribute(k_bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int it = 0; it < 2; ++it) k_bench<MODE><<<blocks, 256, smem>>>(d_ids, per_warp, d_mat, d_sink);
    CK(cudaEventRecord(a));
    const int iters = 5;
    for (int it = 0; it < iters; ++it) k_bench<MODE><<<blocks, 256, smem>>>(d_ids, per_warp, d_mat, d_sink);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double upd = (double)blocks * 8 * (per_warp / 128 * 128) * iters;
    printf("%-44s %8.3f ms/launch  %.3e updates/s  (%.1f GB/s id stream)\n", name, ms / iters, upd / (ms / 1e3),
           upd * 4 / (ms / 1e3) / 1e9);
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    const int blocks = pr.multiProcessorCount * 7;           // 56 warps / SM like the real kernel at tile 1024
    const size_t per_warp = 1 << 16;                          // ids per warp
    const size_t total = (size_t)blocks * 8 * per_warp;       // ~540 M ids = 2.2 GB > L2
    printf("device %s, %d SMs, %zu ids (%.2f GB)\n", pr.name, pr.multiProcessorCount, total, total * 4 / 1e9);
    std::vector<uint32_t> h(total);
    uint32_t* d_ids; CK(cudaMalloc(&d_ids, total * 4));
    uint32_t *d_mat, *d_sink; CK(cudaMalloc(&d_mat, 977 * TILE * 4)); CK(cudaMalloc(&d_sink, 4));
    CK(cudaMemset(d_mat, 0, 977 * TILE * 4));
    for (int pattern = 0; pattern < 3; ++pattern) {
        uint64_t s = 88172645463325252ull;
        auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
        if (pattern == 0) {  // consecutive runs: ids k, k+1, ... (like contiguous cluster members)
            for (size_t i = 0; i < total; i += 128) { uint32_t st = rnd() % (TILE - 128); for (int j = 0; j < 128; ++j) h[i + j] = st + j; }
        } else if (pattern == 1) {  // sorted distinct random ids per 128-run (like interleaved clusters)
            std::vector<uint32_t> perm(TILE); for (int j = 0; j < TILE; ++j) perm[j] = j;
            for (size_t i = 0; i < total; i += 128) {
                for (int j = 0; j < 128; ++j) std::swap(perm[j], perm[j + rnd() % (TILE - j)]);
                std::sort(perm.begin(), perm.begin() + 128);
                for (int j = 0; j < 128; ++j) h[i + j] = perm[j];
            }
        } else {  // stride-4 runs (every 4th sample: 4 interleaved clusters) -> 4-way bank reuse pattern
            for (size_t i = 0; i < total; i += 128) { uint32_t st = rnd() % (TILE - 512); for (int j = 0; j < 128; ++j) h[i + j] = st + 4 * j; }
        }
        CK(cudaMemcpy(d_ids, h.data(), total * 4, cudaMemcpyHostToDevice));
        const char* pn[] = {"consecutive", "sorted-random", "stride-4"};
        printf("-- id pattern: %s\n", pn[pattern]);
        char nm[128];
        snprintf(nm, sizeof nm, "warp-private LDS/IADD/STS [%s]", pn[pattern]); run<0>(nm, d_ids, per_warp, blocks, d_mat, d_sink);
        snprintf(nm, sizeof nm, "CTA-shared red.shared.add.u32 [%s]", pn[pattern]); run<1>(nm, d_ids, per_warp, blocks, d_mat, d_sink);
        snprintf(nm, sizeof nm, "red.global.add.u32 (L2 matrix) [%s]", pn[pattern]); run<2>(nm, d_ids, per_warp, blocks, d_mat, d_sink);
    }
    // pure shared-memory reduction rate, ids in registers
    printf("-- ids in registers, red.shared.add.u32 only (last id pattern: stride-4 columns)\n");
    run_regs<32>("1 CTA/SM x 32 warps, 128-row x 256-col tile (128 KB)", d_ids, pr.multiProcessorCount, 128, d_sink);
    run_regs<16>("2 CTA/SM x 16 warps, 96-row x 256-col tile (96 KB)", d_ids, pr.multiProcessorCount * 2, 96, d_sink);
    run_regs<8>("4 CTA/SM x 8 warps, 48-row x 256-col tile (48 KB)", d_ids, pr.multiProcessorCount * 4, 48, d_sink);
    for (size_t i = 0; i < total && i < (1u << 24); ++i) h[i] = (uint32_t)(i & 255);  // consecutive columns
    CK(cudaMemcpy(d_ids, h.data(), (1u << 24) * 4, cudaMemcpyHostToDevice));
    printf("-- same, consecutive columns (conflict-free)\n");
    run_regs<32>("1 CTA/SM x 32 warps, 128-row x 256-col tile (128 KB)", d_ids, pr.multiProcessorCount, 128, d_sink);
    run_regs<16>("2 CTA/SM x 16 warps, 96-row x 256-col tile (96 KB)", d_ids, pr.multiProcessorCount * 2, 96, d_sink);
    return 0;
}
