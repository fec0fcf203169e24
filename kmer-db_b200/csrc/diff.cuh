// Boundary ("difference") form of the dense all2all (included by kdbx.cu; shares its anonymous namespace).
//
// The reference's row_add has a fast path for 16 consecutive sample ids (src/simd/row_add_avx2.cpp:38-75):
// databases built from related genomes hold long runs of consecutive ids in their sample lists — a delta of 1 is
// a single bit of the Elias-gamma stream (src/elias_gamma.h:104-128).  This file takes that idea to its end.
// Adding w to the cells c of a run [s, e] of one matrix row is the same as adding +w at s and -w at e + 1 to the
// row's DIFFERENCE array and taking prefix sums once, when the accumulator tile is flushed: a run costs two
// shared-memory reductions whatever its length.  So a pattern's full list is kept as its sorted list of run
// boundaries  B = [s1, e1+1, s2, e2+1, ...]  (even positions open a run, odd positions close one), a row r of
// the pattern receives the boundaries b < r (a prefix of B: the ids below r are exactly the ids of the list
// before r; a run that contains r is left open — the cells at and above the diagonal are never read), with
// weight +w at even and -w at odd positions, and the flush turns differences into counts.  All arithmetic is
// uint32 modulo 2^32, so the result is bit-identical to the id form.  U stays the unit of account: the number of
// reductions actually issued is reported next to it (kdbx_stats::physical_updates).
//
// B(p) = B(parent) ++ own boundaries of local(p); when local(p) starts right after the parent's last id, the
// parent's final close and the own first open cancel and both are dropped ("joined").
// Used when the lists are resident, there is one column window and all rows are computed (the headline path);
// everything else runs the id form of kdbx.cu.
#pragma once

// nb[p] = entries of B(p): the parent's, minus its final close when joined, plus the own ones; first_id[p] = smallest id
// of the full list (the root's first local id).  One launch per num_samples level, ascending (parents first), like the
// level-order expansion.  slide_cols != 0 (sliding column window, make_plan): the list of p must not reach below the
// window of the row block of p's LAST local id, i.e. first_id + slide_cols >= end of that block; flag 9 otherwise.
__global__ void k_pull_level(uint32_t count, const uint32_t* __restrict__ order, const Node* __restrict__ nodes,
                             const uint32_t* __restrict__ ownb, uint32_t* __restrict__ nb, const uint32_t* __restrict__ loc,
                             uint32_t* __restrict__ first_id, uint32_t slide_cols, uint32_t win_lo, uint32_t rb_shift,
                             int* __restrict__ err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t p = order[i];
    const Node nd = nodes[p];
    const int32_t q = nd.parent;
    const uint32_t own = ownb[p];
    nb[p] = (q >= 0 ? nb[q] : 0u) + (own >> 1) - (own & 1u);
    if (nd.n == 0) return;
    const uint32_t fid = q >= 0 ? first_id[q] : loc[nd.loff];
    first_id[p] = fid;
    if (slide_cols && nd.l) {
        const uint32_t block_end = (((nd.last - win_lo) >> rb_shift) + 1u) << rb_shift;
        if (fid + slide_cols < block_end) atomicExch(err, 9);
    }
}

struct BoundSlots {   // slots of a boundary list: 16-byte aligned like the id lists (ListSlots)
    const uint32_t* p; uint64_t n;
    __host__ __device__ uint64_t operator()(uint64_t i) const { return i < n ? (uint64_t)((p[i] + 3u) & ~3u) : 0ull; }
};

// B(p) = B(parent)[0 .. nb_parent - joined) ++ own boundaries, built from the decoded local ids.
template <uint32_t kLanes>
__global__ void k_expand_level_diff(uint32_t count, const uint32_t* __restrict__ order, const Node* __restrict__ nodes,
                                    const uint64_t* __restrict__ boff, const uint32_t* __restrict__ nb, const uint32_t* __restrict__ ownb,
                                    const uint32_t* __restrict__ loc, uint32_t* bflat, uint64_t capacity, int* __restrict__ err) {
    const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) / kLanes;
    const uint32_t sub = threadIdx.x & (kLanes - 1);
    const uint32_t group_mask = ((kLanes == 32 ? 0u : (1u << kLanes)) - 1u) << ((threadIdx.x & 31u) & ~(kLanes - 1));
    bool have = gid < count;
    Node nd; nd.parent = -1; nd.n = 0; nd.l = 0; nd.last = 0; nd.loff = 0; nd.up2 = nd.up3 = -1;
    uint32_t* dst = bflat;
    uint32_t own = 0, npar = 0;
    if (have) {
        const uint32_t p = order[gid];
        nd = nodes[p];
        own = ownb[p];
        const uint64_t at = boff[p];
        npar = nb[p] - (own >> 1);                 // entries taken over from the parent (its final close dropped when joined)
        if (at + ((nb[p] + 3u) & ~3u) > capacity) { if (sub == 0) atomicExch(err, 8); have = false; }
        dst = bflat + at;
    }
    if (have && nd.parent >= 0 && npar) {
        // both lists start 16-byte aligned; the copy may run up to 3 entries past what is kept — still inside
        // this pattern's slots, overwritten by its own entries below
        const uint4* src = reinterpret_cast<const uint4*>(bflat + boff[nd.parent]);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        const uint32_t n4 = (npar + 3u) >> 2;
        for (uint32_t j = sub; j < n4; j += kLanes) d4[j] = src[j];
    }
    __syncwarp();
    // own entries: id L[j] opens a run unless it continues one (the first id: unless joined), and L[j] + 1 closes
    // the run unless L[j+1] continues it
    const uint32_t joined = own & 1u;
    const uint32_t* L = loc + nd.loff;
    uint32_t run = npar;
    const uint32_t rounds = have ? (nd.l + kLanes - 1) / kLanes : 0u;
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t j = r * kLanes + sub;
        const bool valid = j < nd.l;
        uint32_t cur = 0; bool s = false, e = false;
        if (valid) {
            cur = L[j];
            s = j == 0 ? !joined : (L[j - 1] + 1u != cur);
            e = j + 1 == nd.l || L[j + 1] != cur + 1u;
        }
        const uint32_t cnt = (uint32_t)s + (uint32_t)e;
        uint32_t incl = cnt;
#pragma unroll
        for (uint32_t o = 1; o < kLanes; o <<= 1) {
            const uint32_t up = __shfl_up_sync(group_mask, incl, o, kLanes);
            if (sub >= o) incl += up;
        }
        uint32_t at = run + incl - cnt;
        if (s) dst[at++] = cur;
        if (e) dst[at] = cur + 1u;
        run += __shfl_sync(group_mask, incl, kLanes - 1, kLanes);
    }
}

// ---- job enumeration ------------------------------------------------------------------------------------------
// The jobs of the headline path (one column window, all rows; their number per key was counted by the decoder): a
// pattern x a block of matrix rows = the run of the pattern's local ids that fall into that block.  One lane walks
// one pattern's local ids from first to last, whatever their number (a warp-cooperative walk of long lists, 32
// positions at a time, cost four times the instructions: scans and ballots per round).
//
// Boundary form: the rows are the local ids loc[rows .. rows + k); all of them receive the boundaries [0, c0) (those
// below the first row) and row j those of [c0, c0 + ext) that lie below its id.  Job fields: off/off_hi = start of
// B(p) (40 bits), a = low 32 bits of the row offset in loc, b = c0, A0 = ext | row offset high bits << 8, k, w.
// cnt_j = entries of B(p) below the local id L[j] = cnt0 + sum over i < j of (L[i] opens a run) + (a run closes
// after L[i]): the open L[i] and the close L[i] + 1 (present iff L[i+1] != L[i] + 1) both lie below L[j].
// Id form (same walk, K = false): off/off_hi = start of the full list, a = 0, b = n, A0 = position of the first row.
template <bool kBoundary, class Emit>
__device__ __forceinline__ void walk_runs(uint64_t lo, uint64_t hi, const Node* __restrict__ nodes, const uint64_t* __restrict__ list_off,
                                          const uint32_t* __restrict__ nb, const uint32_t* __restrict__ ownb, const uint32_t* __restrict__ W,
                                          const uint32_t* __restrict__ loc, uint32_t rb_shift, uint64_t stride, uint64_t first_p,
                                          unsigned long long& physical, Emit emit) {
    for (uint64_t p = lo + first_p; p < hi; p += stride) {
        const Node nd = nodes[p];
        if (nd.l == 0) continue;
        const uint32_t w = W[p];
        const uint64_t base = list_off[p];
        uint32_t own = 0, cnt = 0;
        if (kBoundary) { own = ownb[p]; cnt = nb[p] - (own >> 1); }   // boundaries below the first local id
        const uint32_t first = nd.n - nd.l;
        const uint32_t* rows = loc + nd.loff;
        auto make_job = [&](uint32_t run_j, uint32_t k, uint32_t c0, uint32_t ext) {
            Job jb;
            jb.off = (uint32_t)base; jb.off_hi = (uint32_t)(base >> 32);
            if (kBoundary) {
                const uint64_t r = nd.loff + run_j;
                jb.a = (uint32_t)r; jb.b = c0; jb.A0 = ext | ((uint32_t)(r >> 32) << 8);
            } else {
                jb.a = 0; jb.b = nd.n; jb.A0 = first + run_j;
            }
            jb.k = k; jb.w = w; jb.pad = 0;
            return jb;
        };
        uint32_t prev = 0, run_j = 0, run_rb = 0, run_c0 = 0, run_cl = 0;
        bool starts_prev = false;
        for (uint32_t j = 0; j < nd.l; ++j) {
            const uint32_t row = rows[j];
            const bool gap = j != 0 && prev + 1u != row;
            if (kBoundary && j) cnt += (uint32_t)starts_prev + (uint32_t)gap;
            const uint32_t rb = row >> rb_shift;
            if (j == 0) { run_rb = rb; run_c0 = cnt; }
            else if (rb != run_rb) {
                const uint32_t k = j - run_j, i = first + run_j;
                if (w != 0 && (unsigned long long)k * i + k * (k - 1u) / 2u != 0) emit(run_rb, make_job(run_j, k, run_c0, run_cl - run_c0));
                run_j = j; run_rb = rb; run_c0 = cnt;
            }
            run_cl = cnt;
            physical += kBoundary ? cnt : first + j;
            starts_prev = j == 0 ? !(own & 1u) : gap;
            prev = row;
        }
        const uint32_t k = nd.l - run_j, i = first + run_j;
        if (w != 0 && (unsigned long long)k * i + k * (k - 1u) / 2u != 0) emit(run_rb, make_job(run_j, k, run_c0, run_cl - run_c0));
    }
}

// list_off = boff (boundary lists) or noff (id lists); per = patterns per block, the slices the decoder counted by
template <bool kBoundary>
__global__ void __launch_bounds__(kBucketThreads, 4)
k_job_fill_runs(uint64_t P, const Node* __restrict__ nodes, const uint64_t* __restrict__ list_off, const uint32_t* __restrict__ nb,
                const uint32_t* __restrict__ ownb, const uint32_t* __restrict__ W, const uint32_t* __restrict__ loc, uint32_t rb_shift,
                uint32_t nkeys, const uint32_t* __restrict__ blockbase, Job* __restrict__ jobs, uint64_t per,
                unsigned long long* __restrict__ physical_total) {
    __shared__ uint32_t s_next[kDecodeHistKeys];
    const uint32_t* mine = blockbase + (size_t)blockIdx.x * nkeys;
    for (uint32_t k = threadIdx.x; k < nkeys; k += blockDim.x) s_next[k] = mine[k];
    __syncthreads();
    uint64_t lo, hi;
    block_slice(0, P, lo, hi, per);
    unsigned long long physical = 0;
    walk_runs<kBoundary>(lo, hi, nodes, list_off, nb, ownb, W, loc, rb_shift, blockDim.x, threadIdx.x, physical,
                         [&](uint32_t key, const Job& jb) {
                             const uint32_t slot = atomicAdd(&s_next[key], 1u);
                             jobs[slot] = jb;
                         });
    for (int o = 16; o; o >>= 1) physical += __shfl_xor_sync(0xffffffffu, physical, o);
    if ((threadIdx.x & 31) == 0 && physical) atomicAdd(physical_total, physical);
}

// ---- the scatter-add kernel, boundary form --------------------------------------------------------------------
// Same schedule as k_scatter_add (persistent CTAs, work units of one row block, CTA-owned accumulator tile in
// shared memory, a warp per job, one load of a list entry feeds all k rows of the job) with three differences:
// the list entries are run boundaries and carry the weight +w (even positions) or -w (odd positions); the rows'
// accumulator addresses and ids sit in a small per-warp shared-memory table (one broadcast LDS.64 per row instead
// of shuffles); the flush takes prefix sums along every row before it adds the tile into the packed triangle.
constexpr uint32_t kDiffBatch = 8;

// tile[..] += v iff y < lim.  Not a predicated reduction: ptxas turns `@p red.shared` (and an `if` around the asm
// statement alike) into a branch with a convergence barrier, six instructions per guarded reduction; adding 0 instead
// takes four (compare, select, address, reduction).  Entries that are absent hold this lane's padding column (every
// accumulator row is followed by kRowPad padding words), so the address is always inside the tile.
__device__ __forceinline__ void red_shared_add_below(uint32_t saddr, uint32_t v, uint32_t y, uint32_t lim) {
    red_shared_add(saddr, y < lim ? v : 0u);
}

template <uint32_t F, uint32_t G>
__device__ __forceinline__ void last_rows_diff(uint32_t k, uint32_t rows_saddr, uint32_t stride, uint32_t row_base, uint32_t x0, uint32_t x1,
                                               uint32_t x2, uint32_t y0, uint32_t y1, uint32_t y2, uint32_t wl) {
#pragma unroll 2
    for (uint32_t j = 0; j < k; ++j) {
        uint32_t lim;   // 4 * (row id): entries below it count; the row's accumulators start at lim * stride + row_base
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lim) : "r"(rows_saddr + j * 4u));
        const uint32_t ro = lim * stride + row_base;
        if (F > 0) red_shared_add(ro + x0, wl);
        if (F > 1) red_shared_add(ro + x1, wl);
        if (F > 2) red_shared_add(ro + x2, wl);
        if (G > 0) red_shared_add_below(ro + y0, wl, y0, lim);
        if (G > 1) red_shared_add_below(ro + y1, wl, y1, lim);
        if (G > 2) red_shared_add_below(ro + y2, wl, y2, lim);
    }
}

// (64 registers = 1024 threads per SM: the accumulator tile allows one CTA per SM anyway; with __launch_bounds__
// alone ptxas settles on 32 registers and spills)
__global__ void __maxnreg__(64)
k_scatter_diff(const Unit* __restrict__ units, const uint32_t* __restrict__ n_units_ptr, const Job* __restrict__ jobs,
               const uint32_t* __restrict__ bflat, const uint32_t* __restrict__ loc, uint32_t* __restrict__ tri,
               uint32_t id_lo, uint32_t num_rows, uint32_t tile_cols, uint32_t rb_shift, uint32_t* __restrict__ unit_counter) {
    extern __shared__ uint4 tile4[];
    uint32_t* tile = reinterpret_cast<uint32_t*>(tile4);
    __shared__ uint32_t s_unit, s_next_job;
    // per warp: 4 * (row id) of the current job's rows — one broadcast LDS per row; the address of the row's accumulators
    // follows by one multiply-add (a 64-bit entry holding both costs two shared-memory wavefronts per row)
    __shared__ uint32_t s_rows[32][32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_units = *n_units_ptr;
    const uint32_t R = 1u << rb_shift;
    const uint32_t stride = tile_cols + kRowPad;   // padding words: see red_shared_add_below
    const uint32_t tile_saddr = (uint32_t)__cvta_generic_to_shared(tile);
    const uint32_t rows_saddr = (uint32_t)__cvta_generic_to_shared(&s_rows[warp][0]);
    constexpr uint32_t kNone = 0xFFFFFFFFu;
    uint32_t cur_key = kNone;
    for (;;) {
        if (threadIdx.x == 0) s_unit = atomicAdd(unit_counter, 1u);
        __syncthreads();
        const uint32_t u = s_unit;
        const bool done = u >= n_units;
        Unit un; un.key = kNone; un.job_begin = un.job_end = 0; un.pad = 0;
        if (!done) un = units[u];
        if (un.key != cur_key) {
            if (cur_key != kNone) {
                // flush: prefix sums along each row (differences -> counts), then the row goes into the packed triangle as
                // ONE bulk reduction (cp.reduce.async.bulk ... .add.u32: the TMA engine adds up to 6 KB of shared memory into
                // global memory; the warp issues one instruction instead of 32 x nc / 32 atomics).  A bulk operation needs
                // 16-byte aligned addresses on both sides, and a row of the packed triangle starts wherever s(s-1)/2 falls:
                // the scanned row is therefore written back `sh` words to the right of where it was read (sh = the row's
                // global word offset modulo 4; the reads run one 32-column chunk ahead of the writes), which gives the
                // shared and the global address of a column the same alignment.  The up to three columns before the first
                // and after the last aligned one go by ordinary reductions.  A warp per row; only the columns below the
                // diagonal exist.
                const uint32_t row0 = cur_key << rb_shift;
                const uint32_t col0 = row0 + R > tile_cols ? row0 + R - tile_cols : 0u;   // sliding window (make_plan): the tile's first column
                for (uint32_t r = warp; r < R; r += (blockDim.x >> 5)) {
                    const uint32_t row = row0 + r;                 // relative to id_lo, like the columns
                    const uint32_t nc = min(tile_cols, row - min(row, col0));
                    if (nc == 0 || row >= num_rows) continue;   // (the last row block may reach past the matrix: nothing was added there)
                    const uint64_t out0 = tri_offset((uint64_t)row + id_lo) + id_lo + col0;
                    const uint32_t sh = (uint32_t)out0 & 3u;
                    uint32_t* src = tile + r * stride;
                    uint32_t carry = 0;
                    uint32_t v = lane < nc ? src[lane] : 0u;
                    for (uint32_t c0 = 0; c0 < nc; c0 += 32) {
                        const uint32_t cn = c0 + 32 + lane;
                        const uint32_t v_next = cn < nc ? src[cn] : 0u;   // read before the shifted write below can reach it
                        for (int o = 1; o < 32; o <<= 1) { const uint32_t up = __shfl_up_sync(0xffffffffu, v, o); if ((int)lane >= o) v += up; }
                        v += carry;
                        carry = __shfl_sync(0xffffffffu, v, 31);
                        __syncwarp();
                        if (c0 + lane < nc) src[c0 + lane + sh] = v;
                        v = v_next;
                    }
                    // columns [a0, a1) are 16-byte aligned on both sides
                    const uint32_t a0 = min(nc, (4u - sh) & 3u);
                    const uint32_t a1 = a0 + ((nc - a0) & ~3u);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the writes above, before the TMA engine reads them
                    __syncwarp();
                    if (lane == 0 && a1 > a0) {
                        const uint32_t saddr = tile_saddr + (r * stride + a0 + sh) * 4u;
                        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u32 [%0], [%1], %2;" ::"l"(tri + out0 + a0), "r"(saddr),
                                     "r"((a1 - a0) * 4u)
                                     : "memory");
                    }
                    if (lane < a0) { const uint32_t x = src[lane + sh]; if (x) atomicAdd(&tri[out0 + lane], x); }
                    if (a1 + lane < nc) { const uint32_t x = src[a1 + lane + sh]; if (x) atomicAdd(&tri[out0 + a1 + lane], x); }
                }
                // the tile is zeroed next: the bulk reductions must have read it
                if (lane == 0) {
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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
        const uint32_t row0 = un.key << rb_shift;
        const uint32_t col0 = row0 + R > tile_cols ? row0 + R - tile_cols : 0u;   // first column of the tile (sliding window; 0 when whole rows fit)
        const uint32_t row_base = tile_saddr - row0 * 4u * stride - col0 * 4u;   // cell (r, c) sits at (4 r) * stride + 4 c + row_base
        const uint32_t pad_col = col0 + tile_cols + lane;                          // this lane's padding word, as a column
        for (;;) {
            uint32_t jb = 0;
            if (lane == 0) jb = atomicAdd(&s_next_job, kDiffBatch);
            jb = __shfl_sync(0xffffffffu, jb, 0);
            if (jb >= un.job_end) break;
            const uint32_t cnt = min(kDiffBatch, un.job_end - jb);
            uint4 part = make_uint4(0, 0, 0, 0);  // lane 2q: (off, rows, c0, ext|rows_hi<<8) of job q; lane 2q+1: (k, w, off_hi, -)
            if (lane < 2 * cnt) part = ldg_nc_v4(reinterpret_cast<const uint4*>(jobs + jb) + lane);
            // One job of look-ahead: the row ids and the first list entries of job q+1 are requested before job q's
            // reductions are issued, so a warp does not sit out a DRAM round trip per job (the lists of a bucket's jobs
            // are scattered over HBM; with 8 warps per scheduler that latency was the largest stall of the id form).
            struct Pending {
                const uint32_t* list;
                uint32_t c0, e, k, w, row;
                uint32_t v0, v1, v2, v3, v4, v5;   // c0 >= 128: the first 128-entry slice (v0..v3); else the F full groups (v0..v2) and the G guarded ones (v3..v5)
            };
            auto request = [&](uint32_t q, Pending& J) {
                const uint32_t off = __shfl_sync(0xffffffffu, part.x, 2 * q), rows_lo = __shfl_sync(0xffffffffu, part.y, 2 * q);
                const uint32_t c0 = __shfl_sync(0xffffffffu, part.z, 2 * q), ex = __shfl_sync(0xffffffffu, part.w, 2 * q);
                const uint32_t k = __shfl_sync(0xffffffffu, part.x, 2 * q + 1), w = __shfl_sync(0xffffffffu, part.y, 2 * q + 1);
                const uint32_t off_hi = __shfl_sync(0xffffffffu, part.z, 2 * q + 1);
                J.list = bflat + (((uint64_t)off_hi << 32) | off);
                J.c0 = c0; J.e = c0 + (ex & 0xFFu); J.k = k; J.w = w;
                J.row = 0;
                if (lane < k) J.row = ldg_nc_u32(loc + (((uint64_t)(ex >> 8) << 32) | rows_lo) + lane);
                const uint32_t* p = J.list + lane;
                J.v0 = J.v1 = J.v2 = 0u; J.v3 = J.v4 = J.v5 = pad_col;
                if (c0 >= 128) {
                    J.v0 = ldg_nc_u32(p); J.v1 = ldg_nc_u32(p + 32); J.v2 = ldg_nc_u32(p + 64); J.v3 = ldg_nc_u32(p + 96);
                } else {
                    const uint32_t full = c0 >> 5, rem = J.e - full * 32;   // rem: 0..93 entries in the guarded groups
                    if (full > 0) J.v0 = ldg_nc_u32(p);
                    if (full > 1) J.v1 = ldg_nc_u32(p + 32);
                    if (full > 2) J.v2 = ldg_nc_u32(p + 64);
                    const uint32_t* pg = p + full * 32;
                    if (lane < rem) J.v3 = ldg_nc_u32(pg);
                    if (lane + 32 < rem) J.v4 = ldg_nc_u32(pg + 32);
                    if (lane + 64 < rem) J.v5 = ldg_nc_u32(pg + 64);
                }
            };
            // the pass over the rows for up to 3 full groups below c0 (unconditional) and up to 3 groups that reach into
            // [c0, e) — entries of the job's own rows, taken by row j iff they lie below its id (absent entries hold the
            // lane's padding column, which is not below any row)
            auto last_pass = [&](uint32_t k, uint32_t full, uint32_t groups, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t y0, uint32_t y1,
                                 uint32_t y2, uint32_t wl) {
                switch (full * 4 + groups) {
                    case 0: break;
                    case 1: last_rows_diff<0, 1>(k, rows_saddr, stride, row_base, x0, x1, x2, y0, y1, y2, wl); break;
                    case 2: last_rows_diff<0, 2>(k, rows_saddr, stride, row_base, x0, x1, x2, y0, y1, y2, wl); break;
                    case 3: last_rows_diff<0, 3>(k, rows_saddr, stride, row_base, x0, x1, x2, y0, y1, y2, wl); break;
                    case 4: last_rows_diff<1, 0>(k, rows_saddr, stride, row_base, x0, x1, x2, y0, y1, y2, wl); break;
                    case 5: last_rows_diff<1, 1>(k, rows_saddr, stride, row_base, x0, x1, x2, y0, y1, y2, wl); break;
                    case 6: last_rows_diff<1, 2>(k, rows_saddr, stride, row_base, x0, x1, x2, y0, y1, y2, wl); break;
                    case 7: last_rows_diff<1, 3>(k, rows_saddr, stride, row_base, x0, x1, x2, y0, y1, y2, wl); break;
                    case 8: last_rows_diff<2, 0>(k, rows_saddr, stride, row_base, x0, x1, x2, y0, y1, y2, wl); break;
                    case 9: last_rows_diff<2, 1>(k, rows_saddr, stride, row_base, x0, x1, x2, y0, y1, y2, wl); break;
                    case 10: last_rows_diff<2, 2>(k, rows_saddr, stride, row_base, x0, x1, x2, y0, y1, y2, wl); break;
                    case 11: last_rows_diff<2, 3>(k, rows_saddr, stride, row_base, x0, x1, x2, y0, y1, y2, wl); break;
                    case 12: last_rows_diff<3, 0>(k, rows_saddr, stride, row_base, x0, x1, x2, y0, y1, y2, wl); break;
                    case 13: last_rows_diff<3, 1>(k, rows_saddr, stride, row_base, x0, x1, x2, y0, y1, y2, wl); break;
                    case 14: last_rows_diff<3, 2>(k, rows_saddr, stride, row_base, x0, x1, x2, y0, y1, y2, wl); break;
                    default: last_rows_diff<3, 3>(k, rows_saddr, stride, row_base, x0, x1, x2, y0, y1, y2, wl); break;
                }
            };
            Pending cur, nxt;
            request(0, cur);
            nxt = cur;
#pragma unroll 1
            for (uint32_t q = 0; q < cnt; ++q) {
                if (q + 1 < cnt) request(q + 1, nxt);
                const uint32_t k = cur.k, c0 = cur.c0, e = cur.e;
                __syncwarp();   // the previous job's reads of the row table are done
                if (lane < k) s_rows[warp][lane] = cur.row * 4u;
                __syncwarp();
                const uint32_t wl = (lane & 1u) ? 0u - cur.w : cur.w;   // slices start at even positions: the lane's parity is the entry's
                if (c0 < 128) {
                    const uint32_t full = c0 >> 5;
                    last_pass(k, full, (e - full * 32 + 31u) >> 5, cur.v0 * 4u, cur.v1 * 4u, cur.v2 * 4u, cur.v3 * 4u, cur.v4 * 4u, cur.v5 * 4u, wl);
                } else {
                    // whole 128-entry slices of the boundaries every row receives (the first one is already here) ...
                    const uint32_t* list = cur.list;
                    uint32_t x0 = cur.v0 * 4u, x1 = cur.v1 * 4u, x2 = cur.v2 * 4u, x3 = cur.v3 * 4u;
                    uint32_t c = 0;
                    for (;;) {
                        for (uint32_t j = 0; j < k; ++j) {
                            uint32_t lim;
                            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lim) : "r"(rows_saddr + j * 4u));
                            const uint32_t ro = lim * stride + row_base;
                            red_shared_add(ro + x0, wl); red_shared_add(ro + x1, wl);
                            red_shared_add(ro + x2, wl); red_shared_add(ro + x3, wl);
                        }
                        c += 128;
                        if (c + 128 > c0) break;
                        const uint32_t* p = list + c + lane;
                        x0 = ldg_nc_u32(p) * 4u; x1 = ldg_nc_u32(p + 32) * 4u; x2 = ldg_nc_u32(p + 64) * 4u; x3 = ldg_nc_u32(p + 96) * 4u;
                    }
                    // ... then the rest
                    const uint32_t full = (c0 - c) >> 5;                 // 0..3
                    const uint32_t cg = c + full * 32;                   // first entry of the guarded groups
                    const uint32_t rem = e - cg;                          // 0..93
                    const uint32_t* p = list + c + lane;
                    x0 = x1 = x2 = 0;
                    if (full > 0) x0 = ldg_nc_u32(p) * 4u;
                    if (full > 1) x1 = ldg_nc_u32(p + 32) * 4u;
                    if (full > 2) x2 = ldg_nc_u32(p + 64) * 4u;
                    const uint32_t* pg = list + cg + lane;
                    uint32_t y0 = pad_col * 4u, y1 = pad_col * 4u, y2 = pad_col * 4u;
                    if (lane < rem) y0 = ldg_nc_u32(pg) * 4u;
                    if (lane + 32 < rem) y1 = ldg_nc_u32(pg + 32) * 4u;
                    if (lane + 64 < rem) y2 = ldg_nc_u32(pg + 64) * 4u;
                    last_pass(k, full, (rem + 31u) >> 5, x0, x1, x2, y0, y1, y2, wl);
                }
                cur = nxt;
            }
        }
        __syncthreads();
    }
}
