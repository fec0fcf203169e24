/*
 * kdbx.h — C ABI of the B200-native common-k-mer counting path (libkdbx.so).
 *
 * This is the drop-in boundary for ONE path of refresh-bio/kmer-db (v2.3.1): the
 * similarity engine `SimilarityCalculator` over `PrefixKmerDb`'s pattern trie.  The
 * reference has no FFI; the seam it offers is the C++ class declared in
 *     src/similarity_calculator.h:4-16
 * whose methods the mode drivers call once per run
 *     src/console_all2all.cpp:21,34      (all2all)
 *     src/console_all2all_sparse.cpp:30,44 (all2all_sp)
 *     src/console_new2all.cpp:26,78,82   (one2all / one2all_sp)
 * Every entry point below names the reference interface it replaces.  Plain pointers and
 * sizes only: no C++ types, no torch types.  All functions return 0 (KDBX_OK) or a negative
 * error code and never throw; `kdbx_last_error` gives the text.  The host-side C++ mirror
 * (kmer-db_b200/host/similarity_calculator.h) rethrows std::runtime_error so that the CLI
 * keeps the reference's error behaviour (src/main.cpp:51-59).
 *
 * There is NO CPU fallback behind this ABI: every compute entry point fails with
 * KDBX_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef KDBX_H
#define KDBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KDBX_ABI_VERSION 4

enum {
    KDBX_OK = 0,
    KDBX_ERR_ARG = -1,    /* bad argument / malformed trie view            */
    KDBX_ERR_CUDA = -2,   /* CUDA runtime error, or no usable device       */
    KDBX_ERR_STATE = -3,  /* call order violated (e.g. no patterns loaded) */
    KDBX_ERR_NOMEM = -4   /* device or pinned-host allocation failed       */
};

typedef struct kdbx_ctx kdbx_ctx;

/* Replaces the constructor arguments of SimilarityCalculator(num_threads, cacheBufferMb)
 * (src/similarity_calculator.h:6).  `-buffer` (the 8 MB CPU cache block,
 * src/similarity_calculator.cpp:51) has no meaning on the GPU; its analogue is `chunk_ids`,
 * the number of decoded sample ids expanded per pass (sized to stay L2-resident). */
typedef struct kdbx_config {
    int32_t device;        /* CUDA ordinal; -1 = current device                          */
    uint32_t flags;        /* KDBX_FLAG_*                                                */
    uint64_t chunk_ids;    /* decoded ids per chunk; 0 = default                         */
    uint32_t tile_cols;    /* columns of a CTA's shared-memory accumulator tile (multiple
                              of 32); 0 = default                                        */
    uint32_t unit_updates; /* target updates per work unit; 0 = default                  */
    uint64_t sparse_block_cells; /* kdbx_all2all_sparse: dense cells accumulated per row block;
                                    0 = a quarter of the free HBM                          */
    uint64_t query_batch_kmers;  /* kdbx_new2all_batch: k-mers per device pass; 0 = 2^28      */
    uint32_t tile_rows;          /* matrix rows per accumulator tile (power of two <= 32); 0 = default */
    uint32_t scatter_threads;    /* threads per CTA of the scatter-add kernel; 0 = default     */
    uint64_t upload_chunk_bytes; /* KDBX_FLAG_ASYNC_UPLOAD: bytes of Elias-gamma payload per chunk of the transfer (the
                                    decoder starts on a chunk as soon as it has arrived); 0 = 96 MB, at most 8 chunks */
} kdbx_config;

#define KDBX_FLAG_NONE 0u
/* Expand full sample-id lists chunk by chunk (parent-chain walk) even when the lists of all
 * patterns would fit in HBM at once; the default picks the resident, level-ordered expansion
 * whenever 4 * sum(num_samples) bytes take at most 40 % of the free device memory. */
#define KDBX_FLAG_CHUNKED_LISTS 1u
/* kdbx_load_patterns returns as soon as the copies are enqueued (headers on the compute stream, the
 * Elias-gamma payload on a second stream) instead of waiting for them, so that the first stages of
 * the next compute call overlap the transfer: the headers travel first, the payload follows in chunks
 * (kdbx_config::upload_chunk_bytes), and the decoder works on a chunk while the next ones are on the link.
 * The caller's arrays must then stay valid and unchanged until that compute call has returned. */
#define KDBX_FLAG_ASYNC_UPLOAD 2u
/* Form of the full sample lists in the dense all2all.  Default: the library looks at the decoded local lists and
 * keeps every list as its sorted RUN BOUNDARIES (start of each run of consecutive ids, one past its end) when that
 * holds at most 3/4 of the entries the plain id lists would — the analogue of the reference's fast path for 16
 * consecutive ids (src/simd/row_add_avx2.cpp:38-75): a run then costs two accumulator updates whatever its length,
 * and prefix sums along the rows restore the counts when a tile is flushed (uint32 arithmetic: same bits).
 * KDBX_FLAG_ID_LISTS forces plain id lists, KDBX_FLAG_BOUNDARY_LISTS forces run boundaries wherever that form is
 * implemented (one column window, all rows, lists resident in HBM). */
#define KDBX_FLAG_ID_LISTS 4u
#define KDBX_FLAG_BOUNDARY_LISTS 8u

/* Borrowed, read-only SoA view of `std::vector<pattern_t>` (src/pattern.h:42-55) as
 * PrefixKmerDb::getPatterns() exposes it (src/prefix_kmer_db.h:87-175).  One entry per trie
 * node, index = pattern id; entry 0 is the empty sentinel (src/prefix_kmer_db.cpp:24).
 * parent_id[p] < p or -1.  `payload` holds the Elias-gamma coded deltas of each node's LOCAL
 * sample ids (src/elias_gamma.h:104-128), MSB-first in 64-bit words, node p starting at word
 * payload_off[p] and spanning ceil(num_bits[p]/128)*2 words (src/pattern.h:80-82).
 * The library copies what it needs; the view is NOT mutated (the reference's all2all
 * mutates num_kmers in place, src/similarity_calculator.cpp:64-72 — we do not). */
typedef struct kdbx_trie_view {
    uint64_t num_patterns;
    uint32_t num_samples;
    uint32_t _pad;
    const int64_t* num_kmers;          /* pattern_t::num_kmers                          */
    const int64_t* parent_id;          /* pattern_t::parent_id                          */
    const uint32_t* num_samples_full;  /* pattern_t::num_samples (node + ancestors), n  */
    const uint32_t* num_local_samples; /* pattern_t::num_local_samples, l               */
    const uint32_t* last_sample_id;    /* pattern_t::last_sample_id                     */
    const uint32_t* num_bits;          /* pattern_t::num_bits                           */
    const uint64_t* payload_off;       /* word offset of pattern_t::data; NULL = densely
                                          packed in pattern order                       */
    const uint64_t* payload;           /* concatenated pattern_t::data                  */
    uint64_t payload_words;
    /* Optional 32-bit mirrors of the two 64-bit arrays (both or neither; NULL = not given): parent_id as int32 and
     * num_kmers as uint32, for callers that keep them (every value must fit, i.e. equal the 64-bit one).  When present
     * they are what travels to the device — 24 instead of 40 bytes of header per pattern over PCIe — and are widened
     * there. */
    const int32_t* parent_id32;
    const uint32_t* num_kmers32;
} kdbx_trie_view;

/* Counters and per-stage device times of the last compute call (CUDA events on the
 * library's stream).  U is a property of the input trie:
 * U = sum_p l_p (2 n_p - l_p - 1) / 2 = the number of `row[col] += w` executions of the
 * reference's dense path (src/simd/row_add_avx2.cpp:57-72). */
typedef struct kdbx_stats {
    uint64_t updates;        /* U restricted to the requested rows                       */
    uint64_t jobs;           /* (pattern, local position, tile) work items emitted       */
    uint64_t flat_ids;       /* sum_p n_p expanded                                       */
    uint64_t local_ids;      /* sum_p l_p decoded                                        */
    uint64_t units;          /* work units scheduled on warps                            */
    uint32_t chunks;
    uint32_t kernel_launches;/* kernels of this library launched by the call             */
    float ms_upload;         /* H2D of the trie view (kdbx_load_patterns)                */
    float ms_prepare;        /* scans, node packing, W accumulation, gamma decode        */
    float ms_expand;         /* full-list expansion, all chunks                          */
    float ms_bucket;         /* job histogram + scan + scatter + unit build, all chunks  */
    float ms_scatter;        /* the scatter-add kernel, all chunks                       */
    float ms_total;          /* whole compute call on the device                         */
    float ms_download;       /* D2H of the result                                        */
    uint32_t scatter_launches;
    uint32_t _pad;
    uint64_t probes;         /* new2all: k-mers looked up                                */
    uint64_t hits;           /* new2all: k-mers found in the database                    */
    float ms_probe;          /* new2all: hash probe kernel                               */
    float ms_compact;        /* sparse: filter + compaction kernels                      */
    uint64_t physical_updates; /* shared-memory reductions actually issued: = updates with id lists, fewer with
                                  boundary lists (a run of consecutive ids costs two whatever its length)   */
    uint32_t list_form;      /* 0 = sample-id lists, 1 = run-boundary lists (KDBX_FLAG_BOUNDARY_LISTS)      */
    float ms_collective;     /* multi-GPU: the reduce-scatter of the partial matrices                       */
} kdbx_stats;

/* Library / device management ---------------------------------------------------------- */
int kdbx_abi_version(void);
int kdbx_device_count(void);   /* number of sm_100 devices, <0 on CUDA error */
int kdbx_open(const kdbx_config* cfg, kdbx_ctx** out);
void kdbx_close(kdbx_ctx* ctx);
/* Error text of the last failed call on ctx (ctx may be NULL for kdbx_open failures). */
const char* kdbx_last_error(const kdbx_ctx* ctx);

/* Page-locked host memory for the trie view / result so that H2D and D2H run at link rate. */
int kdbx_host_alloc(void** out, size_t bytes);
void kdbx_host_free(void* p);

/* Stage the trie in HBM.  Replaces handing `PrefixKmerDb&` to the calculator
 * (src/console_all2all.cpp:26,34).  Only copies; all decoding happens in the compute calls. */
int kdbx_load_patterns(kdbx_ctx* ctx, const kdbx_trie_view* view);

/* Optional hint after kdbx_load_patterns: every sample id of the staged trie lies in [lo, hi).  The part of a
 * sharded database that one GPU holds (kdbxh_partition) covers a narrow band of sample ids; with the band declared
 * the dense all2all lays its accumulator tiles over that band only, i.e. it plans like a database of hi - lo
 * samples (whole rows in one accumulator tile up to 1728 samples, a sliding window beyond).  The result is unchanged — rows and columns outside the band
 * are zero — and an id outside the band is an error (KDBX_ERR_ARG), not silently dropped.  Reset to [0, N) by
 * every kdbx_load_patterns.  No analogue in the reference (src/similarity_calculator.cpp:290-329 blocks by
 * pattern count only). */
int kdbx_set_sample_window(kdbx_ctx* ctx, uint32_t lo, uint32_t hi);

/* Number of `row[col] += w` updates per matrix row (length num_samples), the quantity
 * row-block sharding is balanced on; sum = U.  Computed on the device. */
int kdbx_row_updates(kdbx_ctx* ctx, uint64_t* out_updates_per_row);

/* Replaces SimilarityCalculator::all2all (src/similarity_calculator.cpp:42-438) including
 * its W accumulation (:64-72), decode (:110-163), counting sort by row (:166-203,244-280)
 * and row_add loop (:206-241; src/simd/row_add.h:16).  Fills the packed lower-triangular
 * matrix exactly as LowerTriangularMatrix<uint32_t> lays it out (src/array.h:140,156-159):
 * row s starts at s(s-1)/2 and has s cells; out_tri has N(N-1)/2 cells.  Bit-exact
 * (uint32 wrap-around included).  `out_tri` is HOST memory. */
int kdbx_all2all_dense(kdbx_ctx* ctx, uint32_t* out_tri, kdbx_stats* stats);

/* Row-block shard of the same matrix for multi-GPU runs: only rows [row_begin,row_end)
 * are computed; out_rows receives their cells, i.e. the packed range
 * [row_begin(row_begin-1)/2, row_end(row_end-1)/2).  Analogue of the reference's row
 * ownership between threads (src/similarity_calculator.cpp:371-395). */
int kdbx_all2all_dense_rows(kdbx_ctx* ctx, uint32_t row_begin, uint32_t row_end,
                            uint32_t* out_rows, kdbx_stats* stats);

/* The step after the path, on the device: the decimal text of the dense table's cells.  For every row s in
 * [row_begin, row_end) the s cells of the packed row, each followed by ',' — exactly what the reference prints
 * after "<name>,<total-kmers>," (src/console_all2all.cpp:65-78 -> LowerTriangularMatrix::saveRow, src/array.h:254-258,
 * plain decimal by src/conversion.h:99-165) — formatted from the matrix that the last kdbx_all2all_dense /
 * kdbx_all2all_dense_rows call on this context left in device memory (the rows must lie inside that call's range).
 * row_off receives rows + 1 byte offsets into the text; *bytes the total.  text == NULL: sizes only.  HOST pointers. */
int kdbx_csv_dense_rows(kdbx_ctx* ctx, uint32_t row_begin, uint32_t row_end, char* text, uint64_t capacity,
                        uint64_t* row_off, uint64_t* bytes);

/* `distance` on the device: DistanceConsole::run (src/console_distance.cpp:74-200) for a dense triangular table.  For
 * every row s in [row_begin, row_end) the s cells measure(common, total_kmers[s], total_kmers[c]) printed with six
 * decimals exactly as num2str(double) does (src/conversion.h:167-219,254-260: "0" for zero, else (uint64)(v * 10^6 + 0.5)
 * as <int>.<6 digits>), each followed by ','.  Measures: KDBX_METRIC_JACCARD / MIN / MAX / COSINE / NUM_KMERS
 * (src/params.cpp:14-42) — one or two correctly rounded IEEE operations, so the bytes are the reference's; the
 * logarithm-based ones stay with the caller.  Works on the resident matrix: the one the last kdbx_all2all_dense[_rows]
 * call left, or a packed triangle staged with kdbx_stage_matrix (what the `distance` mode parsed from a table).
 * sample_kmers: uint32[row_end] "total-kmers", none of them 0.  text == NULL: sizes only.  HOST pointers. */
int kdbx_stage_matrix(kdbx_ctx* ctx, const uint32_t* tri, uint32_t num_samples);
int kdbx_distance_dense_rows(kdbx_ctx* ctx, int metric, const uint32_t* sample_kmers, uint32_t row_begin, uint32_t row_end,
                             char* text, uint64_t capacity, uint64_t* row_off, uint64_t* bytes);

/* Same, result left in DEVICE memory (`d_out_rows` is a CUDA device pointer on ctx's
 * device, e.g. a torch tensor's data_ptr); no D2H inside the call. */
int kdbx_all2all_dense_rows_device(kdbx_ctx* ctx, uint32_t row_begin, uint32_t row_end,
                                   void* d_out_rows, kdbx_stats* stats);

/* ---- sparse all2all ------------------------------------------------------------------- */

/* Output filter evaluated ON THE DEVICE while rows are compacted.  Mirrors CombinedFilter
 * (src/sparse_filters.h:33-61): a cell (row, col, common) is kept iff common != 0,
 * min_common <= common <= max_common (KmerFilter, :26-30) and every metric bound holds
 * (MetricFilter, :12-23: lo <= metric(common, cnt[row], cnt[col]) <= hi in IEEE double).
 * Only measures whose arithmetic is exactly reproducible on the GPU (one division / one sqrt,
 * round-to-nearest) are offered here; the log-based ones (mash, ani, ...) stay with the caller,
 * who filters the returned rows on the host (kmer-db_b200/host does). */
enum { KDBX_METRIC_JACCARD = 0, KDBX_METRIC_MIN = 1, KDBX_METRIC_MAX = 2, KDBX_METRIC_COSINE = 3,
       KDBX_METRIC_NUM_KMERS = 4 /* kdbx_distance_dense_rows only */ };
typedef struct kdbx_metric_bound {
    int32_t metric;     /* KDBX_METRIC_*                                                   */
    int32_t _pad;
    double lo, hi;      /* inclusive; use -DBL_MAX / DBL_MAX for an open side               */
} kdbx_metric_bound;
typedef struct kdbx_filter {
    uint32_t min_common, max_common;      /* KmerFilter bounds; 0 / UINT32_MAX = open        */
    uint32_t num_metric_bounds;           /* <= 4                                            */
    uint32_t _pad;
    kdbx_metric_bound metric_bounds[4];
    const uint32_t* sample_kmers;         /* uint32[num_samples] "total-kmers" of each sample
                                             (src/kmer_db.h:38); required iff bounds are given */
} kdbx_filter;

/* Rows of sorted (col, val) pairs, val > 0 and passing the filter: the content of the
 * reference's SparseMatrix after compact2 (src/array.h:391-446).  Arrays are host memory owned
 * by the library (page-locked when the result came from one row block); release with
 * kdbx_free_csr. */
typedef struct kdbx_csr {
    uint32_t num_rows;      /* = num_samples                                               */
    uint32_t _pad;          /* owned by the library (allocation kind)                       */
    uint64_t nnz;
    uint64_t* row_ptr;      /* num_rows + 1; row s = [row_ptr[s], row_ptr[s+1])             */
    uint32_t* col;          /* ascending within a row, all < s                              */
    uint32_t* val;
} kdbx_csr;

/* Replaces SimilarityCalculator::all2all_sp + SparseMatrix::compact2 + CBubbleHelper
 * (src/similarity_calculator.cpp:442-657, src/array.h:391-446, src/bubble_helper.h): the same
 * matrix as kdbx_all2all_dense, delivered as sparse rows.  The reference keeps one hash map per
 * row and defers patterns with >= bubbleSize samples; here blocks of rows are accumulated
 * densely in HBM by the same scatter-add kernels (180 GB holds the whole triangle up to
 * N ~ 2*10^5) and compacted by a filter + prefix-sum kernel, so bubbles have no analogue.
 * `filter` may be NULL (keep every non-zero cell). */
int kdbx_all2all_sparse(kdbx_ctx* ctx, const kdbx_filter* filter, kdbx_csr* out, kdbx_stats* stats);
/* Rows [row_begin, row_end) of the same result (the other rows come back empty).  The unit of a multi-GPU sparse
 * run: every device stages the whole trie, takes a block of rows balanced on kdbx_row_updates, and the caller
 * concatenates the rows — the grid of the reference's all2all-parts (src/console_all2all_parts.cpp:143-331) with the
 * database replicated instead of split, and no exchange between the devices. */
int kdbx_all2all_sparse_rows(kdbx_ctx* ctx, uint32_t row_begin, uint32_t row_end, const kdbx_filter* filter,
                             kdbx_csr* out, kdbx_stats* stats);
void kdbx_free_csr(kdbx_csr* csr);

/* ---- new2all: query samples against the database --------------------------------------- */

/* The database's prefix-bucketed k-mer tables, PrefixKmerDb::hashtables (src/prefix_kmer_db.h:198),
 * in the raw in-memory form of hash_map_lp<uint32_t suffix, int32_t pattern_id>
 * (src/hashmap_lp.h:69-99): table t occupies slots[slot_off[t] .. slot_off[t+1]), its size is a
 * power of two, a slot is the 8-byte item {uint32 key; int32 val} and val == INT32_MAX marks an
 * empty slot; a key lives at or after fmix32(key) & (size-1) with linear probing (:308-333). */
typedef struct kdbx_tables_view {
    uint64_t num_tables;        /* 2^max(8, k*bits_per_symbol - 32) (src/prefix_kmer_db.cpp:54-62) */
    const uint64_t* slot_off;   /* num_tables + 1                                           */
    const uint64_t* slots;
} kdbx_tables_view;

/* Stage the tables in HBM (replaces handing PrefixKmerDb::getHashtables() to one2all,
 * src/similarity_calculator.cpp:815). */
int kdbx_load_hashtables(kdbx_ctx* ctx, const kdbx_tables_view* view);

/* Replaces the per-query calls of SimilarityCalculator::one2all<false> / one2all_sp that
 * New2AllConsole issues from its worker threads (src/console_new2all.cpp:64-94,
 * src/similarity_calculator.cpp:810-925, 929-1051), batched: query q owns the ascending,
 * unique k-mers kmers[q_off[q] .. q_off[q+1]) (KmerHelper::unique, src/kmer_extract.h:113-119).
 * out[q * num_samples + s] = number of the query's k-mers present in database sample s.
 * Needs kdbx_load_patterns and kdbx_load_hashtables.  `out` is HOST memory, n_queries x N. */
int kdbx_new2all_batch(kdbx_ctx* ctx, const uint64_t* kmers, const uint64_t* q_off, uint32_t n_queries,
                       uint32_t* out, kdbx_stats* stats);

/* ---- database against database: one cell of the grid of all2all-parts ---------------------- */

/* Replaces SimilarityCalculator::db2db_sp + SparseMatrix::compact2 (src/similarity_calculator.cpp:1225-1540,
 * src/array.h:391-446) as All2AllPartsConsole::run calls them for every cell (row part i, column part j < i) of the
 * grid of partial databases (src/console_all2all_parts.cpp:163-254): out row s1 (a sample of the ROW database) holds
 * the ascending (s2, common) pairs of the samples s2 of the COLUMN database that share k-mers with it and pass the
 * filter; common = number of k-mers present in both samples.  Both databases must be staged — kdbx_load_patterns
 * AND kdbx_load_hashtables — on two contexts of the same device, with equal k-mer length and alphabet (equal numbers
 * of prefix buckets).  The k-mers both databases hold are found by probing the column database's tables with every
 * k-mer of the row database's tables (the reference merges the sorted buckets, :1262-1287), the (pattern, pattern)
 * pairs are sorted and counted (:1310-1321), and every distinct pair adds its count to the cells of
 * list(pattern1) x list(pattern2) (:1434-1517) in a dense block of rows in HBM, compacted like kdbx_all2all_sparse.
 * filter->sample_kmers are the ROW database's "total-kmers", cols_sample_kmers the COLUMN database's (both needed iff
 * metric bounds are given; CombinedFilter(row counts, column counts), src/console_all2all_parts.cpp:181-186).
 * out->num_rows = samples of the row database; column ids are local to the column database.  stats->updates = number
 * of cell updates, probes = k-mers of the row database, hits = k-mers found in both. */
int kdbx_db2db_sparse(kdbx_ctx* rows_db, kdbx_ctx* cols_db, const kdbx_filter* filter, const uint32_t* cols_sample_kmers,
                      kdbx_csr* out, kdbx_stats* stats);

/* Pattern-sharded variant for multi-GPU runs: the trie is cut into chunks of patterns (the
 * library's unit of streaming, kdbx_config::chunk_ids); this call executes the chunks c with
 * c % num_parts == part and leaves a PARTIAL matrix (whole packed triangle, N(N-1)/2 cells) in
 * device memory.  The element-wise uint32 sum of the partial matrices of all parts is the matrix
 * of kdbx_all2all_dense — the caller adds them with one collective (bench.py: NCCL all-reduce /
 * reduce-scatter over NVLink).  stats->updates counts the executed chunks only.  The reference's
 * nearest analogue is the per-thread partition of each cache block, src/similarity_calculator.cpp:
 * 302-327. */
int kdbx_all2all_dense_part_device(kdbx_ctx* ctx, uint32_t part, uint32_t num_parts, void* d_out_tri,
                                   kdbx_stats* stats);

/* ---- several GPUs of one node -------------------------------------------------------------- */

/* The reference is one process; its template for sharding is the grid of partial databases of all2all-parts
 * (src/console_all2all_parts.cpp:143-331).  Here the database is cut into sub-tries (kdbxh_partition: the matrix is
 * linear in num_kmers, so the parts' matrices add up), one context and one GPU per part, and ONE collective —
 * ncclReduceScatter(uint32, sum) over NVLink — adds the partial matrices and leaves rank r with the cells
 * [r B, (r+1) B) of the packed triangle, B = ceil(N(N-1)/2 / ranks).  NCCL is bound at run time (libnccl.so.2).
 *
 * One process per GPU: rank 0 calls kdbx_comm_unique_id, the launcher hands the 128 bytes to every rank, and every
 * rank calls kdbx_comm_init_rank (collective: all ranks must call).  One process, several contexts:
 * kdbx_comm_init_all; the compute calls below must then be issued from one host thread per context. */
#define KDBX_COMM_ID_BYTES 128
int kdbx_comm_unique_id(void* id_out);
int kdbx_comm_init_rank(kdbx_ctx* ctx, int nranks, int rank, const void* id);
int kdbx_comm_init_all(kdbx_ctx* const* ctxs, int n);
void kdbx_comm_destroy(kdbx_ctx* ctx);

/* The dense all2all of the sub-trie staged on this context, reduce-scattered over the communicator (a context without
 * one is a communicator of one rank).  The rank's block of the packed triangle — *num_cells cells starting at cell
 * *first_cell (src/array.h:140 layout) — goes to d_block (device memory, room for B cells) or out_block (host). */
int kdbx_all2all_dense_reduce_scatter_device(kdbx_ctx* ctx, void* d_block, uint64_t* first_cell, uint64_t* num_cells,
                                             kdbx_stats* stats);
int kdbx_all2all_dense_reduce_scatter(kdbx_ctx* ctx, uint32_t* out_block, uint64_t* first_cell, uint64_t* num_cells,
                                      kdbx_stats* stats);

/* ---- build: database construction on the device --------------------------------------------- */

/* Replaces, per sample, what BuildConsole::run drives on host threads (src/console_build.cpp:91-118):
 * KmerHelper::extract + MinHashFilter (src/kmer_extract.h:13-97, src/filter.h:40-115), sort + unique of
 * the sample's k-mers (src/kmer_extract.h:101-119), and PrefixKmerDb::addKmers (src/prefix_kmer_db.cpp:
 * 244-434): prefix histogram / hashtable find-or-insert (:67-178) and the pattern extend-or-split
 * (:181-240).  Semantics: SURVEY.md §A.4.  New pattern ids are handed out in ascending order of the
 * pattern they split from (the reference's order within one sample depends on its thread timing, :219). */
typedef struct kdbx_builder kdbx_builder;

typedef struct kdbx_build_params {
    uint32_t kmer_length;        /* -k                                                            */
    uint32_t bits_per_symbol;    /* Alphabet::bitsPerSymbol (src/alphabet.h:32-58)                */
    uint32_t alphabet_size;      /* symbols; the complement of symbol s is size-1-s               */
    uint32_t preserve_strand;    /* non-zero: no canonical form (-preserve-strand, amino acids)   */
    double fraction;             /* -f: minhash fraction, >= 1 keeps every k-mer                  */
    double fraction_start;       /* -f-start                                                      */
    int8_t symbol_map[256];      /* byte -> symbol, < 0 = not in the alphabet                     */
    uint64_t table_capacity_hint;/* expected number of distinct k-mers; 0 = grow on demand        */
} kdbx_build_params;

typedef struct kdbx_build_result {
    uint64_t num_patterns;       /* trie nodes, sentinel included                                 */
    uint64_t payload_words;      /* 64-bit words of Elias-gamma payload                           */
    uint64_t num_tables;         /* 2^max(8, k*bits - 32)                                         */
    uint64_t total_slots;        /* slots of all raw prefix tables                                */
    uint64_t kmers_count;        /* distinct k-mers in the database                               */
    uint64_t sum_local_samples;  /* sum of num_local_samples                                      */
    uint64_t table_capacity;     /* slots of the device-side k-mer table                          */
    uint32_t num_samples;
    uint32_t kernel_launches;
    uint32_t table_growths;
    float ms_finish;             /* gamma coding + table export on the device                     */
    uint64_t reserved[2];
} kdbx_build_result;

/* Caller-owned HOST arrays kdbx_builder_export fills (sizes from kdbx_build_result; NULL = skip):
 * the trie in the layout of kdbx_trie_view and the k-mer tables in the layout of kdbx_tables_view,
 * i.e. exactly what PrefixKmerDb::serialize writes (src/prefix_kmer_db.cpp:438-574). */
typedef struct kdbx_build_arrays {
    int64_t* num_kmers;           /* [num_patterns]      */
    int64_t* parent_id;           /* [num_patterns]      */
    uint32_t* num_samples_full;   /* [num_patterns]      */
    uint32_t* num_local_samples;  /* [num_patterns]      */
    uint32_t* last_sample_id;     /* [num_patterns]      */
    uint32_t* num_bits;           /* [num_patterns]      */
    uint64_t* payload_off;        /* [num_patterns]      */
    uint64_t* payload;            /* [payload_words]     */
    uint64_t* slot_off;           /* [num_tables + 1]    */
    uint64_t* slots;              /* [total_slots]       */
    uint64_t* table_filled;       /* [num_tables]        */
} kdbx_build_arrays;

int kdbx_builder_open(kdbx_ctx* ctx, const kdbx_build_params* params, kdbx_builder** out);
void kdbx_builder_close(kdbx_builder* b);
/* `build -extend` (src/console_build.cpp:48-57): continue the database staged on the context with
 * kdbx_load_patterns + kdbx_load_hashtables.  Must precede the first sample. */
int kdbx_builder_adopt(kdbx_builder* b);
/* One sample from its sequence: `symbols` are the sample's records back to back, separated by any
 * byte outside the alphabet (k-mers never span records).  unique_kmers (may be NULL) receives the
 * sample's "total-kmers" (src/kmer_db.h:127-129).  Samples get consecutive ids in call order. */
int kdbx_builder_add_sequence(kdbx_builder* b, const char* symbols, uint64_t len, uint64_t* unique_kmers);
/* One sample from k-mers that are already canonical, shifted, filtered, ascending and unique. */
int kdbx_builder_add_kmers(kdbx_builder* b, const uint64_t* kmers, uint64_t count);
/* Elias-gamma coding of the local sample lists and conversion of the k-mer table into the
 * reference's prefix-bucketed raw tables, on the device; reports the sizes to allocate. */
int kdbx_builder_finish(kdbx_builder* b, kdbx_build_result* out);
int kdbx_builder_export(kdbx_builder* b, const kdbx_build_arrays* arrays);

/* new2all from the queries' sequences: the k-mer extraction, minhash, sort and unique that New2AllConsole's
 * loader threads run per query on the host (src/console_new2all.cpp:64-94, src/kmer_extract.h:13-119) happen on
 * the device with `params` (the database's k, alphabet and fraction — as for kdbx_builder_open), then the same
 * probe / count / scatter as kdbx_new2all_batch.  Query q = symbols[q_off[q] .. q_off[q+1]) (records separated by
 * a byte outside the alphabet); unique_kmers[q] (may be NULL) receives its number of distinct k-mers, the
 * "total-kmers" column of the reference's table.  `out` is HOST memory, n_queries x N. */
int kdbx_new2all_sequences(kdbx_ctx* ctx, const kdbx_build_params* params, const char* symbols, const uint64_t* q_off,
                           uint32_t n_queries, uint32_t* out, uint64_t* unique_kmers, kdbx_stats* stats);

/* Debug / test taps (not used by the product path): copy intermediate device arrays of the
 * last compute call to the host.  what: 0 = W (uint32[P]), 1 = decoded local ids
 * (uint32[sum l]), 2 = local offsets (uint64[P+1]).  Returns elements written or <0. */
int64_t kdbx_debug_fetch(kdbx_ctx* ctx, int what, void* out, uint64_t max_elems);

#ifdef __cplusplus
}
#endif
#endif /* KDBX_H */
