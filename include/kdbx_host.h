/*
 * kdbx_host.h — C ABI of the host-side companion library (libkdbx_host.so, no CUDA code).
 *
 * It holds what sits either side of the GPU path in the reference's mode drivers and that a
 * caller in another language needs in order to drive libkdbx.so end to end:
 *   - the .db reader/writer          (PrefixKmerDb::deserialize / serialize,
 *                                      src/prefix_kmer_db.cpp:578-748 / 438-574)
 *   - the byte-exact CSV emitters    (All2AllConsole::run, src/console_all2all.cpp:40-78)
 *   - FASTA ingest and the database builder (LoaderEx, KmerHelper, PrefixKmerDb::addKmers)
 *   - the synthetic database generator used by bench.py and the tests (ours).
 * Everything returns 0 or a negative code; kdbxh_last_error() gives the (thread-local) text.
 */
#ifndef KDBX_HOST_H
#define KDBX_HOST_H

#include "kdbx.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kdbxh_trie kdbxh_trie;

typedef struct kdbxh_synth_params {
    uint32_t num_samples;
    uint32_t num_clusters;
    uint64_t genome_kmers;   /* distinct k-mers per genome (~ length in bp) */
    uint32_t k;
    int32_t interleaved;     /* 0: clusters contiguous in sample order; 1: round-robin */
    double mutation_rate;    /* substitutions per base vs the template genome */
    uint64_t seed;
    int32_t threads;         /* 0 = all host cores */
    int32_t _pad;
    double cluster_skew;     /* 0: equal clusters; s in (0,1): sizes spread over [1-s, 1+s] x the mean */
} kdbxh_synth_params;

typedef struct kdbxh_totals {
    uint64_t num_patterns;
    uint64_t num_samples;
    uint64_t updates;        /* U = sum l(2n-l-1)/2 */
    uint64_t sum_n;
    uint64_t sum_l;
    uint64_t payload_bytes;
    uint64_t kmers_count;
    uint32_t kmer_length;
    uint32_t _pad;
    double fraction;
} kdbxh_totals;

const char* kdbxh_last_error(void);

/* pinned != 0: arrays live in page-locked memory (needs libkdbx.so + a CUDA device). */
kdbxh_trie* kdbxh_trie_new(int pinned);
void kdbxh_trie_free(kdbxh_trie* t);

int kdbxh_read_db(kdbxh_trie* t, const char* path);
int kdbxh_write_db(const kdbxh_trie* t, const char* path);
int kdbxh_synth(kdbxh_trie* t, const kdbxh_synth_params* params);

/* Structural checks a valid kmer-db trie satisfies (parent < child, n = n_parent + l,
 * ascending local lists that continue the parent's list, payload bounds).  0 = valid. */
int kdbxh_validate(const kdbxh_trie* t);

/* dst := the sub-database of the first `num_samples` samples.  Only defined when those samples
 * share no pattern with later ones and their patterns form a prefix of the pattern order (true
 * at cluster boundaries of cluster-contiguous generated databases); anything else is an error.
 * Used to cut a bounded sample of a workload for the CPU baseline. */
int kdbxh_prefix(const kdbxh_trie* src, uint32_t num_samples, kdbxh_trie* dst);

/* Sharding for multi-GPU runs (ours; the reference is single-process).  dst := part `part` of
 * `num_parts`: a valid database over the same samples that holds a contiguous piece of the trie's
 * depth-first preorder (balanced on the GPU pipeline's cost model) plus the ancestors of that piece
 * with num_kmers = 0.  The shared-k-mer matrix is linear in num_kmers, so the matrices of the parts
 * sum (uint32, wrapping) to the matrix of src: one GPU per part + one NCCL all-reduce.
 * owned_updates (may be NULL) receives U of the patterns the part owns. */
int kdbxh_partition(const kdbxh_trie* src, uint32_t num_parts, uint32_t part, kdbxh_trie* dst, uint64_t* owned_updates);
/* All parts of one cut without repeating the preorder walk: src must outlive the partitioner.  window (may be
 * NULL) receives the band of sample ids [lo, hi) that the part's lists lie in, for kdbx_set_sample_window. */
typedef struct kdbxh_partitioner kdbxh_partitioner;
kdbxh_partitioner* kdbxh_partitioner_new(const kdbxh_trie* src, uint32_t num_parts);
void kdbxh_partitioner_free(kdbxh_partitioner* p);
int kdbxh_partitioner_part(const kdbxh_partitioner* p, uint32_t part, kdbxh_trie* dst, uint64_t* owned_updates, uint32_t* window);
/* The same for all parts at once, each written to "<prefix><part>of<num_parts>.db" by its own host thread.  The
 * arrays (may be NULL) receive per part: U of the owned patterns, the sample window (2 entries), U of the part as a
 * database of its own (owned + ancestor chain), its number of patterns. */
int kdbxh_partition_write_all(const kdbxh_trie* src, uint32_t num_parts, const char* prefix, uint64_t* owned_updates,
                              uint32_t* windows, uint64_t* part_updates, uint64_t* part_patterns);
/* Shifts every sample id of t by `offset` inside a sample table of `new_total` entries (the other
 * entries are empty samples); the trie keeps its shape.  Lays shards of a workload side by side. */
int kdbxh_relabel(kdbxh_trie* t, uint32_t offset, uint32_t new_total);

int kdbxh_view(const kdbxh_trie* t, kdbx_trie_view* out);
int kdbxh_totals_of(const kdbxh_trie* t, kdbxh_totals* out);
/* name of sample i (NUL-terminated, owned by the trie) and its total-kmers count */
const char* kdbxh_sample_name(const kdbxh_trie* t, uint32_t i);
uint64_t kdbxh_sample_kmers(const kdbxh_trie* t, uint32_t i);

/* Like kdbxh_read_db, but also loads the k-mer tables (DeserializationMode::Everything,
 * src/kmer_db.h:55-60) — what new2all and build -extend need. */
int kdbxh_read_db_full(kdbxh_trie* t, const char* path);
/* Flattened view of the trie's k-mer tables for kdbx_load_hashtables (one contiguous slot array,
 * built on first use and owned by the trie).  Error when the trie has no tables. */
int kdbxh_tables_view(kdbxh_trie* t, kdbx_tables_view* out);

/* Incremental construction of a database from per-sample k-mer sets: the restatement of
 * PrefixKmerDb::addKmers (src/prefix_kmer_db.cpp:244-434) that the `build` mode drives.
 * kmers: ascending, unique, nt alphabet, already canonical / shifted (see kdbxh_samples_load). */
typedef struct kdbxh_builder kdbxh_builder;
kdbxh_builder* kdbxh_builder_new(int threads);
void kdbxh_builder_free(kdbxh_builder* b);
int kdbxh_builder_add_sample(kdbxh_builder* b, const char* name, const uint64_t* kmers, uint64_t count, uint32_t k, double fraction);
int kdbxh_builder_finish(kdbxh_builder* b, kdbxh_trie* out);

/* FASTA ingest (LoaderEx + KmerHelper::extract + MinHashFilter, src/loader_ex.cpp, src/kmer_extract.h:13-97,
 * src/filter.h:40-115): list_arg is a sample list file or one FASTA file; every sample's k-mers come
 * back ascending and unique.  alphabet_id: enum AlphabetType (src/alphabet.h:10-18), 0 = nt. */
typedef struct kdbxh_samples kdbxh_samples;
kdbxh_samples* kdbxh_samples_load(const char* list_arg, uint32_t k, double fraction, double fraction_start, int32_t alphabet_id,
                                  int multisample, int threads);
void kdbxh_samples_free(kdbxh_samples* s);
uint32_t kdbxh_samples_count(const kdbxh_samples* s);
const char* kdbxh_samples_name(const kdbxh_samples* s, uint32_t i);
const uint64_t* kdbxh_samples_kmers(const kdbxh_samples* s, uint32_t i, uint64_t* count);

/* tri: packed lower-triangular uint32 matrix, N(N-1)/2 cells (src/array.h:140). */
int kdbxh_write_all2all_csv(const kdbxh_trie* t, const uint32_t* tri, const char* path, int sparse);

/* The one2all table (src/console_one2all.cpp:82-92): the database's two header lines and one row
 * `<sample>,<kmers>,<sims[0]>,...,<sims[N-1]>,` without a newline after it. */
int kdbxh_write_one2all_csv(const kdbxh_trie* t, const char* sample, uint64_t kmers, const uint32_t* sims, const char* path);

/* The all2all-sp / all2all-parts table from sparse rows (src/console_all2all_sparse.cpp:50-98, SparseMatrix::saveRowSparse
 * src/array.h:625-637), with the reference's output options evaluated on the host:
 *   filters      the command-line words of -min / -max, e.g. "-min jaccard:0.9 -max 100" (NULL = none; src/params.cpp:418-455)
 *   sample_rows  "<criterion>:<count>" of -sample-rows (NULL = off; src/sampler.h, src/array.h:451-541): every sample keeps
 *                its <count> best neighbours of the symmetric matrix
 * The matrix comes as a grid of cells: rows of cells[i] are the samples row_shifts[i] + r of `t`, its columns the samples
 * col_shifts[i] + c (one cell with both shifts 0 = all2all-sp; the cells of all2all-parts otherwise, which needs
 * sample_rows: without it that mode writes its rows as they are computed).  saved (optional) = pairs written. */
int kdbxh_write_sparse_csv(const kdbxh_trie* t, const kdbx_csr* cells, const uint32_t* row_shifts, const uint32_t* col_shifts,
                           uint32_t num_cells, const char* filters, const char* sample_rows, const char* path, uint64_t* saved);

#ifdef __cplusplus
}
#endif
#endif /* KDBX_HOST_H */
