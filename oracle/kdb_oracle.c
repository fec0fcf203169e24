/*
 * kdb_oracle.c — CPU restatement of kmer-db's dense all2all path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product (kmer-db_b200/) never links, imports or executes anything under
 * oracle/.  Plain scalar C, O(U) time, written from the behaviour of refresh-bio/kmer-db v2.3.1:
 *
 *   .db parsing            PrefixKmerDb::deserialize(SkipHashtables)  src/prefix_kmer_db.cpp:578-748
 *                          pattern_t::unpack                          src/pattern.cpp:50-94
 *                          hash_map_lp::deserialize(skipData)         src/hashmap_lp.h:546-566
 *   Elias-gamma decode     CEliasGamma::decode / Decode               src/elias_gamma.h:133-256,371-378
 *   local ids of a node    pattern_t::decodeSamples                   src/pattern.cpp:99-109
 *   W accumulation         SimilarityCalculator::all2all              src/similarity_calculator.cpp:64-72
 *   full list of a node    decomp worker (ancestors' lists, root first) src/similarity_calculator.cpp:130-152
 *   jobs and row updates   sample2pattern + matrix workers            src/similarity_calculator.cpp:154-157,191-196,214-231
 *   row_add                row[ids[k]] += to_add                      src/simd/row_add_avx2.cpp:57-72
 *   packed triangle        LowerTriangularMatrix::operator[]          src/array.h:140
 *   CSV bytes              All2AllConsole::run                        src/console_all2all.cpp:40-78
 *                          num2str / num2str_sparse                   src/conversion.h:248-299
 *   sparse all2all         all2all_sp: every node a flat clique with its own num_kmers, no W
 *                          accumulation (oracle_all2all_bruteforce)   src/similarity_calculator.cpp:442-657
 *   k-mer tables           hash_map_lp raw form, find()               src/hashmap_lp.h:52-64,308-333,546-605
 *   query vs database      one2all<false>                             src/similarity_calculator.cpp:810-925
 *   k-mer extraction       KmerHelper::extract, MinHashFilter         src/kmer_extract.h:13-97, src/filter.h:40-115
 *   FASTA records          GenomeInputFile::extractSubsequences       src/genome_input_file.h:287-337
 *   new2all CSV            New2AllConsole::run                        src/console_new2all.cpp:98-161
 *   database vs database   db2db_sp (k-mer matching, pair counts, list x list adds)  src/similarity_calculator.cpp:1225-1540
 *   all2all-parts          All2AllPartsConsole::run (grid of cells, shifted columns) src/console_all2all_parts.cpp:11-371
 *
 * Parity is PINNED: tests/test_oracle.py checks this file against the reference's own golden
 * CSVs (test/virus/k18.csv, k18.sparse.csv, k24.csv, k18.frac.csv, k18.n2a.csv, k18.n2a.sparse.csv,
 * test/synth/a2a, a2a-sparse, n2a, n2a-sparse; committed under tests/golden/; all2all-parts: the reference's CI check
 * `all2all-parts(part1, part2) == k18.sparse.csv`, .github/workflows/self-hosted.yml:357-363) and against the unmodified reference binary built by
 * oracle/build_ref.sh (oracle/_ref/kmer-db) on generated databases.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t num_patterns;
    uint32_t num_samples;
    uint32_t kmer_length;
    double fraction;
    int64_t* num_kmers;
    int64_t* parent_id;
    uint32_t* n;
    uint32_t* l;
    uint32_t* last;
    uint32_t* bits;
    uint64_t* payload_off; /* in 64-bit words */
    uint64_t* payload;
    uint64_t payload_words;
    char** names;
    uint64_t* sample_kmers;
    /* k-mer tables (only with oracle_db_read_full): table t = slots[table_off[t] .. table_off[t+1]) */
    double start_fraction;
    int32_t alphabet;
    uint64_t num_tables;
    uint64_t* table_off;
    uint64_t* slots; /* {u32 key; i32 val} as one little-endian u64; val == INT32_MAX: empty */
} oracle_db;

/* ---- Elias-gamma: (b-1) ones, a zero, then b-1 low bits, MSB-first in u64 words -------- */
static uint32_t get_bit(const uint64_t* w, uint32_t pos) { return (uint32_t)((w[pos >> 6] >> (63 - (pos & 63))) & 1u); }

static uint32_t gamma_decode_one(const uint64_t* w, uint32_t* pos) {
    uint32_t ones = 0, v = 1;
    while (get_bit(w, *pos)) { ++ones; ++*pos; }
    ++*pos; /* the zero */
    for (uint32_t i = 0; i < ones; ++i) { v = (v << 1) | get_bit(w, *pos); ++*pos; }
    return v;
}

/* local ids of one node, ascending (src/pattern.cpp:99-109) */
void oracle_decode_local(const uint64_t* words, uint32_t l, uint32_t last, uint32_t* out) {
    if (l == 0) return;
    uint32_t pos = 0;
    out[l - 1] = last;
    for (uint32_t i = 0; i + 1 < l; ++i) out[i] = gamma_decode_one(words, &pos); /* deltas */
    for (int64_t i = (int64_t)l - 2; i >= 0; --i) out[i] = out[i + 1] - out[i];
}

/* ---- dense all2all over SoA arrays -------------------------------------------------------
 * tri: N(N-1)/2 uint32 cells, zeroed here.  Returns U (number of row[col] += w executions),
 * or UINT64_MAX on allocation failure.  Does not modify its inputs. */
uint64_t oracle_all2all(uint64_t P, uint32_t N, const int64_t* num_kmers, const int64_t* parent_id, const uint32_t* n,
                        const uint32_t* l, const uint32_t* last, const uint32_t* bits, const uint64_t* payload_off,
                        const uint64_t* payload, uint32_t* tri) {
    (void)bits;
    const uint64_t cells = N ? (uint64_t)N * (N - 1) / 2 : 0;
    memset(tri, 0, cells * sizeof(uint32_t));
    int64_t* W = (int64_t*)malloc(P * sizeof(int64_t));
    uint32_t* full = (uint32_t*)malloc(((size_t)N + 1) * sizeof(uint32_t));
    if (!W || !full) { free(W); free(full); return UINT64_MAX; }
    memcpy(W, num_kmers, P * sizeof(int64_t));
    for (uint64_t i = P; i-- > 1;) /* children before parents: parent_id < own id */
        if (parent_id[i] >= 0) W[parent_id[i]] += W[i];
    uint64_t U = 0;
    for (uint64_t p = 0; p < P; ++p) {
        if (l[p] == 0) continue;
        /* full list, written back to front while walking to the root */
        uint32_t* out = full + n[p];
        for (int64_t q = (int64_t)p; q >= 0; q = parent_id[q]) {
            out -= l[q];
            oracle_decode_local(payload + payload_off[q], l[q], last[q], out);
        }
        const uint32_t to_add = (uint32_t)W[p];
        for (uint32_t i = n[p] - l[p]; i < n[p]; ++i) { /* one job per LOCAL position */
            const uint64_t s = full[i];
            uint32_t* row = tri + s * (s - 1) / 2; /* i == 0 implies no columns; row unused */
            for (uint32_t k = 0; k < i; ++k) row[full[k]] += to_add;
            U += i;
        }
    }
    free(W); free(full);
    return U;
}

/* Pattern-sharded variant (multi-GPU tests): only the jobs of patterns p with (p / chunk) % parts == part
 * are executed, with W from the WHOLE trie; the partial matrices of all parts sum to oracle_all2all's. */
uint64_t oracle_all2all_part(uint64_t P, uint32_t N, const int64_t* num_kmers, const int64_t* parent_id, const uint32_t* n,
                             const uint32_t* l, const uint32_t* last, const uint64_t* payload_off, const uint64_t* payload,
                             uint64_t chunk, uint32_t part, uint32_t parts, uint32_t* tri) {
    const uint64_t cells = N ? (uint64_t)N * (N - 1) / 2 : 0;
    memset(tri, 0, cells * sizeof(uint32_t));
    int64_t* W = (int64_t*)malloc(P * sizeof(int64_t));
    uint32_t* full = (uint32_t*)malloc(((size_t)N + 1) * sizeof(uint32_t));
    if (!W || !full) { free(W); free(full); return UINT64_MAX; }
    memcpy(W, num_kmers, P * sizeof(int64_t));
    for (uint64_t i = P; i-- > 1;)
        if (parent_id[i] >= 0) W[parent_id[i]] += W[i];
    uint64_t U = 0;
    for (uint64_t p = 0; p < P; ++p) {
        if (l[p] == 0 || (p / chunk) % parts != part) continue;
        uint32_t* out = full + n[p];
        for (int64_t q = (int64_t)p; q >= 0; q = parent_id[q]) {
            out -= l[q];
            oracle_decode_local(payload + payload_off[q], l[q], last[q], out);
        }
        for (uint32_t i = n[p] - l[p]; i < n[p]; ++i) {
            const uint64_t s = full[i];
            uint32_t* row = tri + s * (s - 1) / 2;
            for (uint32_t k = 0; k < i; ++k) row[full[k]] += (uint32_t)W[p];
            U += i;
        }
    }
    free(W); free(full);
    return U;
}

/* Independent second opinion used by the tests on small tries: M[s][t] = sum of num_kmers over
 * nodes whose full list contains both s and t (a k-mer of node p lies in exactly the samples
 * of p's full list).  O(sum n^2), no W accumulation, no job rule. */
/* The same matrix by the column-side regrouping DESIGN.md §8 proposes as the next formulation (ours; nothing in
 * the reference does this — it is here so that the claim "fewer updates, identical bits" is checked, and as the
 * restatement a future kernel would be pinned on).  With W_p the subtree sums (src/similarity_calculator.cpp:64-72)
 * and S_q(r) = sum of W_p over the STRICT descendants p of q that hold sample r locally:
 *     M[r][c] += S_q(r)   for every c in local(q), r in supp(S_q)      (r > c: descendants' ids come later)
 *     M[r][c] += W_q      for r, c in local(q), c < r                  (the triangle inside the node)
 * S is carried up the trie: S_parent += S_q + W_q * 1[local(q)].  All sums are modulo 2^32, like the reference's
 * uint32 adds.  Returns the number of matrix updates plus merge operations performed (UINT64_MAX on allocation failure). */
uint64_t oracle_all2all_regrouped(uint64_t P, uint32_t N, const int64_t* num_kmers, const int64_t* parent_id, const uint32_t* n,
                                  const uint32_t* l, const uint32_t* last, const uint64_t* payload_off, const uint64_t* payload,
                                  uint32_t* tri) {
    (void)n;
    const uint64_t cells = N ? (uint64_t)N * (N - 1) / 2 : 0;
    memset(tri, 0, cells * sizeof(uint32_t));
    int64_t* W = (int64_t*)malloc(P * sizeof(int64_t));
    uint32_t** rows = (uint32_t**)calloc(P, sizeof(uint32_t*));   /* per node: (row, weight) pairs pushed up by its children */
    uint64_t* cnt = (uint64_t*)calloc(P, sizeof(uint64_t));
    uint64_t* cap = (uint64_t*)calloc(P, sizeof(uint64_t));
    uint32_t* acc = (uint32_t*)calloc((size_t)N + 1, sizeof(uint32_t));   /* dense scratch over rows */
    uint8_t* seen = (uint8_t*)calloc((size_t)N + 1, 1);
    uint32_t* touched = (uint32_t*)malloc(((size_t)N + 1) * sizeof(uint32_t));
    uint32_t* loc = (uint32_t*)malloc(((size_t)N + 1) * sizeof(uint32_t));
    uint64_t ops = 0;
    int fail = !W || !rows || !cnt || !cap || !acc || !seen || !touched || !loc;
    if (!fail) {
        memcpy(W, num_kmers, P * sizeof(int64_t));
        for (uint64_t i = P; i-- > 1;)
            if (parent_id[i] >= 0) W[parent_id[i]] += W[i];
    }
    for (uint64_t q = P; !fail && q-- > 0;) {   /* children have larger ids: S_q is complete when q is reached */
        uint32_t nt = 0;
        for (uint64_t e = 0; e < cnt[q]; ++e) {   /* compact the pushed pairs into S_q */
            const uint32_t r = rows[q][2 * e];
            if (!seen[r]) { seen[r] = 1; touched[nt++] = r; }
            acc[r] += rows[q][2 * e + 1];
        }
        ops += cnt[q];
        free(rows[q]); rows[q] = NULL;
        if (l[q]) oracle_decode_local(payload + payload_off[q], l[q], last[q], loc);
        const uint32_t wq = (uint32_t)W[q];
        for (uint32_t t = 0; t < nt; ++t) {       /* rectangle: rows of the strict descendants x local(q) */
            const uint64_t r = touched[t];
            uint32_t* row = tri + r * (r - 1) / 2;
            for (uint32_t j = 0; j < l[q]; ++j) row[loc[j]] += acc[r];
            ops += l[q];
        }
        for (uint32_t i = 1; i < l[q]; ++i) {      /* triangle inside the node */
            const uint64_t r = loc[i];
            uint32_t* row = tri + r * (r - 1) / 2;
            for (uint32_t j = 0; j < i; ++j) row[loc[j]] += wq;
            ops += i;
        }
        const int64_t par = parent_id[q];
        if (par >= 0) {                            /* T_q = S_q + W_q * 1[local(q)] goes up */
            const uint64_t need = cnt[par] + nt + l[q];
            if (need > cap[par]) {
                const uint64_t ncap = need * 2 + 8;
                uint32_t* np = (uint32_t*)realloc(rows[par], ncap * 2 * sizeof(uint32_t));
                if (!np) { fail = 1; break; }
                rows[par] = np; cap[par] = ncap;
            }
            uint64_t at = cnt[par];
            for (uint32_t t = 0; t < nt; ++t) { rows[par][2 * at] = touched[t]; rows[par][2 * at + 1] = acc[touched[t]]; ++at; }
            for (uint32_t j = 0; j < l[q]; ++j) { rows[par][2 * at] = loc[j]; rows[par][2 * at + 1] = wq; ++at; }
            cnt[par] = at;
        }
        for (uint32_t t = 0; t < nt; ++t) { acc[touched[t]] = 0; seen[touched[t]] = 0; }
    }
    if (rows) for (uint64_t q = 0; q < P; ++q) free(rows[q]);
    free(W); free(rows); free(cnt); free(cap); free(acc); free(seen); free(touched); free(loc);
    return fail ? UINT64_MAX : ops;
}

/* The same matrix by the run-boundary ("difference") form the CUDA library uses when sample lists hold runs of
 * consecutive ids (kmer-db_b200/csrc/diff.cuh; ours — the reference's nearest relative is the 16-consecutive-id fast
 * path of row_add, src/simd/row_add_avx2.cpp:38-75).  A pattern's full list is kept as its sorted run boundaries
 * B = [s1, e1+1, s2, e2+1, ...]; a local row r receives +W_p at the even and -W_p at the odd entries of B that lie
 * below r, into a difference matrix; prefix sums along every row (columns below the diagonal only) restore the
 * counts.  All sums are modulo 2^32.  Returns the number of difference updates performed, reports U through
 * *u_out (UINT64_MAX on allocation failure). */
uint64_t oracle_all2all_boundary(uint64_t P, uint32_t N, const int64_t* num_kmers, const int64_t* parent_id, const uint32_t* n,
                                 const uint32_t* l, const uint32_t* last, const uint64_t* payload_off, const uint64_t* payload,
                                 uint32_t* tri, uint64_t* u_out) {
    const uint64_t cells = N ? (uint64_t)N * (N - 1) / 2 : 0;
    memset(tri, 0, cells * sizeof(uint32_t));
    int64_t* W = (int64_t*)malloc(P * sizeof(int64_t));
    uint32_t* full = (uint32_t*)malloc(((size_t)N + 1) * sizeof(uint32_t));
    uint32_t* B = (uint32_t*)malloc((2 * (size_t)N + 2) * sizeof(uint32_t));
    if (!W || !full || !B) { free(W); free(full); free(B); return UINT64_MAX; }
    memcpy(W, num_kmers, P * sizeof(int64_t));
    for (uint64_t i = P; i-- > 1;)
        if (parent_id[i] >= 0) W[parent_id[i]] += W[i];
    uint64_t U = 0, phys = 0;
    for (uint64_t p = 0; p < P; ++p) {
        if (l[p] == 0) continue;
        uint32_t* out = full + n[p];
        for (int64_t q = (int64_t)p; q >= 0; q = parent_id[q]) {
            out -= l[q];
            oracle_decode_local(payload + payload_off[q], l[q], last[q], out);
        }
        uint32_t nb = 0;   /* run boundaries of the full list */
        for (uint32_t i = 0; i < n[p]; ++i) {
            if (i == 0 || full[i] != full[i - 1] + 1) B[nb++] = full[i];
            if (i + 1 == n[p] || full[i + 1] != full[i] + 1) B[nb++] = full[i] + 1;
        }
        const uint32_t w = (uint32_t)W[p];
        for (uint32_t i = n[p] - l[p]; i < n[p]; ++i) {
            const uint64_t r = full[i];
            uint32_t* row = tri + r * (r - 1) / 2;
            for (uint32_t e = 0; e < nb && B[e] < r; ++e) { row[B[e]] += (e & 1) ? 0u - w : w; ++phys; }
            U += i;
        }
    }
    for (uint64_t r = 1; r < N; ++r) {   /* differences -> counts */
        uint32_t* row = tri + r * (r - 1) / 2;
        for (uint64_t c = 1; c < r; ++c) row[c] += row[c - 1];
    }
    free(W); free(full); free(B);
    if (u_out) *u_out = U;
    return phys;
}

int oracle_all2all_bruteforce(uint64_t P, uint32_t N, const int64_t* num_kmers, const int64_t* parent_id, const uint32_t* n,
                              const uint32_t* l, const uint32_t* last, const uint64_t* payload_off, const uint64_t* payload,
                              uint32_t* tri) {
    const uint64_t cells = N ? (uint64_t)N * (N - 1) / 2 : 0;
    memset(tri, 0, cells * sizeof(uint32_t));
    uint32_t* full = (uint32_t*)malloc(((size_t)N + 1) * sizeof(uint32_t));
    if (!full) return -1;
    for (uint64_t p = 0; p < P; ++p) {
        if (n[p] == 0 || num_kmers[p] == 0) continue;
        uint32_t* out = full + n[p];
        for (int64_t q = (int64_t)p; q >= 0; q = parent_id[q]) {
            out -= l[q];
            oracle_decode_local(payload + payload_off[q], l[q], last[q], out);
        }
        for (uint32_t a = 1; a < n[p]; ++a)
            for (uint32_t b = 0; b < a; ++b) {
                const uint64_t s = full[a];
                tri[s * (s - 1) / 2 + full[b]] += (uint32_t)num_kmers[p];
            }
    }
    free(full);
    return 0;
}

/* ---- .db reader (hashtables skipped) ---------------------------------------------------- */
static int rd(FILE* f, void* dst, size_t bytes) { return bytes == 0 || fread(dst, 1, bytes, f) == bytes; }

void oracle_db_free(oracle_db* db) {
    if (!db) return;
    free(db->num_kmers); free(db->parent_id); free(db->n); free(db->l); free(db->last); free(db->bits);
    free(db->payload_off); free(db->payload); free(db->sample_kmers); free(db->table_off); free(db->slots);
    if (db->names) for (uint32_t i = 0; i < db->num_samples; ++i) free(db->names[i]);
    free(db->names);
    free(db);
}

static oracle_db* db_read(const char* path, int with_tables);
oracle_db* oracle_db_read(const char* path) { return db_read(path, 0); }
oracle_db* oracle_db_read_full(const char* path) { return db_read(path, 1); }

static oracle_db* db_read(const char* path, int with_tables) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    oracle_db* db = (oracle_db*)calloc(1, sizeof(oracle_db));
    uint64_t format_word, kmers_count, nsamples, ntables, P;
    uint8_t inited;
    int ok = rd(f, &format_word, 8) && rd(f, &db->kmer_length, 4) && rd(f, &db->fraction, 8) && rd(f, &db->start_fraction, 8) &&
             rd(f, &db->alphabet, 4) && rd(f, &inited, 1) && rd(f, &kmers_count, 8) && rd(f, &nsamples, 8);
    if (!ok) goto fail;
    db->num_samples = (uint32_t)nsamples;
    db->names = (char**)calloc(nsamples + 1, sizeof(char*));
    db->sample_kmers = (uint64_t*)calloc(nsamples + 1, 8);
    for (uint64_t i = 0; i < nsamples; ++i) {
        uint64_t len;
        if (!rd(f, &db->sample_kmers[i], 8) || !rd(f, &len, 8)) goto fail;
        db->names[i] = (char*)calloc(len + 1, 1);
        if (!rd(f, db->names[i], len)) goto fail;
    }
    if (!rd(f, &ntables, 8)) goto fail;
    uint64_t slot_cap = 0;
    if (with_tables) {
        if (!(format_word & 1)) goto fail;
        db->num_tables = ntables;
        db->table_off = (uint64_t*)calloc(ntables + 1, 8);
    }
    for (uint64_t i = 0; i < ntables; ++i) {
        if (format_word & 1) { /* raw: f64 + 7 u64 header, bit-vector, filled slots */
            uint64_t hdr[8];
            if (!rd(f, hdr, sizeof hdr)) goto fail;
            const uint64_t filled = hdr[1], allocated = hdr[2];
            if (!with_tables) {
                if (fseeko(f, (off_t)(((allocated + 63) / 64) * 8 + filled * 8), SEEK_CUR)) goto fail;
                continue;
            }
            /* rebuild the slot array: bit i of the bit-vector says slot i is used; the used slots
             * follow in slot order (src/hashmap_lp.h:481-528) */
            const uint64_t base = db->table_off[i], bvw = (allocated + 63) / 64;
            db->table_off[i + 1] = base + allocated;
            if (base + allocated > slot_cap) {
                slot_cap = (base + allocated) * 2;
                db->slots = (uint64_t*)realloc(db->slots, slot_cap * 8);
            }
            uint64_t* bv = (uint64_t*)malloc((bvw + 1) * 8);
            if (!rd(f, bv, bvw * 8)) { free(bv); goto fail; }
            for (uint64_t sidx = 0; sidx < allocated; ++sidx) {
                if ((bv[sidx >> 6] >> (sidx & 63)) & 1) { if (!rd(f, &db->slots[base + sidx], 8)) { free(bv); goto fail; } }
                else db->slots[base + sidx] = (uint64_t)0x7FFFFFFFu << 32;
            }
            free(bv);
        } else {
            uint64_t total, seen = 0, portion;
            if (!rd(f, &total, 8)) goto fail;
            while (seen < total) {
                if (!rd(f, &portion, 8) || fseeko(f, (off_t)(portion * 8), SEEK_CUR)) goto fail;
                seen += portion;
            }
        }
    }
    if (!rd(f, &P, 8)) goto fail;
    db->num_patterns = P;
    db->num_kmers = (int64_t*)malloc((P + 1) * 8); db->parent_id = (int64_t*)malloc((P + 1) * 8);
    db->n = (uint32_t*)malloc((P + 1) * 4); db->l = (uint32_t*)malloc((P + 1) * 4);
    db->last = (uint32_t*)malloc((P + 1) * 4); db->bits = (uint32_t*)malloc((P + 1) * 4);
    db->payload_off = (uint64_t*)malloc((P + 1) * 8);
    {
        uint64_t cap = 1024, used = 0, pid = 0;
        db->payload = (uint64_t*)malloc(cap * 8);
        while (pid < P) {
            uint64_t block;
            if (!rd(f, &block, 8)) goto fail;
            unsigned char* buf = (unsigned char*)malloc(block + 1);
            if (!rd(f, buf, block)) { free(buf); goto fail; }
            uint64_t at = 0;
            while (at < block && pid < P) { /* 40-byte header + ceil(bits/128)*16 payload bytes */
                memcpy(&db->num_kmers[pid], buf + at, 8); memcpy(&db->parent_id[pid], buf + at + 8, 8);
                memcpy(&db->n[pid], buf + at + 16, 4); memcpy(&db->l[pid], buf + at + 20, 4);
                memcpy(&db->last[pid], buf + at + 24, 4); memcpy(&db->bits[pid], buf + at + 28, 4);
                at += 40;
                const uint64_t words = db->bits[pid] ? ((uint64_t)(db->bits[pid] + 127) / 128) * 2 : 0;
                if (used + words + 2 > cap) { while (used + words + 2 > cap) cap *= 2; db->payload = (uint64_t*)realloc(db->payload, cap * 8); }
                db->payload_off[pid] = used;
                memcpy(db->payload + used, buf + at, words * 8);
                used += words; at += words * 8;
                ++pid;
            }
            free(buf);
        }
        db->payload[used] = 0; db->payload[used + 1] = 0;
        db->payload_words = used;
    }
    fclose(f);
    return db;
fail:
    fclose(f);
    oracle_db_free(db);
    return NULL;
}

/* ---- CSV (src/console_all2all.cpp:40-78) ------------------------------------------------- */
int oracle_write_csv(const oracle_db* db, const uint32_t* tri, const char* path, int sparse) {
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    fprintf(f, "kmer-length: %u fraction: %g ,db-samples ,", db->kmer_length, db->fraction);
    for (uint32_t i = 0; i < db->num_samples; ++i) fprintf(f, "%s,", db->names[i]);
    fprintf(f, "\nquery-samples,total-kmers,");
    for (uint32_t i = 0; i < db->num_samples; ++i) fprintf(f, "%llu,", (unsigned long long)db->sample_kmers[i]);
    fprintf(f, "\n");
    for (uint64_t s = 0; s < db->num_samples; ++s) {
        fprintf(f, "%s,%llu,", db->names[s], (unsigned long long)db->sample_kmers[s]);
        const uint32_t* row = tri + s * (s - 1) / 2;
        for (uint64_t c = 0; c < s; ++c) {
            if (!sparse) fprintf(f, "%u,", row[c]);
            else if (row[c]) fprintf(f, "%llu:%u,", (unsigned long long)(c + 1), row[c]);
        }
        fprintf(f, "\n");
    }
    return fclose(f);
}

/* whole path on a file: returns U or UINT64_MAX */
uint64_t oracle_all2all_file(const char* db_path, const char* csv_path, int sparse) {
    oracle_db* db = oracle_db_read(db_path);
    if (!db) return UINT64_MAX;
    const uint64_t N = db->num_samples;
    uint32_t* tri = (uint32_t*)calloc(N * (N ? N - 1 : 0) / 2 + 1, 4);
    uint64_t U = oracle_all2all(db->num_patterns, db->num_samples, db->num_kmers, db->parent_id, db->n, db->l, db->last, db->bits,
                                db->payload_off, db->payload, tri);
    if (U != UINT64_MAX && csv_path && oracle_write_csv(db, tri, csv_path, sparse)) U = UINT64_MAX;
    free(tri);
    oracle_db_free(db);
    return U;
}


/* ---- query vs database: one2all<false> (src/similarity_calculator.cpp:810-925) ----------- */
static uint32_t fmix32(uint32_t h) { h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h; }

/* pattern id stored for `kmer`, or -1 (hash_map_lp::find, src/hashmap_lp.h:308-333) */
int64_t oracle_lookup(const oracle_db* db, uint64_t kmer) {
    const uint64_t prefix = kmer >> 32;
    const uint32_t suffix = (uint32_t)kmer;
    if (prefix >= db->num_tables) return -1;
    const uint64_t* t = db->slots + db->table_off[prefix];
    const uint64_t mask = db->table_off[prefix + 1] - db->table_off[prefix] - 1;
    for (uint64_t h = fmix32(suffix) & mask;; h = (h + 1) & mask) {
        const uint32_t val = (uint32_t)(t[h] >> 32);
        if (val == 0x7FFFFFFFu) return -1;
        if ((uint32_t)t[h] == suffix) return (int64_t)val;
    }
}

/* out[s] (N cells, zeroed here) = number of the query's k-mers present in sample s.  kmers must be
 * unique (KmerHelper::unique, src/kmer_extract.h:113-119).  Returns the number of hits. */
uint64_t oracle_one2all(const oracle_db* db, const uint64_t* kmers, uint64_t count, uint32_t* out) {
    const uint32_t N = db->num_samples;
    memset(out, 0, (size_t)N * 4);
    int32_t* hit = (int32_t*)calloc(db->num_patterns + 1, 4); /* per-pattern hit counts (the unordered_map) */
    uint32_t* full = (uint32_t*)malloc(((size_t)N + 1) * 4);
    uint64_t hits = 0;
    for (uint64_t i = 0; i < count; ++i) {
        const int64_t pid = oracle_lookup(db, kmers[i]);
        if (pid < 0 || db->num_kmers[pid] == 0) continue; /* :846-847 */
        ++hit[pid];
        ++hits;
    }
    for (uint64_t p = 0; p < db->num_patterns; ++p) {
        if (!hit[p]) continue;
        uint32_t* o = full + db->n[p];
        for (int64_t q = (int64_t)p; q >= 0; q = db->parent_id[q]) { /* :890-900 */
            o -= db->l[q];
            oracle_decode_local(db->payload + db->payload_off[q], db->l[q], db->last[q], o);
        }
        for (uint32_t k = 0; k < db->n[p]; ++k) out[full[k]] += (uint32_t)hit[p];
    }
    free(hit); free(full);
    return hits;
}

/* ---- database against database: db2db_sp (src/similarity_calculator.cpp:1225-1540) --------------------------
 * out (N1 x N2 cells, zeroed here): out[s1 * N2 + s2] = number of k-mers present in sample s1 of db1 and in sample
 * s2 of db2.  The reference merges the sorted (suffix, pattern) pairs of each prefix bucket of the two databases
 * (:1262-1287; keys are unique inside a bucket, so the merge finds exactly the suffixes both buckets hold — restated
 * here as a lookup of every used slot of db1 in db2's bucket), sorts and counts the (pattern1, pattern2) pairs
 * (:1310-1321), and adds every distinct pair's count to the cells of list(pattern1) x list(pattern2) (:1434-1517;
 * bubbles only defer the same additions to compact2, src/array.h:409-415).  Returns the number of cell updates, or
 * UINT64_MAX when the databases do not fit together. */
static int cmp_pair(const void* a, const void* b) { const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return x < y ? -1 : x > y; }

static void full_list(const oracle_db* db, uint64_t p, uint32_t* buf /* n[p] ids */) {
    uint32_t* o = buf + db->n[p];
    for (int64_t q = (int64_t)p; q >= 0; q = db->parent_id[q]) {
        o -= db->l[q];
        oracle_decode_local(db->payload + db->payload_off[q], db->l[q], db->last[q], o);
    }
}

uint64_t oracle_db2db(const oracle_db* db1, const oracle_db* db2, uint32_t* out) {
    const uint32_t N1 = db1->num_samples, N2 = db2->num_samples;
    if (db1->num_tables != db2->num_tables || db1->kmer_length != db2->kmer_length) return UINT64_MAX;
    memset(out, 0, (size_t)N1 * N2 * 4);
    uint64_t cap = 1024, np = 0;
    uint64_t* pairs = (uint64_t*)malloc(cap * 8);
    for (uint64_t t = 0; t < db1->num_tables; ++t) {
        for (uint64_t i = db1->table_off[t]; i < db1->table_off[t + 1]; ++i) {
            const uint32_t pid1 = (uint32_t)(db1->slots[i] >> 32);
            if (pid1 == 0x7FFFFFFFu) continue;
            const int64_t pid2 = oracle_lookup(db2, (t << 32) | (uint32_t)db1->slots[i]);
            if (pid2 < 0) continue;
            if (np == cap) { cap *= 2; pairs = (uint64_t*)realloc(pairs, cap * 8); }
            pairs[np++] = ((uint64_t)pid1 << 32) | (uint32_t)pid2;
        }
    }
    qsort(pairs, np, 8, cmp_pair);
    uint32_t* l1 = (uint32_t*)malloc(((size_t)N1 + 1) * 4);
    uint32_t* l2 = (uint32_t*)malloc(((size_t)N2 + 1) * 4);
    uint64_t updates = 0;
    for (uint64_t i = 0; i < np;) {
        uint64_t j = i;
        while (j < np && pairs[j] == pairs[i]) ++j;
        const uint64_t p1 = pairs[i] >> 32, p2 = (uint32_t)pairs[i];
        const uint32_t cnt = (uint32_t)(j - i);
        full_list(db1, p1, l1);
        full_list(db2, p2, l2);
        for (uint32_t a = 0; a < db1->n[p1]; ++a)
            for (uint32_t b = 0; b < db2->n[p2]; ++b) out[(size_t)l1[a] * N2 + l2[b]] += cnt;
        updates += (uint64_t)db1->n[p1] * db2->n[p2];
        i = j;
    }
    free(pairs); free(l1); free(l2);
    return updates;
}

/* all2all-parts on a list of database files (src/console_all2all_parts.cpp:11-371), no filters: header lines over the
 * samples of all parts, then for grid row i the lines of part i's samples — the cells (i, 0..i-1) by oracle_db2db and the
 * diagonal cell by the all2all restatement (all2all_sp yields the same matrix), non-zero cells as <global col + 1>:<val>,
 * (SparseMatrix::saveRowSparse(row, out, idx_shift), src/array.h:625-637).  Returns the number of pairs, or -1. */
int64_t oracle_all2all_parts_file(const char* list_path, const char* csv_path) {
    FILE* lf = fopen(list_path, "r");
    if (!lf) return -1;
    char name[4096];
    char** files = NULL;
    uint32_t parts = 0;
    while (fscanf(lf, "%4095s", name) == 1) {
        files = (char**)realloc(files, (parts + 1) * sizeof(char*));
        files[parts++] = strdup(name);
    }
    fclose(lf);
    oracle_db** dbs = (oracle_db**)calloc(parts + 1, sizeof(oracle_db*));
    int64_t saved = 0;
    FILE* out = NULL;
    for (uint32_t i = 0; i < parts; ++i) {
        dbs[i] = oracle_db_read_full(files[i]);
        if (!dbs[i] || dbs[i]->kmer_length != dbs[0]->kmer_length || dbs[i]->fraction != dbs[0]->fraction) { saved = -1; goto done; }
    }
    out = fopen(csv_path, "wb");
    if (!out) { saved = -1; goto done; }
    if (parts) {
        fprintf(out, "kmer-length: %u fraction: %g ,db-samples ,", dbs[0]->kmer_length, dbs[0]->fraction);
        for (uint32_t i = 0; i < parts; ++i) for (uint32_t s = 0; s < dbs[i]->num_samples; ++s) fprintf(out, "%s,", dbs[i]->names[s]);
        fprintf(out, "\nquery-samples,total-kmers,");
        for (uint32_t i = 0; i < parts; ++i) for (uint32_t s = 0; s < dbs[i]->num_samples; ++s) fprintf(out, "%llu,", (unsigned long long)dbs[i]->sample_kmers[s]);
        fprintf(out, "\n");
    }
    for (uint32_t i = 0; i < parts; ++i) {
        const oracle_db* r = dbs[i];
        const uint64_t N = r->num_samples;
        uint32_t** rect = (uint32_t**)calloc(i + 1, sizeof(uint32_t*));
        for (uint32_t j = 0; j < i; ++j) {
            rect[j] = (uint32_t*)malloc((size_t)N * dbs[j]->num_samples * 4 + 4);
            if (oracle_db2db(r, dbs[j], rect[j]) == UINT64_MAX) saved = -1;
        }
        uint32_t* tri = (uint32_t*)calloc(N * (N ? N - 1 : 0) / 2 + 1, 4);
        if (oracle_all2all(r->num_patterns, r->num_samples, r->num_kmers, r->parent_id, r->n, r->l, r->last, r->bits, r->payload_off, r->payload,
                           tri) == UINT64_MAX) saved = -1;
        for (uint64_t s = 0; s < N && saved >= 0; ++s) {
            fprintf(out, "%s,%llu,", r->names[s], (unsigned long long)r->sample_kmers[s]);
            uint64_t shift = 0;
            for (uint32_t j = 0; j < i; ++j) {
                const uint32_t N2 = dbs[j]->num_samples;
                for (uint32_t c = 0; c < N2; ++c)
                    if (rect[j][s * N2 + c]) { fprintf(out, "%llu:%u,", (unsigned long long)(shift + c + 1), rect[j][s * N2 + c]); ++saved; }
                shift += N2;
            }
            const uint32_t* row = tri + s * (s - 1) / 2;
            for (uint64_t c = 0; c < s; ++c)
                if (row[c]) { fprintf(out, "%llu:%u,", (unsigned long long)(shift + c + 1), row[c]); ++saved; }
            fprintf(out, "\n");
        }
        for (uint32_t j = 0; j < i; ++j) free(rect[j]);
        free(rect); free(tri);
        if (saved < 0) break;
    }
done:
    if (out) fclose(out);
    for (uint32_t i = 0; i < parts; ++i) { oracle_db_free(dbs[i]); free(files[i]); }
    free(dbs); free(files);
    return saved;
}

/* ---- k-mer extraction, written window by window (no rolling state) so that it is independent of
 * the product's scanner: src/kmer_extract.h:13-97, alphabets src/alphabet.h:79-86 ----------------- */
static const char* alphabet_groups(int32_t id, int* preserve) {
    static const char* g[] = {"A,C,G,TU", "A,C,G,TU", "K,R,E,D,Q,N,C,G,H,I,L,V,M,F,Y,W,P,S,T,A", "KREDQN,C,G,H,ILV,M,F,Y,W,P,STA",
                              "AST,C,DN,EQ,FY,G,H,IV,KR,LM,P,W", "STPAG,NDEQ,HRK,MILV,FYW,C"};
    *preserve = id != 0;
    return (id >= 0 && id <= 5) ? g[id] : NULL;
}

static uint64_t minhash_of(uint64_t x, uint32_t k) { /* src/filter.h:96-115 */
    const uint64_t kd4 = (k + 3) / 4;
    uint64_t h = x * 0x87c37b91114253d5ull;
    h = (h << 31) | (h >> 33);
    h *= 0x4cf5ad432745937full;
    uint64_t h1 = (42 ^ h) ^ kd4, h2 = 42 ^ kd4;
    h1 += h2; h2 += h1;
    h1 ^= h1 >> 33; h1 *= 0xff51afd7ed558ccdull; h1 ^= h1 >> 33; h1 *= 0xc4ceb9fe1a85ec53ull; h1 ^= h1 >> 33;
    h2 ^= h2 >> 33; h2 *= 0xff51afd7ed558ccdull; h2 ^= h2 >> 33; h2 *= 0xc4ceb9fe1a85ec53ull; h2 ^= h2 >> 33;
    h1 += h2; h2 += h1;
    return h1 ^ h2;
}

static int cmp_u64(const void* a, const void* b) { const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return x < y ? -1 : x > y; }

/* k-mers of one sequence appended to out (capacity: len).  Returns the new count. */
static uint64_t extract_seq(const char* seq, uint64_t len, uint32_t k, int32_t alphabet, double fraction, double start, uint64_t* out,
                            uint64_t cnt) {
    int preserve, map[256], nsym = 1, bits = 0;
    const char* g = alphabet_groups(alphabet, &preserve);
    for (int i = 0; i < 256; ++i) map[i] = -1;
    for (const char* c = g; *c; ++c) {
        if (*c == ',') { ++nsym; continue; }
        map[(unsigned char)*c] = nsym - 1;
        map[(unsigned char)(*c - 'A' + 'a')] = nsym - 1;
    }
    while ((1 << bits) < nsym) ++bits;
    const int prefix_bits = (int)k * bits - 32;
    const uint32_t shift = prefix_bits < 8 ? (uint32_t)(8 - prefix_bits) : 0;
    const uint64_t lo = (uint64_t)((double)UINT64_MAX * start), hi = (uint64_t)((double)UINT64_MAX * (start + fraction));
    for (uint64_t i = 0; i + k <= len; ++i) {
        uint64_t fwd = 0, rev = 0;
        int ok = 1;
        for (uint32_t j = 0; j < k; ++j) {
            const int sy = map[(unsigned char)seq[i + j]];
            if (sy < 0) { ok = 0; break; }
            fwd = (fwd << bits) | (uint64_t)sy;
            rev |= (uint64_t)(nsym - 1 - sy) << (bits * j); /* complement, reversed */
        }
        if (!ok) continue;
        uint64_t can = (preserve || fwd < rev) ? fwd : rev;
        can = (can << shift) | (can & ((1ull << shift) - 1));
        if (fraction < 1.0) { const uint64_t h = minhash_of(can, k); if (!(h >= lo && h < hi)) continue; }
        out[cnt++] = can;
    }
    return cnt;
}

/* new2all on files: database (with tables) x FASTA list -> CSV (src/console_new2all.cpp:12-174).
 * list_path: whitespace-separated entries, opened as given or with .fa/.fna/.fasta appended (plain
 * text only here).  multisample: every record is a query named by its header up to the first
 * space; otherwise one query per file named by the entry's last path component.  Returns the
 * number of queries or -1. */
int64_t oracle_new2all_file(const char* db_path, const char* list_path, int multisample, const char* csv_path, int sparse) {
    oracle_db* db = oracle_db_read_full(db_path);
    if (!db) return -1;
    FILE* lf = fopen(list_path, "rb");
    FILE* out = fopen(csv_path, "wb");
    if (!lf || !out) { if (lf) fclose(lf); if (out) fclose(out); oracle_db_free(db); return -1; }
    fprintf(out, "kmer-length: %u fraction: %g ,db-samples ,", db->kmer_length, db->fraction);
    for (uint32_t i = 0; i < db->num_samples; ++i) fprintf(out, "%s,", db->names[i]);
    fprintf(out, "\nquery-samples,total-kmers,");
    for (uint32_t i = 0; i < db->num_samples; ++i) fprintf(out, "%llu,", (unsigned long long)db->sample_kmers[i]);
    fprintf(out, "\n");
    uint32_t* sims = (uint32_t*)malloc(((size_t)db->num_samples + 1) * 4);
    char entry[4096];
    int64_t nq = 0;
    while (fscanf(lf, "%4000s", entry) == 1) {
        static const char* exts[] = {"", ".fa", ".fna", ".fasta"};
        FILE* f = NULL;
        char path[4200];
        for (int e = 0; e < 4 && !f; ++e) { snprintf(path, sizeof path, "%s%s", entry, exts[e]); f = fopen(path, "rb"); }
        if (!f) continue;
        fseeko(f, 0, SEEK_END);
        const uint64_t size = (uint64_t)ftello(f);
        fseeko(f, 0, SEEK_SET);
        char* data = (char*)malloc(size + 2);
        if (fread(data, 1, size, f) != size) { fclose(f); free(data); continue; }
        fclose(f);
        data[size] = 0;
        uint64_t* kmers = (uint64_t*)malloc((size + 1) * 8);
        uint64_t cnt = 0;
        const char* base = strrchr(entry, '/');
        const char* qname = base ? base + 1 : entry;
        char* p = strchr(data, '>');
        char header[1024] = "";
        while (p) {
            char* eol = strchr(p, '\n');
            if (!eol) eol = data + size;
            size_t hl = (size_t)(eol - (p + 1));
            if (hl && p[hl] == '\r') --hl;
            if (hl > sizeof header - 1) hl = sizeof header - 1;
            memcpy(header, p + 1, hl); header[hl] = 0;
            char* sp = strchr(header, ' ');
            if (sp) *sp = 0;
            char* seq = (*eol) ? eol + 1 : eol;
            char* next = strchr(seq, '>');
            char* end = next ? next : data + size;
            uint64_t w = 0;
            for (char* c = seq; c < end; ++c) if (*c != '\n' && *c != '\r') seq[w++] = *c;
            cnt = extract_seq(seq, w, db->kmer_length, db->alphabet, db->fraction, db->start_fraction, kmers, cnt);
            if (multisample || !next) {
                qsort(kmers, cnt, 8, cmp_u64);
                uint64_t u = 0;
                for (uint64_t i = 0; i < cnt; ++i) if (i == 0 || kmers[i] != kmers[i - 1]) kmers[u++] = kmers[i];
                oracle_one2all(db, kmers, u, sims);
                fprintf(out, "%s,%llu,", multisample ? header : qname, (unsigned long long)u);
                for (uint32_t c = 0; c < db->num_samples; ++c) {
                    if (!sparse) fprintf(out, "%u,", sims[c]);
                    else if (sims[c]) fprintf(out, "%u:%u,", c + 1, sims[c]);
                }
                fprintf(out, "\n");
                ++nq;
                cnt = 0;
            }
            p = next;
        }
        free(kmers); free(data);
    }
    free(sims);
    fclose(lf); fclose(out);
    oracle_db_free(db);
    return nq;
}

#ifdef ORACLE_MAIN
int main(int argc, char** argv) {
    int sparse = 0, a = 1;
    if (argc > 1 && !strcmp(argv[1], "-sparse")) { sparse = 1; a = 2; }
    if (argc - a != 2) { fprintf(stderr, "usage: kdb_oracle [-sparse] <db> <out.csv>\n"); return 2; }
    const uint64_t U = oracle_all2all_file(argv[a], argv[a + 1], sparse);
    if (U == UINT64_MAX) { fprintf(stderr, "oracle failed\n"); return 1; }
    fprintf(stderr, "U=%llu\n", (unsigned long long)U);
    return 0;
}
#endif
