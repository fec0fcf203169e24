// Elias-gamma parsing of a pattern's local sample list in two passes over ONE walk of the bits (included by kdbx.cu;
// plain host/device C++, so the same code is compiled by g++ for the CPU test tests/test_host.py::test_gamma_tokens_*).
//
// The stream holds the l - 1 deltas between consecutive local ids in append order, and only the LAST id is stored
// (src/pattern.cpp:99-109, src/elias_gamma.h:104-128,133-256): the first id is known only when every delta has been
// read.  A delta of 1 (consecutive sample ids) is the single bit 0, and the lists of related genomes are mostly such
// bits, so the parser takes a whole run of zero bits at once and parks ONE token for it — kRunToken | run length in
// the slot of the run's first id — instead of one delta per id; a delta >= 2 is parked as itself.  The second pass
// turns tokens into ids front to back: a run is a store-only loop, and the row blocks it crosses (the job counts of
// the decoder, kdbx.cu: DecodeHist) follow from arithmetic on its first id instead of a test per id.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define KDBX_HD __host__ __device__ __forceinline__
#else
#define KDBX_HD inline
#endif

KDBX_HD uint32_t kdbx_clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__clzll((long long)x);
#else
    return x ? (uint32_t)__builtin_clzll(x) : 64u;
#endif
}
KDBX_HD uint32_t kdbx_min_u32(uint32_t a, uint32_t b) { return a < b ? a : b; }

constexpr uint32_t kRunToken = 0x80000000u;

// Pass 1.  w: the pattern's payload (MSB first in 64-bit words; the word after the last one must be readable), nb: its
// number of bits, l: ids in the list (>= 2).  Writes the tokens of ids 1 .. l-1 to out[1 .. l) (slots inside a run,
// after its first, are left untouched); sum = the sum of all deltas, runs = 1 + the number of deltas >= 2 (the maximal
// runs of consecutive ids).  Returns 0, or 1 when the stream does not hold exactly l - 1 well-formed codes in nb bits,
// or 2 when a delta does not fit 31 bits (no sample id can be that large).
template <class Out>
KDBX_HD int gamma_parse_tokens(const uint64_t* __restrict__ w, uint32_t nb, uint32_t l, Out out, uint64_t& sum, uint32_t& runs) {
    uint32_t i = 1, pos = 0, cur_idx = 0;
    uint64_t w0 = 0, w1 = 0;
    if (nb) { w0 = w[0]; w1 = w[1]; }
    sum = 0; runs = 1;
    while (i < l && pos < nb) {
        const uint32_t idx = pos >> 6, off = pos & 63u;
        if (idx != cur_idx) { w0 = idx == cur_idx + 1u ? w1 : w[idx]; w1 = w[idx + 1u]; cur_idx = idx; }
        const uint64_t win = off ? (w0 << off) | (w1 >> (64u - off)) : w0;   // the next 64 bits of the stream
        const uint32_t rem = nb - pos;
        if (!(win >> 63)) {   // zero bits: deltas of 1
            const uint32_t z = kdbx_min_u32(kdbx_min_u32(kdbx_clz64(win), rem), l - i);
            out[i] = kRunToken | z;
            i += z; pos += z; sum += z;
        } else {              // `ones` one bits, a zero bit, `ones` low bits: the value (1 << ones) | low  (>= 2)
            const uint32_t ones = kdbx_clz64(~win);
            const uint32_t len = 2u * ones + 1u;
            if (ones > 31u || len > rem) return 1;
            const uint32_t v = (1u << ones) | ((uint32_t)(win >> (64u - len)) & ((1u << ones) - 1u));
            if (v & kRunToken) return 2;
            out[i++] = v;
            pos += len; sum += v; ++runs;
        }
    }
    return (i == l && pos == nb) ? 0 : 1;
}

// Pass 2 without job counting: out[0] = first (already stored by the caller); ids 1 .. l-1 from the tokens.
template <class Out>
KDBX_HD void tokens_to_ids(Out out, uint32_t l, uint32_t first) {
    uint32_t cur = first, i = 1;
    while (i < l) {
        const uint32_t v = out[i];
        if (v & kRunToken) {
            const uint32_t z = v & ~kRunToken;
            for (uint32_t t = 0; t < z; ++t) out[i + t] = cur + 1u + t;
            cur += z; i += z;
        } else {
            cur += v;
            out[i++] = cur;
        }
    }
}

// Pass 2 with job counting: additionally calls close(row_block, j, k) for every maximal stretch of k ids starting at list
// position j that fall into one block of 1 << sh matrix rows — the same stretches, in the same order, as a test of every
// id against its predecessor's block would find.
template <class Out, class Close>
KDBX_HD void tokens_to_ids_blocks(Out out, uint32_t l, uint32_t first, uint32_t sh, Close&& close) {
    uint32_t cur = first, i = 1, run_j = 0, run_rb = first >> sh;
    while (i < l) {
        const uint32_t v = out[i];
        if (v & kRunToken) {
            const uint32_t z = v & ~kRunToken;
            for (uint32_t t = 0; t < z; ++t) out[i + t] = cur + 1u + t;
            uint32_t id = cur + 1u, k = i, rem = z;   // ids id .. id + rem - 1 at positions k ..
            while (rem) {
                const uint32_t rb = id >> sh;
                if (rb != run_rb) { close(run_rb, run_j, k - run_j); run_j = k; run_rb = rb; }
                const uint32_t take = kdbx_min_u32(rem, ((rb + 1u) << sh) - id);
                id += take; k += take; rem -= take;
            }
            cur += z; i += z;
        } else {
            cur += v;
            out[i] = cur;
            const uint32_t rb = cur >> sh;
            if (rb != run_rb) { close(run_rb, run_j, i - run_j); run_j = i; run_rb = rb; }
            ++i;
        }
    }
    close(run_rb, run_j, l - run_j);
}
