// Elias-gamma bit-stream codec for local sample-id lists.  Wire format as produced by the
// reference (src/elias_gamma.h:104-128): a value v >= 1 with b = bitlen(v)
// is (b-1) one-bits, a zero bit, then the low b-1 bits of v; codes are packed MSB-first into
// consecutive 64-bit words.  The implementation is ours (clz based, no lookup tables).
#pragma once
#include <cstdint>

namespace kdbx {

// Appends the code of `v` (v >= 1) at bit position `nbits` of `words` (which must be zeroed
// beyond nbits and large enough); advances nbits.
inline void gamma_put(uint64_t* words, uint32_t& nbits, uint32_t v) {
    const uint32_t b = 32 - (uint32_t)__builtin_clz(v);  // bit length
    const uint32_t len = 2 * b - 1;
    // code value, right-aligned in `len` bits: (b-1) ones, 0, low b-1 bits of v
    const uint64_t ones = (b > 1) ? (((uint64_t)1 << (b - 1)) - 1) : 0;
    const uint64_t code = (ones << b) | (uint64_t)(v - ((uint32_t)1 << (b - 1)));
    const uint32_t w = nbits >> 6, used = nbits & 63, room = 64 - used;
    if (len <= room) {
        words[w] |= code << (room - len);
    } else {
        const uint32_t rest = len - room;
        words[w] |= code >> rest;
        words[w + 1] |= code << (64 - rest);
    }
    nbits += len;
}

inline uint32_t gamma_code_len(uint32_t v) { return 2 * (32 - (uint32_t)__builtin_clz(v)) - 1; }

// Reads one bit-field of `cnt` (<= 32) bits starting at absolute bit `pos`.
inline uint32_t gamma_bits(const uint64_t* words, uint32_t pos, uint32_t cnt) {
    if (cnt == 0) return 0;
    const uint32_t w = pos >> 6, off = pos & 63;
    uint64_t x = words[w] << off;
    if (off + cnt > 64) x |= words[w + 1] >> (64 - off);
    return (uint32_t)(x >> (64 - cnt));
}

// Decodes one value at bit position pos; advances pos.
inline uint32_t gamma_get(const uint64_t* words, uint32_t& pos) {
    uint32_t ones = 0;
    for (;;) {  // count the unary prefix across word boundaries
        const uint32_t w = pos >> 6, off = pos & 63;
        const uint64_t x = ~(words[w] << off);  // leading ones -> leading zeros of ~x
        const uint32_t avail = 64 - off;
        uint32_t run = x ? (uint32_t)__builtin_clzll(x) : 64;
        if (run >= avail) { ones += avail; pos += avail; continue; }
        ones += run; pos += run + 1;  // skip the terminating zero
        break;
    }
    const uint32_t low = gamma_bits(words, pos, ones);
    pos += ones;
    return ((uint32_t)1 << ones) | low;
}

// Local sample ids of one pattern, ascending, into out[0..l) (src/pattern.cpp:99-109).
inline void decode_local_ids(const uint64_t* words, uint32_t l, uint32_t last, uint32_t* out) {
    if (l == 0) return;
    uint32_t pos = 0;
    for (uint32_t i = 0; i + 1 < l; ++i) out[i] = gamma_get(words, pos);  // delta(i -> i+1)
    uint32_t cur = last;
    for (uint32_t i = l; i-- > 0;) {
        const uint32_t d = (i > 0) ? out[i - 1] : 0;
        out[i] = cur;
        cur -= d;
    }
}

}  // namespace kdbx
