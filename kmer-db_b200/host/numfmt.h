// Number formatting of the CSV tables: plain decimal integers and doubles with exactly six
// decimals, rounded by adding 0.5 at the sixth place and truncating; an exact zero prints as
// "0" (behaviour of num2str / Double2PChar, src/conversion.h:167-219,254-260).
#pragma once
#include <cstdint>
#include <cstring>

namespace kdbx {

inline char* put_u64(char* p, uint64_t v) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

inline char* put_f6(char* p, double val) {
    if (val == 0) { *p++ = '0'; return p; }
    if (val < 0) { *p++ = '-'; val = -val; }
    const uint64_t x = (uint64_t)(val * 1000000.0 + 0.5);
    p = put_u64(p, x / 1000000u);
    *p++ = '.';
    uint32_t frac = (uint32_t)(x % 1000000u);
    for (int i = 5; i >= 0; --i) { p[i] = (char)('0' + frac % 10); frac /= 10; }
    return p + 6;
}

}  // namespace kdbx
