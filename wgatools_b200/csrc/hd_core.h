// hd_core.h -- per-thread building blocks of the column walk, usable from device code (nvcc) and
// from a plain g++ build (tests/test_hd_core.py compiles them on the CPU box, where there is no GPU,
// to check the byte-SIMD logic against a scalar loop before any GPU time is spent).
//
// Columns are handled 4 at a time as bytes of a 32-bit word (byte 0 = lowest address = earliest
// column).  A "flag word" carries one bit per column at bit 7 of each byte (0x80808080 mask).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define WGA_HD __host__ __device__ __forceinline__
#else
#define WGA_HD static inline
#endif

namespace wga {

constexpr uint32_t kHi = 0x80808080u;
constexpr uint32_t kGap4 = 0x2D2D2D2Du;  // "----"

// 0x80 in every byte of x that is zero (exact, carry-free)
WGA_HD uint32_t zero_bytes(uint32_t x) {
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & kHi;
}

// 0x80 in every byte that is ASCII whitespace: 0x09..0x0D or 0x20 (Rust char::is_whitespace on ASCII,
// the split_whitespace used by parse_sline, src/parser/maf.rs:139)
WGA_HD uint32_t ws_bytes(uint32_t w) {
    uint32_t lo = w & 0x7F7F7F7Fu;
    uint32_t ge9 = lo + 0x77777777u;   // bit7 set iff (b&0x7f) >= 9
    uint32_t gt13 = lo + 0x72727272u;  // bit7 set iff (b&0x7f) >= 14
    uint32_t ctl = ge9 & ~gt13 & ~w & kHi;
    return ctl | zero_bytes(w ^ 0x20202020u);
}
WGA_HD uint32_t nl_bytes(uint32_t w) { return zero_bytes(w ^ 0x0A0A0A0Au); }

WGA_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
// (hi:lo) >> (8*sh) low word, sh in 0..3
WGA_HD uint32_t funnel_r_bytes(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh * 8);
#else
    return sh ? (lo >> (8 * sh)) | (hi << (32 - 8 * sh)) : lo;
#endif
}
// previous-column flags aligned to the current column: (cur << 8) | (prev >> 24)
WGA_HD uint32_t prev_flags(uint32_t prev_word, uint32_t cur_word) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(prev_word, cur_word, 8);
#else
    return (cur_word << 8) | (prev_word >> 24);
#endif
}

// Column classes.  EXT = cigar_cat_ext (src/parser/cigar.rs:298-308): equal bytes are '=' even when
// both are '-'.  CALLER = cigar_cat_ext_caller (cigar.rs:314-328): both '-' is 'W'.
enum : int { CLS_EQ = 0, CLS_X = 1, CLS_I = 2, CLS_D = 3, CLS_W = 4 };

struct Flags4 {  // one flag word per class for 4 columns
    uint32_t eq, x, i, d, w;
};

template <bool CALLER>
WGA_HD Flags4 classify4(uint32_t t, uint32_t q, uint32_t valid /* 0x80 per in-record byte */) {
    uint32_t e = zero_bytes(t ^ q);
    uint32_t tg = zero_bytes(t ^ kGap4);
    uint32_t qg = zero_bytes(q ^ kGap4);
    Flags4 f;
    if (CALLER) {
        f.w = tg & qg & valid;
        f.i = tg & ~qg & valid;
        f.d = qg & ~tg & valid;
        f.eq = e & ~tg & valid;
    } else {
        f.w = 0;
        f.i = tg & ~e & valid;
        f.d = qg & ~e & valid;
        f.eq = e & valid;
    }
    f.x = ~(e | tg | qg) & valid;
    return f;
}

// 0x80 flags for the bytes of a word whose bits are set in the low nibble of n (bit i -> byte i): one multiply spreads the
// four bits to bit 0 of the four bytes (bit i lands at i, i+7, i+14, i+21; only i + 7i = 8i survives the mask)
WGA_HD uint32_t nibble_flags(uint32_t n) { return (((n & 0xFu) * 0x00204081u) & 0x01010101u) << 7; }

// class code (CLS_*) of four columns, one per byte (bits 0-2), from their class flags: '=' is 0
WGA_HD uint32_t class_codes(const Flags4 &f) { return (f.x >> 7) | (f.i >> 6) | (f.d >> 7) | (f.d >> 6) | (f.w >> 5); }

// class flags of a single (t,q) byte pair placed at bit 31 (as the "previous word" of prev_flags)
template <bool CALLER>
WGA_HD Flags4 classify_prev(uint32_t tb, uint32_t qb, bool valid) {
    Flags4 f = classify4<CALLER>(tb << 24, (qb << 24) | 0x00010203u /* low bytes differ from t's zeros */,
                                 valid ? 0x80000000u : 0u);
    return f;
}

// decimal digits of v (>= 1)
WGA_HD int ndigits_u64(uint64_t v) {
    if (v <= 0xFFFFFFFFull) {  // coordinates and sizes: compares only, no 64-bit division
        const uint32_t x = (uint32_t)v;
        return x < 100000u ? (x < 100u ? (x < 10u ? 1 : 2) : (x < 1000u ? 3 : (x < 10000u ? 4 : 5)))
                           : (x < 10000000u ? (x < 1000000u ? 6 : 7) : (x < 100000000u ? 8 : (x < 1000000000u ? 9 : 10)));
    }
    int n = 1;
    if (v >= 10000000000ull) { v /= 10000000000ull; n += 10; }
    if (v >= 100000ull) { v /= 100000ull; n += 5; }
    if (v >= 1000ull) { n += 3; v /= 1000ull; }
    if (v >= 100ull) n += 2; else if (v >= 10ull) n += 1;
    return n;
}
// writes v in decimal at dst[0..nd) where nd == ndigits_u64(v); returns nd
WGA_HD int write_u64(char *dst, uint64_t v, int nd) {
    int i = nd - 1;
    for (; i >= 0 && v > 0xFFFFFFFFull; --i) {
        dst[i] = (char)('0' + (int)(v % 10));
        v /= 10;
    }
    uint32_t x = (uint32_t)v;  // 32-bit division by a constant = one multiply-high
    for (; i >= 0; --i) {
        const uint32_t q = x / 10u;
        dst[i] = (char)('0' + (int)(x - q * 10u));
        x = q;
    }
    return nd;
}

// str::parse::<u64>() (src/utils.rs:69-74): optional '+', at least one digit, no overflow
WGA_HD bool parse_u64_bytes(const unsigned char *p, uint64_t n, uint64_t *out) {
    uint64_t i = 0;
    if (n == 0) return false;
    if (p[0] == '+') i = 1;
    if (i == n) return false;
    uint64_t v = 0;
    for (; i < n; ++i) {
        unsigned d = (unsigned)p[i] - '0';
        if (d > 9) return false;
        if (v > (0xFFFFFFFFFFFFFFFFull - d) / 10) return false;
        v = v * 10 + d;
    }
    *out = v;
    return true;
}

WGA_HD uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace wga
