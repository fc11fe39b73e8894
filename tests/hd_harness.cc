// CPU harness around wgatools_b200/csrc/hd_core.h and host_fmt.hpp (test infrastructure): lets the
// no-GPU test suite check the byte-SIMD column classifier and the host formatters.
#include <cstring>
#include <string>
#include "../wgatools_b200/csrc/hd_core.h"
#include "../wgatools_b200/csrc/host_fmt.hpp"
using namespace wga;
extern "C" {
uint32_t hd_zero_bytes(uint32_t x) { return zero_bytes(x); }
uint32_t hd_ws_bytes(uint32_t x) { return ws_bytes(x); }
void hd_classify4(uint32_t t, uint32_t q, uint32_t valid, int caller, uint32_t *out5) {
    Flags4 f = caller ? classify4<true>(t, q, valid) : classify4<false>(t, q, valid);
    out5[0] = f.eq; out5[1] = f.x; out5[2] = f.i; out5[3] = f.d; out5[4] = f.w;
}
void hd_classify_prev(uint32_t tb, uint32_t qb, int caller, uint32_t *out5) {
    Flags4 f = caller ? classify_prev<true>(tb, qb, true) : classify_prev<false>(tb, qb, true);
    out5[0] = f.eq; out5[1] = f.x; out5[2] = f.i; out5[3] = f.d; out5[4] = f.w;
}
int hd_ndigits(uint64_t v) { return ndigits_u64(v); }
int hd_write_u64(char *dst, uint64_t v) { int n = ndigits_u64(v); write_u64(dst, v, n); return n; }
int hd_parse_u64(const char *s, uint64_t n, uint64_t *out) { return parse_u64_bytes((const unsigned char *)s, n, out); }
uint32_t hd_prev_flags(uint32_t p, uint32_t c) { return prev_flags(p, c); }
uint32_t hd_funnel(uint32_t lo, uint32_t hi, uint32_t sh) { return funnel_r_bytes(lo, hi, sh); }
void hd_fmt_f32(float v, char *dst) { std::string s = fmt_f32(v); memcpy(dst, s.c_str(), s.size() + 1); }
int hd_natord(const char *a, const char *b) { return natord_cmp(a, b); }
uint32_t hd_nibble_flags(uint32_t n) { return nibble_flags(n); }
uint32_t hd_class_codes(uint32_t t, uint32_t q, uint32_t valid, int caller) {
    return class_codes(caller ? classify4<true>(t, q, valid) : classify4<false>(t, q, valid));
}
}
