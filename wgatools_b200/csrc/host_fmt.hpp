// host_fmt.hpp -- host-side formatting used by the stat accumulator and the CLI: the text layouts of
// third-party crates the reference prints through (restated from their published algorithms):
//   ryu 1.0.14  Buffer::format(f32)   -> fmt_f32      (csv serialises f32 through ryu)
//   natord 1.0.9 compare              -> natord_cmp   (src/tools/stat.rs:116)
//   csv 1.2.2   QuoteStyle::Necessary -> csv_field
#pragma once
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <string_view>

namespace wga {

// ryu's Buffer::format(f32) text (what csv prints): shortest round-trip digits, laid out as ryu's pretty printer does
// (d2s/f2s "format32": plain decimals for exponents in [-5, 13), scientific otherwise).  Writes at most 24 bytes, returns
// the length.
inline size_t fmt_f32_to(char *out, float v) {
    char *o = out;
    if (std::isnan(v)) { memcpy(o, "NaN", 3); return 3; }
    if (std::isinf(v)) {
        if (v < 0) { memcpy(o, "-inf", 4); return 4; }
        memcpy(o, "inf", 3);
        return 3;
    }
    if (std::signbit(v)) { *o++ = '-'; v = -v; }
    if (v == 0.0f) { memcpy(o, "0.0", 3); return (size_t)(o - out) + 3; }
    // shortest round-trip digits in scientific form: d[.ddd]e[+-]XX
    char buf[32];
    auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
    char digits[16];
    int length = 0;
    const char *p = buf;
    for (; p < res.ptr && *p != 'e'; ++p)
        if (*p >= '0' && *p <= '9') digits[length++] = *p;
    int e10 = 0;
    {
        ++p;  // 'e'
        bool neg = *p == '-';
        if (*p == '-' || *p == '+') ++p;
        for (; p < res.ptr; ++p) e10 = e10 * 10 + (*p - '0');
        if (neg) e10 = -e10;
    }
    while (length > 1 && digits[length - 1] == '0') --length;
    const int k = e10 - (length - 1), kk = length + k;
    if (0 <= k && kk <= 13) {  // 1234e7 -> 12340000000.0
        memcpy(o, digits, (size_t)length); o += length;
        for (int i = length; i < kk; ++i) *o++ = '0';
        *o++ = '.'; *o++ = '0';
    } else if (0 < kk && kk <= 13) {  // 1234e-2 -> 12.34
        memcpy(o, digits, (size_t)kk); o += kk;
        *o++ = '.';
        memcpy(o, digits + kk, (size_t)(length - kk)); o += length - kk;
    } else if (-6 < kk && kk <= 0) {  // 1234e-6 -> 0.001234
        *o++ = '0'; *o++ = '.';
        for (int i = 0; i < -kk; ++i) *o++ = '0';
        memcpy(o, digits, (size_t)length); o += length;
    } else {  // 1e30, 1.234e33
        *o++ = digits[0];
        if (length > 1) { *o++ = '.'; memcpy(o, digits + 1, (size_t)(length - 1)); o += length - 1; }
        *o++ = 'e';
        auto r2 = std::to_chars(o, o + 8, kk - 1);
        o = r2.ptr;
    }
    return (size_t)(o - out);
}
inline std::string fmt_f32(float v) {
    char b[32];
    return std::string(b, fmt_f32_to(b, v));
}

inline bool nat_ws(char c) { return c == ' ' || (c >= 9 && c <= 13); }

// natord::compare: whitespace skipped, digit runs compared numerically (right-aligned) unless one
// starts with '0' (left-aligned, "fractional"), everything else by char
inline int natord_cmp(std::string_view a, std::string_view b) {
    size_t i = 0, j = 0;
    auto dig = [](std::string_view s, size_t p) -> int {
        return p < s.size() && s[p] >= '0' && s[p] <= '9' ? s[p] - '0' : -1;
    };
    for (;;) {
        while (i < a.size() && nat_ws(a[i])) ++i;
        while (j < b.size() && nat_ws(b[j])) ++j;
        bool ea = i >= a.size(), eb = j >= b.size();
        if (ea || eb) return ea && eb ? 0 : (ea ? -1 : 1);
        int da = dig(a, i), db = dig(b, j);
        if (da < 0 || db < 0) {
            if (a[i] != b[j]) return (unsigned char)a[i] < (unsigned char)b[j] ? -1 : 1;
            ++i; ++j;
            continue;
        }
        bool fractional = da == 0 || db == 0;
        int first_diff = da == db ? 0 : (da < db ? -1 : 1);
        if (fractional && first_diff) return first_diff;
        for (;;) {
            ++i; ++j;
            da = dig(a, i); db = dig(b, j);
            if (da >= 0 && db >= 0) {
                int c = da == db ? 0 : (da < db ? -1 : 1);
                if (fractional) { if (c) return c; }
                else if (!first_diff) first_diff = c;
            } else if (da >= 0) return 1;
            else if (db >= 0) return -1;
            else break;
        }
        if (!fractional && first_diff) return first_diff;
    }
}

inline void csv_field(std::string &out, std::string_view f) {
    bool need = false;
    for (char c : f) need |= c == '\t' || c == '"' || c == '\r' || c == '\n';
    if (!need) { out.append(f.data(), f.size()); return; }
    out.push_back('"');
    for (char c : f) { if (c == '"') out.push_back('"'); out.push_back(c); }
    out.push_back('"');
}

}  // namespace wga
