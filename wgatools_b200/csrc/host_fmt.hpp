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

inline std::string fmt_f32(float v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    std::string out;
    if (std::signbit(v)) { out.push_back('-'); v = -v; }
    if (v == 0.0f) return out + "0.0";
    // shortest round-trip digits in scientific form: d[.ddd]e[+-]XX
    char buf[48];
    auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
    std::string sci(buf, res.ptr);
    size_t epos = sci.find('e');
    std::string digits;
    for (size_t i = 0; i < epos; ++i)
        if (sci[i] >= '0' && sci[i] <= '9') digits.push_back(sci[i]);
    int e10 = std::stoi(sci.substr(epos + 1));
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    int length = (int)digits.size(), k = e10 - (length - 1), kk = length + k;
    if (0 <= k && kk <= 13) {  // 1234e7 -> 12340000000.0
        out += digits;
        out.append((size_t)(kk - length), '0');
        out += ".0";
    } else if (0 < kk && kk <= 13) {  // 1234e-2 -> 12.34
        out += digits.substr(0, (size_t)kk);
        out.push_back('.');
        out += digits.substr((size_t)kk);
    } else if (-6 < kk && kk <= 0) {  // 1234e-6 -> 0.001234
        out += "0.";
        out.append((size_t)(-kk), '0');
        out += digits;
    } else if (length == 1) {  // 1e30
        out += digits;
        out += "e" + std::to_string(kk - 1);
    } else {  // 1234e30 -> 1.234e33
        out.push_back(digits[0]);
        out.push_back('.');
        out += digits.substr(1);
        out += "e" + std::to_string(kk - 1);
    }
    return out;
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
