// paf.cuh -- `stat -f paf`: on-device PAF tokenisation + CIGAR counters.
//
// Replaces src/parser/paf.rs:22-31 (csv reader: tab-delimited, no header, flexible, '#' comments),
// :48-65 PafRecord (12 typed columns + tags), :122-140 get_cigar_string (first cg:Z: tag), and
// src/parser/cigar.rs:43-75 (tokenizer) + :629-707 parse_paf_to_cigar (M,= -> matched; X; I/D by strand; any
// other op is CigarOpInvalid).
//
// The delimiter index ('\t', '\n') reuses the single-pass slot scan of tokenize.cuh with a PAF byte class.
// CIGAR text is walked one record per thread with 16-byte loads (records are short -- a few hundred ops --
// and the per-op work, a decimal parse, dominates; see DESIGN.md for the bound).
// A record without cg:Z: but with a cs:Z: tag gets its CIGAR from k_cs_convert (paf.rs:159-218 cs_to_cigar), written to a
// device arena; cig_off then points there (as a wrapped 64-bit distance from the slab text, so every consumer's
// `text + cig_off` keeps working).
// Quoted csv fields: see k_unquote_lines below (simple quoting is normalised away, the rest is reported, never silently wrong).
#pragma once
#include "ctx.cuh"
#include "stat.cuh"
#include "tokenize.cuh"

namespace wga {

struct PafRecs {
    uint64_t *q_name_off, *t_name_off, *cig_off, *cig_len;
    uint32_t *q_name_len, *t_name_len;
    uint64_t *q_size, *t_size, *q_start, *t_start;
    uint8_t *neg;
};

// flags: bit0 = record line (non-empty, not a '#' comment).  Lines end at '\r' or '\n' (tokenize.cuh CLS 1); a comment runs
// to the next '\n' (csv-core InComment), so a line is also a comment when a line before it -- reached only through
// '\r' line ends -- starts with '#'.
__global__ void k_paf_line_flags(LineView lv, uint8_t *__restrict__ flags, uint32_t *__restrict__ isrec) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > lv.n_lines) return;
    if (k == lv.n_lines) { isrec[k] = 0; return; }
    const uint64_t s = lv.start(k), e = lv.end(k);
    bool rec = e > s && lv.text[s] != '#';     // csv: empty lines skipped, comment(Some(b'#'))
    for (uint64_t j = k; rec && j > 0 && lv.text[lv.end(j - 1)] == '\r'; --j) {
        const uint64_t ps = lv.start(j - 1);
        if (lv.end(j - 1) > ps && lv.text[ps] == '#') rec = false;
    }
    flags[k] = rec ? 1 : 0;
    isrec[k] = rec ? 1u : 0u;
}

// csv deserializer integer rule (csv 1.2.2 deserializer.rs): "0x" prefix => radix 16, else str::parse::<u64>()
__device__ __forceinline__ bool csv_u64(const uint8_t *p, uint64_t n, uint64_t *out) {
    if (n >= 2 && p[0] == '0' && p[1] == 'x') {
        if (n == 2) return false;
        uint64_t i = 2, v = 0;
        if (p[i] == '+') { ++i; if (i == n) return false; }
        for (; i < n; ++i) {
            uint8_t c = p[i];
            uint32_t d = c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : 99u;
            if (d > 15 || (v >> 60)) return false;
            v = v * 16 + d;
        }
        *out = v;
        return true;
    }
    return parse_u64_bytes(p, n, out);
}

// err key: record index << 32 | position (24 bits) << 8 | code   (lowest record, then earliest position wins)
__device__ __forceinline__ void paf_report(unsigned long long *err_key, uint64_t rec, uint64_t pos, int code) {
    if (pos > 0xFFFFFFull) pos = 0xFFFFFFull;
    atomicMin(err_key, (unsigned long long)((rec << 32) | (pos << 8) | (uint64_t)code));
}

// ---- quoted csv fields (csv reader default: quoting on).  The fast path refuses a field that starts with '"'
// (k_paf_parse reports WGA_E_UNSUPPORTED); the entry points then normalise the text and run again: a field "abc" whose
// content holds no quote, tab or line break -- quoting that was not needed -- is replaced by its content, which is what
// the reader would hand to the deserializer.  Content that holds a '"' (escaped ""), a tab or a line break, text behind
// the closing quote, or a first field whose content starts with '#', stay refused.
// PASS 0: byte length of every normalised line (terminator included); PASS 1: the bytes.
template <int PASS>
__global__ void k_unquote_lines(LineView lv, const uint8_t *__restrict__ flags, uint64_t *__restrict__ new_len,
                                const uint64_t *__restrict__ new_off, uint8_t *__restrict__ out, unsigned long long *err_key) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > lv.n_lines) return;
    if (k == lv.n_lines) { if (PASS == 0) new_len[k] = 0; return; }
    const uint64_t S = lv.start(k), E = lv.end(k);
    const uint64_t term = k < lv.n_nl ? 1 : 0;  // the line's own terminator byte ('\n' or '\r') is kept
    uint8_t *o = PASS == 1 ? out + new_off[k] : nullptr;
    if (!flags[k]) {  // empty line or comment: verbatim
        if (PASS == 0) new_len[k] = (E - S) + term;
        else for (uint64_t i = S; i < E + term; ++i) *o++ = lv.text[i];
        return;
    }
    const uint64_t wlo = k == 0 ? 0 : lv.nl_ws_idx[k - 1] + 1, whi = k < lv.n_nl ? lv.nl_ws_idx[k] : lv.n_ws;
    uint64_t fb = S, w = wlo, n = 0;
    bool complex = false;
    for (uint64_t nf = 0;; ++nf) {
        uint64_t fe = w < whi ? (lv.ws_ent[w] & kPosMask) : E;
        if (fe > E) fe = E;
        uint64_t cb = fb, ce = fe;  // content
        if (fe > fb && lv.text[fb] == '"') {
            bool simple = fe - fb >= 2 && lv.text[fe - 1] == '"';
            for (uint64_t i = fb + 1; simple && i + 1 < fe; ++i) simple = lv.text[i] != '"';
            if (simple && nf == 0 && fe - fb > 2 && lv.text[fb + 1] == '#') simple = false;  // would turn into a comment line
            if (!simple) complex = true;
            cb = fb + 1; ce = fe - 1;
        }
        if (PASS == 0) n += ce - cb;
        else for (uint64_t i = cb; i < ce; ++i) *o++ = lv.text[i];
        if (w >= whi) break;
        if (PASS == 0) n += 1;
        else *o++ = '\t';
        fb = fe + 1;
        ++w;
    }
    if (PASS == 0) {
        if (n == 0) complex = true;  // a record of one empty quoted field is not an empty line
        if (complex) paf_report(err_key, k, 0, WGA_E_UNSUPPORTED);
        new_len[k] = n + term;
    } else if (term) *o = lv.text[E];
}

// one PAF line -> 12 typed columns + cg:Z: span (paf.rs:48-65,122-140)
// tag_key: where a missing cg:Z: tag is reported (== err_key for commands that look every record's CIGAR up in file order)
// cs_total: summed length of the cs:Z: tags that stand in for a missing cg:Z: (arena bound for k_cs_convert)
__global__ void k_paf_parse(LineView lv, const uint8_t *__restrict__ flags, const uint32_t *__restrict__ rec_scan,
                            PafRecs R, unsigned long long *err_key, unsigned long long *tag_key,
                            unsigned long long *cs_total) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= lv.n_lines || !flags[k]) return;
    const uint64_t r = rec_scan[k];
    uint64_t S = lv.start(k), E = lv.end(k);
    if (E > S && lv.text[E - 1] == '\r') --E;
    uint64_t wlo = k == 0 ? 0 : lv.nl_ws_idx[k - 1] + 1;
    uint64_t whi = k < lv.n_nl ? lv.nl_ws_idx[k] : lv.n_ws;
    // a record with fewer than 6 fields never reaches its names: keep them defined (later passes hash / compare them)
    R.q_name_off[r] = S; R.q_name_len[r] = 0; R.t_name_off[r] = S; R.t_name_len[r] = 0;
    // fields 0..11, then tags
    uint64_t fb = S, w = wlo;
    uint64_t v[12];
    uint8_t neg = 0;
    int code = 0;
    uint64_t cg_off = 0, cg_len = 0, cs_off = 0, cs_len = 0;
    bool have_cg = false, have_cs = false, quoted = false;
    uint64_t nf = 0;
    for (;; ++nf) {
        uint64_t fe = w < whi ? (lv.ws_ent[w] & kPosMask) : E;
        if (fe > E) fe = E;
        const uint8_t *f = lv.text + fb;
        const uint64_t fl = fe - fb;
        if (fl && f[0] == '"') quoted = true;
        if (nf < 12) {
            if (nf == 0) { R.q_name_off[r] = fb; R.q_name_len[r] = (uint32_t)fl; }
            else if (nf == 5) { R.t_name_off[r] = fb; R.t_name_len[r] = (uint32_t)fl; }
            else if (nf == 4) {
                if (fl == 1 && f[0] == '+') neg = 0;
                else if (fl == 1 && f[0] == '-') neg = 1;
                else if (!code) code = WGA_E_CSV;
            } else if (!csv_u64(f, fl, &v[nf])) { if (!code) code = WGA_E_CSV; }
        } else if (!have_cg && fl >= 5 && f[0] == 'c' && f[2] == ':' && f[3] == 'Z' && f[4] == ':') {
            if (f[1] == 'g') { have_cg = true; cg_off = fb + 5; cg_len = fl - 5; }
            else if (f[1] == 's' && !have_cs) { have_cs = true; cs_off = fb + 5; cs_len = fl - 5; }  // the FIRST cs:Z: tag
        }
        if (w >= whi) { ++nf; break; }
        fb = fe + 1;
        ++w;
    }
    if (quoted) { paf_report(err_key, r, 0, WGA_E_UNSUPPORTED); code = WGA_E_UNSUPPORTED; }
    if (!code && nf < 12) code = WGA_E_CSV;
    if (code) {
        if (code != WGA_E_UNSUPPORTED) paf_report(err_key, r, 0, code);
        R.cig_len[r] = 0; R.cig_off[r] = 0; R.neg[r] = 0;
        R.q_size[r] = R.t_size[r] = R.q_start[r] = R.t_start[r] = 0;
        return;
    }
    R.q_size[r] = v[1]; R.q_start[r] = v[2];
    R.t_size[r] = v[6]; R.t_start[r] = v[7];
    R.neg[r] = neg;
    if (!have_cg) {  // no tag at all is the reference's error; a cs:Z: tag is converted by k_cs_convert (paf.rs:129-135)
        if (have_cs) {
            cg_off = cs_off; cg_len = cs_len;
            atomicAdd(cs_total, (unsigned long long)cs_len + 1ull);
        } else {
            paf_report(tag_key, r, 0x800000, WGA_E_CIGAR_TAG_NOT_FOUND);
            cg_len = 0;
            cg_off = 0;
        }
        R.neg[r] = neg | 2 | (have_cs ? 4 : 0);  // bit 1 marks "no cg:Z:" (bit 0 stays the strand), bit 2 "but a cs:Z:"
    }
    R.cig_off[r] = cg_off;
    R.cig_len[r] = cg_len;
}

// cs:Z: -> CIGAR (paf.rs:159-218 cs_to_cigar): the regex (:[0-9]+|\*[a-z][a-z]|[=\+\-][A-Za-z]+) finds leftmost
// non-overlapping tokens (anything else is skipped); `:N` extends or opens an M run, `*xy` an X run, `-seq` / `+seq`
// close the open run and emit their own op, `=seq` is matched but ignored (the match arm `_ => {}`).  The output is
// never longer than the input (every token yields at most its own byte count), so record r owns cs_len + 1 arena
// bytes handed out by an atomic cursor; the bytes behind the CIGAR are tab-filled, as in PAF text.  A `:N` beyond u64 makes the reference's parse().unwrap() panic.
// One thread per cs record: cs:Z: without cg:Z: is the rare spelling (minimap2 --cs without -c).
__device__ __forceinline__ uint32_t cs_put_u64(uint8_t *o, unsigned long long v) {
    uint8_t tmp[20];
    uint32_t n = 0;
    do { tmp[n++] = (uint8_t)('0' + v % 10); v /= 10; } while (v);
    for (uint32_t i = 0; i < n; ++i) o[i] = tmp[n - 1 - i];
    return n;
}
__global__ void k_cs_convert(uint64_t n_rec, const uint8_t *__restrict__ text, PafRecs R, uint8_t *__restrict__ arena,
                             unsigned long long *__restrict__ cursor, unsigned long long *tag_key) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint8_t negf = R.neg[r];
    if ((negf & 6) != 6) return;
    const uint64_t n = R.cig_len[r];
    const uint8_t *cs = text + R.cig_off[r];
    uint8_t *o = arena + atomicAdd(cursor, (unsigned long long)n + 1ull);
    uint64_t w = 0, i = 0;
    uint8_t last_op = 'M';
    unsigned long long last_len = 0;
    auto is_alpha = [](uint8_t c) { return (uint8_t)((c | 0x20) - 'a') < 26u; };
    auto flush = [&]() { if (last_len) { w += cs_put_u64(o + w, last_len); o[w++] = last_op; } };
    while (i < n) {
        const uint8_t c = cs[i];
        if (c == ':' && i + 1 < n && (uint8_t)(cs[i + 1] - '0') < 10u) {
            unsigned long long v = 0;
            uint64_t j = i + 1;
            for (; j < n && (uint8_t)(cs[j] - '0') < 10u; ++j) {
                const unsigned long long d = cs[j] - '0';
                if (v > (0xFFFFFFFFFFFFFFFFull - d) / 10) { paf_report(tag_key, r, 0x800000, WGA_E_REF_PANIC); return; }
                v = v * 10 + d;
            }
            if (last_op == 'M') last_len += v;
            else { flush(); last_op = 'M'; last_len = v; }
            i = j;
        } else if (c == '*' && i + 2 < n && (uint8_t)(cs[i + 1] - 'a') < 26u && (uint8_t)(cs[i + 2] - 'a') < 26u) {
            if (last_op == 'X') last_len += 1;
            else { flush(); last_op = 'X'; last_len = 1; }
            i += 3;
        } else if ((c == '=' || c == '+' || c == '-') && i + 1 < n && is_alpha(cs[i + 1])) {
            uint64_t j = i + 1;
            while (j < n && is_alpha(cs[j])) ++j;
            if (c != '=') {
                flush();
                w += cs_put_u64(o + w, j - (i + 1));
                o[w++] = c == '-' ? 'D' : 'I';
                last_len = 0;
                last_op = 'M';
            }
            i = j;
        } else {
            ++i;
        }
    }
    flush();
    for (uint64_t k = w; k <= n; ++k) o[k] = '\t';  // what follows a CIGAR in PAF text: never a digit, never stale bytes
    R.cig_off[r] = (uint64_t)(o - text);  // wraps when the arena lies below the slab: text + cig_off is the arena address
    R.cig_len[r] = w;
    R.neg[r] = negf & 1;                  // from here on the record is indistinguishable from one that carried cg:Z:
}

// CIGAR text -> the 11 counters of struct Cigar (cigar.rs:16-29), fast path: one WARP per record, lane i of a step
// owns the aligned 16-byte piece 32*step+i.  A token belongs to the piece that holds its op char.  Phase 1 (SWAR):
// every lane parses the digit run at the END of its piece (<= 8 digits: flag words + the 3-multiply 8-digit parse)
// and hands (value, digits) to the lane on its right by shuffle (lane 31 -> lane 0 of the next step).  Phase 2: a
// branch-free pass over the 16 bytes that starts from the left neighbour's trailing number.  Anything that is not
// `digits+ one-of(M=XID)` -- i.e. every input on which the reference raises an error -- and any length of 9+ digits
// only sets slow[r]; those records are redone by the exact one-thread kernel below, which owns all error reporting.
__device__ __forceinline__ uint32_t digit_flags(uint32_t w) {  // bit 7 of every byte in '0'..'9'
    const uint32_t x = w ^ 0x30303030u;
    return ~(((x & 0x7F7F7F7Fu) + 0x76767676u) | x) & 0x80808080u;
}
__global__ void __launch_bounds__(128, 8) k_paf_cigar_counts_warp(uint64_t n_rec, const uint8_t *__restrict__ text, PafRecs R,
                                                               unsigned long long *__restrict__ cnt11, uint8_t *__restrict__ slow) {
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (r >= n_rec) return;
    const uint8_t negf = R.neg[r];
    const uint64_t len = R.cig_len[r];
    if ((negf & 2) || len == 0) { if (lane == 0) slow[r] = 1; return; }
    const uint8_t *p = text + R.cig_off[r];
    const uint64_t a = reinterpret_cast<uintptr_t>(p) & 15u;
    const uint8_t *base = p - a;
    const uint64_t nvec = (a + len + 15) >> 4;
    unsigned long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;  // M/=, X, I, D lengths
    uint32_t e_i = 0, e_d = 0;
    bool bad = false;
    uint32_t carry_val = 0, carry_nd = 0;  // trailing number of the previous step's last piece
    for (uint64_t v0 = 0; v0 < nvec; v0 += 32) {
        const uint64_t vi = v0 + lane;
        const bool have = vi < nvec;
        uint4 cur = make_uint4(0x3A3A3A3Au, 0x3A3A3A3Au, 0x3A3A3A3Au, 0x3A3A3A3Au);
        if (have) cur = ldg16(base + 16 * vi);
        // ---- phase 1: digit run at the end of the piece
        uint32_t t_nd, t_val = 0;
        {
            const unsigned long long F = ((unsigned long long)digit_flags(cur.w) << 32) | digit_flags(cur.z);
            const unsigned long long nF = ~F & 0x8080808080808080ull;
            t_nd = nF ? (uint32_t)__clzll((long long)nF) >> 3 : 8u;
            if (t_nd == 8u && (digit_flags(cur.y) & 0x80000000u)) t_nd = 9u;  // longer than the window: flagged by the consumer
            if (t_nd) {
                const unsigned long long H = ((unsigned long long)cur.w << 32) | cur.z;
                const unsigned long long keep = t_nd >= 8u ? ~0ull : ~0ull << (8u * (8u - t_nd));
                unsigned long long v = ((H & keep) | (0x3030303030303030ull & ~keep)) - 0x3030303030303030ull;
                v = v * 10 + (v >> 8);
                v = (((v & 0x000000FF000000FFull) * 0x000F424000000064ull) + (((v >> 16) & 0x000000FF000000FFull) * 0x0000271000000001ull)) >> 32;
                t_val = (uint32_t)v;
            }
        }
        uint32_t val = __shfl_up_sync(0xFFFFFFFFu, t_val, 1), nd = __shfl_up_sync(0xFFFFFFFFu, t_nd, 1);
        if (lane == 0) { val = carry_val; nd = carry_nd; }
        carry_val = __shfl_sync(0xFFFFFFFFu, t_val, 31);
        carry_nd = __shfl_sync(0xFFFFFFFFu, t_nd, 31);
        if (!have) continue;
        // ---- phase 2: the 16 bytes, branch-free
        const int64_t idx0 = (int64_t)(16 * vi) - (int64_t)a;  // CIGAR index of byte 0 of this piece
        const int klo = idx0 < 0 ? (int)-idx0 : 0;
        const int khi = (uint64_t)(idx0 + 16) > len ? (int)((int64_t)len - idx0) : 16;
        if (idx0 <= 0) { val = 0; nd = 0; }  // nothing belongs to the CIGAR on the left of its first piece
        const uint32_t cw[4] = {cur.x, cur.y, cur.z, cur.w};
        uint32_t am = 0, ax = 0, ai = 0, ad = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t ch = (cw[k >> 2] >> (8 * (k & 3))) & 0xFFu;
            const uint32_t dg = ch - '0';
            const bool in = k >= klo && k < khi;
            const bool isd = in && dg <= 9u, iso = in && dg > 9u;
            const bool m = ch == 'M' || ch == '=', x = ch == 'X', ii = ch == 'I', dd = ch == 'D';
            bad = bad || (iso && (nd == 0u || nd > 8u || !(m || x || ii || dd)));
            am += (iso && m) ? val : 0u; ax += (iso && x) ? val : 0u;
            ai += (iso && ii) ? val : 0u; ad += (iso && dd) ? val : 0u;
            e_i += (iso && ii) ? 1u : 0u; e_d += (iso && dd) ? 1u : 0u;
            val = isd ? val * 10u + dg : (iso ? 0u : val);
            nd = isd ? nd + 1u : (iso ? 0u : nd);
        }
        if ((uint64_t)(idx0 + 16) >= len && nd != 0u) bad = true;  // the CIGAR ends inside a number
        c0 += am; c1 += ax; c2 += ai; c3 += ad;
    }
    bad = __any_sync(0xFFFFFFFFu, bad);
    unsigned long long ei = e_i, ed = e_d;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        c0 += __shfl_xor_sync(0xFFFFFFFFu, c0, d); c1 += __shfl_xor_sync(0xFFFFFFFFu, c1, d);
        c2 += __shfl_xor_sync(0xFFFFFFFFu, c2, d); c3 += __shfl_xor_sync(0xFFFFFFFFu, c3, d);
        ei += __shfl_xor_sync(0xFFFFFFFFu, ei, d); ed += __shfl_xor_sync(0xFFFFFFFFu, ed, d);
    }
    if (lane == 0) {
        slow[r] = bad ? 1 : 0;
        if (!bad) {
            const bool neg = negf == 1;
            unsigned long long *o = cnt11 + r * 11;
            o[0] = c0; o[1] = c1;
            o[2] = neg ? 0 : ei; o[3] = neg ? 0 : c2; o[4] = neg ? 0 : ed; o[5] = neg ? 0 : c3;
            o[6] = neg ? ei : 0; o[7] = neg ? c2 : 0; o[8] = neg ? ed : 0; o[9] = neg ? c3 : 0;
            o[10] = neg ? 1 : 0;
        }
    }
}

// ---- flattened CIGAR counters (the default): the 16-byte pieces of ALL records form one stream, lane i of a warp step
// owns piece g0 + i whatever record it belongs to, so short records do not leave lanes idle and neighbouring lanes read
// neighbouring memory.  k_cig_pieces counts the pieces of every record (>= 1: a record without a usable CIGAR keeps a
// dummy piece and goes to the exact kernel), an exclusive scan gives pbase[], k_cig_chunk_rec notes the record every
// chunk of kCigChunk pieces starts in, and k_paf_cigar_flat walks its chunk: the record of a lane falls out of one
// __reduce_or_sync over the next 32 record boundaries; inside a piece the lane loops over its op characters (tokens),
// not over its bytes: the digits in front of an op are parsed with two multiplies (<= 4 digits) and the digits a token
// has in the previous piece arrive by shuffle as before.  Sums go to acc6[r] = {M/=, X, I, D lengths, I, D events} with
// one atomic per non-zero value and run of lanes; k_cig_finish spreads them into the 11 counters by strand.
constexpr uint32_t kCigChunk = 512;  // pieces per warp (8 KB of CIGAR text)

__global__ void k_cig_pieces(uint64_t n_rec, const uint8_t *__restrict__ text, PafRecs R, uint64_t *__restrict__ nvec,
                             uint8_t *__restrict__ slow) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rec) return;
    if (r == n_rec) { nvec[r] = 0; return; }
    const uint64_t len = R.cig_len[r];
    if ((R.neg[r] & 2) || len == 0) { nvec[r] = 1; slow[r] = 1; return; }
    const uint64_t a = reinterpret_cast<uintptr_t>(text + R.cig_off[r]) & 15u;
    nvec[r] = (a + len + 15) >> 4;
    slow[r] = 0;
}
__global__ void k_cig_chunk_rec(uint64_t n_rec, const uint64_t *__restrict__ pbase, uint32_t *__restrict__ chunk_rec) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint64_t lo = pbase[r], hi = pbase[r + 1];
    for (uint64_t c = (lo + kCigChunk - 1) / kCigChunk; c * kCigChunk < hi; ++c) chunk_rec[c] = (uint32_t)r;
}

// digit run at the END of a piece: (value, digits); 9 digits = "longer than the 8-byte window" (flagged by the consumer)
__device__ __forceinline__ uint32_t cig_trailing(const uint4 &cur, uint32_t &t_val) {
    const unsigned long long F = ((unsigned long long)digit_flags(cur.w) << 32) | digit_flags(cur.z);
    const unsigned long long nF = ~F & 0x8080808080808080ull;
    uint32_t t_nd = nF ? (uint32_t)__clzll((long long)nF) >> 3 : 8u;
    if (t_nd == 8u && (digit_flags(cur.y) & 0x80000000u)) t_nd = 9u;
    t_val = 0;
    if (t_nd) {
        const unsigned long long H = ((unsigned long long)cur.w << 32) | cur.z;
        const unsigned long long keep = t_nd >= 8u ? ~0ull : ~0ull << (8u * (8u - t_nd));
        unsigned long long v = ((H & keep) | (0x3030303030303030ull & ~keep)) - 0x3030303030303030ull;
        v = v * 10 + (v >> 8);
        v = (((v & 0x000000FF000000FFull) * 0x000F424000000064ull) + (((v >> 16) & 0x000000FF000000FFull) * 0x0000271000000001ull)) >> 32;
        t_val = (uint32_t)v;
    }
    return t_nd;
}
__device__ __forceinline__ uint32_t movemask4(uint32_t hi_flags) {  // bit 7 of byte k -> bit k
    return (((hi_flags >> 7) & 0x01010101u) * 0x01020408u) >> 24;
}
__device__ __forceinline__ uint32_t sel4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t k) {
    return k == 0 ? w0 : k == 1 ? w1 : k == 2 ? w2 : k == 3 ? w3 : 0x3A3A3A3Au;
}

__global__ void __launch_bounds__(128) k_paf_cigar_flat(uint64_t n_rec, const uint8_t *__restrict__ text, PafRecs R,
                                                        const uint64_t *__restrict__ pbase, const uint32_t *__restrict__ chunk_rec,
                                                        unsigned long long *__restrict__ acc6, uint8_t *__restrict__ slow) {
    const uint64_t c = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t P = pbase[n_rec];
    uint64_t g0 = c * kCigChunk;
    if (g0 >= P) return;
    const uint64_t g_end = g0 + kCigChunk < P ? g0 + kCigChunk : P;
    uint64_t r0 = chunk_rec[c];
    uint32_t carry_val = 0, carry_nd = 0;  // trailing number of the piece before the step
    if (g0 > pbase[r0]) {  // the chunk starts inside record r0: every lane reads the piece before it (one address)
        const uint8_t *p = text + R.cig_off[r0];
        const uint8_t *base = p - (reinterpret_cast<uintptr_t>(p) & 15u);
        const uint4 prev = ldg16(base + 16 * (g0 - pbase[r0] - 1));
        carry_nd = cig_trailing(prev, carry_val);
    }
    for (; g0 < g_end; g0 += 32) {
        // ---- which record does piece g0 + lane belong to: boundaries pbase[r0 + 1 ..] that fall into (g0, g0 + 32]
        const uint64_t bi = r0 + 1 + lane;
        const uint64_t rel = (bi <= n_rec ? pbase[bi] : ~0ull) - g0;
        const uint32_t M = __reduce_or_sync(0xFFFFFFFFu, (rel >= 1 && rel <= 32) ? 1u << (uint32_t)(rel - 1) : 0u);
        const uint64_t r = r0 + __popc(M & ((1u << lane) - 1u));
        r0 += __popc(M);
        const uint64_t g = g0 + lane;
        const bool have = g < g_end;
        uint4 cur = make_uint4(0x3A3A3A3Au, 0x3A3A3A3Au, 0x3A3A3A3Au, 0x3A3A3A3Au);
        uint64_t vi = 0, len = 0, a = 0;
        bool dummy = true;
        if (have) {
            len = R.cig_len[r];
            dummy = (R.neg[r] & 2) || len == 0;
            if (!dummy) {
                const uint8_t *p = text + R.cig_off[r];
                a = reinterpret_cast<uintptr_t>(p) & 15u;
                vi = g - pbase[r];
                cur = ldg16(p - a + 16 * vi);
            }
        }
        // ---- phase 1: digit run at the end of the piece, handed to the right neighbour
        uint32_t t_val;
        const uint32_t t_nd = cig_trailing(cur, t_val);
        uint32_t val = __shfl_up_sync(0xFFFFFFFFu, t_val, 1), nd = __shfl_up_sync(0xFFFFFFFFu, t_nd, 1);
        if (lane == 0) { val = carry_val; nd = carry_nd; }
        carry_val = __shfl_sync(0xFFFFFFFFu, t_val, 31);
        carry_nd = __shfl_sync(0xFFFFFFFFu, t_nd, 31);
        uint32_t am = 0, ax = 0, ai = 0, ad = 0, e_i = 0, e_d = 0;
        bool bad = false;
        if (have && !dummy) {
            // ---- phase 2: the 16 bytes, branch-free and fully unrolled (CIGAR text runs ~2.5 bytes per token: a loop over
            //      the tokens diverges and costs ~100 instructions each, this costs ~20 per byte).  Bytes outside the CIGAR
            //      are replaced by 'M' (adds the running value 0 to the total); structure is checked on bit masks afterwards.
            const int64_t idx0 = (int64_t)(16 * vi) - (int64_t)a;  // CIGAR index of byte 0 of this piece
            const uint32_t klo = idx0 < 0 ? (uint32_t)-idx0 : 0u;
            const uint32_t khi = (uint64_t)(idx0 + 16) > len ? (uint32_t)((int64_t)len - idx0) : 16u;
            if (vi == 0) { val = 0; nd = 0; }  // nothing belongs to the CIGAR on the left of its first piece
            uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
            const uint32_t rm = ((1u << khi) - 1u) & ~((1u << klo) - 1u);
            if (rm != 0xFFFFu) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t in4 = (rm >> (4 * q)) & 0xFu;  // bit b: byte b of word q is inside the CIGAR
                    const uint32_t bm = ((in4 & 1u) * 0xFFu) | ((in4 & 2u) * 0x7F80u) | ((in4 & 4u) * 0x3FC000u) | ((in4 & 8u) * 0x1FE00000u);
                    w[q] = (w[q] & bm) | (0x4D4D4D4Du & ~bm);
                }
            }
            uint32_t tot = 0, ev = 0, nok = 0, v = val;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint32_t ch = __byte_perm(w[k >> 2], 0u, 0x4440u | (uint32_t)(k & 3));
                // predicated adds spelled in PTX (nvcc turns the C++ `if (c) x += v` into SEL + IADD, twice the instructions):
                //   any non-digit closes the running number into `tot`, X / I / D also into their own sums, I / D count events
                asm("{\n\t.reg .pred pd, px, pi, pe;\n\t.reg .u32 dg, ok;\n\t"
                    "sub.u32 dg, %7, 48;\n\t"
                    "setp.lt.u32 pd, dg, 10;\n\t"
                    "setp.eq.u32 px, %7, 88;\n\t"
                    "setp.eq.u32 pi, %7, 73;\n\t"
                    "setp.eq.u32 pe, %7, 68;\n\t"
                    "@!pd add.u32 %0, %0, %5;\n\t"
                    "@px add.u32 %1, %1, %5;\n\t"
                    "@pi add.u32 %2, %2, %5;\n\t"
                    "@pe add.u32 %3, %3, %5;\n\t"
                    "@pi add.u32 %4, %4, 1;\n\t"
                    "@pe add.u32 %4, %4, 65536;\n\t"
                    "sub.u32 ok, %7, 61;\n\t"           // '=' 'D' 'I' 'M' 'X' = 0x3D + {0, 7, 12, 16, 27}; shr clamps: other bytes give 0
                    "shr.u32 ok, 0x08011081, ok;\n\t"
                    "and.b32 ok, ok, 1;\n\t"
                    "add.u32 %6, %6, ok;\n\t"
                    "mad.lo.u32 %5, %5, 10, dg;\n\t"
                    "@!pd mov.u32 %5, 0;\n\t}"
                    : "+r"(tot), "+r"(ax), "+r"(ai), "+r"(ad), "+r"(ev), "+r"(v), "+r"(nok)
                    : "r"(ch));
            }
            am = tot - ax - ai - ad;
            e_i = ev & 0xFFFFu; e_d = ev >> 16;
            const uint32_t dmask = movemask4(digit_flags(w[0])) | (movemask4(digit_flags(w[1])) << 4) | (movemask4(digit_flags(w[2])) << 8) |
                                   (movemask4(digit_flags(w[3])) << 12);
            const uint32_t nd16 = ~dmask & 0xFFFFu;
            // every non-digit byte must be one of M = X I D (the bytes outside the CIGAR were replaced by 'M'); no op right
            // after a non-digit (no digits of its own: the byte before the CIGAR's first byte counts as one, the left
            // neighbour's last byte is a non-digit iff it hands over 0 digits)
            const uint32_t prev_nd = (nd16 << 1) | ((vi == 0 || nd == 0u) ? 1u : 0u);
            bad = nok != (uint32_t)__popc(nd16) || (nd16 & prev_nd & rm);
            // more than 8 digits: nine in a row inside the piece, or the run that continues the left neighbour's
            uint32_t r8 = dmask & (dmask >> 1);
            r8 &= r8 >> 2;
            r8 &= r8 >> 4;
            const uint32_t lead = (uint32_t)__ffs(~dmask) - 1u;  // digits at the start of the piece (dmask has 16 bits: < 32)
            bad = bad || (r8 & (dmask >> 8)) || (nd + lead > 8u && nd != 0u);
            // the CIGAR must not end in a digit
            if ((uint64_t)(idx0 + 16) >= len && ((dmask >> (khi - 1u)) & 1u)) bad = true;
        }
        // ---- per-record sums: lanes of one record are contiguous; one warp-uniform pass per run of the step, seven
        //      full-mask REDUX each.  A lane's sums are < 2^28 (two eight-digit tokens per piece at most): the low 24 bits of
        //      the four sums are added as they are (32 x 2^24 fits), the high 4 bits and the two event counts ride in
        //      10-bit fields (32 x 15 and 32 x 8 fit).
        uint32_t heads = (M << 1) | 1u;  // bit i: lane i starts a run
        const uint32_t hiA = (ax >> 24) | ((ai >> 24) << 10) | ((ad >> 24) << 20);
        const uint32_t hiB = (am >> 24) | (e_i << 10) | (e_d << 20);
        while (heads) {
            const uint32_t h = __ffs(heads) - 1;
            heads &= heads - 1;
            const uint32_t e = heads ? (uint32_t)__ffs(heads) - 1u : 32u;
            const bool in = lane >= h && lane < e;
            const uint32_t m_lo = __reduce_add_sync(0xFFFFFFFFu, in ? am & 0xFFFFFFu : 0u);
            const uint32_t x_lo = __reduce_add_sync(0xFFFFFFFFu, in ? ax & 0xFFFFFFu : 0u);
            const uint32_t i_lo = __reduce_add_sync(0xFFFFFFFFu, in ? ai & 0xFFFFFFu : 0u);
            const uint32_t d_lo = __reduce_add_sync(0xFFFFFFFFu, in ? ad & 0xFFFFFFu : 0u);
            const uint32_t sA = __reduce_add_sync(0xFFFFFFFFu, in ? hiA : 0u), sB = __reduce_add_sync(0xFFFFFFFFu, in ? hiB : 0u);
            const uint32_t s_bad = __reduce_or_sync(0xFFFFFFFFu, in && bad ? 1u : 0u);
            if (lane == h && have) {
                unsigned long long *o6 = acc6 + r * 6;
                const unsigned long long sm = ((unsigned long long)(sB & 0x3FFu) << 24) + m_lo, sx = ((unsigned long long)(sA & 0x3FFu) << 24) + x_lo,
                                         si = ((unsigned long long)((sA >> 10) & 0x3FFu) << 24) + i_lo,
                                         sd = ((unsigned long long)((sA >> 20) & 0x3FFu) << 24) + d_lo;
                const uint32_t s_ei = (sB >> 10) & 0x3FFu, s_ed = (sB >> 20) & 0x3FFu;
                if (sm) atomicAdd(o6 + 0, sm);
                if (sx) atomicAdd(o6 + 1, sx);
                if (si) atomicAdd(o6 + 2, si);
                if (sd) atomicAdd(o6 + 3, sd);
                if (s_ei) atomicAdd(o6 + 4, (unsigned long long)s_ei);
                if (s_ed) atomicAdd(o6 + 5, (unsigned long long)s_ed);
                if (s_bad) slow[r] = 1;
            }
        }
    }
}

// acc6 -> the 11 counters of struct Cigar by strand (records flagged slow are redone by k_paf_cigar_counts)
__global__ void k_cig_finish(uint64_t n_rec, PafRecs R, const unsigned long long *__restrict__ acc6, const uint8_t *__restrict__ slow,
                             unsigned long long *__restrict__ cnt11) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec || slow[r]) return;
    const unsigned long long *a = acc6 + r * 6;
    const bool neg = (R.neg[r] & 1) != 0;
    unsigned long long *o = cnt11 + r * 11;
    o[0] = a[0]; o[1] = a[1];
    o[2] = neg ? 0 : a[4]; o[3] = neg ? 0 : a[2]; o[4] = neg ? 0 : a[5]; o[5] = neg ? 0 : a[3];
    o[6] = neg ? a[4] : 0; o[7] = neg ? a[2] : 0; o[8] = neg ? a[5] : 0; o[9] = neg ? a[3] : 0;
    o[10] = neg ? 1 : 0;
}

// The same warp-per-record scan producing the ordered OP TABLE of every record (paf2chain; the op tables of paf2maf /
// call -f paf / pafpseudo still use their one-thread kernels).  PASS 0: ops per record + the slow flag (malformed or
// 9+ digit lengths: redone by the exact kernel, which also raises the reference's error); PASS 1: (length, type) of op
// k at op_base[r] + k, type 0 = M/=/X, 1 = I, 2 = D.
template <int PASS>
__global__ void __launch_bounds__(128, 8) k_paf_ops_warp(uint64_t n_rec, const uint8_t *__restrict__ text, PafRecs R,
                                                         uint64_t *__restrict__ n_ops, uint8_t *__restrict__ slow,
                                                         const uint64_t *__restrict__ op_base, uint64_t *__restrict__ op_len,
                                                         uint8_t *__restrict__ op_type) {
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (r >= n_rec) return;
    const uint8_t negf = R.neg[r];
    const uint64_t len = R.cig_len[r];
    if ((negf & 2) || len == 0) { if (PASS == 0 && lane == 0) { slow[r] = 1; n_ops[r] = 0; } return; }
    if (PASS == 1 && slow[r]) return;
    const uint8_t *p = text + R.cig_off[r];
    const uint64_t a = reinterpret_cast<uintptr_t>(p) & 15u;
    const uint8_t *base = p - a;
    const uint64_t nvec = (a + len + 15) >> 4;
    bool bad = false;
    uint64_t done = 0;  // ops of the previous steps
    uint32_t carry_val = 0, carry_nd = 0;
    for (uint64_t v0 = 0; v0 < nvec; v0 += 32) {
        const uint64_t vi = v0 + lane;
        const bool have = vi < nvec;
        uint4 cur = make_uint4(0x3A3A3A3Au, 0x3A3A3A3Au, 0x3A3A3A3Au, 0x3A3A3A3Au);
        if (have) cur = ldg16(base + 16 * vi);
        uint32_t t_nd, t_val = 0;
        {
            const unsigned long long F = ((unsigned long long)digit_flags(cur.w) << 32) | digit_flags(cur.z);
            const unsigned long long nF = ~F & 0x8080808080808080ull;
            t_nd = nF ? (uint32_t)__clzll((long long)nF) >> 3 : 8u;
            if (t_nd == 8u && (digit_flags(cur.y) & 0x80000000u)) t_nd = 9u;
            if (PASS == 1 && t_nd) {
                const unsigned long long H = ((unsigned long long)cur.w << 32) | cur.z;
                const unsigned long long keep = t_nd >= 8u ? ~0ull : ~0ull << (8u * (8u - t_nd));
                unsigned long long v = ((H & keep) | (0x3030303030303030ull & ~keep)) - 0x3030303030303030ull;
                v = v * 10 + (v >> 8);
                v = (((v & 0x000000FF000000FFull) * 0x000F424000000064ull) + (((v >> 16) & 0x000000FF000000FFull) * 0x0000271000000001ull)) >> 32;
                t_val = (uint32_t)v;
            }
        }
        uint32_t val = __shfl_up_sync(0xFFFFFFFFu, t_val, 1), nd = __shfl_up_sync(0xFFFFFFFFu, t_nd, 1);
        if (lane == 0) { val = carry_val; nd = carry_nd; }
        carry_val = __shfl_sync(0xFFFFFFFFu, t_val, 31);
        carry_nd = __shfl_sync(0xFFFFFFFFu, t_nd, 31);
        const int64_t idx0 = (int64_t)(16 * vi) - (int64_t)a;
        const int klo = idx0 < 0 ? (int)-idx0 : 0;
        const int khi = !have ? 0 : (uint64_t)(idx0 + 16) > len ? (int)((int64_t)len - idx0) : 16;
        if (idx0 <= 0) { val = 0; nd = 0; }
        const uint32_t cw[4] = {cur.x, cur.y, cur.z, cur.w};
        // ops in this piece -> position of the piece's first op inside the record
        uint32_t mine = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t ch = (cw[k >> 2] >> (8 * (k & 3))) & 0xFFu;
            mine += (k >= klo && k < khi && (ch - '0') > 9u) ? 1u : 0u;
        }
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += y;
        }
        const uint32_t step_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        uint64_t o = (PASS == 1 ? op_base[r] : 0) + done + (incl - mine);
        done += step_total;
        if (!have) continue;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t ch = (cw[k >> 2] >> (8 * (k & 3))) & 0xFFu;
            const uint32_t dg = ch - '0';
            const bool in = k >= klo && k < khi;
            const bool isd = in && dg <= 9u, iso = in && dg > 9u;
            const bool m = ch == 'M' || ch == '=' || ch == 'X', ii = ch == 'I', dd = ch == 'D';
            if (PASS == 0) bad = bad || (iso && (nd == 0u || nd > 8u || !(m || ii || dd)));
            if (PASS == 1 && iso) {
                op_len[o] = val;
                op_type[o] = ii ? 1 : dd ? 2 : 0;
                ++o;
            }
            val = isd ? val * 10u + dg : (iso ? 0u : val);
            nd = isd ? nd + 1u : (iso ? 0u : nd);
        }
        if (PASS == 0 && (uint64_t)(idx0 + 16) >= len && nd != 0u) bad = true;
    }
    if (PASS == 0) {
        bad = __any_sync(0xFFFFFFFFu, bad);
        if (lane == 0) { slow[r] = bad ? 1 : 0; n_ops[r] = bad ? 0 : done; }
    }
}

// Exact path, one record per thread: the tokenizer of cigar.rs:43-75 with the reference's errors.  only != nullptr:
// just the records the warp kernel flagged.
__global__ void k_paf_cigar_counts(uint64_t n_rec, const uint8_t *__restrict__ text, PafRecs R,
                                   unsigned long long *__restrict__ cnt11, unsigned long long *err_key,
                                   const uint8_t *__restrict__ only) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    if (only && !only[r]) return;
    unsigned long long *o = cnt11 + r * 11;
    unsigned long long c_m = 0, c_x = 0, c_ie = 0, c_i = 0, c_de = 0, c_d = 0;
    const uint8_t negf = R.neg[r];
    const uint64_t len = R.cig_len[r];
    const uint8_t *p = text + R.cig_off[r];
    if (negf & 2) { for (int i = 0; i < 11; ++i) o[i] = 0; return; }
    if (len == 0) {  // fold_many1 on "" -> nom error -> e.input[..10] panics (errors.rs:88-97)
        paf_report(err_key, r, 0x800001, WGA_E_REF_PANIC);
        for (int i = 0; i < 11; ++i) o[i] = 0;
        return;
    }
    uint64_t val = 0;
    uint32_t nd = 0;        // digits in the current number
    bool ovf = false;
    uint32_t op_run = 0;    // non-digit chars since the last digit
    uint8_t op = 0;
    int code = 0;
    uint64_t epos = 0;
    // tokens: digits* then nondigits*; a token is closed when a digit follows a non-digit or at the end
    auto close_token = [&](uint64_t pos) {
        if (code) return;
        if (op_run != 1) { code = WGA_E_CIGAR_OP_INVALID; epos = pos; return; }     // cst2cu: op must be one char
        if (nd == 0 || ovf) { code = WGA_E_PARSE_INT; epos = pos; return; }          // parse_str2u64
        switch (op) {
            case 'M': case '=': c_m += val; break;
            case 'X': c_x += val; break;
            case 'I': c_ie += 1; c_i += val; break;
            case 'D': c_de += 1; c_d += val; break;
            default: code = WGA_E_CIGAR_OP_INVALID; epos = pos; break;
        }
    };
    const uint64_t a = reinterpret_cast<uintptr_t>(p) & 15u;
    const uint8_t *base = p - a;
    const uint64_t nvec = (a + len + 15) >> 4;
    bool in_op = false;
    for (uint64_t vi = 0; vi < nvec && !code; ++vi) {
        uint4 vv = ldg16(base + 16 * vi);
        uint32_t w[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int64_t idx = (int64_t)(16 * vi + k) - (int64_t)a;
            if (idx < 0 || (uint64_t)idx >= len) continue;
            const uint8_t ch = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
            const bool dig = ch >= '0' && ch <= '9';
            if (dig) {
                if (in_op) {  // a new token starts: close the previous one
                    close_token((uint64_t)idx);
                    val = 0; nd = 0; ovf = false; op_run = 0; in_op = false;
                }
                const uint32_t d = ch - '0';
                if (val > (0xFFFFFFFFFFFFFFFFull - d) / 10) ovf = true;
                val = val * 10 + d;
                ++nd;
            } else {
                in_op = true;
                ++op_run;
                op = ch;
            }
        }
    }
    if (!code) close_token(len);  // last token (digits with no op char -> op_run == 0 -> CigarOpInvalid(""))
    if (code) paf_report(err_key, r, 0x800002 + (epos > 0x7FFFF0ull ? 0x7FFFF0ull : epos), code);
    const bool neg = negf == 1;
    o[0] = c_m; o[1] = c_x;
    o[2] = neg ? 0 : c_ie; o[3] = neg ? 0 : c_i; o[4] = neg ? 0 : c_de; o[5] = neg ? 0 : c_d;
    o[6] = neg ? c_ie : 0; o[7] = neg ? c_i : 0; o[8] = neg ? c_de : 0; o[9] = neg ? c_d : 0;
    o[10] = neg ? 1 : 0;
}

}  // namespace wga
