// chain_in.cuh -- on-device UCSC chain reader + chain2paf rows + the op table chain2maf feeds to the paf2maf expansion
// (the chain -> PAF / MAF direction).
//
// Replaces src/parser/chain.rs:27-75 (ChainReader / ChainRecords), :206-383 (parse_header, line_not_chain,
// parse_line_to_cdl, parse_chain_data_line, chain_parser), :430-452 (ChainRecord::convert2paf),
// src/parser/cigar.rs:554-626 (parse_chain_to_cigar), src/converter.rs:360-388 (parse_chain_to_insert),
// :391-416 (chain2paf).
//
// The reference applies one nom parser to "what is left of the file", record after record.  Its combinators reduce to
// a rule that is LOCAL per line, which is what makes the reader parallel:
//   * every line that holds a 'c' must START with "chain" -- it is a header (tag("chain")), anything else ends in a
//     nom error whichever state the sequential parser is in (is_not("chain\n") stops at the letter, take_while stops at
//     the first 'c', and tag("chain") fails there);
//   * the data lines of a record are the maximal run of lines right after its header that are non-empty, hold none of
//     the letters c,h,a,i,n and END IN A NEWLINE (terminated(is_not("chain\n"), line_ending)); fold_many1 needs one;
//   * everything between that run and the next header is skipped by take_while(|x| x != 'c') -- blank lines, comments
//     without a 'c', and a last data line that lacks its newline (silently dropped by the reference, and here);
//   * the file must start with a header (no skipping before the first record).
// A data line is split_whitespace(): size [query_diff [target_diff]] (chain.rs:333-350: in THAT order).
// Errors are keyed (record << 32 | stage << 8 | code) so that the smallest key is the error the sequential parser
// meets first.  One deviation: a line that holds "chain" somewhere after other text (the reference would parse a header
// out of the middle of that line) is refused with WGA_E_UNSUPPORTED.
#pragma once
#include "ctx.cuh"
#include "emit_paf.cuh"
#include "paf2maf.cuh"
#include "tokenize.cuh"

namespace wga {

enum : uint8_t { CHL_OTHER = 0, CHL_HEADER = 1, CHL_DATA = 2, CHL_BAD = 3 };

__device__ __forceinline__ void chain_report(unsigned long long *err_key, uint64_t rec, uint64_t stage, int code) {
    if (stage > 0xFFFFFFull) stage = 0xFFFFFFull;
    atomicMin(err_key, (unsigned long long)((rec << 32) | (stage << 8) | (uint64_t)code));
}
// nom errors are converted with e.input[..10] (errors.rs:88-97): fewer than 10 bytes left = a panic instead
__device__ __forceinline__ int nom_code(uint64_t rest) { return rest < 10 ? WGA_E_REF_PANIC : WGA_E_NOM; }

// ---- 1. line classes
__global__ void k_chain_classify(LineView lv, uint8_t *__restrict__ cls, uint32_t *__restrict__ is_hdr,
                                 uint32_t *__restrict__ nonvalid, uint64_t *__restrict__ first_c) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > lv.n_lines) return;
    if (k == lv.n_lines) { is_hdr[k] = 0; nonvalid[k] = 0; return; }
    const uint64_t S = lv.start(k), E = lv.end(k);
    const uint8_t *p = lv.text;
    uint64_t fc = ~0ull;
    bool letters = false;
    for (uint64_t i = S; i < E; ++i) {
        const uint8_t c = p[i];
        if (c == 'c') { if (fc == ~0ull) fc = i; letters = true; }
        else if (c == 'h' || c == 'a' || c == 'i' || c == 'n') letters = true;
    }
    uint8_t c = CHL_OTHER;
    if (E - S >= 5 && p[S] == 'c' && p[S + 1] == 'h' && p[S + 2] == 'a' && p[S + 3] == 'i' && p[S + 4] == 'n') c = CHL_HEADER;
    else if (fc != ~0ull) c = CHL_BAD;
    else if (E > S && !letters && k < lv.n_nl) c = CHL_DATA;
    cls[k] = c;
    first_c[k] = fc;
    is_hdr[k] = c == CHL_HEADER;
    nonvalid[k] = c != CHL_DATA;
}

// header line list and non-data line list (sorted by construction)
__global__ void k_chain_lists(uint64_t n_lines, const uint8_t *__restrict__ cls, const uint32_t *__restrict__ hdr_scan,
                              const uint32_t *__restrict__ nv_scan, uint64_t *__restrict__ hdr_line, uint64_t *__restrict__ nv_line) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_lines) return;
    const uint8_t c = cls[k];
    if (c == CHL_HEADER) hdr_line[hdr_scan[k]] = k;
    if (c != CHL_DATA) nv_line[nv_scan[k]] = k;
}

// Rust's str::parse::<f64>(): [+-] (digits [. digits*] | . digits+) [(e|E) [+-] digits+] | [+-] (inf | infinity | nan), words
// case-insensitive.  Only validity matters (the score is never printed).
__device__ __forceinline__ bool rust_f64_ok(const uint8_t *s, uint64_t n) {
    uint64_t i = 0;
    if (n == 0) return false;
    if (s[0] == '+' || s[0] == '-') i = 1;
    if (i == n) return false;
    const uint64_t m = n - i;
    if (m == 3 || m == 8) {
        const char *w3a = "inf", *w3b = "nan", *w8 = "infinity";
        bool a = m == 3, b = m == 3, c = m == 8;
        for (uint64_t k = 0; k < m; ++k) {
            const uint8_t ch = s[i + k] | 0x20u;
            if (m == 3) { a &= ch == (uint8_t)w3a[k]; b &= ch == (uint8_t)w3b[k]; }
            else c &= ch == (uint8_t)w8[k];
        }
        if (a || b || c) return true;
    }
    uint64_t nd = 0;
    while (i < n && s[i] >= '0' && s[i] <= '9') { ++i; ++nd; }
    if (i < n && s[i] == '.') { ++i; while (i < n && s[i] >= '0' && s[i] <= '9') { ++i; ++nd; } }
    if (nd == 0) return false;
    if (i < n && (s[i] == 'e' || s[i] == 'E')) {
        ++i;
        if (i < n && (s[i] == '+' || s[i] == '-')) ++i;
        uint64_t ne = 0;
        while (i < n && s[i] >= '0' && s[i] <= '9') { ++i; ++ne; }
        if (ne == 0) return false;
    }
    return i == n;
}

// whitespace-separated tokens of text[from, E) of line k, at most `cap`; returns the count
__device__ __forceinline__ int line_tokens(const LineView &lv, uint64_t k, uint64_t from, uint64_t E, uint64_t *tb, uint64_t *te, int cap) {
    const uint64_t wlo = k == 0 ? 0 : lv.nl_ws_idx[k - 1] + 1, whi = k < lv.n_nl ? lv.nl_ws_idx[k] : lv.n_ws;
    int nt = 0;
    uint64_t p = from;
    for (uint64_t w = wlo; w <= whi && nt < cap; ++w) {
        const uint64_t x = w < whi ? (lv.ws_ent[w] & kPosMask) : E;
        if (x < from) continue;
        if (x > p) { tb[nt] = p; te[nt] = x; ++nt; }
        p = x + 1;
    }
    return nt;
}

struct ChainRecs {  // per chain record (indexed like PafRecs so that the paf2maf expansion can run on it)
    PafRecs R;       // names, sizes, starts, neg (bit0 = query '-'); cig_* unused
    uint64_t *t_end, *q_end;
    uint8_t *t_neg;
    uint64_t *first_data, *n_data;  // first data line (line index) and number of data lines
};

// ---- 2. headers (one thread per record): parse_header (chain.rs:206-315), line_ending, fold_many1's "at least one"
__global__ void k_chain_headers(LineView lv, uint64_t n_rec, const uint64_t *__restrict__ hdr_line, const uint32_t *__restrict__ nv_scan,
                                const uint64_t *__restrict__ nv_line, uint64_t n_nv, ChainRecs C, unsigned long long *err_key) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint64_t h = hdr_line[r];
    const uint64_t S = lv.start(h), E = lv.end(h);
    uint64_t tb[12], te[12];
    const int nt = line_tokens(lv, h, S + 5, E, tb, te, 12);
    uint64_t u[12] = {0};
    int code = 0;
    uint8_t tneg = 0, qneg = 0;
    for (int k = 0; k < 12 && !code; ++k) {
        if (k >= nt) { code = WGA_E_CHAIN_FIELD_MISSING; break; }
        const uint8_t *s = lv.text + tb[k];
        const uint64_t n = te[k] - tb[k];
        if (k == 0) { if (!rust_f64_ok(s, n)) code = WGA_E_PARSE_FLOAT; }
        else if (k == 3 || k == 8) {
            if (n == 1 && s[0] == '+') { }
            else if (n == 1 && s[0] == '-') { if (k == 3) tneg = 1; else qneg = 1; }
            else code = WGA_E_PARSE_STRAND;
        } else if (k != 1 && k != 6) { if (!parse_u64_bytes(s, n, &u[k])) code = WGA_E_PARSE_INT; }
    }
    // data lines: up to the next line that is not one
    const uint64_t j = (uint64_t)nv_scan[h] + 1;  // h itself is entry nv_scan[h] of the non-data list
    const uint64_t next = j < n_nv ? nv_line[j] : lv.n_lines;
    const uint64_t nd = next - h - 1;
    C.first_data[r] = h + 1;
    C.n_data[r] = code ? 0 : nd;
    if (code) chain_report(err_key, r, 1, code);
    else if (h >= lv.n_nl) chain_report(err_key, r, 2, WGA_E_REF_PANIC);  // line_ending on an empty rest (CrLf, then input[..10])
    else if (nd == 0) chain_report(err_key, r, 2, nom_code(h + 1 < lv.n_lines ? lv.len - lv.start(h + 1) : 0));  // Many1
    PafRecs &R = C.R;
    const bool ok = code == 0;
    R.t_name_off[r] = ok ? tb[1] : 0; R.t_name_len[r] = ok ? (uint32_t)(te[1] - tb[1]) : 0;
    R.q_name_off[r] = ok ? tb[6] : 0; R.q_name_len[r] = ok ? (uint32_t)(te[6] - tb[6]) : 0;
    R.t_size[r] = u[2]; R.t_start[r] = u[4]; C.t_end[r] = u[5];
    R.q_size[r] = u[7]; R.q_start[r] = u[9]; C.q_end[r] = u[10];
    R.neg[r] = qneg; C.t_neg[r] = tneg;
    R.cig_off[r] = 0; R.cig_len[r] = 0;
}

// ---- 3. data lines + stray lines (one thread per line)
// v[3k..3k+2] = size, query_diff, target_diff; dl_cig[k] = bytes of "<size>M[<t>I][<q>D]"; dl_ops[k] = ops of the line
__global__ void k_chain_data(LineView lv, const uint8_t *__restrict__ cls, const uint32_t *__restrict__ hdr_scan,
                             const uint32_t *__restrict__ nv_scan, const uint64_t *__restrict__ hdr_line,
                             const uint64_t *__restrict__ first_c, uint64_t *__restrict__ v, uint32_t *__restrict__ dl_cig,
                             uint32_t *__restrict__ dl_ops, unsigned long long *err_key) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > lv.n_lines) return;
    if (k == lv.n_lines) { dl_cig[k] = 0; dl_ops[k] = 0; v[3 * k] = v[3 * k + 1] = v[3 * k + 2] = 0; return; }
    const uint8_t c = cls[k];
    uint64_t size = 0, qd = 0, td = 0;
    uint32_t cig = 0, ops = 0;
    const uint64_t n_hdr_before = hdr_scan[k];  // headers strictly before line k
    if (k == 0 && c != CHL_HEADER) {
        chain_report(err_key, 0, 0, nom_code(lv.len));  // tag("chain") at the very start of the file
    } else if (c == CHL_BAD && k > 0 && cls[k - 1] != CHL_HEADER) {
        // reached by take_while(|x| x != 'c') after record n_hdr_before-1: tag("chain") fails at the first 'c' ...
        const uint64_t fc = first_c[k];
        const uint8_t *p = lv.text + fc;
        const bool mid_chain = lv.len - fc >= 5 && p[1] == 'h' && p[2] == 'a' && p[3] == 'i' && p[4] == 'n';
        chain_report(err_key, n_hdr_before, 0, mid_chain ? WGA_E_UNSUPPORTED : nom_code(lv.len - fc));  // ... unless "chain" starts there
    } else if (c == CHL_DATA && n_hdr_before > 0) {
        const uint64_t r = n_hdr_before - 1, h = hdr_line[r];
        if ((uint64_t)nv_scan[k] == (uint64_t)nv_scan[h] + 1) {  // no non-data line between the header and this line
            uint64_t tb[3], te[3];
            const int nt = line_tokens(lv, k, lv.start(k), lv.end(k), tb, te, 3);
            int code = 0;
            if (nt < 1) code = WGA_E_CHAIN_FIELD_MISSING;
            else if (!parse_u64_bytes(lv.text + tb[0], te[0] - tb[0], &size)) code = WGA_E_PARSE_INT;
            else if (nt >= 2 && !parse_u64_bytes(lv.text + tb[1], te[1] - tb[1], &qd)) code = WGA_E_PARSE_INT;
            else if (nt >= 3 && !parse_u64_bytes(lv.text + tb[2], te[2] - tb[2], &td)) code = WGA_E_PARSE_INT;
            if (code) {
                chain_report(err_key, r, 3 + (k - h), code);
                size = qd = td = 0;
            } else {
                cig = ndigits_u64(size) + 1 + (td ? ndigits_u64(td) + 1 : 0) + (qd ? ndigits_u64(qd) + 1 : 0);
                ops = 1 + (td ? 1 : 0) + (qd ? 1 : 0);
            }
        }
    }
    v[3 * k] = size; v[3 * k + 1] = qd; v[3 * k + 2] = td;
    dl_cig[k] = cig;
    dl_ops[k] = ops;
}

__global__ void k_chain_split3(uint64_t n, const uint64_t *__restrict__ v, uint64_t *__restrict__ a, uint64_t *__restrict__ b) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= n) { a[k] = v[3 * k]; b[k] = v[3 * k + 1]; }
}

// ---- 4. chain2paf rows (chain.rs:430-451 + the csv writer of converter.rs:396-414): 12 columns + "cg:Z:<cigar>"
struct C2pView {
    uint64_t n_rec;
    const uint8_t *text;
    ChainRecs C;
    const uint64_t *S_size, *S_qd;   // exclusive scans over lines of size / query_diff
    const uint64_t *S_cig;           // exclusive scan of dl_cig
    const uint64_t *v;
    uint64_t *row_len;
    uint32_t *row_fixed;
    const uint64_t *row_off;
    char *out;
};
template <int PASS>
__global__ void k_c2p_rows(C2pView V) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > V.n_rec) return;
    if (r == V.n_rec) { if (PASS == 0) V.row_len[r] = 0; return; }
    const PafRecs &R = V.C.R;
    const uint64_t l0 = V.C.first_data[r], l1 = l0 + V.C.n_data[r];
    const uint64_t match = V.S_size[l1] - V.S_size[l0], del = V.S_qd[l1] - V.S_qd[l0], cig = V.S_cig[l1] - V.S_cig[l0];
    const uint8_t *qn = V.text + R.q_name_off[r], *tn = V.text + R.t_name_off[r];
    if (PASS == 0) {
        const uint32_t fixed = csv_name_len(qn, R.q_name_len[r]) + 1 + ndigits_u64(R.q_size[r]) + 1 + ndigits_u64(R.q_start[r]) + 1 +
                               ndigits_u64(V.C.q_end[r]) + 3 + csv_name_len(tn, R.t_name_len[r]) + 1 + ndigits_u64(R.t_size[r]) + 1 +
                               ndigits_u64(R.t_start[r]) + 1 + ndigits_u64(V.C.t_end[r]) + 1 + ndigits_u64(match) + 1 +
                               ndigits_u64(match + del) + 1 + 3 + 1 + 5;
        V.row_fixed[r] = fixed;
        V.row_len[r] = (uint64_t)fixed + cig + 1;
    } else {
        char *o = V.out + V.row_off[r];
        o = csv_name_put(o, qn, R.q_name_len[r]); *o++ = '\t';
        o = put_u64(o, R.q_size[r]); *o++ = '\t'; o = put_u64(o, R.q_start[r]); *o++ = '\t'; o = put_u64(o, V.C.q_end[r]); *o++ = '\t';
        *o++ = (R.neg[r] & 1) ? '-' : '+'; *o++ = '\t';
        o = csv_name_put(o, tn, R.t_name_len[r]); *o++ = '\t';
        o = put_u64(o, R.t_size[r]); *o++ = '\t'; o = put_u64(o, R.t_start[r]); *o++ = '\t'; o = put_u64(o, V.C.t_end[r]); *o++ = '\t';
        o = put_u64(o, match); *o++ = '\t'; o = put_u64(o, match + del); *o++ = '\t';
        *o++ = '2'; *o++ = '5'; *o++ = '5'; *o++ = '\t';
        *o++ = 'c'; *o++ = 'g'; *o++ = ':'; *o++ = 'Z'; *o++ = ':';
        o[cig] = '\n';
    }
}
// the CIGAR tokens, one thread per data line: "<size>M" then "<target_diff>I" then "<query_diff>D" (cigar.rs:582-611)
__global__ void k_c2p_cigar(uint64_t n_lines, const uint32_t *__restrict__ dl_cig, const uint32_t *__restrict__ hdr_scan, C2pView V) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_lines || dl_cig[k] == 0) return;
    const uint64_t r = (uint64_t)hdr_scan[k] - 1;
    char *o = V.out + V.row_off[r] + V.row_fixed[r] + (V.S_cig[k] - V.S_cig[V.C.first_data[r]]);
    const uint64_t size = V.v[3 * k], qd = V.v[3 * k + 1], td = V.v[3 * k + 2];
    o = put_u64(o, size); *o++ = 'M';
    if (td) { o = put_u64(o, td); *o++ = 'I'; }
    if (qd) { o = put_u64(o, qd); *o++ = 'D'; }
}

// ---- 5. chain2maf: FASTA spans and the op table of the paf2maf expansion (converter.rs:268-388)
__global__ void k_c2m_spans(uint64_t n_rec, const uint8_t *__restrict__ text, ChainRecs C, P2mRecs X, FaView T, FaView Q,
                            const uint64_t *__restrict__ S_ops, uint64_t *__restrict__ op_base, unsigned long long *be_key) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rec) return;
    if (r == n_rec) { op_base[r] = S_ops[C.first_data[n_rec - 1] + C.n_data[n_rec - 1]]; return; }
    const PafRecs &R = C.R;
    op_base[r] = S_ops[C.first_data[r]];
    X.q_end[r] = C.q_end[r]; X.t_end[r] = C.t_end[r]; X.mapq[r] = 255;
    X.LT[r] = X.LQ[r] = X.t_src[r] = X.q_src[r] = 0;
    const uint64_t ti = fa_find(T, text + R.t_name_off[r], R.t_name_len[r]);
    const uint64_t qi = ti == ~0ull ? ~0ull : fa_find(Q, text + R.q_name_off[r], R.q_name_len[r]);
    if (ti == ~0ull || qi == ~0ull) { paf_report(be_key, r, 1, WGA_E_NOM); return; }  // faidx: sequence not found
    uint64_t first, count;
    fa_clip(T.meta[ti * kFaMeta + 4], R.t_start[r], C.t_end[r] - 1, &first, &count);  // u64 wrap on 0 - 1 as in release Rust
    X.t_src[r] = T.meta[ti * kFaMeta + 3] + first;
    X.LT[r] = count;
    fa_clip(Q.meta[qi * kFaMeta + 4], R.q_start[r], C.q_end[r] - 1, &first, &count);
    X.q_src[r] = Q.meta[qi * kFaMeta + 3] + first;
    X.LQ[r] = count;
}
// ops of one data line: M(size) always, I(target_diff) and D(query_diff) when non-zero -- the order parse_chain_to_insert
// applies them in (converter.rs:366-386)
__global__ void k_c2m_ops(uint64_t n_lines, const uint32_t *__restrict__ dl_ops, const uint32_t *__restrict__ hdr_scan,
                          const uint64_t *__restrict__ S_ops, const uint64_t *__restrict__ v, uint64_t *__restrict__ op_len,
                          uint8_t *__restrict__ op_type, uint32_t *__restrict__ op_rec, uint32_t *__restrict__ op_pos) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_lines || dl_ops[k] == 0) return;
    const uint32_t r = hdr_scan[k] - 1;
    uint64_t o = S_ops[k];
    const uint64_t size = v[3 * k], qd = v[3 * k + 1], td = v[3 * k + 2];
    op_len[o] = size; op_type[o] = 0; op_rec[o] = r; op_pos[o] = 0; ++o;
    if (td) { op_len[o] = td; op_type[o] = 1; op_rec[o] = r; op_pos[o] = 0; ++o; }
    if (qd) { op_len[o] = qd; op_type[o] = 2; op_rec[o] = r; op_pos[o] = 0; }
}

}  // namespace wga
