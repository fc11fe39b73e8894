// validate.cuh -- `validate [-f fixed.paf]`: expected query/target ends from the CIGAR counters (SURVEY 8(f) item 3,
// a free rider on the PAF CIGAR kernel).
//
// Replaces src/tools/validate.rs:44-147: exp_query_end = query_start + matched + mismatched + ins + inv_ins,
// exp_target_end = target_start + matched + mismatched + del + inv_del; a record whose end differs is counted,
// listed as "{name}:{start}-{end}" and (with --fix) re-serialised with the expected end.  The reference folds the
// records with rayon in unspecified order; the lists and the fixed records come in FILE order here.
// `rec.get_stat().unwrap()` (validate.rs:79) makes every CIGAR / missing-tag error a panic -> WGA_E_REF_PANIC.
#pragma once
#include "ctx.cuh"
#include "emit_paf.cuh"
#include "paf.cuh"
#include "paf2maf.cuh"

namespace wga {

struct ValCols {  // typed columns the stat path does not keep
    uint64_t *q_end, *t_end, *matches, *block_len, *mapq;
    uint64_t *tags_off;   // first byte after the 12th field's tab; line end + 1 when the record has no tags
    uint64_t *line_end;
};

__global__ void k_val_cols(LineView lv, const uint8_t *__restrict__ flags, const uint32_t *__restrict__ rec_scan, ValCols C) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= lv.n_lines || !flags[k]) return;
    const uint64_t r = rec_scan[k];
    uint64_t S = lv.start(k), E = lv.end(k);
    if (E > S && lv.text[E - 1] == '\r') --E;
    uint64_t wlo = k == 0 ? 0 : lv.nl_ws_idx[k - 1] + 1, whi = k < lv.n_nl ? lv.nl_ws_idx[k] : lv.n_ws;
    uint64_t fb = S, w = wlo, v[12] = {0};
    uint64_t tags = E + 1;
    for (uint64_t nf = 0; nf < 12; ++nf) {
        uint64_t fe = w < whi ? (lv.ws_ent[w] & kPosMask) : E;
        if (fe > E) fe = E;
        if (nf == 3 || nf >= 8) csv_u64(lv.text + fb, fe - fb, &v[nf]);
        if (w >= whi) break;
        fb = fe + 1;
        ++w;
        if (nf == 11) tags = fb;
    }
    C.q_end[r] = v[3]; C.t_end[r] = v[8]; C.matches[r] = v[9]; C.block_len[r] = v[10]; C.mapq[r] = v[11];
    C.tags_off[r] = tags;
    C.line_end[r] = E;
}

struct ValView {
    uint64_t n_rec;
    const uint8_t *text;
    PafRecs R;
    ValCols C;
    const unsigned long long *cnt11;
    int fix;
    uint64_t *q_len, *t_len, *row_len;          // [n_rec+1] bytes of the record's list lines / fixed row
    const uint64_t *q_off, *t_off, *row_off;
    char *q_out, *t_out, *row_out;
    unsigned long long *n_bad;                  // [2]: query / target invalid counts
};

__device__ __forceinline__ uint32_t tags_csv_len(const uint8_t *p, uint64_t n) {  // fields separated by tabs, csv-quoted
    uint64_t b = 0, len = 0;
    for (uint64_t i = 0; i <= n; ++i)
        if (i == n || p[i] == '\t') { len += csv_name_len(p + b, (uint32_t)(i - b)) + 1; b = i + 1; }
    return (uint32_t)len;  // every field is preceded by one tab
}
__device__ __forceinline__ char *tags_csv_put(char *o, const uint8_t *p, uint64_t n) {
    uint64_t b = 0;
    for (uint64_t i = 0; i <= n; ++i)
        if (i == n || p[i] == '\t') { *o++ = '\t'; o = csv_name_put(o, p + b, (uint32_t)(i - b)); b = i + 1; }
    return o;
}

// ---- one PAF row by one WARP (csv serialisation of a PafRecord: 12 typed columns, then the tags).  Lane k < 12 owns
// column k: it sizes its field, a warp scan places it, the lane writes it; the tags -- a CIGAR, i.e. nearly all of the
// row -- are counted (quotes) / copied by the whole warp with aligned word stores.  Tags that hold a '"' take lane 0's
// exact csv path.  Returns the row's byte length; PASS 1 also writes it to `out`.
struct PafRowIn {
    const uint8_t *qn, *tn;
    uint32_t qnl, tnl;
    uint64_t v[12];     // the integer columns by column number (0, 4, 5 unused)
    bool neg;
    const uint8_t *tags;  // bytes after the 12th column's tab
    uint64_t tags_n;
    bool has_tags;
};
__device__ __forceinline__ uint64_t warp_count_byte(const uint8_t *p, uint64_t n, uint8_t c, uint32_t lane) {
    // occurrences of c in p[0, n): aligned 16-byte pieces, one per lane and iteration
    const uint64_t a = reinterpret_cast<uintptr_t>(p) & 15u;
    const uint8_t *base = p - a;
    const uint64_t nvec = (a + n + 15) >> 4;
    const uint32_t pat = 0x01010101u * c;
    uint32_t cnt = 0;
    for (uint64_t v = lane; v < nvec; v += 32) {
        const uint4 x = ldg16(base + 16 * v);
        const int64_t i0 = (int64_t)(16 * v) - (int64_t)a;
        const int lo = i0 < 0 ? (int)-i0 : 0, hi = (uint64_t)(i0 + 16) > n ? (int)((int64_t)n - i0) : 16;
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) cnt += __popc(zero_bytes(w[k] ^ pat) & range_flags(lo - 4 * k, hi - 4 * k));
    }
    return __reduce_add_sync(0xFFFFFFFFu, cnt);
}
template <int PASS>
__device__ __forceinline__ uint64_t warp_paf_row(const PafRowIn &I, char *out, uint32_t lane) {
    uint32_t flen = 0;
    if (lane == 0) flen = csv_name_len(I.qn, I.qnl);
    else if (lane == 5) flen = csv_name_len(I.tn, I.tnl);
    else if (lane == 4) flen = 1;
    else if (lane < 12) flen = ndigits_u64(I.v[lane]);
    uint32_t incl = lane < 12 ? flen + 1 : 0;  // every column is followed by a tab -- or, the last one, by the tags / the newline
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (uint32_t)d) incl += y;
    }
    const uint32_t fixed = __shfl_sync(0xFFFFFFFFu, incl, 11) - 1;  // 12 columns and the 11 tabs between them
    uint64_t tags_len = 0;
    bool quoted = false;
    if (I.has_tags) {
        quoted = warp_count_byte(I.tags, I.tags_n, '"', lane) != 0;
        if (!quoted) tags_len = I.tags_n + 1;
        else {
            if (lane == 0) tags_len = tags_csv_len(I.tags, I.tags_n);
            tags_len = __shfl_sync(0xFFFFFFFFu, tags_len, 0);
        }
    }
    const uint64_t total = (uint64_t)fixed + tags_len + 1;
    if (PASS == 0) return total;
    if (lane < 12) {
        char *o = out + (incl - flen - 1);
        if (lane == 0) o = csv_name_put(o, I.qn, I.qnl);
        else if (lane == 5) o = csv_name_put(o, I.tn, I.tnl);
        else if (lane == 4) *o++ = I.neg ? '-' : '+';
        else o = put_u64(o, I.v[lane]);
        if (lane < 11) *o = '\t';
    }
    if (I.has_tags) {
        char *t = out + fixed;
        if (!quoted) {
            if (lane == 0) t[0] = '\t';
            warp_copy(t + 1, I.tags, I.tags_n, lane);
        } else if (lane == 0) tags_csv_put(t, I.tags, I.tags_n);
    }
    if (lane == 0) out[total - 1] = '\n';
    return total;
}

// warp-per-record twin of k_val_rows below (same ValView, same bytes)
template <int PASS>
__global__ void __launch_bounds__(128) k_val_rows_warp(ValView V) {
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (r > V.n_rec) return;
    if (r == V.n_rec) { if (PASS == 0 && lane == 0) { V.q_len[r] = V.t_len[r] = V.row_len[r] = 0; } return; }
    const unsigned long long *c = V.cnt11 + r * 11;
    const uint64_t qs = V.R.q_start[r], ts = V.R.t_start[r], qe = V.C.q_end[r], te = V.C.t_end[r];
    const uint64_t exp_q = qs + c[0] + c[1] + c[3] + c[7], exp_t = ts + c[0] + c[1] + c[5] + c[9];  // validate.rs:82-86,103-107
    const bool qbad = exp_q != qe, tbad = exp_t != te;
    const uint8_t *qn = V.text + V.R.q_name_off[r], *tn = V.text + V.R.t_name_off[r];
    const uint32_t qnl = V.R.q_name_len[r], tnl = V.R.t_name_len[r];
    if (lane == 0) {  // the two lists: only records that fail
        if (PASS == 0) {
            V.q_len[r] = qbad ? qnl + 1 + ndigits_u64(qs) + 1 + ndigits_u64(qe) + 1 : 0;
            V.t_len[r] = tbad ? tnl + 1 + ndigits_u64(ts) + 1 + ndigits_u64(te) + 1 : 0;
            if (qbad) atomicAdd(V.n_bad, 1ull);
            if (tbad) atomicAdd(V.n_bad + 1, 1ull);
        } else {
            if (qbad) {
                char *o = V.q_out + V.q_off[r];
                for (uint32_t i = 0; i < qnl; ++i) *o++ = (char)qn[i];
                *o++ = ':'; o = put_u64(o, qs); *o++ = '-'; o = put_u64(o, qe); *o++ = '\n';
            }
            if (tbad) {
                char *o = V.t_out + V.t_off[r];
                for (uint32_t i = 0; i < tnl; ++i) *o++ = (char)tn[i];
                *o++ = ':'; o = put_u64(o, ts); *o++ = '-'; o = put_u64(o, te); *o++ = '\n';
            }
        }
    }
    if (!V.fix) { if (PASS == 0 && lane == 0) V.row_len[r] = 0; return; }
    PafRowIn I;
    I.qn = qn; I.tn = tn; I.qnl = qnl; I.tnl = tnl; I.neg = V.R.neg[r] & 1;
    I.v[1] = V.R.q_size[r]; I.v[2] = qs; I.v[3] = exp_q; I.v[6] = V.R.t_size[r]; I.v[7] = ts; I.v[8] = exp_t;
    I.v[9] = V.C.matches[r]; I.v[10] = V.C.block_len[r]; I.v[11] = V.C.mapq[r];
    I.has_tags = V.C.tags_off[r] <= V.C.line_end[r];
    I.tags_n = I.has_tags ? V.C.line_end[r] - V.C.tags_off[r] : 0;
    I.tags = V.text + V.C.tags_off[r];
    const uint64_t n = warp_paf_row<PASS>(I, PASS == 1 ? V.row_out + V.row_off[r] : nullptr, lane);
    if (PASS == 0 && lane == 0) V.row_len[r] = n;
}

template <int PASS>
__global__ void k_val_rows(ValView V) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > V.n_rec) return;
    if (r == V.n_rec) { if (PASS == 0) { V.q_len[r] = V.t_len[r] = V.row_len[r] = 0; } return; }
    const unsigned long long *c = V.cnt11 + r * 11;
    const uint64_t qs = V.R.q_start[r], ts = V.R.t_start[r], qe = V.C.q_end[r], te = V.C.t_end[r];
    const uint64_t exp_q = qs + c[0] + c[1] + c[3] + c[7], exp_t = ts + c[0] + c[1] + c[5] + c[9];  // validate.rs:82-86,103-107
    const bool qbad = exp_q != qe, tbad = exp_t != te;
    const uint8_t *qn = V.text + V.R.q_name_off[r], *tn = V.text + V.R.t_name_off[r];
    const uint32_t qnl = V.R.q_name_len[r], tnl = V.R.t_name_len[r];
    const bool neg = V.R.neg[r] & 1;
    const bool has_tags = V.C.tags_off[r] <= V.C.line_end[r];
    const uint64_t tags_n = has_tags ? V.C.line_end[r] - V.C.tags_off[r] : 0;
    if (PASS == 0) {
        V.q_len[r] = qbad ? qnl + 1 + ndigits_u64(qs) + 1 + ndigits_u64(qe) + 1 : 0;
        V.t_len[r] = tbad ? tnl + 1 + ndigits_u64(ts) + 1 + ndigits_u64(te) + 1 : 0;
        if (qbad) atomicAdd(V.n_bad, 1ull);
        if (tbad) atomicAdd(V.n_bad + 1, 1ull);
        uint64_t row = 0;
        if (V.fix) {
            row = csv_name_len(qn, qnl) + 1 + ndigits_u64(V.R.q_size[r]) + 1 + ndigits_u64(qs) + 1 + ndigits_u64(exp_q) + 1 + 1 + 1 +
                  csv_name_len(tn, tnl) + 1 + ndigits_u64(V.R.t_size[r]) + 1 + ndigits_u64(ts) + 1 + ndigits_u64(exp_t) + 1 +
                  ndigits_u64(V.C.matches[r]) + 1 + ndigits_u64(V.C.block_len[r]) + 1 + ndigits_u64(V.C.mapq[r]) +
                  (has_tags ? tags_csv_len(V.text + V.C.tags_off[r], tags_n) : 0) + 1;
        }
        V.row_len[r] = row;
    } else {
        if (qbad) {
            char *o = V.q_out + V.q_off[r];
            for (uint32_t i = 0; i < qnl; ++i) *o++ = (char)qn[i];
            *o++ = ':'; o = put_u64(o, qs); *o++ = '-'; o = put_u64(o, qe); *o++ = '\n';
        }
        if (tbad) {
            char *o = V.t_out + V.t_off[r];
            for (uint32_t i = 0; i < tnl; ++i) *o++ = (char)tn[i];
            *o++ = ':'; o = put_u64(o, ts); *o++ = '-'; o = put_u64(o, te); *o++ = '\n';
        }
        if (V.fix) {  // validate.rs:136-144: csv serialisation of the (corrected) PafRecord
            char *o = V.row_out + V.row_off[r];
            o = csv_name_put(o, qn, qnl); *o++ = '\t';
            o = put_u64(o, V.R.q_size[r]); *o++ = '\t';
            o = put_u64(o, qs); *o++ = '\t';
            o = put_u64(o, exp_q); *o++ = '\t';
            *o++ = neg ? '-' : '+'; *o++ = '\t';
            o = csv_name_put(o, tn, tnl); *o++ = '\t';
            o = put_u64(o, V.R.t_size[r]); *o++ = '\t';
            o = put_u64(o, ts); *o++ = '\t';
            o = put_u64(o, exp_t); *o++ = '\t';
            o = put_u64(o, V.C.matches[r]); *o++ = '\t';
            o = put_u64(o, V.C.block_len[r]); *o++ = '\t';
            o = put_u64(o, V.C.mapq[r]);
            if (has_tags) o = tags_csv_put(o, V.text + V.C.tags_off[r], tags_n);
            *o++ = '\n';
        }
    }
}

}  // namespace wga
