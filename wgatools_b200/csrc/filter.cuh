// filter.cuh -- `filter -f paf`: record filter and the all-to-all pair-sum filter.
//
// Replaces src/tools/filter.rs:41-63 filter_paf + :92-106 filter_alignrec (a record is dropped when
// target_end - target_start < min_block_size or query_length < min_query_size) and :108-160 filter_paf_align_pair
// (`-a`: per (query_name, target_name) sum of target_end - target_start over the WHOLE file; a record is kept when its
// pair's sum >= the cutoff; min_block_size / min_query_size are ignored then, src/utils.rs:559-563).  Kept records are
// re-serialised by the csv writer (tab-delimited, fields quoted only when they hold a quote, tab, CR or LF).
// The pair sums reuse the open-addressing table of stat.cuh keyed by the two names; u64 arithmetic wraps as the
// reference's release build does.  The reference collects the records through rayon's par_bridge, so its output order
// is unspecified; rows come in FILE order here (the same convention as validate's lists).
#pragma once
#include "ctx.cuh"
#include "emit_paf.cuh"
#include "paf.cuh"
#include "validate.cuh"

namespace wga {

// -b / -q: keep[r] = !(block < min_block | query_length < min_query)
__global__ void k_flt_keep(uint64_t n_rec, PafRecs R, const uint64_t *__restrict__ t_end, uint64_t min_block, uint64_t min_query,
                           uint8_t *__restrict__ keep) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint64_t block = t_end[r] - R.t_start[r];  // target_align_size (paf.rs:142-145)
    keep[r] = !((block < min_block) | (R.q_size[r] < min_query));
}

// -a, step 1: the record's contribution in counter 0 of an otherwise empty counter row (input of k_pair_insert)
__global__ void k_flt_contrib(uint64_t n_rec, PafRecs R, const uint64_t *__restrict__ t_end, unsigned long long *__restrict__ cnt11) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    unsigned long long *o = cnt11 + r * 11;
    o[0] = t_end[r] - R.t_start[r];
#pragma unroll
    for (int i = 1; i < 11; ++i) o[i] = 0;
}
// -a, step 2: keep[r] = sum of the record's pair >= cutoff
__global__ void k_flt_keep_pair(uint64_t n_rec, const uint64_t *__restrict__ rec_slot, const unsigned long long *__restrict__ vals,
                                uint64_t cutoff, uint8_t *__restrict__ keep) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    keep[r] = vals[rec_slot[r] * kPairVals] >= cutoff;
}

struct FltView {
    uint64_t n_rec;
    const uint8_t *text;
    PafRecs R;
    ValCols C;
    const uint8_t *keep;
    uint64_t *row_len;         // [n_rec + 1]
    const uint64_t *row_off;
    char *out;
};

// csv serialisation of a kept PafRecord (12 typed columns, then the tags one by one)
template <int PASS>
__global__ void k_flt_rows(FltView V) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > V.n_rec) return;
    if (r == V.n_rec) { if (PASS == 0) V.row_len[r] = 0; return; }
    if (!V.keep[r]) { if (PASS == 0) V.row_len[r] = 0; return; }
    const uint8_t *qn = V.text + V.R.q_name_off[r], *tn = V.text + V.R.t_name_off[r];
    const uint32_t qnl = V.R.q_name_len[r], tnl = V.R.t_name_len[r];
    const uint64_t qs = V.R.q_start[r], ts = V.R.t_start[r], qe = V.C.q_end[r], te = V.C.t_end[r];
    const bool has_tags = V.C.tags_off[r] <= V.C.line_end[r];
    const uint64_t tags_n = has_tags ? V.C.line_end[r] - V.C.tags_off[r] : 0;
    if (PASS == 0) {
        V.row_len[r] = csv_name_len(qn, qnl) + 1 + ndigits_u64(V.R.q_size[r]) + 1 + ndigits_u64(qs) + 1 + ndigits_u64(qe) + 1 + 1 + 1 +
                       csv_name_len(tn, tnl) + 1 + ndigits_u64(V.R.t_size[r]) + 1 + ndigits_u64(ts) + 1 + ndigits_u64(te) + 1 +
                       ndigits_u64(V.C.matches[r]) + 1 + ndigits_u64(V.C.block_len[r]) + 1 + ndigits_u64(V.C.mapq[r]) +
                       (has_tags ? tags_csv_len(V.text + V.C.tags_off[r], tags_n) : 0) + 1;
    } else {
        char *o = V.out + V.row_off[r];
        o = csv_name_put(o, qn, qnl); *o++ = '\t';
        o = put_u64(o, V.R.q_size[r]); *o++ = '\t';
        o = put_u64(o, qs); *o++ = '\t';
        o = put_u64(o, qe); *o++ = '\t';
        *o++ = (V.R.neg[r] & 1) ? '-' : '+'; *o++ = '\t';
        o = csv_name_put(o, tn, tnl); *o++ = '\t';
        o = put_u64(o, V.R.t_size[r]); *o++ = '\t';
        o = put_u64(o, ts); *o++ = '\t';
        o = put_u64(o, te); *o++ = '\t';
        o = put_u64(o, V.C.matches[r]); *o++ = '\t';
        o = put_u64(o, V.C.block_len[r]); *o++ = '\t';
        o = put_u64(o, V.C.mapq[r]);
        if (has_tags) o = tags_csv_put(o, V.text + V.C.tags_off[r], tags_n);
        *o++ = '\n';
    }
}

// warp-per-record twin of k_flt_rows (validate.cuh warp_paf_row)
template <int PASS>
__global__ void __launch_bounds__(128) k_flt_rows_warp(FltView V) {
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (r > V.n_rec) return;
    if (r == V.n_rec || !V.keep[r]) { if (PASS == 0 && lane == 0) V.row_len[r] = 0; return; }
    PafRowIn I;
    I.qn = V.text + V.R.q_name_off[r]; I.tn = V.text + V.R.t_name_off[r]; I.qnl = V.R.q_name_len[r]; I.tnl = V.R.t_name_len[r];
    I.neg = V.R.neg[r] & 1;
    I.v[1] = V.R.q_size[r]; I.v[2] = V.R.q_start[r]; I.v[3] = V.C.q_end[r]; I.v[6] = V.R.t_size[r]; I.v[7] = V.R.t_start[r];
    I.v[8] = V.C.t_end[r]; I.v[9] = V.C.matches[r]; I.v[10] = V.C.block_len[r]; I.v[11] = V.C.mapq[r];
    I.has_tags = V.C.tags_off[r] <= V.C.line_end[r];
    I.tags_n = I.has_tags ? V.C.line_end[r] - V.C.tags_off[r] : 0;
    I.tags = V.text + V.C.tags_off[r];
    const uint64_t n = warp_paf_row<PASS>(I, PASS == 1 ? V.out + V.row_off[r] : nullptr, lane);
    if (PASS == 0 && lane == 0) V.row_len[r] = n;
}

}  // namespace wga
