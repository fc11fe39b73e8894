// plot.cuh -- `dotplot -m base-level` data: the base-level plot segments of every record.
//
// Replaces src/parser/cigar.rs:807-914 (reserve_query_start_end, emit_baseplotdatas), :917-952
// parse_cigar_to_base_plotdata and :955-985 parse_maf_to_base_plotdata, plus the csv rendering of
// src/tools/dotplot.rs:283-289 (struct BasePlotdata :181-190: ref_start, ref_end, query_start, query_end, cigar,
// ref_chro, query_chro).  The reference's state machine, restated per record over a run list that already exists
// (the EXT-classified runs of the dense MAF walk, or the op table of the PAF CIGAR):
//   * an I / D op longer than skip_cutoff is a segment of its own and ends the open M segment;
//   * every other op belongs to a stretch between two such ops: the first M / = / X op of the stretch opens ONE M
//     segment at the current offsets, and every later op of the stretch (M-type, short I, short D) moves the
//     segment's end to the offsets behind it; short I / D ops BEFORE the first M-type op only advance the offsets;
//   * ops that are none of M = X I D are ignored altogether (no offset moves): the `_ => {}` arm;
//   * on the '-' strand query_start and query_end change places (offsets still count up from query_start).
// Rows come in file order (the reference's par_bridge order is unspecified).  One thread per record, two passes
// (sizes, rows) like the chain writer.
#pragma once
#include "chain.cuh"

namespace wga {

__device__ __forceinline__ uint32_t csvc_name_len(const uint8_t *p, uint32_t n) {  // csv::Writer, ',' delimited
    uint32_t q = 0;
    bool need = false;
    for (uint32_t i = 0; i < n; ++i) { q += p[i] == '"'; need |= p[i] == '"' || p[i] == ',' || p[i] == '\r' || p[i] == '\n'; }
    return need ? n + q + 2 : n;
}
__device__ __forceinline__ char *csvc_name_put(char *o, const uint8_t *p, uint32_t n) {
    bool need = false;
    for (uint32_t i = 0; i < n; ++i) need |= p[i] == '"' || p[i] == ',' || p[i] == '\r' || p[i] == '\n';
    if (need) *o++ = '"';
    for (uint32_t i = 0; i < n; ++i) {
        if (p[i] == '"') *o++ = '"';
        *o++ = (char)p[i];
    }
    if (need) *o++ = '"';
    return o;
}

template <int PASS>
__global__ void k_plot_rows(ChainView V, uint64_t cutoff) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > V.n_rec) return;
    if (r == V.n_rec) { if (PASS == 0) V.row_len[r] = 0; return; }
    if (V.skip && V.skip[r]) { if (PASS == 0) V.row_len[r] = 0; return; }
    const uint64_t j0 = V.run_first[r], j1 = V.run_first[r + 1];
    const bool neg = V.neg[r] & 1;
    const uint8_t *tn = V.text + V.t_name_off[r], *qn = V.text + V.q_name_off[r];
    const uint32_t tnl = V.t_name_len[r], qnl = V.q_name_len[r];
    const uint32_t tail = 3 + csvc_name_len(tn, tnl) + 1 + csvc_name_len(qn, qnl) + 1;  // ",C,ref,query\n"
    char *o = PASS == 1 ? V.out + V.row_off[r] : nullptr;
    uint64_t bytes = 0;
    auto row = [&](uint64_t rs, uint64_t re, uint64_t qs, uint64_t qe, char c) {
        if (neg) { const uint64_t t = qs; qs = qe; qe = t; }
        if (PASS == 0) { bytes += ndigits_u64(rs) + 1 + ndigits_u64(re) + 1 + ndigits_u64(qs) + 1 + ndigits_u64(qe) + tail; return; }
        o = put_u64(o, rs); *o++ = ',';
        o = put_u64(o, re); *o++ = ',';
        o = put_u64(o, qs); *o++ = ',';
        o = put_u64(o, qe); *o++ = ',';
        *o++ = c; *o++ = ',';
        o = csvc_name_put(o, tn, tnl); *o++ = ',';
        o = csvc_name_put(o, qn, qnl); *o++ = '\n';
    };
    uint64_t ref = V.t_start[r], qry = V.q_start[r];
    bool open = false;                       // last_m: an M segment is the last one pushed
    uint64_t m_rs = 0, m_qs = 0, m_re = 0, m_qe = 0;
    for (uint64_t j = j0; j < j1; ++j) {
        const uint32_t t = chain_type(V, j);
        const uint64_t l = V.run_len[j];
        if (t == 0) {
            if (!open) { open = true; m_rs = ref; m_qs = qry; }
            ref += l; qry += l;
            m_re = ref; m_qe = qry;
        } else if (t == 1) {
            if (l > cutoff) {
                if (open) { row(m_rs, m_re, m_qs, m_qe, 'M'); open = false; }
                row(ref, ref, qry, qry + l, 'I');
            } else if (open) m_qe = qry + l;
            qry += l;
        } else if (t == 2) {
            if (l > cutoff) {
                if (open) { row(m_rs, m_re, m_qs, m_qe, 'M'); open = false; }
                row(ref, ref + l, qry, qry, 'D');
            } else if (open) m_re = ref + l;
            ref += l;
        }
    }
    if (open) row(m_rs, m_re, m_qs, m_qe, 'M');
    if (PASS == 0) V.row_len[r] = bytes;
}

// ---- the same rows by one WARP per record (the default; k_plot_rows behind WGA_PLOT_SEQ=1).  While an M segment is open its end
// IS the current pair of offsets (every op of the stretch moves it there), so the machine reduces to: the long I / D ops cut
// the op list into stretches; a stretch with at least one M-type op yields one M row from the offsets in front of its first
// M-type op to the offsets behind its last op; every long op yields its own row.  Lane i of a step owns op 32*step + i;
// offsets in front of an op = two u64 prefix sums per step, and a step without a long op and without a segment to open
// (the common one) only needs the two sums.  Carried between steps: the offsets, and the open segment's start.
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}
template <int PASS>
__global__ void __launch_bounds__(128) k_plot_rows_warp(ChainView V, uint64_t cutoff) {
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (r > V.n_rec) return;
    if (r == V.n_rec) { if (PASS == 0 && lane == 0) V.row_len[r] = 0; return; }
    if (V.skip && V.skip[r]) { if (PASS == 0 && lane == 0) V.row_len[r] = 0; return; }
    const uint64_t j0 = V.run_first[r], j1 = V.run_first[r + 1];
    const bool neg = V.neg[r] & 1;
    const uint8_t *tn = V.text + V.t_name_off[r], *qn = V.text + V.q_name_off[r];
    const uint32_t tnl = V.t_name_len[r], qnl = V.q_name_len[r];
    const uint32_t tail = 3 + csvc_name_len(tn, tnl) + 1 + csvc_name_len(qn, qnl) + 1;  // ",C,ref,query\n"
    auto row_size = [&](uint64_t rs, uint64_t re, uint64_t qs, uint64_t qe) -> uint32_t {  // the strand swap keeps the digits
        return (uint32_t)(ndigits_u64(rs) + 1 + ndigits_u64(re) + 1 + ndigits_u64(qs) + 1 + ndigits_u64(qe)) + tail;
    };
    auto row_put = [&](char *o, uint64_t rs, uint64_t re, uint64_t qs, uint64_t qe, char c) -> char * {
        if (neg) { const uint64_t t = qs; qs = qe; qe = t; }
        o = put_u64(o, rs); *o++ = ',';
        o = put_u64(o, re); *o++ = ',';
        o = put_u64(o, qs); *o++ = ',';
        o = put_u64(o, qe); *o++ = ',';
        *o++ = c; *o++ = ',';
        o = csvc_name_put(o, tn, tnl); *o++ = ',';
        o = csvc_name_put(o, qn, qnl); *o++ = '\n';
        return o;
    };
    unsigned long long ref = V.t_start[r], qry = V.q_start[r], m_rs = 0, m_qs = 0, bytes = 0;
    bool open = false;
    char *o = PASS == 1 ? V.out + V.row_off[r] : nullptr;
    for (uint64_t base = j0; base < j1; base += 32) {
        const uint64_t j = base + lane;
        const bool have = j < j1;
        const uint32_t t = have ? chain_type(V, j) : 3u;
        const unsigned long long l = have ? V.run_len[j] : 0ull;
        const bool isM = t == 0u, lng = (t == 1u || t == 2u) && l > cutoff;
        const uint32_t bM = __ballot_sync(0xFFFFFFFFu, isM), bL = __ballot_sync(0xFFFFFFFFu, lng);
        const unsigned long long ar = (t == 0u || t == 2u) ? l : 0ull, aq = (t == 0u || t == 1u) ? l : 0ull;
        if (bL == 0u && (open || bM == 0u)) {  // no row, no segment to open: the offsets move on
            ref += warp_sum_u64(ar);
            qry += warp_sum_u64(aq);
            continue;
        }
        const unsigned long long PR = scan_u64(ar, lane), PQ = scan_u64(aq, lane);
        const unsigned long long rb = ref + PR - ar, qb = qry + PQ - aq;  // offsets in front of this lane's op
        // the stretch this lane's op closes (if it is a long op): M-type ops between the previous long op and here
        const uint32_t below = (1u << lane) - 1u;
        const uint32_t prevL = bL & below;
        const uint32_t behind_prev = prevL ? ~((2u << (31u - (uint32_t)__clz(prevL))) - 1u) : 0xFFFFFFFFu;
        const uint32_t mBetween = bM & below & behind_prev;
        const uint32_t fm = mBetween ? (uint32_t)__ffs(mBetween) - 1u : 0u;
        const unsigned long long s_r = shfl_u64(rb, fm), s_q = shfl_u64(qb, fm);
        const bool carried = prevL == 0u && open;  // the segment came into the step open
        const bool seg_open = carried || mBetween != 0u;
        const unsigned long long seg_rs = carried ? m_rs : s_r, seg_qs = carried ? m_qs : s_q;
        uint32_t mine = 0;
        if (lng) {
            if (seg_open) mine += row_size(seg_rs, rb, seg_qs, qb);
            mine += t == 1u ? row_size(rb, rb, qb, qb + l) : row_size(rb, rb + l, qb, qb);
        }
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += y;
        }
        const uint32_t step_bytes = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (PASS == 1 && lng) {
            char *q = o + (incl - mine);
            if (seg_open) q = row_put(q, seg_rs, rb, seg_qs, qb, 'M');
            if (t == 1u) row_put(q, rb, rb, qb, qb + l, 'I');
            else row_put(q, rb, rb + l, qb, qb, 'D');
        }
        if (PASS == 1) o += step_bytes;
        bytes += step_bytes;
        // state behind the step's last op
        const uint32_t after_last = bL ? ~((2u << (31u - (uint32_t)__clz(bL))) - 1u) : 0xFFFFFFFFu;
        const uint32_t mAfter = bM & after_last;
        const uint32_t fm2 = mAfter ? (uint32_t)__ffs(mAfter) - 1u : 0u;
        const unsigned long long n_rs = shfl_u64(rb, fm2), n_qs = shfl_u64(qb, fm2);
        if (bL) { open = mAfter != 0u; m_rs = n_rs; m_qs = n_qs; }
        else if (!open && bM) { open = true; m_rs = n_rs; m_qs = n_qs; }
        ref += shfl_u64(PR, 31);
        qry += shfl_u64(PQ, 31);
    }
    if (lane != 0) return;
    if (open) {
        if (PASS == 1) row_put(o, m_rs, ref, m_qs, qry, 'M');
        bytes += row_size(m_rs, ref, m_qs, qry);
    }
    if (PASS == 0) V.row_len[r] = bytes;
}

}  // namespace wga
