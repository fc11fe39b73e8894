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

}  // namespace wga
