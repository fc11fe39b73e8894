// chain.cuh -- MAF / PAF -> UCSC chain (maf2chain, paf2chain).
//
// Replaces src/converter.rs:57-92 (maf2chain) and :148-173 (paf2chain); src/parser/chain.rs:92-204 (ChainDataLine and
// ChainHeader Display, TryFrom<&MAFRecord> / TryFrom<&PafRecord> including the negative-strand quirk: query.end is
// computed from the already updated query.start, chain.rs:133-138,174-179); src/parser/cigar.rs:155-245 (head / tail
// indel trim) and :252-295,435-490 (parse_cigar_to_chain, parse_maf_seq_to_chain, cigar_unit_chain).
//
// Both commands are sparse passes over a run list that already exists: the caller-independent run heads of the dense
// MAF walk (classes = X I D, '=' and 'X' merge into the chain's ungapped blocks) or the op table of the PAF CIGAR.
#pragma once
#include "ctx.cuh"
#include "emit_paf.cuh"

namespace wga {

struct ChainView {
    uint64_t n_rec, id_base;
    const uint8_t *text;
    const uint64_t *t_name_off, *q_name_off;
    const uint32_t *t_name_len, *q_name_len;
    const uint64_t *t_size, *q_size, *t_start, *t_end, *q_start, *q_end;  // query coordinates on the forward strand
    const uint8_t *neg, *skip;   // skip may be null
    const uint64_t *run_first;   // [n_rec+1]
    const uint64_t *run_len;
    const uint8_t *run_type;
    int maf;                     // run_type holds walk classes (= X I D) instead of op types (M I D)
    uint64_t *row_len;           // [n_rec+1]
    const uint64_t *row_off;
    char *out;
};

__device__ __forceinline__ uint32_t chain_type(const ChainView &V, uint64_t j) {  // 0 ungapped, 1 I, 2 D
    const uint32_t c = V.run_type[j];
    return V.maf ? (c <= 1u ? 0u : c - 1u) : c;
}

// "chain 255 tName tSize + tStart tEnd qName qSize qStrand qStart qEnd id", then "\n size qdiff tdiff" per ungapped
// block that is followed by a gap, "\n size" for the last block, and a blank line
template <int PASS>
__global__ void k_chain_rows(ChainView V) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > V.n_rec) return;
    if (r == V.n_rec) { if (PASS == 0) V.row_len[r] = 0; return; }
    if (V.skip && V.skip[r]) { if (PASS == 0) V.row_len[r] = 0; return; }
    const uint64_t j0 = V.run_first[r], j1 = V.run_first[r + 1];
    // head / tail indel trim (cigar.rs:155-245): leading I/D runs add up, the tail keeps the LAST I and the LAST D run
    uint64_t hi = 0, hd = 0, ti = 0, td = 0;
    bool head = true;
    uint64_t size = 0, qd = 0, tdf = 0, body = 0;  // body = bytes of the data lines (PASS 0)
    for (uint64_t j = j0; j < j1; ++j) {
        const uint32_t t = chain_type(V, j);
        const uint64_t l = V.run_len[j];
        if (t == 0) {
            ti = 0; td = 0; head = false;
            if (PASS == 0) {
                if (size != 0 && tdf + qd != 0) { body += 1 + ndigits_u64(size) + 1 + ndigits_u64(qd) + 1 + ndigits_u64(tdf); size = 0; }
                size += l; tdf = 0; qd = 0;
            }
        } else if (t == 1) { if (head) hi += l; ti = l; if (PASS == 0) tdf += l; }
        else { if (head) hd += l; td = l; if (PASS == 0) qd += l; }
    }
    const bool neg = V.neg[r] & 1;
    const uint64_t ts = V.t_start[r] + hd, te = V.t_end[r] - td;
    uint64_t qs, qe;
    if (!neg) { qs = V.q_start[r] + hi; qe = V.q_end[r] - ti; }
    else { qs = V.q_size[r] - (V.q_end[r] - hi); qe = V.q_size[r] - (qs + ti); }  // chain.rs:133-138 (sic)
    const uint8_t *tn = V.text + V.t_name_off[r], *qn = V.text + V.q_name_off[r];
    const uint32_t tnl = V.t_name_len[r], qnl = V.q_name_len[r];
    const uint64_t id = V.id_base + r;
    if (PASS == 0) {
        const uint64_t hdr = 10 + tnl + 1 + ndigits_u64(V.t_size[r]) + 3 + ndigits_u64(ts) + 1 + ndigits_u64(te) + 1 + qnl + 1 +
                             ndigits_u64(V.q_size[r]) + 3 + ndigits_u64(qs) + 1 + ndigits_u64(qe) + 1 + ndigits_u64(id);
        V.row_len[r] = hdr + body + 1 + ndigits_u64(size) + 2;
        return;
    }
    char *o = V.out + V.row_off[r];
    const char h[10] = {'c', 'h', 'a', 'i', 'n', '\t', '2', '5', '5', '\t'};
    for (int k = 0; k < 10; ++k) *o++ = h[k];
    for (uint32_t k = 0; k < tnl; ++k) *o++ = (char)tn[k];
    *o++ = '\t'; o = put_u64(o, V.t_size[r]); *o++ = '\t'; *o++ = '+'; *o++ = '\t';
    o = put_u64(o, ts); *o++ = '\t'; o = put_u64(o, te); *o++ = '\t';
    for (uint32_t k = 0; k < qnl; ++k) *o++ = (char)qn[k];
    *o++ = '\t'; o = put_u64(o, V.q_size[r]); *o++ = '\t'; *o++ = neg ? '-' : '+'; *o++ = '\t';
    o = put_u64(o, qs); *o++ = '\t'; o = put_u64(o, qe); *o++ = '\t'; o = put_u64(o, id);
    size = 0; qd = 0; tdf = 0;
    for (uint64_t j = j0; j < j1; ++j) {  // cigar.rs:458-490 cigar_unit_chain
        const uint32_t t = chain_type(V, j);
        const uint64_t l = V.run_len[j];
        if (t == 0) {
            if (size != 0 && tdf + qd != 0) {
                *o++ = '\n'; o = put_u64(o, size); *o++ = '\t'; o = put_u64(o, qd); *o++ = '\t'; o = put_u64(o, tdf);
                size = 0;
            }
            size += l; tdf = 0; qd = 0;
        } else if (t == 1) tdf += l;
        else qd += l;
    }
    *o++ = '\n'; o = put_u64(o, size); *o++ = '\n'; *o++ = '\n';
}

// per-record header fields of a MAF record (target = s-line 0, query = the selected s-line; maf.rs:424-460)
__global__ void k_chain_meta_maf(uint64_t n_rec, SLines sl, const uint64_t *__restrict__ tl, const uint64_t *__restrict__ ql,
                                 uint64_t *__restrict__ tno, uint64_t *__restrict__ qno, uint32_t *__restrict__ tnl,
                                 uint32_t *__restrict__ qnl, uint64_t *__restrict__ m6 /* t_size q_size t_start t_end q_start q_end */) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint64_t t = tl[r], q = ql[r];
    tno[r] = sl.name_off[t]; tnl[r] = sl.name_len[t];
    qno[r] = sl.name_off[q]; qnl[r] = sl.name_len[q];
    uint64_t qs, qe;
    query_fwd(sl, q, &qs, &qe);
    m6[r] = sl.src[t]; m6[n_rec + r] = sl.src[q]; m6[2 * n_rec + r] = sl.start[t]; m6[3 * n_rec + r] = sl.start[t] + sl.ali[t];
    m6[4 * n_rec + r] = qs; m6[5 * n_rec + r] = qe;
}

}  // namespace wga
