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
__device__ __forceinline__ void chain_row_seq(const ChainView &V, uint64_t r) {
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
template <int PASS>
__global__ void k_chain_rows(ChainView V) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > V.n_rec) return;
    if (r == V.n_rec) { if (PASS == 0) V.row_len[r] = 0; return; }
    if (V.skip && V.skip[r]) { if (PASS == 0) V.row_len[r] = 0; return; }
    chain_row_seq<PASS>(V, r);
}

// ---- the same rows by one WARP per record (the default).  With all lengths > 0 the chain writer is a function of the
// op list's run structure: ungapped ops (M / = / X) merge into runs, I / D ops into gap groups; every gap group that
// lies BETWEEN two ungapped runs yields one line "size qdiff tdiff" (size = the run before it, qdiff / tdiff = the D / I
// sums of the group), leading and trailing gap groups only move the header coordinates (head / tail trim), and the last
// run's size closes the record.  Lane i of a step owns op 32*step + i: three u64 prefix sums per step (ungapped, I, D
// lengths), the run / group an op belongs to falls out of two ballots, the lane of the first ungapped op AFTER an inner gap
// group writes that group's line.  Carried between steps: the run (or the run before the open gap group) and the open
// group's sums.  A record that holds a zero-length op (where the sequential rules differ) is done by lane 0 with
// chain_row_seq.  hdr_len[r]: bytes of the header line (PASS 0 -> PASS 1: the lines are written before the trims are known).
__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, uint32_t src) { return __shfl_sync(0xFFFFFFFFu, v, src); }
__device__ __forceinline__ unsigned long long scan_u64(unsigned long long v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= (uint32_t)d) v += y;
    }
    return v;
}
template <int PASS>
__global__ void __launch_bounds__(128) k_chain_rows_warp(ChainView V, uint32_t *__restrict__ hdr_len, uint8_t *__restrict__ irregular) {
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (r > V.n_rec) return;
    if (r == V.n_rec) { if (PASS == 0 && lane == 0) V.row_len[r] = 0; return; }
    if (V.skip && V.skip[r]) { if (PASS == 0 && lane == 0) V.row_len[r] = 0; return; }
    if (PASS == 1 && irregular[r]) { if (lane == 0) chain_row_seq<1>(V, r); return; }
    const uint64_t j0 = V.run_first[r], j1 = V.run_first[r + 1];
    bool seenM = false, inGap = false, zero = false;
    unsigned long long curM = 0, gI = 0, gD = 0, hi = 0, hd = 0, ti = 0, td = 0, body = 0;
    char *o = PASS == 1 ? V.out + V.row_off[r] + hdr_len[r] : nullptr;
    for (uint64_t base = j0; base < j1; base += 32) {
        const uint64_t j = base + lane;
        const bool have = j < j1;
        const unsigned long long l = have ? V.run_len[j] : 0ull;
        const uint32_t t = have ? chain_type(V, j) : 3u;
        const bool isM = t == 0u, isI = t == 1u, isD = t == 2u;
        zero = zero || (have && l == 0ull);
        const uint32_t bM = __ballot_sync(0xFFFFFFFFu, isM), bI = __ballot_sync(0xFFFFFFFFu, isI), bD = __ballot_sync(0xFFFFFFFFu, isD);
        const uint32_t bG = bI | bD;
        const uint32_t cnt = j1 - base < 32 ? (uint32_t)(j1 - base) : 32u, last = cnt - 1u;
        const unsigned long long PM = scan_u64(isM ? l : 0ull, lane), PI = scan_u64(isI ? l : 0ull, lane), PD = scan_u64(isD ? l : 0ull, lane);
        // ---- lines: the first ungapped op after an inner gap group
        const uint32_t below = (1u << lane) - 1u;
        const uint32_t mBefore = bM & below;
        const bool prevGap = lane == 0 ? inGap : ((bG >> (lane - 1u)) & 1u) != 0u;
        const bool emit = isM && prevGap && (seenM || mBefore != 0u);
        const uint32_t pm = mBefore ? 31u - (uint32_t)__clz(mBefore) : 0u;           // last ungapped op before this lane
        const uint32_t gBefore = bG & ((1u << pm) - 1u);
        const uint32_t pg = gBefore ? 31u - (uint32_t)__clz(gBefore) : 0u;           // last gap op before that run
        const unsigned long long PM_pm = shfl_u64(PM, pm), PI_pm = shfl_u64(PI, pm), PD_pm = shfl_u64(PD, pm), PM_pg = shfl_u64(PM, pg);
        unsigned long long size, qd, tdf;
        if (mBefore) {
            tdf = PI - PI_pm; qd = PD - PD_pm;
            size = gBefore ? PM_pm - PM_pg : (inGap ? 0ull : curM) + PM_pm;
        } else {  // the group began before this step (or right at its start, behind the carried run)
            tdf = (inGap ? gI : 0ull) + PI; qd = (inGap ? gD : 0ull) + PD;
            size = curM;
        }
        const uint32_t mine = emit ? 1u + (uint32_t)ndigits_u64(size) + 1u + (uint32_t)ndigits_u64(qd) + 1u + (uint32_t)ndigits_u64(tdf) : 0u;
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += y;
        }
        const uint32_t step_bytes = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (PASS == 1 && emit) {
            char *q = o + (incl - mine);
            *q++ = '\n'; q = put_u64(q, size); *q++ = '\t'; q = put_u64(q, qd); *q++ = '\t'; q = put_u64(q, tdf);
        }
        if (PASS == 1) o += step_bytes;
        body += step_bytes;
        // ---- head trim: I / D lengths in front of the record's first ungapped op
        if (!seenM) {
            const uint32_t fm = bM ? (uint32_t)__ffs(bM) - 1u : cnt;
            if (fm) { hi += shfl_u64(PI, fm - 1u); hd += shfl_u64(PD, fm - 1u); }
        }
        // ---- tail trim: the LAST I and the LAST D op behind the last ungapped op
        {
            const uint32_t after = bM ? ~((2u << (31u - (uint32_t)__clz(bM))) - 1u) : 0xFFFFFFFFu;
            const uint32_t ia = bI & after, da = bD & after;
            const unsigned long long li = shfl_u64(l, ia ? 31u - (uint32_t)__clz(ia) : 0u), ld = shfl_u64(l, da ? 31u - (uint32_t)__clz(da) : 0u);
            ti = ia ? li : (bM ? 0ull : ti);
            td = da ? ld : (bM ? 0ull : td);
        }
        // ---- state behind the step's last op
        const bool endGap = ((bG >> last) & 1u) != 0u;
        const unsigned long long PM_l = shfl_u64(PM, last), PI_l = shfl_u64(PI, last), PD_l = shfl_u64(PD, last);
        if (endGap) {
            if (bM) {  // the run in front of the open group ends at the step's last ungapped op
                const uint32_t lm = 31u - (uint32_t)__clz(bM);
                const uint32_t gB = bG & ((1u << lm) - 1u);
                const unsigned long long PM_lm = shfl_u64(PM, lm), PI_lm = shfl_u64(PI, lm), PD_lm = shfl_u64(PD, lm);
                const unsigned long long PM_g = shfl_u64(PM, gB ? 31u - (uint32_t)__clz(gB) : 0u);
                curM = gB ? PM_lm - PM_g : (inGap ? 0ull : curM) + PM_lm;
                gI = PI_l - PI_lm; gD = PD_l - PD_lm;
            } else {
                gI = (inGap ? gI : 0ull) + PI_l; gD = (inGap ? gD : 0ull) + PD_l;
            }
        } else {
            const unsigned long long PM_g = shfl_u64(PM, bG ? 31u - (uint32_t)__clz(bG) : 0u);
            curM = bG ? PM_l - PM_g : (inGap ? 0ull : curM) + PM_l;
            gI = 0; gD = 0;
        }
        inGap = endGap;
        seenM = seenM || bM != 0u;
    }
    zero = __any_sync(0xFFFFFFFFu, zero);
    if (zero) {  // the sequential rules (size != 0 && gaps != 0) are not the run structure: lane 0 redoes the record
        if (PASS == 0 && lane == 0) { irregular[r] = 1; chain_row_seq<0>(V, r); }
        return;
    }
    if (lane != 0) return;
    const bool neg = V.neg[r] & 1;
    const uint64_t ts = V.t_start[r] + hd, te = V.t_end[r] - td;
    uint64_t qs, qe;
    if (!neg) { qs = V.q_start[r] + hi; qe = V.q_end[r] - ti; }
    else { qs = V.q_size[r] - (V.q_end[r] - hi); qe = V.q_size[r] - (qs + ti); }  // chain.rs:133-138 (sic)
    const uint32_t tnl = V.t_name_len[r], qnl = V.q_name_len[r];
    const uint64_t id = V.id_base + r;
    const uint64_t size = seenM ? curM : 0ull;
    if (PASS == 0) {
        const uint64_t hdr = 10 + tnl + 1 + ndigits_u64(V.t_size[r]) + 3 + ndigits_u64(ts) + 1 + ndigits_u64(te) + 1 + qnl + 1 +
                             ndigits_u64(V.q_size[r]) + 3 + ndigits_u64(qs) + 1 + ndigits_u64(qe) + 1 + ndigits_u64(id);
        hdr_len[r] = (uint32_t)hdr;
        irregular[r] = 0;
        V.row_len[r] = hdr + body + 1 + ndigits_u64(size) + 2;
        return;
    }
    const uint8_t *tn = V.text + V.t_name_off[r], *qn = V.text + V.q_name_off[r];
    char *h = V.out + V.row_off[r];
    const char hh[10] = {'c', 'h', 'a', 'i', 'n', '\t', '2', '5', '5', '\t'};
    for (int k = 0; k < 10; ++k) *h++ = hh[k];
    for (uint32_t k = 0; k < tnl; ++k) *h++ = (char)tn[k];
    *h++ = '\t'; h = put_u64(h, V.t_size[r]); *h++ = '\t'; *h++ = '+'; *h++ = '\t';
    h = put_u64(h, ts); *h++ = '\t'; h = put_u64(h, te); *h++ = '\t';
    for (uint32_t k = 0; k < qnl; ++k) *h++ = (char)qn[k];
    *h++ = '\t'; h = put_u64(h, V.q_size[r]); *h++ = '\t'; *h++ = neg ? '-' : '+'; *h++ = '\t';
    h = put_u64(h, qs); *h++ = '\t'; h = put_u64(h, qe); *h++ = '\t'; h = put_u64(h, id);
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
