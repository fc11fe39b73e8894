// paf2maf.cuh -- PAF + 2 FASTA -> MAF blocks.
//
// Replaces src/converter.rs:176-265 paf2maf: faidx fetch of target [t_start, t_end-1] and query
// [q_start, q_end-1] (rust-htslib fetch_seq_string, inclusive end, clipped to the sequence), reverse_complement
// of '-' strand queries (src/utils.rs:83-101), parse_cigar_to_insert (src/parser/cigar.rs:492-551: M/=/X advance,
// I inserts '-' into the target row, D into the query row) and MAFWriter::write_record (src/parser/maf.rs:566-581).
//
// The reference edits Strings (insert_str is O(L) per indel => quadratic).  Here a row is DEFINED by its runs:
//   target row = fetched target bases with a gap run of len_k inserted at gapped offset o_k for every I op,
//   query  row = (reverse-complemented) query bases with a gap run at o_k for every D op,
// so every CIGAR op owns a disjoint slice of each output row and all slices are written in parallel from the
// FASTA resident in HBM (newline-stripped once at load time).  Row length = fetched bases + inserted gaps, which
// also reproduces the reference on CIGARs that consume fewer / more bases than were fetched (left-over bases
// trail the row; an insert past the end of the row is the reference's insert_str panic -> WGA_E_REF_PANIC).
#pragma once
#include "ctx.cuh"
#include "paf.cuh"

struct wga_fasta {  // device-resident FASTA: sequences concatenated without line breaks
    int device = 0;
    uint8_t *d_seq = nullptr;      // all bases
    uint8_t *d_names = nullptr;    // concatenated names
    uint64_t *d_meta = nullptr;    // per sequence (sorted by name hash): hash, name_off, name_len, seq_base, seq_len
    uint64_t n_seq = 0, total = 0;
};

namespace wga {

constexpr int kFaMeta = 5;

__global__ void k_fasta_unwrap(uint64_t n_seq, const uint8_t *__restrict__ fa, uint64_t fa_len,
                               const uint64_t *__restrict__ ent /* per seq: offset, linebases, linewidth, seq_base, len */,
                               uint8_t *__restrict__ dst) {
    const uint64_t sidx = blockIdx.y;
    if (sidx >= n_seq) return;
    const uint64_t off = ent[sidx * 5], lb = ent[sidx * 5 + 1], lw = ent[sidx * 5 + 2], base = ent[sidx * 5 + 3], len = ent[sidx * 5 + 4];
    if (lb == 0) return;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < len; b += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t p = off + b / lb * lw + b % lb;
        dst[base + b] = p < fa_len ? fa[p] : (uint8_t)'N';
    }
}

__device__ __forceinline__ uint64_t fnv1a(const uint8_t *p, uint32_t n) {
    uint64_t h = 0xCBF29CE484222325ull;
    for (uint32_t i = 0; i < n; ++i) h = (h ^ p[i]) * 0x100000001B3ull;
    return h;
}

struct FaView {
    const uint8_t *seq, *names;
    const uint64_t *meta;
    uint64_t n_seq;
};
// sequence index by name (binary search on the name hash, then byte compare); ~0 when absent
__device__ __forceinline__ uint64_t fa_find(const FaView &F, const uint8_t *name, uint32_t len) {
    const uint64_t h = fnv1a(name, len);
    uint64_t lo = 0, hi = F.n_seq;
    while (lo < hi) { uint64_t m = (lo + hi) >> 1; if (F.meta[m * kFaMeta] < h) lo = m + 1; else hi = m; }
    for (; lo < F.n_seq && F.meta[lo * kFaMeta] == h; ++lo) {
        if (F.meta[lo * kFaMeta + 2] != len) continue;
        const uint8_t *nm = F.names + F.meta[lo * kFaMeta + 1];
        bool eq = true;
        for (uint32_t i = 0; i < len; ++i) eq &= nm[i] == name[i];
        if (eq) return lo;
    }
    return ~0ull;
}
// htslib faidx_adjust_position(end_adjust = 1) + fai_retrieve: [beg, end] inclusive -> (first base, count)
__device__ __forceinline__ void fa_clip(uint64_t len_u, uint64_t beg_u, uint64_t end_u, uint64_t *first, uint64_t *count) {
    int64_t beg = (int64_t)beg_u, end = (int64_t)end_u, len = (int64_t)len_u;
    if (end < beg) beg = end;
    if (beg < 0) beg = 0; else if (len <= beg) beg = len;
    if (end < 0) end = 0; else if (len <= end) end = len - 1;
    int64_t want = end + 1 - beg;
    *first = (uint64_t)beg;
    *count = want > 0 ? (uint64_t)want : 0;
}

struct P2mRecs {  // per PAF record
    uint64_t *q_end, *t_end, *mapq;
    uint64_t *t_src, *q_src;  // absolute index of the first fetched base in the FASTA's d_seq
    uint64_t *LT, *LQ;        // fetched base counts
    uint64_t *n_ops;
};

// extra typed columns + FASTA spans (converter.rs:196-225)
// end_adj: 1 = paf2maf / pafpseudo fetch [start, end-1]; 0 = `call -f paf` fetches [start, end] (paf.rs:221-237, Q13)
__global__ void k_p2m_spans(LineView lv, const uint8_t *__restrict__ flags, const uint32_t *__restrict__ rec_scan, PafRecs R,
                            P2mRecs X, FaView T, FaView Q, uint64_t end_adj, int lookup, unsigned long long *err_key) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= lv.n_lines || !flags[k]) return;
    const uint64_t r = rec_scan[k];
    // re-split the line for columns 3, 8, 11 (k_paf_parse validated all 12 already)
    uint64_t S = lv.start(k), E = lv.end(k);
    if (E > S && lv.text[E - 1] == '\r') --E;
    uint64_t wlo = k == 0 ? 0 : lv.nl_ws_idx[k - 1] + 1, whi = k < lv.n_nl ? lv.nl_ws_idx[k] : lv.n_ws;
    uint64_t fb = S, w = wlo, v3 = 0, v8 = 0, v11 = 0;
    for (uint64_t nf = 0; nf < 12; ++nf) {
        uint64_t fe = w < whi ? (lv.ws_ent[w] & kPosMask) : E;
        if (fe > E) fe = E;
        uint64_t val = 0;
        if (nf == 3 || nf == 8 || nf == 11) csv_u64(lv.text + fb, fe - fb, &val);
        if (nf == 3) v3 = val; else if (nf == 8) v8 = val; else if (nf == 11) v11 = val;
        if (w >= whi) break;
        fb = fe + 1;
        ++w;
    }
    X.q_end[r] = v3; X.t_end[r] = v8; X.mapq[r] = v11;
    X.LT[r] = X.LQ[r] = X.t_src[r] = X.q_src[r] = 0;
    if (!lookup) return;  // pafpseudo looks sequences up later, for the records it keeps
    const uint64_t ti = fa_find(T, lv.text + R.t_name_off[r], R.t_name_len[r]);
    const uint64_t qi = ti == ~0ull ? ~0ull : fa_find(Q, lv.text + R.q_name_off[r], R.q_name_len[r]);
    if (ti == ~0ull || qi == ~0ull) { paf_report(err_key, r, 1, WGA_E_NOM); return; }  // faidx: sequence not found
    uint64_t first, count;
    fa_clip(T.meta[ti * kFaMeta + 4], R.t_start[r], v8 - end_adj, &first, &count);
    X.t_src[r] = T.meta[ti * kFaMeta + 3] + first;
    X.LT[r] = count;
    fa_clip(Q.meta[qi * kFaMeta + 4], R.q_start[r], v3 - end_adj, &first, &count);
    X.q_src[r] = Q.meta[qi * kFaMeta + 3] + first;
    X.LQ[r] = count;
}

// reverse_complement validity of '-' strand queries (utils.rs:83-101): the reference walks the string from its END
__global__ void k_p2m_revcomp_check(uint64_t n_rec, PafRecs R, P2mRecs X, const uint8_t *__restrict__ qseq,
                                    unsigned long long *err_key) {
    const uint64_t r = blockIdx.x;
    if (r >= n_rec || !(R.neg[r] & 1)) return;
    const uint8_t *p = qseq + X.q_src[r];
    const uint64_t n = X.LQ[r];
    for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint8_t c = p[n - 1 - i];
        const uint8_t u = c & 0xDFu;
        const bool ok = (u == 'A' || u == 'C' || u == 'G' || u == 'T' || u == 'N') && c >= 'A';
        if (!ok) { paf_report(err_key, r, 2 + (i > 0x7FFFF0ull ? 0x7FFFF0ull : i), WGA_E_VCF_BASE); return; }
    }
}

// CIGAR -> op count (PASS 0) or op table (PASS 1).  Tokenizer semantics of cigar.rs:43-75 / parse_cigar_to_insert.
template <int PASS>
__global__ void k_p2m_ops(uint64_t n_rec, const uint8_t *__restrict__ text, PafRecs R, uint64_t *__restrict__ n_ops,
                          const uint64_t *__restrict__ op_base, uint64_t *__restrict__ op_len, uint8_t *__restrict__ op_type,
                          uint32_t *__restrict__ op_rec, uint32_t *__restrict__ op_pos, unsigned long long *err_key,
                          const uint8_t *__restrict__ only = nullptr, bool other_ok = false) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rec) return;
    if (r == n_rec) { if (PASS == 0) n_ops[r] = 0; return; }
    if (only && !only[r]) return;  // the warp kernel (paf.cuh k_paf_ops_warp) already did this record
    const uint64_t len = R.cig_len[r];
    const uint8_t *p = text + R.cig_off[r];
    if ((R.neg[r] & 2) || len == 0) {
        if (PASS == 0) {
            n_ops[r] = 0;
            if (!(R.neg[r] & 2) && len == 0) paf_report(err_key, r, 0x800001, WGA_E_REF_PANIC);  // fold_many1 on "" (errors.rs:88-97)
        }
        return;
    }
    uint64_t o = PASS == 1 ? op_base[r] : 0, cnt = 0;
    uint64_t val = 0;
    uint32_t nd = 0, op_run = 0;
    bool ovf = false, in_op = false, failed = false;
    uint8_t op = 0;
    auto close_token = [&](uint64_t pos) {
        if (failed) return;
        int code = 0;
        if (op_run != 1) code = WGA_E_CIGAR_OP_INVALID;
        else if (nd == 0 || ovf) code = WGA_E_PARSE_INT;
        const bool known = op == 'M' || op == '=' || op == 'X' || op == 'I' || op == 'D';
        if (!code && !known && !other_ok) code = WGA_E_CIGAR_OP_INVALID;  // other_ok: type 3, ignored by the consumer (plot.cuh)
        if (code) {
            failed = true;
            if (PASS == 0) paf_report(err_key, r, 0x800002 + (pos > 0x7FFFF0ull ? 0x7FFFF0ull : pos), code);
            return;
        }
        if (PASS == 1) {
            op_len[o] = val;
            op_type[o] = (op == 'I') ? 1 : (op == 'D') ? 2 : known ? 0 : 3;
            if (op_rec) op_rec[o] = (uint32_t)r;
            if (op_pos) op_pos[o] = (uint32_t)(pos > 0x7FFFF0ull ? 0x7FFFF0ull : pos);
            ++o;
        }
        ++cnt;
    };
    for (uint64_t i = 0; i < len && !failed; ++i) {
        const uint8_t ch = p[i];
        if (ch >= '0' && ch <= '9') {
            if (in_op) { close_token(i); val = 0; nd = 0; ovf = false; op_run = 0; in_op = false; }
            const uint32_t d = ch - '0';
            if (val > (0xFFFFFFFFFFFFFFFFull - d) / 10) ovf = true;
            val = val * 10 + d;
            ++nd;
        } else { in_op = true; ++op_run; op = ch; }
    }
    if (!failed) close_token(len);
    if (PASS == 0) n_ops[r] = failed ? 0 : cnt;
}

// per-op contributions for the prefix sums: all = op length, tcon = target bases consumed (M/=/X/D), gaps in target (I)
__global__ void k_p2m_contrib(uint64_t n_ops, const uint64_t *__restrict__ op_len, const uint8_t *__restrict__ op_type,
                              uint64_t *__restrict__ c_all, uint64_t *__restrict__ c_i, uint64_t *__restrict__ c_d) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n_ops) return;
    if (k == n_ops) { c_all[k] = c_i[k] = c_d[k] = 0; return; }
    const uint64_t l = op_len[k];
    const uint8_t t = op_type[k];
    c_all[k] = l;
    c_i[k] = t == 1 ? l : 0;
    c_d[k] = t == 2 ? l : 0;
}

struct P2mView {
    uint64_t n_rec, n_ops;
    const uint8_t *text;
    PafRecs R;
    P2mRecs X;
    const uint64_t *op_base;                // [n_rec+1]
    const uint64_t *op_len;
    const uint8_t *op_type;
    const uint32_t *op_rec, *op_pos;
    const uint64_t *S_all, *S_i, *S_d;      // exclusive scans over all ops [n_ops+1]
    const uint8_t *tseq, *qseq;
    uint64_t *blk_len;                      // [n_rec+1]
    const uint64_t *blk_off;
    uint32_t *fix_t, *fix_q;                // bytes before each row's sequence
    char *out;
    const uint8_t *t_neg = nullptr;         // chain2maf prints the chain header's own target strand; null = '+'
};

// insert_str past the end of the row = the reference's panic (cigar.rs:507,513)
__global__ void k_p2m_check(P2mView V, unsigned long long *err_key) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= V.n_ops) return;
    const uint8_t t = V.op_type[k];
    if (t == 0) return;
    const uint32_t r = V.op_rec[k];
    const uint64_t b = V.op_base[r];
    const uint64_t o = V.S_all[k] - V.S_all[b], gi = V.S_i[k] - V.S_i[b], gd = V.S_d[k] - V.S_d[b];
    // row length right before this op: fetched bases + gaps inserted so far
    const bool bad = t == 1 ? o > V.X.LT[r] + gi : o > V.X.LQ[r] + gd;
    if (bad) paf_report(err_key, r, 0x800002 + V.op_pos[k], WGA_E_REF_PANIC);  // same position scale as the token errors
}

__device__ __forceinline__ uint8_t comp_base(uint8_t c) {
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
        default: return c;  // N / n (validated by k_p2m_revcomp_check)
    }
}

// PASS 0: block byte lengths; PASS 1: the fixed text of each block (a-line, s-line prefixes, line ends)
template <int PASS>
__global__ void k_p2m_blocks(P2mView V) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > V.n_rec) return;
    if (r == V.n_rec) { if (PASS == 0) V.blk_len[r] = 0; return; }
    const uint64_t b0 = V.op_base[r], b1 = V.op_base[r + 1];
    const uint64_t tot_i = V.S_i[b1] - V.S_i[b0], tot_d = V.S_d[b1] - V.S_d[b0];
    const uint64_t trow = V.X.LT[r] + tot_i, qrow = V.X.LQ[r] + tot_d;
    const uint64_t t_st = V.R.t_start[r], t_al = V.X.t_end[r] - t_st, q_st = V.R.q_start[r], q_al = V.X.q_end[r] - q_st;
    const bool neg = V.R.neg[r] == 1;
    const uint64_t q_maf = neg ? V.R.q_size[r] - V.X.q_end[r] : q_st;  // converter.rs:213-216
    const uint32_t tnl = V.R.t_name_len[r], qnl = V.R.q_name_len[r];
    if (PASS == 0) {
        // "a score={mapq}\n" "s\t{t}\t{st}\t{al}\t+\t{size}\t" row "\n" "s\t{q}\t{st}\t{al}\t{strand}\t{size}\t" row "\n" "\n"
        const uint32_t a_len = 8 + ndigits_u64(V.X.mapq[r]) + 1;
        const uint32_t ft = 2 + tnl + 1 + ndigits_u64(t_st) + 1 + ndigits_u64(t_al) + 3 + ndigits_u64(V.R.t_size[r]) + 1;
        const uint32_t fq = 2 + qnl + 1 + ndigits_u64(q_maf) + 1 + ndigits_u64(q_al) + 3 + ndigits_u64(V.R.q_size[r]) + 1;
        V.fix_t[r] = a_len + ft;
        V.fix_q[r] = fq;
        V.blk_len[r] = (uint64_t)a_len + ft + trow + 1 + fq + qrow + 2;
    } else {
        char *o = V.out + V.blk_off[r];
        const char a[8] = {'a', ' ', 's', 'c', 'o', 'r', 'e', '='};
        for (int i = 0; i < 8; ++i) *o++ = a[i];
        o = put_u64(o, V.X.mapq[r]); *o++ = '\n';
        *o++ = 's'; *o++ = '\t';
        for (uint32_t i = 0; i < tnl; ++i) *o++ = (char)V.text[V.R.t_name_off[r] + i];
        *o++ = '\t'; o = put_u64(o, t_st); *o++ = '\t'; o = put_u64(o, t_al); *o++ = '\t';
        *o++ = (V.t_neg && V.t_neg[r]) ? '-' : '+'; *o++ = '\t';
        o = put_u64(o, V.R.t_size[r]); *o++ = '\t';
        o += trow; *o++ = '\n';
        *o++ = 's'; *o++ = '\t';
        for (uint32_t i = 0; i < qnl; ++i) *o++ = (char)V.text[V.R.q_name_off[r] + i];
        *o++ = '\t'; o = put_u64(o, q_maf); *o++ = '\t'; o = put_u64(o, q_al); *o++ = '\t'; *o++ = neg ? '-' : '+'; *o++ = '\t';
        o = put_u64(o, V.R.q_size[r]); *o++ = '\t';
        o += qrow; *o++ = '\n'; *o++ = '\n';
    }
}

// warp-wide copy of n bytes: destination brought to 4-byte alignment, then one 32-bit store per lane and iteration fed
// by two aligned loads and a funnel shift (source and destination are misaligned against each other), byte tails
__device__ __forceinline__ void warp_copy(char *dst, const uint8_t *src, uint64_t n, uint32_t lane) {
    if (n < 96) {
        for (uint64_t i = lane; i < n; i += 32) dst[i] = (char)src[i];
        return;
    }
    const uint32_t head = (4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u;
    if (lane < head) dst[lane] = (char)src[lane];
    dst += head; src += head; n -= head;
    const uint32_t sm = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(src - sm);
    uint64_t nw = n >> 2;
    if (sm && nw) --nw;  // the funnel reads word w + 1: keep every load inside [src - sm, src + n)
    uint32_t *dw = reinterpret_cast<uint32_t *>(dst);
    for (uint64_t w = lane; w < nw; w += 32) {
        const uint32_t lo = __ldg(sw + w);
        dw[w] = sm ? __funnelshift_r(lo, __ldg(sw + w + 1), sm * 8u) : lo;
    }
    const uint64_t done = nw << 2;
    for (uint64_t i = done + lane; i < n; i += 32) dst[i] = (char)src[i];
}

// ---- 16-byte variant for long copies: dst brought to 16-byte alignment by a byte head, then every lane stores one uint4
// per iteration, assembled from two aligned 16-byte loads (the source is misaligned against the destination by sm bytes)
__device__ __forceinline__ uint4 shift16(const uint4 &a, const uint4 &b, uint32_t ws, uint32_t sh) {
    // bytes [sm, sm + 16) of the 32-byte pair (a, b), sm = 4 * ws + sh / 8; ws is warp-uniform
    uint32_t w0, w1, w2, w3, w4;
    switch (ws) {
        case 0: w0 = a.x; w1 = a.y; w2 = a.z; w3 = a.w; w4 = b.x; break;
        case 1: w0 = a.y; w1 = a.z; w2 = a.w; w3 = b.x; w4 = b.y; break;
        case 2: w0 = a.z; w1 = a.w; w2 = b.x; w3 = b.y; w4 = b.z; break;
        default: w0 = a.w; w1 = b.x; w2 = b.y; w3 = b.z; w4 = b.w; break;
    }
    return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}
__device__ __forceinline__ void warp_copy16(char *dst, const uint8_t *src, uint64_t n, uint32_t lane) {
    if (n < 512) { warp_copy(dst, src, n, lane); return; }
    const uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
    if (lane < head) dst[lane] = (char)src[lane];
    dst += head; src += head; n -= head;
    const uint32_t sm = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
    const uint4 *sv = reinterpret_cast<const uint4 *>(src - sm);
    uint64_t nv = n >> 4;
    if (sm && nv) --nv;  // the realign reads vector v + 1: keep every load inside [src - sm, src + n)
    uint4 *dv = reinterpret_cast<uint4 *>(dst);
    const uint32_t ws = sm >> 2, sh = (sm & 3u) * 8u;
    for (uint64_t v = lane; v < nv; v += 32) {
        const uint4 a = __ldg(sv + v);
        dv[v] = sm ? shift16(a, __ldg(sv + v + 1), ws, sh) : a;
    }
    const uint64_t done = nv << 4;
    for (uint64_t i = done + lane; i < n; i += 32) dst[i] = (char)src[i];
}
// complement of 4 bases at once (ACGTNacgtn only -- validated by k_p2m_revcomp_check): A<->T differ by 0x15, C<->G by 0x04;
// bit 1 set and bit 3 clear = C/G, bit 1 clear = A/T, bit 3 set = N (unchanged)
__device__ __forceinline__ uint32_t comp4(uint32_t w) {
    const uint32_t b1 = (w >> 1) & 0x01010101u, b3 = (w >> 3) & 0x01010101u;
    const uint32_t cg = b1 & ~b3, at = ~b1 & 0x01010101u;
    return w ^ ((cg << 2) | at | (at << 2) | (at << 4));
}
// dst[i] = complement(src_end[-1 - i]) for i in [0, n): the reverse complement of the n bytes that END at src_end
__device__ __forceinline__ void warp_revcomp16(char *dst, const uint8_t *src_end, uint64_t n, uint32_t lane) {
    if (n < 512) {
        for (uint64_t i = lane; i < n; i += 32) dst[i] = (char)comp_base(src_end[-1 - (int64_t)i]);
        return;
    }
    const uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
    if (lane < head) dst[lane] = (char)comp_base(src_end[-1 - (int64_t)lane]);
    dst += head; src_end -= head; n -= head;
    // destination vector v <- source block [src_end - 16 (v + 1), src_end - 16 v), reversed
    const uint32_t sm = (uint32_t)(reinterpret_cast<uintptr_t>(src_end) & 15u);  // same misalignment for every block
    const uint4 *top = reinterpret_cast<const uint4 *>(src_end - sm);          // aligned vector that holds src_end[0]
    uint64_t nv = n >> 4;
    uint4 *dv = reinterpret_cast<uint4 *>(dst);
    const uint32_t ws = sm >> 2, sh = (sm & 3u) * 8u;
    for (uint64_t v = lane; v < nv; v += 32) {
        // block start = src_end - 16 (v + 1) = (top - (v + 1)) + sm bytes
        const uint4 a = __ldg(top - (v + 1));
        const uint4 blk = sm ? shift16(a, __ldg(top - v), ws, sh) : a;
        dv[v] = make_uint4(comp4(__byte_perm(blk.w, 0, 0x0123)), comp4(__byte_perm(blk.z, 0, 0x0123)),
                           comp4(__byte_perm(blk.y, 0, 0x0123)), comp4(__byte_perm(blk.x, 0, 0x0123)));
    }
    const uint64_t done = nv << 4;
    for (uint64_t i = done + lane; i < n; i += 32) dst[i] = (char)comp_base(src_end[-1 - (int64_t)i]);
}

// ---- row expansion by RUNS (the default): a row is a plain copy of fetched bases between the gap runs inserted into IT
// (I ops for the target row, D ops for the query row), shifted by the gaps inserted so far.  One warp per gap op = its
// '-' fill + the copy run that follows it up to the next gap op of the same row (or to the end of the fetched bases when
// it is the row's last), plus one warp per record and row for the run before the first gap op.  Runs are ~10 ops long on
// whole-genome CIGARs, so the copies are kilobytes, stored as 16-byte vectors.
__global__ void k_p2m_gap_flags(uint64_t n_ops, const uint8_t *__restrict__ op_type, uint32_t *__restrict__ fI, uint32_t *__restrict__ fD) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n_ops) return;
    const uint8_t t = k < n_ops ? op_type[k] : (uint8_t)0;
    fI[k] = t == 1;
    fD[k] = t == 2;
}
__global__ void k_p2m_gap_lists(uint64_t n_ops, const uint8_t *__restrict__ op_type, const uint64_t *__restrict__ idxI,
                                const uint64_t *__restrict__ idxD, uint64_t *__restrict__ Ilist, uint64_t *__restrict__ Dlist) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_ops) return;
    const uint8_t t = op_type[k];
    if (t == 1) Ilist[idxI[k]] = k;
    else if (t == 2) Dlist[idxD[k]] = k;
}
__global__ void k_p2m_runs(P2mView V, uint64_t n_I, uint64_t n_D, const uint64_t *__restrict__ idxI, const uint64_t *__restrict__ Ilist,
                           const uint64_t *__restrict__ idxD, const uint64_t *__restrict__ Dlist) {
    uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_t = n_I + V.n_rec;
    if (w >= n_t + n_D + V.n_rec) return;
    const bool qrow = w >= n_t;
    if (qrow) w -= n_t;
    const uint64_t n_gap = qrow ? n_D : n_I;
    const uint64_t *gl = qrow ? Dlist : Ilist, *gidx = qrow ? idxD : idxI;
    uint32_t r;
    uint64_t a, next, b0, b1, gap_o = 0, gap_len = 0;
    if (w < n_gap) {  // a gap op and the run after it
        const uint64_t k = gl[w];
        r = V.op_rec[k];
        b0 = V.op_base[r]; b1 = V.op_base[r + 1];
        gap_o = V.S_all[k] - V.S_all[b0];
        gap_len = V.op_len[k];
        a = k + 1;
        next = (w + 1 < n_gap && gl[w + 1] < b1) ? gl[w + 1] : b1;
    } else {          // the run before the record's first gap op
        r = (uint32_t)(w - n_gap);
        b0 = V.op_base[r]; b1 = V.op_base[r + 1];
        a = b0;
        const uint64_t j = gidx[b0];
        next = (j < n_gap && gl[j] < b1) ? gl[j] : b1;
    }
    const uint64_t tot_i = V.S_i[b1] - V.S_i[b0];
    const uint64_t L = qrow ? V.X.LQ[r] : V.X.LT[r];
    char *trow = V.out + V.blk_off[r] + V.fix_t[r];
    char *row = qrow ? trow + (V.X.LT[r] + tot_i) + 1 + V.fix_q[r] : trow;
    for (uint64_t i = lane; i < gap_len; i += 32) row[gap_o + i] = '-';
    // the run [a, next): its first output column and the first base it copies
    const uint64_t o = V.S_all[a] - V.S_all[b0];
    const uint64_t gaps = qrow ? V.S_d[a] - V.S_d[b0] : V.S_i[a] - V.S_i[b0];
    const uint64_t base = o - gaps;
    if (base >= L) return;
    uint64_t n = L - base;                                  // the row's last run also carries the bases no op covers
    if (next < b1) { const uint64_t rl = V.S_all[next] - V.S_all[a]; n = rl < n ? rl : n; }
    if (!qrow) warp_copy16(row + o, V.tseq + V.X.t_src[r] + base, n, lane);
    else if (V.R.neg[r] != 1) warp_copy16(row + o, V.qseq + V.X.q_src[r] + base, n, lane);
    else warp_revcomp16(row + o, V.qseq + V.X.q_src[r] + (L - base), n, lane);  // base b of the row = complement of Q[L - 1 - b]
}

// one warp per CIGAR op (+ one virtual op per record for the bases left over after the last op): writes the op's
// slice of the target row and of the query row
__global__ void k_p2m_rows(P2mView V) {
    const uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (wid >= V.n_ops + V.n_rec) return;
    uint32_t r;
    uint64_t o, tb, qb, len;
    uint8_t type;
    uint64_t b0, b1;
    if (wid < V.n_ops) {
        const uint64_t k = wid;
        r = V.op_rec[k];
        b0 = V.op_base[r]; b1 = V.op_base[r + 1];
        o = V.S_all[k] - V.S_all[b0];
        const uint64_t gi = V.S_i[k] - V.S_i[b0], gd = V.S_d[k] - V.S_d[b0];
        tb = o - gi;  // target bases consumed before this op
        qb = o - gd;
        len = V.op_len[k];
        type = V.op_type[k];
    } else {  // tail: bases fetched but not covered by the CIGAR trail the rows
        r = (uint32_t)(wid - V.n_ops);
        b0 = V.op_base[r]; b1 = V.op_base[r + 1];
        o = V.S_all[b1] - V.S_all[b0];
        tb = o - (V.S_i[b1] - V.S_i[b0]);
        qb = o - (V.S_d[b1] - V.S_d[b0]);
        len = ~0ull;
        type = 3;
    }
    const uint64_t LT = V.X.LT[r], LQ = V.X.LQ[r];
    const uint64_t tot_i = V.S_i[b1] - V.S_i[b0];
    const bool neg = V.R.neg[r] == 1;
    char *trow = V.out + V.blk_off[r] + V.fix_t[r];
    char *qrow = trow + (LT + tot_i) + 1 + V.fix_q[r];
    const uint8_t *T = V.tseq + V.X.t_src[r], *Q = V.qseq + V.X.q_src[r];
    // target row
    if (type == 1) {
        for (uint64_t i = lane; i < len; i += 32) trow[o + i] = '-';
    } else if (tb < LT) {
        uint64_t n = LT - tb < len ? LT - tb : len;
        warp_copy(trow + o, T + tb, n, lane);  // base b sits at b + (gaps inserted before it) = o
    }
    // query row
    if (type == 2) {
        for (uint64_t i = lane; i < len; i += 32) qrow[o + i] = '-';
    } else if (qb < LQ) {
        uint64_t n = LQ - qb < len ? LQ - qb : len;
        if (!neg) warp_copy(qrow + o, Q + qb, n, lane);
        else for (uint64_t i = lane; i < n; i += 32) qrow[o + i] = (char)comp_base(Q[LQ - 1 - (qb + i)]);
    }
}

}  // namespace wga
