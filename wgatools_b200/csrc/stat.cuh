// stat.cuh -- device side of `stat`: pair grouping + counter reduction (src/tools/stat.rs:167-250).
// The reference buffers one PairStat per record and groups them in a HashMap on the host; here the
// per-record counters never leave the GPU: every record hashes its Pair key
// (ref_name, ref_size, query_name, query_size) into an open-addressing table and adds its counters with
// 64-bit atomics; only the (small) table and the ordered list of '-'-strand records -- needed to
// reproduce the order-dependent f32 sum of inv_size -- go back to the host.
#pragma once
#include "ctx.cuh"
#include "tokenize.cuh"

namespace wga {

constexpr int kPairVals = 13;  // 11 counters (struct Cigar order) + ref_start_min + query_start_min

struct RecMeta {  // per-record view of the Pair key and coordinates (MAF: via s-lines; PAF: direct)
    const uint8_t *text;
    const uint64_t *t_name_off, *q_name_off;
    const uint32_t *t_name_len, *q_name_len;
    const uint64_t *t_size, *q_size, *t_start, *q_start;  // q_start already forward-strand
    const uint8_t *neg, *skip;
};

__device__ __forceinline__ uint64_t hash_bytes(const uint8_t *p, uint32_t n, uint64_t h) {
    for (uint32_t i = 0; i < n; ++i) h = (h ^ p[i]) * 0x100000001B3ull;  // FNV-1a
    return h;
}
__device__ __forceinline__ uint64_t pair_key(const RecMeta &m, uint64_t r) {
    uint64_t h = 0xCBF29CE484222325ull;
    h = hash_bytes(m.text + m.t_name_off[r], m.t_name_len[r], h);
    h = mix64(h ^ m.t_size[r]);
    h = hash_bytes(m.text + m.q_name_off[r], m.q_name_len[r], h ^ 0x9E3779B97F4A7C15ull);
    h = mix64(h ^ (m.q_size[r] * 0xD1342543DE82EF95ull));
    return h ? h : 1;
}

// counters11: per-record 11 counters in struct Cigar order
__global__ void k_pair_insert(uint64_t n_rec, RecMeta m, const unsigned long long *__restrict__ cnt11,
                              uint64_t cap_mask, unsigned long long *__restrict__ keys,
                              unsigned long long *__restrict__ first_rec, unsigned long long *__restrict__ vals,
                              uint64_t *__restrict__ rec_slot) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    if (m.skip && m.skip[r]) { rec_slot[r] = ~0ull; return; }
    unsigned long long key = pair_key(m, r);
    uint64_t slot = key & cap_mask;
    for (;;) {
        unsigned long long cur = keys[slot];
        if (cur == key) break;
        if (cur == 0ull) {
            unsigned long long old = atomicCAS(&keys[slot], 0ull, key);
            if (old == 0ull || old == key) break;
        }
        slot = (slot + 1) & cap_mask;
    }
    rec_slot[r] = slot;
    unsigned long long *v = vals + slot * kPairVals;
    const unsigned long long *c = cnt11 + r * 11;
    // Records of one pair are usually consecutive in the file: when all 32 lanes of a warp landed in the same slot the
    // warp adds up first and issues 13 atomics instead of 416 on one address.
    int same = 0;
    const uint32_t act = __activemask();
    if (act == 0xFFFFFFFFu) __match_all_sync(0xFFFFFFFFu, (unsigned long long)slot, &same);
    if (same) {
        const uint32_t lane = threadIdx.x & 31;
        unsigned long long first = r, ts = m.t_start[r], qs = m.q_start[r];
        unsigned long long cc[11];
#pragma unroll
        for (int i = 0; i < 11; ++i) cc[i] = c[i];
#pragma unroll
        for (int d = 16; d; d >>= 1) {
#pragma unroll
            for (int i = 0; i < 11; ++i) cc[i] += __shfl_xor_sync(0xFFFFFFFFu, cc[i], d);
            const unsigned long long f2 = __shfl_xor_sync(0xFFFFFFFFu, first, d), t2 = __shfl_xor_sync(0xFFFFFFFFu, ts, d),
                                     q2 = __shfl_xor_sync(0xFFFFFFFFu, qs, d);
            first = f2 < first ? f2 : first; ts = t2 < ts ? t2 : ts; qs = q2 < qs ? q2 : qs;
        }
        if (lane == 0) {
            atomicMin(&first_rec[slot], first);
#pragma unroll
            for (int i = 0; i < 11; ++i)
                if (cc[i]) atomicAdd(&v[i], cc[i]);
            atomicMin(&v[11], ts);
            atomicMin(&v[12], qs);
        }
        return;
    }
    atomicMin(&first_rec[slot], (unsigned long long)r);
#pragma unroll
    for (int i = 0; i < 11; ++i)
        if (c[i]) atomicAdd(&v[i], c[i]);
    atomicMin(&v[11], (unsigned long long)m.t_start[r]);
    atomicMin(&v[12], (unsigned long long)m.q_start[r]);
}

// a 64-bit key collision between two different pairs would silently merge them: verify by bytes
__global__ void k_pair_verify(uint64_t n_rec, RecMeta m, const unsigned long long *__restrict__ first_rec,
                              const uint64_t *__restrict__ rec_slot, unsigned long long *err_key) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    uint64_t slot = rec_slot[r];
    if (slot == ~0ull) return;
    uint64_t f = first_rec[slot];
    if (f == r) return;
    bool ok = m.t_size[r] == m.t_size[f] && m.q_size[r] == m.q_size[f] && m.t_name_len[r] == m.t_name_len[f] &&
              m.q_name_len[r] == m.q_name_len[f];
    if (ok) {
        const uint8_t *a = m.text + m.t_name_off[r], *b = m.text + m.t_name_off[f];
        for (uint32_t i = 0; i < m.t_name_len[r]; ++i) ok &= a[i] == b[i];
        a = m.text + m.q_name_off[r];
        b = m.text + m.q_name_off[f];
        for (uint32_t i = 0; i < m.q_name_len[r]; ++i) ok &= a[i] == b[i];
    }
    if (!ok) atomicMin(err_key, (unsigned long long)((r << 8) | WGA_E_INTERNAL));
}

__global__ void k_slot_flags(uint64_t cap, const unsigned long long *__restrict__ keys, uint32_t *__restrict__ flag) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap) flag[s] = keys[s] != 0ull;
}

// compact used slots -> pair arrays; also name lengths (t then q) for the name gather
// pair_sizes = {ref_size, query_size, pair hash} per pair
__global__ void k_pair_gather(uint64_t cap, const uint32_t *__restrict__ flag, const uint32_t *__restrict__ idx,
                              const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ first_rec,
                              const unsigned long long *__restrict__ vals, RecMeta m,
                              uint64_t *__restrict__ pair_first, uint64_t *__restrict__ pair_vals,
                              uint64_t *__restrict__ pair_sizes, uint32_t *__restrict__ name_len) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= cap || !flag[s]) return;
    uint32_t p = idx[s];
    uint64_t f = first_rec[s];
    pair_first[p] = f;
    for (int i = 0; i < kPairVals; ++i) pair_vals[(uint64_t)p * kPairVals + i] = vals[s * kPairVals + i];
    pair_sizes[3 * (uint64_t)p] = m.t_size[f];
    pair_sizes[3 * (uint64_t)p + 1] = m.q_size[f];
    pair_sizes[3 * (uint64_t)p + 2] = keys[s];
    name_len[2 * (uint64_t)p] = m.t_name_len[f];
    name_len[2 * (uint64_t)p + 1] = m.q_name_len[f];
}

__global__ void k_pair_names(uint64_t n_pairs, const uint64_t *__restrict__ pair_first, RecMeta m,
                             const uint64_t *__restrict__ name_off, char *__restrict__ name_buf) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n_pairs) return;
    uint64_t f = pair_first[i >> 1];
    const uint8_t *src = (i & 1) ? m.text + m.q_name_off[f] : m.text + m.t_name_off[f];
    uint32_t n = (i & 1) ? m.q_name_len[f] : m.t_name_len[f];
    char *dst = name_buf + name_off[i];
    for (uint32_t k = 0; k < n; ++k) dst[k] = (char)src[k];
}

// pairs in first-occurrence order: the host merges them front to back, so everything it reads is laid out in that order
// here (perm = compact pair index of sorted position oi)
__global__ void k_pair_permute(uint64_t n_pairs, const uint64_t *__restrict__ perm, const uint64_t *__restrict__ pair_first,
                               const uint64_t *__restrict__ pair_vals, const uint64_t *__restrict__ pair_sizes,
                               const uint32_t *__restrict__ name_len, uint64_t *__restrict__ first2, uint64_t *__restrict__ vals2,
                               uint64_t *__restrict__ sizes2, uint32_t *__restrict__ name_len2, uint32_t *__restrict__ invperm) {
    uint64_t oi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (oi > n_pairs) return;
    if (oi == n_pairs) { name_len2[2 * oi] = 0; return; }
    const uint64_t p = perm[oi];
    first2[oi] = pair_first[p];
    for (int i = 0; i < kPairVals; ++i) vals2[oi * kPairVals + i] = pair_vals[p * kPairVals + i];
    for (int i = 0; i < 3; ++i) sizes2[3 * oi + i] = pair_sizes[3 * p + i];
    name_len2[2 * oi] = name_len[2 * p];
    name_len2[2 * oi + 1] = name_len[2 * p + 1];
    invperm[p] = (uint32_t)oi;
}

// ordered list of '-'-strand records: (pair index in first-occurrence order, aligned + query_aligned) in record order
__global__ void k_inv_flags(uint64_t n_rec, const uint8_t *__restrict__ neg, const uint8_t *__restrict__ skip,
                            uint32_t *__restrict__ flag) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rec) flag[r] = neg[r] && !(skip && skip[r]);
}
__global__ void k_inv_gather(uint64_t n_rec, const uint32_t *__restrict__ flag, const uint64_t *__restrict__ idx,
                             const uint64_t *__restrict__ rec_slot, const uint32_t *__restrict__ slot_idx,
                             const uint32_t *__restrict__ invperm, const unsigned long long *__restrict__ cnt11,
                             uint32_t *__restrict__ inv_pair, uint64_t *__restrict__ inv_num) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec || !flag[r]) return;
    const unsigned long long *c = cnt11 + r * 11;
    // RecStat::from (common.rs:116-140): aligned = m + x + del + inv_del; query_align = m + x + ins + inv_ins
    uint64_t aligned = c[0] + c[1] + c[5] + c[9], qal = c[0] + c[1] + c[3] + c[7];
    uint64_t o = idx[r];
    inv_pair[o] = invperm[slot_idx[rec_slot[r]]];
    inv_num[o] = aligned + qal;
}

// MAF: expand the walk's 5 partial counters into the 11 counters of struct Cigar (cigar.rs:16-29)
__global__ void k_maf_counts11(uint64_t n_rec, const unsigned long long *__restrict__ cnt5,
                               const uint64_t *__restrict__ ncols, const uint8_t *__restrict__ neg,
                               const uint8_t *__restrict__ skip, unsigned long long *__restrict__ cnt11) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    unsigned long long *o = cnt11 + r * 11;
    for (int i = 0; i < 11; ++i) o[i] = 0;
    if (skip[r]) return;
    const unsigned long long *c = cnt5 + r * 5;
    uint64_t X = c[0], eI = c[1], cI = c[2], eD = c[3], cD = c[4];
    o[0] = ncols[r] - X - cI - cD;
    o[1] = X;
    if (neg[r]) { o[6] = eI; o[7] = cI; o[8] = eD; o[9] = cD; o[10] = 1; }
    else { o[2] = eI; o[3] = cI; o[4] = eD; o[5] = cD; }
}

// per-record metadata arrays for MAF records out of the s-line SoA
__global__ void k_maf_meta(uint64_t n_rec, SLines sl, const uint64_t *__restrict__ tl, const uint64_t *__restrict__ ql,
                           uint64_t *t_name_off, uint64_t *q_name_off, uint32_t *t_name_len, uint32_t *q_name_len,
                           uint64_t *t_size, uint64_t *q_size, uint64_t *t_start, uint64_t *q_start) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    uint64_t a = tl[r], b = ql[r];
    t_name_off[r] = sl.name_off[a]; t_name_len[r] = sl.name_len[a];
    q_name_off[r] = sl.name_off[b]; q_name_len[r] = sl.name_len[b];
    t_size[r] = sl.src[a]; q_size[r] = sl.src[b];
    t_start[r] = sl.start[a];
    uint64_t st = sl.start[b], al = sl.ali[b], sz = sl.src[b];
    q_start[r] = sl.neg[b] ? sz - st - al : st;  // maf.rs:433-441
}

}  // namespace wga
