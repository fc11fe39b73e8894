// walk.cuh -- the dense column walk: the only kernel that streams the gapped sequences.
//
// Replaces the per-column group_by of src/parser/cigar.rs:344-432 (parse_maf_seq_to_cigar, classifier
// cigar_cat_ext :298-308) and of src/tools/caller.rs:442-446 (classifier cigar_cat_ext_caller,
// cigar.rs:314-328).
//
// Shape (v2).  Persistent warps; a warp walks one work item (<= kSeg columns of one record) at a time.
// Lane i of step k owns the 16 columns of chunk j = 32k+i: one aligned 16-byte load of the target and a
// funnel-shifted pair of 16-byte loads of the query (the two streams are misaligned against each other
// by an arbitrary number of bytes).  Alignments are >99 % identical columns, so the dense part of the
// kernel is only a 128-bit compare: a chunk whose 16 columns and whose preceding column are all '=' can
// neither contain nor start a run and is skipped.  Every other chunk ("slow": contains a mismatch/gap,
// follows one, or touches the item boundary) is pushed -- data words included -- into a per-warp queue in
// shared memory; whenever 32 slow chunks are queued the warp classifies them in lock-step, one chunk per
// lane, with the byte-SIMD flag words of hd_core.h:
//   MODE_STAT  : X / I / D column counts and I / D run starts (events);
//   MODE_RUNS_*: additionally every run head (class change), counted (EMIT=false) or written as
//                (column, class) entries in column order (EMIT=true; warp scan over the batch).
// Algorithmic traffic: 2 bytes per column (+ 32 B of item descriptor).
#pragma once
#include "ctx.cuh"
#include "tokenize.cuh"

namespace wga {

enum { MODE_STAT = 0, MODE_RUNS_EXT = 1, MODE_RUNS_CALLER = 2 };
constexpr int kWalkThreads = 256;
constexpr int kWalkWarps = kWalkThreads / 32;
constexpr int kWalkCtasPerSm = 4;

struct __align__(16) ItemDesc {  // one work item: columns [seg*kSeg, seg*kSeg+n) of record rec
    uint64_t t_off, q_off;       // byte offsets of the item's first column in target / query
    uint32_t n, rec, seg, single;  // single: the record has exactly one item (plain stores, no atomics)
};

struct WalkArgs {
    const uint8_t *text;
    const ItemDesc *items;
    const uint64_t *ncols;
    uint64_t n_items;
    unsigned long long *rec_cnt;  // [n_rec][5]: X, I_events, I_cols, D_events, D_cols
    uint32_t *item_nheads;        // RUNS, EMIT=false
    const uint64_t *run_base;     // RUNS, EMIT=true
    uint64_t *run_pos;
    uint32_t *run_rec;
    uint8_t *run_cls;
};

__global__ void k_fill_item_desc(uint64_t n_rec, const uint32_t *__restrict__ n_items,
                                 const uint64_t *__restrict__ item_base, const uint64_t *__restrict__ t_off,
                                 const uint64_t *__restrict__ q_off, const uint64_t *__restrict__ ncols,
                                 ItemDesc *__restrict__ items) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    uint64_t b = item_base[r], nc = ncols[r];
    uint32_t n = n_items[r];
    for (uint32_t s = 0; s < n; ++s) {
        ItemDesc d;
        uint64_t c0 = (uint64_t)s * kSeg;
        d.t_off = t_off[r] + c0;
        d.q_off = q_off[r] + c0;
        d.n = (uint32_t)(nc - c0 < (uint64_t)kSeg ? nc - c0 : (uint64_t)kSeg);
        d.rec = (uint32_t)r;
        d.seg = s;
        d.single = n == 1;
        items[b + s] = d;
    }
}

// 5 consecutive words starting at word index ws (0..3) of the 8-word window {a,b}
__device__ __forceinline__ void select5(const uint4 &a, const uint4 &b, uint32_t ws, uint32_t w[5]) {
    switch (ws) {
        case 0: w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; break;
        case 1: w[0] = a.y; w[1] = a.z; w[2] = a.w; w[3] = b.x; w[4] = b.y; break;
        case 2: w[0] = a.z; w[1] = a.w; w[2] = b.x; w[3] = b.y; w[4] = b.z; break;
        default: w[0] = a.w; w[1] = b.x; w[2] = b.y; w[3] = b.z; w[4] = b.w; break;
    }
}
__device__ __forceinline__ uint32_t has_zero_byte(uint32_t v) { return (v - 0x01010101u) & ~v & kHi; }

struct WalkAcc {
    uint32_t cX, cI, cD, eI, eD, nheads;
    uint64_t out;
};

// classify one queued chunk per lane (lanes >= cnt idle), accumulate counts, emit heads in order
template <int MODE, bool EMIT>
__device__ __forceinline__ void walk_batch(const WalkArgs &A, uint32_t (*Q)[64], uint32_t cnt, uint32_t lane,
                                           uint32_t a, uint32_t n, uint64_t c0, uint64_t ncols, uint32_t rec,
                                           WalkAcc &acc) {
    constexpr bool CALLER = MODE == MODE_RUNS_CALLER;
    constexpr bool RUNS = MODE != MODE_STAT;
    constexpr bool COUNTS = MODE == MODE_STAT || (MODE == MODE_RUNS_EXT && !EMIT);  // counted once
    uint32_t H[4] = {0, 0, 0, 0}, hc = 0, j = 0;
    Flags4 F[4];
    if (lane < cnt) {
        j = Q[0][lane];
        const uint32_t pv = Q[1][lane];  // bit16: valid, bits 0-7 target byte, 8-15 query byte
        uint32_t tw[4], qw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { tw[k] = Q[2 + k][lane]; qw[k] = Q[6 + k][lane]; }
        const int64_t lo = (int64_t)a - 16ll * j;                  // bytes below lo precede the item
        const int64_t hi = (int64_t)n + (int64_t)a - 16ll * j;     // bytes from hi on follow the item
        const int64_t glo = c0 == 0 ? lo : -1;                     // bytes below glo precede the record
        const int64_t ghi = (int64_t)(ncols - c0) + (int64_t)a - 16ll * j;  // bytes from ghi on follow it
        const bool interior = lo <= 0 && hi >= 16;
        uint32_t cm[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t ctx = kHi;
            cm[k] = kHi;
            if (!interior) {
                int f0 = (int)(lo - 4 * k), t0 = (int)(hi - 4 * k > 4 ? 4 : hi - 4 * k);
                cm[k] = range_flags(f0 < 0 ? 0 : f0, t0);
                int f1 = glo < 0 ? 0 : (int)(glo - 4 * k > 4 ? 4 : glo - 4 * k);
                int t1 = (int)(ghi - 4 * k > 4 ? 4 : ghi - 4 * k);
                ctx = range_flags(f1 < 0 ? 0 : f1, t1);
            }
            F[k] = classify4<CALLER>(tw[k], qw[k], ctx);
        }
        Flags4 pf = classify_prev<CALLER>(pv & 0xFFu, (pv >> 8) & 0xFFu, (pv >> 16) & 1u);
        uint32_t p_eq = pf.eq, p_x = pf.x, p_i = pf.i, p_d = pf.d, p_w = pf.w;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t Pi = prev_flags(p_i, F[k].i), Pd = prev_flags(p_d, F[k].d);
            if (COUNTS) {
                acc.cX += __popc(F[k].x & cm[k]);
                acc.cI += __popc(F[k].i & cm[k]);
                acc.cD += __popc(F[k].d & cm[k]);
                acc.eI += __popc(F[k].i & ~Pi & cm[k]);
                acc.eD += __popc(F[k].d & ~Pd & cm[k]);
            }
            if (RUNS) {
                uint32_t Pe = prev_flags(p_eq, F[k].eq), Px = prev_flags(p_x, F[k].x);
                uint32_t same = (F[k].eq & Pe) | (F[k].x & Px) | (F[k].i & Pi) | (F[k].d & Pd);
                if (CALLER) same |= F[k].w & prev_flags(p_w, F[k].w);
                H[k] = cm[k] & ~same;
                hc += __popc(H[k]);
            }
            p_eq = F[k].eq; p_x = F[k].x; p_i = F[k].i; p_d = F[k].d; p_w = F[k].w;
        }
    }
    if (RUNS) {
        if (!EMIT) {
            acc.nheads += hc;
        } else {
            uint32_t tot;
            uint32_t off = warp_excl_scan(hc, &tot);
            uint64_t o = acc.out + off;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t m = H[k];
                while (m) {
                    int bit = __ffs(m) - 1;
                    m &= m - 1;
                    uint32_t cls = ((F[k].x >> bit) & 1u) * CLS_X + ((F[k].i >> bit) & 1u) * CLS_I +
                                   ((F[k].d >> bit) & 1u) * CLS_D + ((F[k].w >> bit) & 1u) * CLS_W;
                    A.run_pos[o] = c0 + (uint64_t)(16u * j + 4u * k + (uint32_t)(bit >> 3)) - a;
                    A.run_rec[o] = rec;
                    A.run_cls[o] = (uint8_t)cls;
                    ++o;
                }
            }
            acc.out += tot;
        }
    }
}

template <int MODE, bool EMIT>
__global__ void __launch_bounds__(kWalkThreads, kWalkCtasPerSm) k_walk(WalkArgs A) {
    constexpr bool CALLER = MODE == MODE_RUNS_CALLER;
    constexpr bool RUNS = MODE != MODE_STAT;
    constexpr bool COUNTS = MODE == MODE_STAT || (MODE == MODE_RUNS_EXT && !EMIT);
    __shared__ uint32_t s_queue[kWalkWarps][10][64];  // per warp: j, prev, t[4], q[4] (SoA, 64 slots)
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t (*Q)[64] = s_queue[wid];
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint64_t n_warps = (uint64_t)gridDim.x * kWalkWarps;

    for (uint64_t item = (uint64_t)blockIdx.x * kWalkWarps + wid; item < A.n_items; item += n_warps) {
        const uint4 *dp = reinterpret_cast<const uint4 *>(A.items + item);
        const uint4 d0 = __ldg(dp), d1 = __ldg(dp + 1);
        const uint64_t t_off = ((uint64_t)d0.y << 32) | d0.x, q_off = ((uint64_t)d0.w << 32) | d0.z;
        const uint32_t n = d1.x, rec = d1.y, seg = d1.z, single = d1.w;
        const uint64_t c0 = (uint64_t)seg * kSeg;
        const uint64_t ncols = A.ncols[rec];
        const uint8_t *tp = A.text + t_off;
        const uint8_t *qp = A.text + q_off;
        const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(tp) & 15u);
        const uint8_t *tbase = tp - a;  // 16-byte aligned; chunk j covers item columns [16j-a, 16j-a+16)
        const uint8_t *qv = qp - a;     // query byte stream in the same column frame
        const uint32_t s = (uint32_t)(reinterpret_cast<uintptr_t>(qv) & 15u);
        const uint8_t *qbase = qv - s;
        const uint32_t nch = (n + a + 15u) >> 4;
        const uint32_t q_last = (s + n + a - 1u) >> 4;  // last query vector holding a needed byte
        const uint32_t ws = s >> 2, bs = s & 3u;

        WalkAcc acc;
        acc.cX = acc.cI = acc.cD = acc.eI = acc.eD = acc.nheads = 0;
        acc.out = 0;
        if (RUNS && EMIT) acc.out = A.run_base[item];
        // column preceding chunk 0 (only matters when a == 0; otherwise it sits inside chunk 0)
        uint32_t carry_t = 0, carry_q = 0, first_prev = 0;
        if (a == 0 && seg > 0) first_prev = (1u << 16) | ((uint32_t)qp[-1] << 8) | (uint32_t)tp[-1];
        uint32_t qn = 0;

        // software pipeline: vectors of step k+1 are requested before step k is examined
        uint4 tv = make_uint4(0, 0, 0, 0), q0 = tv, q1 = tv;
        {
            const uint32_t j = lane;
            if (j < nch) {
                tv = ldg16(tbase + 16ull * j);
                q0 = ldg16(qbase + 16ull * j);
                if (s != 0 && j + 1 <= q_last) q1 = ldg16(qbase + 16ull * (j + 1));
            }
        }
        for (uint32_t j0 = 0; j0 < nch; j0 += 32) {
            const uint32_t j = j0 + lane;
            const bool act = j < nch;
            uint4 ntv = make_uint4(0, 0, 0, 0), nq0 = ntv, nq1 = ntv;
            {
                const uint32_t jn = j + 32;
                if (jn < nch) {
                    ntv = ldg16(tbase + 16ull * jn);
                    nq0 = ldg16(qbase + 16ull * jn);
                    if (s != 0 && jn + 1 <= q_last) nq1 = ldg16(qbase + 16ull * (jn + 1));
                }
            }
            uint32_t w5[5];
            select5(q0, q1, ws, w5);
            uint32_t tw[4] = {tv.x, tv.y, tv.z, tv.w}, qw[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) qw[k] = funnel_r_bytes(w5[k], w5[k + 1], bs);
            // last column of this chunk -> neighbour lane (previous column of the next chunk)
            const uint32_t tl = act ? (tw[3] >> 24) : 0u, ql = act ? (qw[3] >> 24) : 0u;
            uint32_t pt = __shfl_up_sync(0xFFFFFFFFu, tl, 1), pq = __shfl_up_sync(0xFFFFFFFFu, ql, 1);
            if (lane == 0) { pt = carry_t; pq = carry_q; }
            carry_t = __shfl_sync(0xFFFFFFFFu, tl, 31);
            carry_q = __shfl_sync(0xFFFFFFFFu, ql, 31);
            uint32_t x = (tw[0] ^ qw[0]) | (tw[1] ^ qw[1]) | (tw[2] ^ qw[2]) | (tw[3] ^ qw[3]);
            bool prev_ne = pt != pq;
            if (CALLER) {
                x |= has_zero_byte(tw[0] ^ kGap4) | has_zero_byte(tw[1] ^ kGap4) | has_zero_byte(tw[2] ^ kGap4) |
                     has_zero_byte(tw[3] ^ kGap4);
                prev_ne |= pt == '-';
            }
            const bool interior = 16u * j >= a && 16u * j + 16u <= n + a;
            const bool slow = act && (j == 0 || !interior || x != 0u || prev_ne);
            const uint32_t mask = __ballot_sync(0xFFFFFFFFu, slow);
            if (mask) {
                if (slow) {
                    const uint32_t slot = qn + __popc(mask & lt_mask);
                    Q[0][slot] = j;
                    Q[1][slot] = j == 0 ? first_prev : ((1u << 16) | (pq << 8) | pt);
#pragma unroll
                    for (int k = 0; k < 4; ++k) { Q[2 + k][slot] = tw[k]; Q[6 + k][slot] = qw[k]; }
                }
                qn += __popc(mask);
                __syncwarp();
                if (qn >= 32) {
                    walk_batch<MODE, EMIT>(A, Q, 32, lane, a, n, c0, ncols, rec, acc);
                    __syncwarp();
                    qn -= 32;
                    uint32_t mv[10];
                    if (lane < qn) {
#pragma unroll
                        for (int f = 0; f < 10; ++f) mv[f] = Q[f][lane + 32];
                    }
                    __syncwarp();
                    if (lane < qn) {
#pragma unroll
                        for (int f = 0; f < 10; ++f) Q[f][lane] = mv[f];
                    }
                    __syncwarp();
                }
            }
            tv = ntv; q0 = nq0; q1 = nq1;
        }
        if (qn) {
            walk_batch<MODE, EMIT>(A, Q, qn, lane, a, n, c0, ncols, rec, acc);
            __syncwarp();
        }
        if (COUNTS) {
            uint32_t cX = warp_sum(acc.cX), cI = warp_sum(acc.cI), cD = warp_sum(acc.cD), eI = warp_sum(acc.eI),
                     eD = warp_sum(acc.eD);
            if (lane == 0) {
                unsigned long long *c = A.rec_cnt + (uint64_t)rec * 5;
                if (single) {
                    c[0] = cX; c[1] = eI; c[2] = cI; c[3] = eD; c[4] = cD;
                } else {
                    if (cX) atomicAdd(c + 0, (unsigned long long)cX);
                    if (eI) atomicAdd(c + 1, (unsigned long long)eI);
                    if (cI) atomicAdd(c + 2, (unsigned long long)cI);
                    if (eD) atomicAdd(c + 3, (unsigned long long)eD);
                    if (cD) atomicAdd(c + 4, (unsigned long long)cD);
                }
            }
        }
        if (RUNS && !EMIT) {
            uint32_t nh = warp_sum(acc.nheads);
            if (lane == 0) A.item_nheads[item] = nh;
        }
    }
}

}  // namespace wga
