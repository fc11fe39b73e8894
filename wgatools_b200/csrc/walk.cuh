// walk.cuh -- the dense column walk: the only kernel that streams the gapped sequences.
//
// Replaces the per-column group_by of src/parser/cigar.rs:344-432 (parse_maf_seq_to_cigar, classifier
// cigar_cat_ext :298-308) and of src/tools/caller.rs:442-446 (classifier cigar_cat_ext_caller,
// cigar.rs:314-328).  One warp walks one work item (= up to kSeg columns of one record); each lane
// handles 16 columns per step from one aligned 16-byte load of the target and a re-aligned pair of
// 16-byte loads of the query, classifies the 16 columns with byte-SIMD flag words and
//   MODE_STAT  : accumulates X / I / D column counts and I / D run starts (events);
//   MODE_RUNS_*: additionally finds every run head (class change) and either counts them (EMIT=false)
//                or writes (column, class) entries in column order (EMIT=true).
// Algorithmic traffic: 2 bytes per column (+ ~64 B of record metadata per item).
#pragma once
#include "ctx.cuh"
#include "tokenize.cuh"

namespace wga {

enum { MODE_STAT = 0, MODE_RUNS_EXT = 1, MODE_RUNS_CALLER = 2 };
constexpr int kWalkThreads = 256;
constexpr int kWalkWarps = kWalkThreads / 32;

struct WalkArgs {
    const uint8_t *text;
    const uint64_t *t_off, *q_off, *ncols;
    const uint32_t *rec_n_items;
    const uint32_t *item_rec, *item_seg;
    uint64_t n_items;
    unsigned long long *rec_cnt;  // [n_rec][5]: X, I_events, I_cols, D_events, D_cols
    uint32_t *item_nheads;        // RUNS, EMIT=false
    const uint64_t *run_base;     // RUNS, EMIT=true
    uint64_t *run_pos;
    uint32_t *run_rec;
    uint8_t *run_cls;
};

// 5 consecutive words starting at word index ws (0..3) of the 8-word window {a,b}
__device__ __forceinline__ void select5(const uint4 &a, const uint4 &b, uint32_t ws, uint32_t w[5]) {
    switch (ws) {
        case 0: w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; break;
        case 1: w[0] = a.y; w[1] = a.z; w[2] = a.w; w[3] = b.x; w[4] = b.y; break;
        case 2: w[0] = a.z; w[1] = a.w; w[2] = b.x; w[3] = b.y; w[4] = b.z; break;
        default: w[0] = a.w; w[1] = b.x; w[2] = b.y; w[3] = b.z; w[4] = b.w; break;
    }
}

template <int MODE, bool EMIT>
__global__ void __launch_bounds__(kWalkThreads) k_walk(WalkArgs A) {
    constexpr bool CALLER = MODE == MODE_RUNS_CALLER;
    constexpr bool RUNS = MODE != MODE_STAT;
    constexpr bool COUNTS = MODE == MODE_STAT || (MODE == MODE_RUNS_EXT && !EMIT);  // counted once
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t item = (uint64_t)blockIdx.x * kWalkWarps + (threadIdx.x >> 5);
    if (item >= A.n_items) return;
    const uint32_t rec = A.item_rec[item], seg = A.item_seg[item];
    const uint64_t ncols = A.ncols[rec];
    const uint64_t c0 = (uint64_t)seg * kSeg;
    const uint32_t n = (uint32_t)(ncols - c0 < (uint64_t)kSeg ? ncols - c0 : (uint64_t)kSeg);
    const uint8_t *tp = A.text + A.t_off[rec] + c0;
    const uint8_t *qp = A.text + A.q_off[rec] + c0;
    const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(tp) & 15u);
    const uint8_t *tbase = tp - a;  // 16-byte aligned; chunk j covers item columns [16j-a, 16j-a+16)
    const uint8_t *qv = qp - a;     // query byte stream in the same column frame
    const uint32_t s = (uint32_t)(reinterpret_cast<uintptr_t>(qv) & 15u);
    const uint8_t *qbase = qv - s;
    const uint32_t nch = (n + a + 15u) >> 4;
    const uint32_t q_last = (s + n + a - 1u) >> 4;  // last query vector holding a needed byte
    // bytes of a chunk that are real columns of this record: global column in [0, ncols)
    const int64_t g_lo = c0 == 0 ? (int64_t)a : -1;  // first chunk: bytes below a precede the record
    const int64_t g_hi = (int64_t)(ncols - c0) + (int64_t)a;

    uint32_t cX = 0, cI = 0, cD = 0, eI = 0, eD = 0, nheads = 0;
    uint64_t out = 0;
    if (RUNS && EMIT) out = A.run_base[item];
    uint32_t carry = 0;  // class bits of the column before this step's first chunk
    if (a == 0 && seg > 0) {  // previous column lives in the previous item of the same record
        Flags4 pf = classify_prev<CALLER>(tp[-1], qp[-1], true);
        carry = (pf.eq >> 31) | ((pf.x >> 31) << 1) | ((pf.i >> 31) << 2) | ((pf.d >> 31) << 3) |
                ((pf.w >> 31) << 4);
    }

    for (uint32_t j0 = 0; j0 < nch; j0 += 32) {
        const uint32_t j = j0 + lane;
        const bool act = j < nch;
        Flags4 F[4];
        uint32_t cnt[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { F[k].eq = F[k].x = F[k].i = F[k].d = F[k].w = 0; cnt[k] = 0; }
        if (act) {
            uint4 tv = ldg16(tbase + 16ull * j);
            uint4 q0 = ldg16(qbase + 16ull * j);
            uint4 q1 = make_uint4(0, 0, 0, 0);
            if (s != 0 && j + 1 <= q_last) q1 = ldg16(qbase + 16ull * (j + 1));
            uint32_t w5[5];
            select5(q0, q1, s >> 2, w5);
            const uint32_t bs = s & 3u;
            uint32_t tw[4] = {tv.x, tv.y, tv.z, tv.w}, qw[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) qw[k] = funnel_r_bytes(w5[k], w5[k + 1], bs);
            const int64_t lo = (int64_t)a - 16ll * j;        // bytes below lo precede the item
            const int64_t hi = (int64_t)n + (int64_t)a - 16ll * j;  // bytes from hi on follow the item
            const int64_t glo = g_lo < 0 ? -1 : g_lo - 16ll * j;
            const int64_t ghi = g_hi - 16ll * j;
            const bool interior = lo <= 0 && hi >= 16;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t ctx = kHi;
                cnt[k] = kHi;
                if (!interior) {
                    int f0 = (int)(lo - 4 * k), t0 = (int)(hi - 4 * k > 4 ? 4 : hi - 4 * k);
                    cnt[k] = range_flags(f0 < 0 ? 0 : f0, t0);
                    int f1 = glo < 0 ? 0 : (int)(glo - 4 * k > 4 ? 4 : glo - 4 * k);
                    int t1 = (int)(ghi - 4 * k > 4 ? 4 : ghi - 4 * k);
                    ctx = range_flags(f1 < 0 ? 0 : f1, t1);
                }
                F[k] = classify4<CALLER>(tw[k], qw[k], ctx);
            }
        }
        // class of the column preceding this chunk: neighbour lane / carry from the previous step
        uint32_t pk = (F[3].eq >> 31) | ((F[3].x >> 31) << 1) | ((F[3].i >> 31) << 2) |
                      ((F[3].d >> 31) << 3) | ((F[3].w >> 31) << 4);
        uint32_t up = __shfl_up_sync(0xFFFFFFFFu, pk, 1);
        if (lane == 0) up = carry;
        carry = __shfl_sync(0xFFFFFFFFu, pk, 31);

        uint32_t H[4] = {0, 0, 0, 0};
        uint32_t hc = 0;
        {
            uint32_t p_eq = (up & 1u) << 31, p_x = ((up >> 1) & 1u) << 31, p_i = ((up >> 2) & 1u) << 31,
                     p_d = ((up >> 3) & 1u) << 31, p_w = ((up >> 4) & 1u) << 31;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t Pi = prev_flags(p_i, F[k].i), Pd = prev_flags(p_d, F[k].d);
                if (COUNTS) {
                    cX += __popc(F[k].x & cnt[k]);
                    cI += __popc(F[k].i & cnt[k]);
                    cD += __popc(F[k].d & cnt[k]);
                    eI += __popc(F[k].i & ~Pi & cnt[k]);
                    eD += __popc(F[k].d & ~Pd & cnt[k]);
                }
                if (RUNS) {
                    uint32_t Pe = prev_flags(p_eq, F[k].eq), Px = prev_flags(p_x, F[k].x);
                    uint32_t same = (F[k].eq & Pe) | (F[k].x & Px) | (F[k].i & Pi) | (F[k].d & Pd);
                    if (CALLER) same |= F[k].w & prev_flags(p_w, F[k].w);
                    H[k] = cnt[k] & ~same;
                    hc += __popc(H[k]);
                }
                p_eq = F[k].eq; p_x = F[k].x; p_i = F[k].i; p_d = F[k].d; p_w = F[k].w;
            }
        }
        if (RUNS) {
            if (!EMIT) {
                nheads += hc;
            } else {
                uint32_t tot;
                uint32_t off = warp_excl_scan(hc, &tot);
                uint64_t o = out + off;
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
                out += tot;
            }
        }
    }
    if (COUNTS) {
        cX = warp_sum(cX); cI = warp_sum(cI); cD = warp_sum(cD); eI = warp_sum(eI); eD = warp_sum(eD);
        if (lane == 0) {
            unsigned long long *c = A.rec_cnt + (uint64_t)rec * 5;
            if (A.rec_n_items[rec] == 1) {
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
        nheads = warp_sum(nheads);
        if (lane == 0) A.item_nheads[item] = nheads;
    }
}

}  // namespace wga
