// walk.cuh -- the dense column walk: the only kernel that streams the gapped sequences.
//
// Replaces the per-column group_by of src/parser/cigar.rs:344-432 (parse_maf_seq_to_cigar, classifier
// cigar_cat_ext :298-308) and of src/tools/caller.rs:442-446 (classifier cigar_cat_ext_caller,
// cigar.rs:314-328).
//
// Shape (v4).  Persistent warps; a warp walks one work item (<= kSeg columns of one record) at a time.
// Lane i of step k owns the 16 columns of chunk j = 32k+i: one aligned 16-byte load of the target and a
// funnel-shifted pair of 16-byte loads of the query (the two streams are misaligned against each other by
// an arbitrary number of bytes).  Alignments are >99 % identical columns, so the dense part of the kernel
// is only a 128-bit compare: a chunk whose 16 columns and whose preceding column are all '=' can neither
// contain nor start a run and is skipped.  Every other chunk ("slow": contains a mismatch/gap, follows
// one, or touches the item boundary) is pushed -- data words and item context included -- into a per-warp
// queue in shared memory.  Whenever 32 slow chunks are queued (from any number of items) the warp
// classifies them in lock-step, one chunk per lane, with the byte-SIMD flag words of hd_core.h:
//   - X / I / D column counts and I / D run starts (events) -> atomics on the record's counter row;
//   - run heads (class changes) -> (column-in-item, class) entries appended, in column order, to the
//     item's slot region (capacity n/8+16 by default; an overflow is reported and the caller re-runs with
//     worst-case capacity), so maf2paf / call need ONE dense pass; a sparse kernel compacts the slots.
// Work distribution is dynamic (batches of 8 items from a global counter) and the warp requests the first vectors
// of its next item during the last step of the current one.
// Algorithmic traffic: 2 bytes per column (+ 48 B of item descriptor).
#pragma once
#include "walk_types.cuh"

namespace wga {

#ifndef WGA_WALK_NO_SPARSE  // the non-template helper kernels are defined by walk_kernels.cu only
// rec_slot_base: exclusive scan over records of the summed capacities of their items
__global__ void k_rec_slot_need(uint64_t n_rec, const uint32_t *__restrict__ n_items, const uint64_t *__restrict__ ncols,
                                uint32_t shift, uint64_t *__restrict__ need) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rec) return;
    uint64_t v = 0;
    if (r < n_rec) {
        uint32_t n = n_items[r];
        if (n) v = (uint64_t)(n - 1) * slot_cap(kSeg, shift) + slot_cap(ncols[r] - (uint64_t)(n - 1) * kSeg, shift);
    }
    need[r] = v;
}

__global__ void k_fill_item_desc(uint64_t n_rec, const uint32_t *__restrict__ n_items,
                                 const uint64_t *__restrict__ item_base, const uint64_t *__restrict__ t_off,
                                 const uint64_t *__restrict__ q_off, const uint64_t *__restrict__ ncols,
                                 const uint64_t *__restrict__ rec_slot_base, uint32_t shift,
                                 ItemDesc *__restrict__ items) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    uint64_t b = item_base[r], nc = ncols[r];
    uint32_t n = n_items[r];
    uint64_t sb = rec_slot_base ? rec_slot_base[r] : 0;
    for (uint32_t s = 0; s < n; ++s) {
        ItemDesc d;
        uint64_t c0 = (uint64_t)s * kSeg;
        d.t_off = t_off[r] + c0;
        d.q_off = q_off[r] + c0;
        d.slot_base = sb + (uint64_t)s * slot_cap(kSeg, shift);
        d.n = (uint32_t)(nc - c0 < (uint64_t)kSeg ? nc - c0 : (uint64_t)kSeg);
        uint64_t tail = nc - c0 - d.n;
        d.rec = (uint32_t)r;
        d.seg = s;
        d.tail = (uint32_t)(tail > 16 ? 16 : tail);
        d.pad0 = d.pad1 = 0;
        items[b + s] = d;
    }
}

#endif  // WGA_WALK_NO_SPARSE

__device__ __forceinline__ uint32_t has_zero_byte(uint32_t v) { return (v - 0x01010101u) & ~v & kHi; }
__device__ __forceinline__ int clamp16(int32_t v) { return v < 0 ? 0 : (v > 16 ? 16 : v); }

struct WarpEmit {  // run-head emission state of the warp: the item whose slot region is still being filled
    uint32_t open_item;   // 0xFFFFFFFF = none
    uint32_t open_count;
};

// classify one queued chunk per lane (lanes >= cnt idle); counts -> atomics, heads -> item slots in order
// RECQ: the record index travels in queue field 11 (false: it is read from the item descriptor instead, which
// saves 256 B of shared memory per warp for the TMA-staged variant)
// queue layouts: walk_batch reads one entry per lane through load()
template <bool RECQ>
struct QueueSoA {  // k_walk: one 64-slot row per field
    uint32_t (*Q)[64];
    static constexpr bool kHasRec = RECQ;
    __device__ __forceinline__ void load(uint32_t lane, uint32_t (&tw)[4], uint32_t (&qw)[4], uint32_t &pv, uint32_t &meta,
                                         uint32_t &item, uint32_t &rec) const {
#pragma unroll
        for (int k = 0; k < 4; ++k) { tw[k] = Q[k][lane]; qw[k] = Q[4 + k][lane]; }
        pv = Q[8][lane];
        meta = Q[9][lane];
        item = Q[10][lane];
        rec = RECQ ? Q[11][lane] : 0u;
    }
};

template <int MODE, class QA>
__device__ __forceinline__ void walk_batch_q(const WalkArgs &A, const QA &qa, uint32_t cnt, uint32_t lane, WarpEmit &E) {
    constexpr bool CALLER = MODE == MODE_RUNS_CALLER || MODE == MODE_RUNS_ALL;
    constexpr bool RUNS = MODE != MODE_STAT;
    constexpr bool COUNTS = MODE != MODE_RUNS_CALLER;
    uint32_t H[4] = {0, 0, 0, 0}, CW[4] = {0, 0, 0, 0}, hc = 0, item = 0xFFFFFFFFu, colrel = 0;  // CW: class code per column
    const bool act = lane < cnt;
    if (act) {
        uint32_t tw[4], qw[4];
        uint32_t pv;    // bit16: valid, bits 0-7 target byte, 8-15 query byte
        uint32_t meta;  // f0 | t0<<5 | t1<<10 | seg0<<15 | colrel<<16 (column of byte 0 + 16, in item)
        uint32_t rec_q;
        qa.load(lane, tw, qw, pv, meta, item, rec_q);
        const int f0 = meta & 31, t0 = (meta >> 5) & 31, t1 = (meta >> 10) & 31;
        const int f1 = ((meta >> 15) & 1u) ? f0 : 0;  // bytes before the record start are not columns
        colrel = meta >> 16;
        uint32_t cX = 0, cI = 0, cD = 0, eI = 0, eD = 0, hw = 0;
        Flags4 pf = classify_prev<CALLER>(pv & 0xFFu, (pv >> 8) & 0xFFu, (pv >> 16) & 1u);
        uint32_t p_eq = pf.eq, p_x = pf.x, p_i = pf.i, p_d = pf.d, p_w = pf.w;
        // bytes [f0, t0) of the chunk are the item's columns, [f1, t1) the record's (context): one bit per byte, a nibble per word
        const uint32_t vm = ((1u << t0) - 1u) & ~((1u << f0) - 1u), vc = ((1u << t1) - 1u) & ~((1u << f1) - 1u);
        Flags4 F[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t cm = nibble_flags(vm >> (4 * k)), ctx = nibble_flags(vc >> (4 * k));
            F[k] = classify4<CALLER>(tw[k], qw[k], ctx);
            uint32_t Pi = prev_flags(p_i, F[k].i), Pd = prev_flags(p_d, F[k].d);
            if (COUNTS) {
                cX += __popc(F[k].x & cm);
                cI += __popc(F[k].i & cm);
                cD += __popc(F[k].d & cm);
                eI += __popc(F[k].i & ~Pi & cm);
                eD += __popc(F[k].d & ~Pd & cm);
            }
            if (RUNS) {
                uint32_t Pe = prev_flags(p_eq, F[k].eq), Px = prev_flags(p_x, F[k].x);
                uint32_t same = (F[k].eq & Pe) | (F[k].x & Px) | (F[k].i & Pi) | (F[k].d & Pd);
                if (CALLER) same |= F[k].w & prev_flags(p_w, F[k].w);
                H[k] = cm & ~same;
                hc += __popc(H[k]);
                if (MODE == MODE_RUNS_ALL) hw |= H[k] & F[k].w;
                CW[k] = class_codes(F[k]);
            }
            p_eq = F[k].eq; p_x = F[k].x; p_i = F[k].i; p_d = F[k].d; p_w = F[k].w;
        }
        if (MODE == MODE_RUNS_ALL && hw) atomicAdd(A.n_w, 1u);  // only "are there any" matters
        if (COUNTS && (cX | cI | cD)) {
            const uint32_t rec = QA::kHasRec ? rec_q : __ldg(&A.items[item].rec);
            unsigned long long *c = A.rec_cnt + (uint64_t)rec * 5;
            if (cX) atomicAdd(c + 0, (unsigned long long)cX);
            if (eI) atomicAdd(c + 1, (unsigned long long)eI);
            if (cI) atomicAdd(c + 2, (unsigned long long)cI);
            if (eD) atomicAdd(c + 3, (unsigned long long)eD);
            if (cD) atomicAdd(c + 4, (unsigned long long)cD);
        }
    }
    if (RUNS) {
        // segmented exclusive scan of hc keyed by item (lanes of one item are contiguous)
        uint32_t x = hc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
            uint32_t it = __shfl_up_sync(0xFFFFFFFFu, item, d);
            if (lane >= (uint32_t)d && it == item) x += y;
        }
        const uint32_t first_item = __shfl_sync(0xFFFFFFFFu, item, 0);
        const uint32_t next_item = __shfl_down_sync(0xFFFFFFFFu, item, 1);
        // the previously open item is complete once the batch starts with another item
        if (E.open_item != 0xFFFFFFFFu && first_item != E.open_item) {
            if (lane == 0) A.item_nheads[E.open_item] = E.open_count;
            E.open_item = 0xFFFFFFFFu;
            E.open_count = 0;
        }
        const uint32_t carry = (item == E.open_item) ? E.open_count : 0u;
        const uint32_t incl = carry + x;  // heads of this item up to and including this lane
        if (act) {
            const uint4 *dp = reinterpret_cast<const uint4 *>(A.items + item);
            const uint4 d1 = __ldg(dp + 1);
            const uint64_t sbase = ((uint64_t)d1.y << 32) | d1.x;
            const uint32_t cap = (uint32_t)slot_cap(d1.z, A.cap_shift);
            uint32_t o = incl - hc;
            if (incl > cap) {
                atomicOr(A.overflow, 1u);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t m = H[k];
                    while (m) {
                        int bit = __ffs(m) - 1;  // 7, 15, 23 or 31: the flag bit of byte bit >> 3
                        m &= m - 1;
                        uint32_t cls = (CW[k] >> (bit - 7)) & 7u;
                        uint32_t col = colrel + 4u * k + (uint32_t)(bit >> 3) - 16u;
                        A.slots[sbase + o] = (col << 3) | cls;
                        ++o;
                    }
                }
            }
            const bool seg_end = lane + 1 == cnt || next_item != item;
            if (seg_end && lane + 1 != cnt) A.item_nheads[item] = incl;  // item complete inside the batch
        }
        // the last lane's item stays open
        const uint32_t last = cnt - 1;
        E.open_item = __shfl_sync(0xFFFFFFFFu, item, last);
        E.open_count = __shfl_sync(0xFFFFFFFFu, incl, last);
    }
}

template <int MODE, bool RECQ = true>
__device__ __forceinline__ void walk_batch(const WalkArgs &A, uint32_t (*Q)[64], uint32_t cnt, uint32_t lane, WarpEmit &E) {
    walk_batch_q<MODE>(A, QueueSoA<RECQ>{Q}, cnt, lane, E);
}

// geometry of one work item in the 16-byte frame of its target stream
struct Geom {
    const uint8_t *tbase, *qbase;  // 16-byte aligned; chunk j covers item columns [16j-a, 16j-a+16)
    uint32_t a, s, n, nch, q_last, rec, seg, tail, item, first_prev;
};

__device__ __forceinline__ Geom make_geom(const WalkArgs &A, const uint4 &d0, const uint4 &d1, const uint4 &d2, uint32_t item) {
    Geom g;
    const uint64_t t_off = ((uint64_t)d0.y << 32) | d0.x, q_off = ((uint64_t)d0.w << 32) | d0.z;
    g.n = d1.z; g.rec = d1.w; g.seg = d2.x; g.tail = d2.y; g.item = item;
    const uint8_t *tp = A.text + t_off, *qp = A.text + q_off;
    g.a = (uint32_t)(reinterpret_cast<uintptr_t>(tp) & 15u);
    g.tbase = tp - g.a;
    const uint8_t *qv = qp - g.a;  // query byte stream in the same column frame
    g.s = (uint32_t)(reinterpret_cast<uintptr_t>(qv) & 15u);
    g.qbase = qv - g.s;
    g.nch = (g.n + g.a + 15u) >> 4;
    g.q_last = (g.s + g.n + g.a - 1u) >> 4;  // last query vector holding a needed byte
    // column preceding chunk 0 (only matters when a == 0; otherwise it sits inside chunk 0)
    g.first_prev = 0;
    if (g.a == 0 && g.seg > 0) g.first_prev = (1u << 16) | ((uint32_t)qp[-1] << 8) | (uint32_t)tp[-1];
    return g;
}

struct Vec3 { uint4 t, q0, q1; };
__device__ __forceinline__ Vec3 load_vec(const Geom &g, uint32_t j) {
    Vec3 v;
    v.t = make_uint4(0, 0, 0, 0); v.q0 = v.t; v.q1 = v.t;
    if (j < g.nch) {
        v.t = ldg16(g.tbase + 16ull * j);
        v.q0 = ldg16(g.qbase + 16ull * j);
        if (g.s != 0 && j + 1 <= g.q_last) v.q1 = ldg16(g.qbase + 16ull * (j + 1));
    }
    return v;
}

template <int WS>
__device__ __forceinline__ void realign(const uint4 &q0, const uint4 &q1, uint32_t sh, uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3) {
    uint32_t w0, w1, w2, w3, w4;
    if (WS == 0) { w0 = q0.x; w1 = q0.y; w2 = q0.z; w3 = q0.w; w4 = q1.x; }
    else if (WS == 1) { w0 = q0.y; w1 = q0.z; w2 = q0.w; w3 = q1.x; w4 = q1.y; }
    else if (WS == 2) { w0 = q0.z; w1 = q0.w; w2 = q1.x; w3 = q1.y; w4 = q1.z; }
    else { w0 = q0.w; w1 = q1.x; w2 = q1.y; w3 = q1.z; w4 = q1.w; }
    o0 = __funnelshift_r(w0, w1, sh); o1 = __funnelshift_r(w1, w2, sh);
    o2 = __funnelshift_r(w2, w3, sh); o3 = __funnelshift_r(w3, w4, sh);
}

// The column before a step's first chunk, kept lazily: its "is not plain '='" bit costs nothing (it falls out of the
// ballot below); its bytes are only needed when lane 0 of the next step turns out slow, so every lane just remembers its
// own last words and the slow branch fetches lane 31's.
struct PrevCol {
    uint32_t bit, tw3, qw3;
};

// one step (32 chunks) of the current item: 128-bit compare, push the slow chunks, run a batch when 32 are queued
// tprev / qprev (TMA-staged variant): shared-memory address of the target / query byte of the column before this lane's
// chunk, valid for every lane but the first chunk of a window (first_in_win); null = take them from the neighbour lane
template <int MODE, bool RECQ = true>
__device__ __forceinline__ void walk_step(const WalkArgs &A, uint32_t (*Q)[64], uint32_t &qn, WarpEmit &E, uint32_t lane,
                                          const Geom &g, uint32_t j0, const Vec3 &v, PrevCol &P, const uint8_t *tprev = nullptr,
                                          const uint8_t *qprev = nullptr, bool first_in_win = true) {
    constexpr bool CALLER = MODE == MODE_RUNS_CALLER || MODE == MODE_RUNS_ALL;
    const uint32_t j = j0 + lane;
    const bool act = j < g.nch;
    const uint4 tv = v.t;
    uint32_t qw0, qw1, qw2, qw3;
    const uint32_t sh = (g.s & 3u) * 8u;
    switch (g.s >> 2) {  // warp-uniform
        case 0: realign<0>(v.q0, v.q1, sh, qw0, qw1, qw2, qw3); break;
        case 1: realign<1>(v.q0, v.q1, sh, qw0, qw1, qw2, qw3); break;
        case 2: realign<2>(v.q0, v.q1, sh, qw0, qw1, qw2, qw3); break;
        default: realign<3>(v.q0, v.q1, sh, qw0, qw1, qw2, qw3); break;
    }
    const uint32_t x3 = tv.w ^ qw3;
    uint32_t x = (tv.x ^ qw0) | (tv.y ^ qw1) | (tv.z ^ qw2) | x3;
    bool ne_last = x3 > 0x00FFFFFFu;  // the chunk's last column differs
    if (CALLER) {
        // a both-gap column is '=' for the byte compare but a run of its own for the caller (cigar.rs:314-328).  '-' is
        // 0x2D; every letter has bit 6 set, so "some target byte has bit 6 clear" is a cheap superset of "holds a gap";
        // the slow path classifies exactly.
        const uint32_t g6 = tv.x & tv.y & tv.z & tv.w;
        x |= ~g6 & 0x40404040u;
        ne_last |= (tv.w & 0x40000000u) == 0u;
    }
    ne_last &= act;
    const uint32_t nb = __ballot_sync(0xFFFFFFFFu, ne_last);
    const bool prev_ne = (((nb << 1) | P.bit) >> lane) & 1u;
    // interior chunks: fully inside the item -> 16j >= a and 16j+16 <= n+a
    const uint32_t jf = (g.a + 15u) >> 4, jl = (g.n + g.a) >> 4;
    const bool edge = j == 0 || j < jf || j >= jl;
    const bool slow = act && (edge || x != 0u || prev_ne);
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, slow);
    if (mask) {
        // bytes of the column before each chunk: the neighbour lane's last column, lane 31 of the previous step for lane 0
        uint32_t pb;
        if (tprev) {
            pb = 0;
            if (first_in_win && (mask & 1u)) pb = __shfl_sync(0xFFFFFFFFu, (P.tw3 >> 24) | ((P.qw3 >> 24) << 8), 31);  // warp-uniform
            if (slow && !(first_in_win && lane == 0)) pb = (uint32_t)tprev[0] | ((uint32_t)qprev[0] << 8);
        } else {
            pb = __shfl_up_sync(0xFFFFFFFFu, (tv.w >> 24) | ((qw3 >> 24) << 8), 1);
            const uint32_t cb = __shfl_sync(0xFFFFFFFFu, (P.tw3 >> 24) | ((P.qw3 >> 24) << 8), 31);
            if (lane == 0) pb = cb;
        }
        if (slow) {
            const uint32_t slot = qn + __popc(mask & ((1u << lane) - 1u));
            Q[0][slot] = tv.x; Q[1][slot] = tv.y; Q[2][slot] = tv.z; Q[3][slot] = tv.w;
            Q[4][slot] = qw0; Q[5][slot] = qw1; Q[6][slot] = qw2; Q[7][slot] = qw3;
            Q[8][slot] = j == 0 ? g.first_prev : ((1u << 16) | pb);
            uint32_t f0 = 0, t0 = 16, t1 = 16;
            if (edge) {  // a <= 15, n <= kSeg, j < nch <= kSeg / 16 + 2: everything fits 32-bit signed arithmetic
                const int32_t lo = (int32_t)g.a - 16 * (int32_t)j, hi = (int32_t)(g.n + g.a) - 16 * (int32_t)j;
                f0 = clamp16(lo);
                t0 = clamp16(hi);
                t1 = clamp16(hi + (int32_t)g.tail);  // columns of the next item are context only
            }
            Q[9][slot] = f0 | (t0 << 5) | (t1 << 10) | ((g.seg == 0 ? 1u : 0u) << 15) | ((16u * j + 16u - g.a) << 16);
            Q[10][slot] = g.item;
            if (RECQ) Q[11][slot] = g.rec;
        }
        qn += __popc(mask);
        __syncwarp();
        if (qn >= 32) {
            walk_batch<MODE, RECQ>(A, Q, 32, lane, E);
            __syncwarp();
            qn -= 32;
            constexpr int NF = RECQ ? 12 : 11;
            uint32_t mv[NF];
            if (lane < qn) {
#pragma unroll
                for (int f = 0; f < NF; ++f) mv[f] = Q[f][lane + 32];
            }
            __syncwarp();
            if (lane < qn) {
#pragma unroll
                for (int f = 0; f < NF; ++f) Q[f][lane] = mv[f];
            }
            __syncwarp();
        }
    }
    P.bit = nb >> 31;
    P.tw3 = tv.w;
    P.qw3 = qw3;
}

template <int MODE>
__global__ void __launch_bounds__(kWalkThreads, kWalkCtasPerSm) k_walk(WalkArgs A) {
    constexpr bool RUNS = MODE != MODE_STAT;
    __shared__ uint32_t s_queue[kWalkWarps][12][64];  // per warp, SoA, 64 slots
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t (*Q)[64] = s_queue[wid];
    uint32_t qn = 0;
    WarpEmit E;
    E.open_item = 0xFFFFFFFFu;
    E.open_count = 0;

    // Dynamic scheduling: warps grab batches of kGrab consecutive items from a global counter (record lengths are
    // log-normal, a static split leaves a 30-50 % tail).  The next batch is always requested one batch ahead.
    // Software pipeline across items: the descriptor of the item after next and the first vectors of the next
    // item are requested while the current item is being walked.
    constexpr uint32_t kGrab = 8;
    const uint32_t n_items = (uint32_t)A.n_items;
    uint32_t s_it = 0, s_end = 0, nb_raw = 0;  // scheduler: last handed-out item, end of its batch, next batch (lane 0)
    if (lane == 0) {
        s_it = atomicAdd(A.sched, kGrab);
        nb_raw = atomicAdd(A.sched, kGrab);
    }
    s_it = __shfl_sync(0xFFFFFFFFu, s_it, 0);
    s_end = s_it + kGrab < n_items ? s_it + kGrab : n_items;
    auto advance = [&]() -> bool {  // moves (s_it) to the next item of this warp; false when the work is exhausted
        ++s_it;
        if (s_it < s_end) return true;
        const uint32_t nb = __shfl_sync(0xFFFFFFFFu, nb_raw, 0);
        if (nb >= n_items) { s_it = n_items; return false; }
        s_it = nb;
        s_end = nb + kGrab < n_items ? nb + kGrab : n_items;
        if (lane == 0) nb_raw = atomicAdd(A.sched, kGrab);
        return true;
    };
    if (s_it < n_items) {
        const uint32_t item = s_it;
        const uint4 *dp = reinterpret_cast<const uint4 *>(A.items + item);
        Geom cur = make_geom(A, __ldg(dp), __ldg(dp + 1), __ldg(dp + 2), item);
        bool has2 = advance();
        uint32_t item2 = s_it;
        uint4 e0 = make_uint4(0, 0, 0, 0), e1 = e0, e2 = e0;
        if (has2) {
            const uint4 *ep = reinterpret_cast<const uint4 *>(A.items + item2);
            e0 = __ldg(ep); e1 = __ldg(ep + 1); e2 = __ldg(ep + 2);
        }
        Vec3 v = load_vec(cur, lane);
        for (;;) {
            Geom nxt = cur;
            if (has2) nxt = make_geom(A, e0, e1, e2, item2);
            const bool has3 = has2 && advance();
            const uint32_t item3 = s_it;
            if (has3) {
                const uint4 *ep = reinterpret_cast<const uint4 *>(A.items + item3);
                e0 = __ldg(ep); e1 = __ldg(ep + 1); e2 = __ldg(ep + 2);
            }
            PrevCol carry = {0u, 0u, 0u};  // the column before this step's first chunk
            for (uint32_t j0 = 0; j0 < cur.nch; j0 += 32) {
                Vec3 nv;
                if (j0 + 32 < cur.nch) nv = load_vec(cur, j0 + 32 + lane);
                else if (has2) nv = load_vec(nxt, lane);
                else { nv.t = make_uint4(0, 0, 0, 0); nv.q0 = nv.t; nv.q1 = nv.t; }
                walk_step<MODE>(A, Q, qn, E, lane, cur, j0, v, carry);
                v = nv;
            }
            if (!has2) break;
            cur = nxt;
            item2 = item3;
            has2 = has3;
        }
    }
    if (qn) {
        walk_batch<MODE>(A, Q, qn, lane, E);
        __syncwarp();
    }
    if (RUNS && E.open_item != 0xFFFFFFFFu && lane == 0) A.item_nheads[E.open_item] = E.open_count;
}

#ifndef WGA_WALK_NO_SPARSE
// sparse: item slot regions -> dense run arrays (run_base = exclusive scan of item_nheads)
__global__ void k_runs_compact(uint64_t n_items, const ItemDesc *__restrict__ items, const uint32_t *__restrict__ item_nheads,
                               const uint64_t *__restrict__ run_base, const uint32_t *__restrict__ slots,
                               uint64_t *__restrict__ run_pos, uint32_t *__restrict__ run_rec, uint8_t *__restrict__ run_cls) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (warp >= n_items) return;
    const ItemDesc d = items[warp];
    const uint32_t nh = item_nheads[warp];
    const uint64_t rb = run_base[warp], c0 = (uint64_t)d.seg * kSeg;
    for (uint32_t i = lane; i < nh; i += 32) {
        uint32_t e = slots[d.slot_base + i];
        run_pos[rb + i] = c0 + (e >> 3);
        run_rec[rb + i] = d.rec;
        run_cls[rb + i] = (uint8_t)(e & 7u);
    }
}
#endif  // WGA_WALK_NO_SPARSE

}  // namespace wga
