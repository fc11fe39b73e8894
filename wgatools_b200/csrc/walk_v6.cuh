// walk.cuh -- the dense column walk: the only kernel that streams the gapped sequences.
//
// Replaces the per-column group_by of src/parser/cigar.rs:344-432 (parse_maf_seq_to_cigar, classifier
// cigar_cat_ext :298-308) and of src/tools/caller.rs:442-446 (classifier cigar_cat_ext_caller,
// cigar.rs:314-328).
//
// Shape (v3).  Persistent warps; a warp walks one work item (<= kSeg columns of one record) at a time.
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
// Algorithmic traffic: 2 bytes per column (+ 48 B of item descriptor).
#pragma once
#include "ctx.cuh"
#include "tokenize.cuh"

namespace wga {

enum { MODE_STAT = 0, MODE_RUNS_EXT = 1, MODE_RUNS_CALLER = 2 };
constexpr int kWalkThreads = 256;
constexpr int kWalkWarps = kWalkThreads / 32;
#ifndef WGA_WALK_CTAS
#define WGA_WALK_CTAS 3
#endif
constexpr int kWalkCtasPerSm = WGA_WALK_CTAS;
constexpr int kQWords = 14;  // queue entry: t[4] q[4] prev meta item rec col_lo col_hi
constexpr uint32_t kSlotPad = 16;

struct __align__(16) ItemDesc {  // one work item: columns [seg*kSeg, seg*kSeg+n) of record rec
    uint64_t t_off, q_off;       // byte offsets of the item's first column in target / query
    uint32_t n, rec, seg, tail;  // tail: columns of the record after this item, clipped to 16
    uint64_t slot_base;          // first slot of the item's run-head region (cold: only read when heads are emitted)
    uint64_t pad;
};
static_assert(sizeof(ItemDesc) == 48, "ItemDesc is read as three 16-byte vectors");

struct WalkArgs {
    const uint8_t *text;
    const ItemDesc *items;
    uint64_t n_items;
    unsigned long long *rec_cnt;  // [n_rec][5]: X, I_events, I_cols, D_events, D_cols (zero-initialised)
    uint32_t *item_nheads;        // [n_items] zero-initialised
    uint32_t *slots;              // run-head slots
    uint32_t cap_shift;           // slot capacity of an item = (n >> cap_shift) + kSlotPad
    uint32_t *overflow;
    unsigned int *sched;          // dynamic work counter (zero-initialised)
};

__device__ __host__ __forceinline__ uint64_t slot_cap(uint64_t n, uint32_t shift) { return (n >> shift) + kSlotPad; }

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
        d.pad = 0;
        items[b + s] = d;
    }
}

__device__ __forceinline__ uint32_t has_zero_byte(uint32_t v) { return (v - 0x01010101u) & ~v & kHi; }
__device__ __forceinline__ int clamp16(int64_t v) { return v < 0 ? 0 : (v > 16 ? 16 : (int)v); }

struct WarpEmit {  // run-head emission state of the warp: the item whose slot region is still being filled
    uint32_t open_item;   // 0xFFFFFFFF = none
    uint32_t open_count;
};

// classify one queued chunk per lane (lanes >= cnt idle); counts -> atomics, heads -> item slots in order
template <int MODE>
__device__ __forceinline__ void walk_batch(const WalkArgs &A, uint32_t (*Q)[64], uint32_t cnt, uint32_t lane, WarpEmit &E) {
    constexpr bool CALLER = MODE == MODE_RUNS_CALLER;
    constexpr bool RUNS = MODE != MODE_STAT;
    constexpr bool COUNTS = MODE != MODE_RUNS_CALLER;
    uint32_t H[4] = {0, 0, 0, 0}, hc = 0, item = 0xFFFFFFFFu, colrel = 0;
    Flags4 F[4];
    const bool act = lane < cnt;
    if (act) {
        uint32_t tw[4], qw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { tw[k] = Q[k][lane]; qw[k] = Q[4 + k][lane]; }
        const uint32_t pv = Q[8][lane];    // bit16: valid, bits 0-7 target byte, 8-15 query byte
        const uint32_t meta = Q[9][lane];  // f0 | t0<<5 | t1<<10 | seg0<<15 | colrel<<16 (column of byte 0 + 16, in item)
        item = Q[10][lane];
        const int f0 = meta & 31, t0 = (meta >> 5) & 31, t1 = (meta >> 10) & 31;
        const int f1 = ((meta >> 15) & 1u) ? f0 : 0;  // bytes before the record start are not columns
        colrel = meta >> 16;
        uint32_t cX = 0, cI = 0, cD = 0, eI = 0, eD = 0;
        Flags4 pf = classify_prev<CALLER>(pv & 0xFFu, (pv >> 8) & 0xFFu, (pv >> 16) & 1u);
        uint32_t p_eq = pf.eq, p_x = pf.x, p_i = pf.i, p_d = pf.d, p_w = pf.w;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t cm = range_flags(f0 - 4 * k, t0 - 4 * k), ctx = range_flags(f1 - 4 * k, t1 - 4 * k);
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
            }
            p_eq = F[k].eq; p_x = F[k].x; p_i = F[k].i; p_d = F[k].d; p_w = F[k].w;
        }
        if (COUNTS && (cX | cI | cD)) {
            const uint32_t rec = __ldg(reinterpret_cast<const uint4 *>(A.items + item) + 1).y;
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
            const uint4 d1 = __ldg(dp + 1), d2 = __ldg(dp + 2);
            const uint64_t sbase = ((uint64_t)d2.y << 32) | d2.x;
            const uint32_t cap = (uint32_t)slot_cap(d1.x, A.cap_shift);
            uint32_t o = incl - hc;
            if (incl > cap) {
                atomicOr(A.overflow, 1u);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t m = H[k];
                    while (m) {
                        int bit = __ffs(m) - 1;
                        m &= m - 1;
                        uint32_t cls = ((F[k].x >> bit) & 1u) * CLS_X + ((F[k].i >> bit) & 1u) * CLS_I +
                                       ((F[k].d >> bit) & 1u) * CLS_D + ((F[k].w >> bit) & 1u) * CLS_W;
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

// push the slow chunks of one step into the warp queue and run a batch when 32 are queued (cold path: kept out of
// line so that the four word-shift instantiations of the streaming loop stay small)
// run one batch of 32 queued chunks and shift the remainder of the queue down (out of line: it runs once per 32
// slow chunks, and keeping it out of the streaming loop keeps that loop's register footprint small)
template <int MODE>
__device__ __noinline__ uint3 walk_flush(const WalkArgs &A, uint32_t (*Q)[64], uint32_t qn, uint32_t open_item, uint32_t open_count) {
    const uint32_t lane = threadIdx.x & 31;
    WarpEmit E;
    E.open_item = open_item;
    E.open_count = open_count;
    walk_batch<MODE>(A, Q, 32, lane, E);
    __syncwarp();
    qn -= 32;
    uint32_t mv[11];
    if (lane < qn) {
#pragma unroll
        for (int f = 0; f < 11; ++f) mv[f] = Q[f][lane + 32];
    }
    __syncwarp();
    if (lane < qn) {
#pragma unroll
        for (int f = 0; f < 11; ++f) Q[f][lane] = mv[f];
    }
    __syncwarp();
    return make_uint3(qn, E.open_item, E.open_count);
}

// geometry of one work item in the 16-byte frame of its target stream (hot fields only: the record id, the slot
// base and the column preceding the item are re-read from the descriptor on the cold path)
struct G {
    const uint8_t *tptr, *qptr;  // lane-adjusted pointers to chunk `lane` (16-byte aligned)
    uint32_t a, s, n, item, ts;  // ts = tail | (seg == 0) << 8
};
__device__ __forceinline__ G make_g(const WalkArgs &A, const uint4 &d0, const uint4 &d1, uint32_t item, uint32_t lane) {
    G g;
    const uint64_t t_off = ((uint64_t)d0.y << 32) | d0.x, q_off = ((uint64_t)d0.w << 32) | d0.z;
    g.n = d1.x; g.item = item;
    g.ts = d1.w | ((d1.z == 0 ? 1u : 0u) << 8);
    const uint8_t *tp = A.text + t_off, *qp = A.text + q_off;
    g.a = (uint32_t)(reinterpret_cast<uintptr_t>(tp) & 15u);
    const uint8_t *qv = qp - g.a;  // query byte stream in the column frame of the aligned target stream
    g.s = (uint32_t)(reinterpret_cast<uintptr_t>(qv) & 15u);
    g.tptr = tp - g.a + 16u * lane;  // chunk j covers item columns [16j-a, 16j-a+16)
    g.qptr = qv - g.s + 16u * lane;
    return g;
}
__device__ __forceinline__ void load_first(const G &g, uint32_t lane, uint4 &tv, uint4 &q0, uint4 &q1) {
    const int jrem = (int)((g.n + g.a + 15u) >> 4) - (int)lane, qrem = (int)((g.s + g.n + g.a - 1u) >> 4) - (int)lane;
    tv = make_uint4(0, 0, 0, 0); q0 = tv; q1 = tv;
    if (jrem > 0) {
        tv = ldg16(g.tptr);
        q0 = ldg16(g.qptr);
        if (g.s != 0 && qrem > 0) q1 = ldg16(g.qptr + 16);
    }
}

// descriptor prefetch through shared memory (cp.async keeps the 32 hot bytes out of the register file)
__device__ __forceinline__ void desc_prefetch(uint32_t *dst_smem, const ItemDesc *src, uint32_t lane) {
    if (lane < 2) {
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem + 4 * lane);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(reinterpret_cast<const uint4 *>(src) + lane) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void desc_wait() {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
}

struct Sched {  // dynamic work distribution: batches of kGrab consecutive items from a global counter
    uint32_t it, end, nb_raw;
};
constexpr uint32_t kGrab = 8;
__device__ __forceinline__ bool sched_next(const WalkArgs &A, Sched &S, uint32_t n_items, uint32_t lane) {
    ++S.it;
    if (S.it < S.end) return true;
    const uint32_t nb = __shfl_sync(0xFFFFFFFFu, S.nb_raw, 0);
    if (nb >= n_items) { S.it = n_items; return false; }
    S.it = nb;
    S.end = nb + kGrab < n_items ? nb + kGrab : n_items;
    if (lane == 0) S.nb_raw = atomicAdd(A.sched, kGrab);  // requested one batch ahead
    return true;
}

// stream one item.  WS = word part of the query misalignment (compile-time word select).
// On its last step the warp turns the prefetched descriptor of its next item into a geometry, requests that
// item's first vectors and starts the descriptor prefetch of the item after it.
template <int MODE, int WS>
__device__ __forceinline__ void walk_item(const WalkArgs &A, uint32_t (*Q)[64], uint32_t &qn, WarpEmit &E, uint32_t lane, G &g,
                                          uint4 &tv, uint4 &q0, uint4 &q1, bool &have_next, uint32_t *dslot, uint32_t &dpar,
                                          Sched &S, uint32_t n_items) {
    constexpr bool CALLER = MODE == MODE_RUNS_CALLER;
    const uint32_t a = g.a, n = g.n;
    const uint32_t nch = (n + a + 15u) >> 4, q_last = (g.s + n + a - 1u) >> 4;
    const uint32_t bs8 = (g.s & 3u) * 8u;
    const bool two = g.s != 0;
    // interior chunks: fully inside the item -> 16j >= a and 16j+16 <= n+a
    const uint32_t jf = (a + 15u) >> 4, jl = (n + a) >> 4;
    const uint8_t *tptr = g.tptr, *qptr = g.qptr;
    int jrem = (int)nch - (int)lane;     // chunk j = j0+lane exists iff jrem > 0
    int qrem = (int)q_last - (int)lane;  // its second query vector is needed iff qrem > 0
    uint32_t carry = 0;                  // (t,q) bytes of the column before this step's first chunk: t | q<<8
    uint32_t j = lane;
    G nxt = g;
    for (;;) {
        const bool act = jrem > 0;
        const bool more = j + 32 - lane < nch;  // warp-uniform: this item has another step
        tptr += 512; qptr += 512; jrem -= 32; qrem -= 32;
        uint4 ntv = make_uint4(0, 0, 0, 0), nq0 = ntv, nq1 = ntv;
        if (more) {
            if (jrem > 0) {
                ntv = ldg16(tptr);
                nq0 = ldg16(qptr);
                if (two && qrem > 0) nq1 = ldg16(qptr + 16);
            }
        } else if (have_next) {
            desc_wait();
            const uint4 e0 = *reinterpret_cast<const uint4 *>(dslot + 8 * dpar);
            const uint4 e1 = *reinterpret_cast<const uint4 *>(dslot + 8 * dpar + 4);
            nxt = make_g(A, e0, e1, S.it, lane);
            load_first(nxt, lane, ntv, nq0, nq1);
            have_next = sched_next(A, S, n_items, lane);
            dpar ^= 1u;
            __syncwarp();
            if (have_next) desc_prefetch(dslot + 8 * dpar, A.items + S.it, lane);
            have_next = true | have_next;  // (the item just turned into `nxt` exists; S.it now names the one after)
        }
        uint32_t w0, w1, w2, w3, w4;
        if (WS == 0) { w0 = q0.x; w1 = q0.y; w2 = q0.z; w3 = q0.w; w4 = q1.x; }
        else if (WS == 1) { w0 = q0.y; w1 = q0.z; w2 = q0.w; w3 = q1.x; w4 = q1.y; }
        else if (WS == 2) { w0 = q0.z; w1 = q0.w; w2 = q1.x; w3 = q1.y; w4 = q1.z; }
        else { w0 = q0.w; w1 = q1.x; w2 = q1.y; w3 = q1.z; w4 = q1.w; }
        const uint32_t qw0 = __funnelshift_r(w0, w1, bs8), qw1 = __funnelshift_r(w1, w2, bs8),
                       qw2 = __funnelshift_r(w2, w3, bs8), qw3 = __funnelshift_r(w3, w4, bs8);
        // last column of this chunk -> neighbour lane (previous column of the next chunk)
        const uint32_t lastb = act ? __byte_perm(tv.w, qw3, 0x4473) & 0xFFFFu : 0u;  // t[15] | q[15]<<8
        uint32_t pb = __shfl_up_sync(0xFFFFFFFFu, lastb, 1);
        if (lane == 0) pb = carry;
        carry = __shfl_sync(0xFFFFFFFFu, lastb, 31);
        uint32_t x = (tv.x ^ qw0) | (tv.y ^ qw1) | (tv.z ^ qw2) | (tv.w ^ qw3);
        bool prev_ne = (pb & 0xFFu) != (pb >> 8);
        if (CALLER) {
            x |= has_zero_byte(tv.x ^ kGap4) | has_zero_byte(tv.y ^ kGap4) | has_zero_byte(tv.z ^ kGap4) |
                 has_zero_byte(tv.w ^ kGap4);
            prev_ne |= (pb & 0xFFu) == '-';
        }
        const bool edge = j == 0 || j < jf || j >= jl;
        const bool slow = act && (edge || x != 0u || prev_ne);
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, slow);
        if (mask) {
            const uint32_t seg0 = g.ts >> 8;
            uint32_t f0 = 0, t0 = 16, t1 = 16, prev = (1u << 16) | pb;
            if (slow) {
                if (edge) {
                    const int64_t lo = (int64_t)a - 16ll * j, hi = (int64_t)n + (int64_t)a - 16ll * j;
                    f0 = clamp16(lo);
                    t0 = clamp16(hi);
                    t1 = clamp16(hi + (int64_t)(g.ts & 0xFFu));  // columns of the next item are context only
                }
                if (j == 0) {  // column preceding chunk 0 (only matters when a == 0; otherwise it sits inside chunk 0)
                    prev = 0;
                    if (a == 0 && !seg0) prev = (1u << 16) | ((uint32_t)g.qptr[(int)g.s - 1] << 8) | (uint32_t)g.tptr[-1];
                }
            }
            const uint32_t meta = f0 | (t0 << 5) | (t1 << 10) | (seg0 << 15) | ((16u * j + 16u - a) << 16);
            if (slow) {
                const uint32_t slot = qn + __popc(mask & ((1u << lane) - 1u));
                Q[0][slot] = tv.x; Q[1][slot] = tv.y; Q[2][slot] = tv.z; Q[3][slot] = tv.w;
                Q[4][slot] = qw0; Q[5][slot] = qw1; Q[6][slot] = qw2; Q[7][slot] = qw3;
                Q[8][slot] = prev; Q[9][slot] = meta; Q[10][slot] = g.item;
            }
            qn += __popc(mask);
            __syncwarp();
            if (qn >= 32) {
                const uint3 st = walk_flush<MODE>(A, Q, qn, E.open_item, E.open_count);
                qn = st.x; E.open_item = st.y; E.open_count = st.z;
            }
        }
        j += 32;
        tv = ntv; q0 = nq0; q1 = nq1;
        if (!more) break;
    }
    g = nxt;
}

template <int MODE>
__global__ void __launch_bounds__(kWalkThreads, kWalkCtasPerSm) k_walk(WalkArgs A) {
    constexpr bool RUNS = MODE != MODE_STAT;
    __shared__ uint32_t s_queue[kWalkWarps][11][64];          // per warp, SoA, 64 slots
    __shared__ __align__(16) uint32_t s_desc[kWalkWarps][16];  // per warp: two 32-byte descriptor slots
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t (*Q)[64] = s_queue[wid];
    uint32_t *dslot = s_desc[wid];
    uint32_t qn = 0, dpar = 0;
    WarpEmit E;
    E.open_item = 0xFFFFFFFFu;
    E.open_count = 0;
    const uint32_t n_items = (uint32_t)A.n_items;
    Sched S;
    S.it = 0; S.end = 0; S.nb_raw = 0;
    if (lane == 0) {
        S.it = atomicAdd(A.sched, kGrab);
        S.nb_raw = atomicAdd(A.sched, kGrab);
    }
    S.it = __shfl_sync(0xFFFFFFFFu, S.it, 0);
    S.end = S.it + kGrab < n_items ? S.it + kGrab : n_items;
    if (S.it < n_items) {
        const uint4 *dp = reinterpret_cast<const uint4 *>(A.items + S.it);
        G cur = make_g(A, __ldg(dp), __ldg(dp + 1), S.it, lane);
        uint4 tv, q0, q1;
        load_first(cur, lane, tv, q0, q1);
        bool have_next = sched_next(A, S, n_items, lane);  // S.it = the warp's next item
        if (have_next) desc_prefetch(dslot + 8 * dpar, A.items + S.it, lane);
        for (;;) {
            const bool had_next = have_next;
            bool hn = have_next;
            switch (cur.s >> 2) {  // warp-uniform
                case 0: walk_item<MODE, 0>(A, Q, qn, E, lane, cur, tv, q0, q1, hn, dslot, dpar, S, n_items); break;
                case 1: walk_item<MODE, 1>(A, Q, qn, E, lane, cur, tv, q0, q1, hn, dslot, dpar, S, n_items); break;
                case 2: walk_item<MODE, 2>(A, Q, qn, E, lane, cur, tv, q0, q1, hn, dslot, dpar, S, n_items); break;
                default: walk_item<MODE, 3>(A, Q, qn, E, lane, cur, tv, q0, q1, hn, dslot, dpar, S, n_items); break;
            }
            if (!had_next) break;   // the item just walked was the warp's last
            have_next = S.it < n_items;  // is there an item after the one now in `cur`?
        }
    }
    if (qn) {
        walk_batch<MODE>(A, Q, qn, lane, E);
        __syncwarp();
    }
    if (RUNS && E.open_item != 0xFFFFFFFFu && lane == 0) A.item_nheads[E.open_item] = E.open_count;
}

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

}  // namespace wga
