// walk_tma.cuh -- TMA-staged variant of the dense column walk (same algorithm and outputs as k_walk, walk.cuh).
//
// k_walk keeps its operands in registers: one step of vectors is in flight per warp plus one prefetched step, and
// the kernel idles on `long_scoreboard` (profiles/r01g_walk_full_ncu.md: 50-54 % of DRAM peak).  Here every warp owns
// a ring of kTmaStages shared-memory windows of kTmaWin chunks (target bytes + query bytes of kTmaWin/32 steps); one
// lane fills them with 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) up to kTmaStages
// windows ahead -- across item boundaries -- and the warp reads its vectors with LDS.128.  The 24 prefetch registers
// of k_walk are gone.  Measured on the bench workload (1 M blocks): the bulk-copy REQUEST rate is what limits this
// path -- 1 KB requests 2.3 TB/s, 1.5 KB 3.7 TB/s, 2 KB 4.2-4.5 TB/s -- and, at equal request size, the number of
// resident warps (12 warps 3.5, 16 warps 4.2, 20 warps 4.5 TB/s).  Shared memory (2 x 4 KB of windows + 3 KB of
// slow-chunk queue per warp) caps the SM at 20 warps: 2 CTAs x 10 warps, 2 stages x 128 chunks.
//
// Round 2 (v6, 5.3 TB/s = 0.81 of the measured HBM peak): the consumer no longer goes through walk_step.  tma_steps<MODE, WS>
// is the step loop instantiated for the four word shifts of the query realignment (dispatched once per window), the
// slow-chunk queue is a ring of 16-byte vectors (3 STS.128 per push), and the kernel has one classification site
// (walk_batch_q).  tma_steps2 (32 columns per lane) is a measured-slower experiment kept behind WGA_TMA_LANE32.
#pragma once
#include "walk.cuh"

namespace wga {

#ifndef WGA_TMA_STAGES
#define WGA_TMA_STAGES 2
#endif
#ifndef WGA_TMA_WIN
#define WGA_TMA_WIN 128
#endif
#ifndef WGA_TMA_CTAS
#define WGA_TMA_CTAS 2
#endif
#ifndef WGA_TMA_THREADS
#define WGA_TMA_THREADS 320
#endif
#ifndef WGA_TMA_PREFETCH
#define WGA_TMA_PREFETCH 0
#endif
#ifndef WGA_TMA_LANE32
#define WGA_TMA_LANE32 0  // 1 = 32 columns per lane and step (tma_steps2): measured SLOWER (1.35 against 1.30 ms, see tma_steps2)
#endif
constexpr int kTmaThreads = WGA_TMA_THREADS, kTmaWarps = kTmaThreads / 32, kTmaCtasPerSm = WGA_TMA_CTAS;
constexpr int kTmaStages = WGA_TMA_STAGES;
constexpr uint32_t kTmaWin = WGA_TMA_WIN;                          // chunks per window (32 per step)
constexpr uint32_t kTmaStageBytes = kTmaWin * 16 + kTmaWin * 16 + 16;  // T | Q (+ one vector for the realign)
constexpr int kTmaGeoms = 3;                                      // item geometries in flight per warp: the (at most) two
                                                                  // items of the windows in flight + the one opened ahead
// Experiment knob (off): request the windows that come kTmaPrefetch ring turns later into L2 already now
// (cp.async.bulk.prefetch.L2, SASS UBLKPF.L2).  Measured on the bench workload (profiles/r02_walk_experiments.md): 2 and 4
// windows ahead make the kernel SLOWER (1.65 / 1.64 ms against 1.52 ms), i.e. the walk is not waiting for DRAM latency
// that more bytes in flight could hide.
constexpr uint32_t kTmaPrefetch = WGA_TMA_PREFETCH;               // windows ahead (0 = off)

struct __align__(16) TmaWin { uint32_t gslot, j0, cnt, pad; };
struct __align__(16) TmaWarp {
    uint8_t stage[kTmaStages][kTmaStageBytes];
    Geom geom[kTmaGeoms];
    TmaWin win[kTmaStages];
    unsigned long long bar[kTmaStages];
};
// slow-chunk queue of a warp: a ring of 64 entries, three 16-byte vectors per entry (3 STS.128 per push, 3 LDS.128 per
// batch lane; the row-per-field layout of k_walk costs 11 of each)
struct __align__(16) TmaQueue {
    uint4 t[64];  // target words of the chunk
    uint4 q[64];  // query words, realigned to the target's columns
    uint4 m[64];  // prev column | meta | item | rec (walk_batch_q)
};
struct QueueVec {
    const TmaQueue *q;
    uint32_t head;
    static constexpr bool kHasRec = true;
    __device__ __forceinline__ void load(uint32_t lane, uint32_t (&tw)[4], uint32_t (&qw)[4], uint32_t &pv, uint32_t &meta,
                                         uint32_t &item, uint32_t &rec) const {
        const uint32_t i = (head + lane) & 63u;
        const uint4 a = q->t[i], b = q->q[i], c = q->m[i];
        tw[0] = a.x; tw[1] = a.y; tw[2] = a.z; tw[3] = a.w;
        qw[0] = b.x; qw[1] = b.y; qw[2] = b.z; qw[3] = b.w;
        pv = c.x; meta = c.y; item = c.z; rec = c.w;
    }
};
struct TmaSmem {
    TmaQueue queue[kTmaWarps];
    TmaWarp w[kTmaWarps];
};
// two CTAs per SM: 2 x (sizeof(TmaSmem) + 1 KB reserved per CTA) must fit the SM's 228 KB
static_assert(kTmaCtasPerSm * (sizeof(TmaSmem) + 1024) <= 233472, "k_walk_tma: shared memory budget of the SM");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint4 lds16(const uint8_t *p) { return *reinterpret_cast<const uint4 *>(p); }

// Steps st, st+32, ... of one window (the algorithm of walk_step, walk.cuh) until the window ends or 32 slow chunks are
// queued.  WS = word part of the query's byte shift (warp-uniform per item): the caller dispatches on it once per call,
// not once per step.  Lane i of a step owns chunk st+i of the window; lanes past the window's end read stale shared
// memory and are inactive.  jlm1 = (first chunk that is not fully inside the item) - 1: edge <=> j == 0 || j >= jl
// <=> (j - 1) >= jlm1 in unsigned arithmetic.
template <int MODE, int WS>
__device__ __forceinline__ uint32_t tma_steps(TmaQueue &QV, uint32_t qhead, uint32_t &qn, uint32_t lane, const Geom &g,
                                              uint32_t jlm1, uint32_t wj0, uint32_t wcnt, uint32_t st, const uint8_t *Ts,
                                              PrevCol &P) {
    constexpr bool CALLER = MODE == MODE_RUNS_CALLER || MODE == MODE_RUNS_ALL;
    const uint32_t sh = (g.s & 3u) * 8u;
    const uint8_t *pT = Ts + 16u * (st + lane);  // this lane's target vector; its query vectors sit kTmaWin * 16 behind
    do {
        const uint4 tv = lds16(pT), q0 = lds16(pT + kTmaWin * 16), q1 = lds16(pT + kTmaWin * 16 + 16);
        uint32_t qw0, qw1, qw2, qw3;
        realign<WS>(q0, q1, sh, qw0, qw1, qw2, qw3);
        const uint32_t rel = st + lane;
        const bool act = rel < wcnt;
        const uint32_t x3 = tv.w ^ qw3;
        uint32_t x = (tv.x ^ qw0) | (tv.y ^ qw1) | (tv.z ^ qw2) | x3;
        bool ne_last = x3 > 0x00FFFFFFu;  // the chunk's last column differs
        if (CALLER) {                     // bit 6 clear in a target byte: superset of "holds a gap" (see walk_step)
            const uint32_t g6 = tv.x & tv.y & tv.z & tv.w;
            x |= ~g6 & 0x40404040u;
            ne_last |= (tv.w & 0x40000000u) == 0u;
        }
        ne_last &= act;
        const uint32_t nb = __ballot_sync(0xFFFFFFFFu, ne_last);
        const bool prev_ne = (((nb << 1) | P.bit) >> lane) & 1u;
        const uint32_t j = wj0 + rel;
        const bool edge = j - 1u >= jlm1;
        const bool slow = act && (edge || x != 0u || prev_ne);
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, slow);
        if (mask) {
            // bytes of the column before the chunk: one byte back in the window; for the window's first chunk the last
            // column of the previous window (lane 31 of its last step keeps it)
            uint32_t pb = 0;
            if (st == 0 && (mask & 1u)) pb = __shfl_sync(0xFFFFFFFFu, (P.tw3 >> 24) | ((P.qw3 >> 24) << 8), 31);  // warp-uniform
            if (slow) {
                if (rel != 0) pb = (uint32_t)pT[-1] | ((uint32_t)pT[kTmaWin * 16 + g.s - 1u] << 8);
                uint32_t f0 = 0, t0 = 16, t1 = 16;
                if (edge) {  // a <= 15, n <= kSeg: everything fits 32-bit signed arithmetic
                    const int32_t lo = (int32_t)g.a - 16 * (int32_t)j, hi = (int32_t)(g.n + g.a) - 16 * (int32_t)j;
                    f0 = clamp16(lo);
                    t0 = clamp16(hi);
                    t1 = clamp16(hi + (int32_t)g.tail);  // columns of the next item are context only
                }
                const uint32_t slot = (qhead + qn + __popc(mask & ((1u << lane) - 1u))) & 63u;
                QV.t[slot] = tv;
                QV.q[slot] = make_uint4(qw0, qw1, qw2, qw3);
                QV.m[slot] = make_uint4(j == 0 ? g.first_prev : ((1u << 16) | pb),
                                        f0 | (t0 << 5) | (t1 << 10) | ((g.seg == 0 ? 1u : 0u) << 15) | ((16u * j + 16u - g.a) << 16),
                                        g.item, g.rec);
            }
            qn += __popc(mask);
        }
        P.bit = nb >> 31;
        P.tw3 = tv.w;
        P.qw3 = qw3;
        st += 32;
        pT += 512;
    } while (st < wcnt && qn < 32);
    return st;
}

// 32 columns per lane: lane i of a double step owns chunks st+2i (A) and st+2i+1 (B) of the window, i.e. 32 contiguous
// target bytes and the 48 query bytes around them.  Everything that tma_steps pays once per 32 chunks (loop, ballots,
// the push bookkeeping) is paid once per 64 here.  The queue takes at most 32 new entries at a time (ring of 64, up to
// 31 waiting): a double step with more slow chunks than fit is pushed in two halves -- lanes 0..15, then, after the
// caller had a chance to classify a batch, lanes 16..31 (`half`, warp-uniform; the step is recomputed, the carry is not
// advanced in between).
// Experiment knob (WGA_TMA_LANE32=1, off): parity-green, but on the bench workload the walk takes 1.35 ms against 1.30 ms
// with tma_steps -- the two-entry push costs what the shared bookkeeping saves (188 SASS instructions per 64 chunks either
// way), the 32-byte lane stride makes every LDS.128 a 2-way bank conflict and the last step of an item idles more lanes.
template <int MODE, int WS>
__device__ __forceinline__ uint32_t tma_steps2(TmaQueue &QV, uint32_t qhead, uint32_t &qn, uint32_t lane, const Geom &g,
                                               uint32_t jlm1, uint32_t wj0, uint32_t wcnt, uint32_t st, uint32_t &half,
                                               const uint8_t *Ts, PrevCol &P) {
    constexpr bool CALLER = MODE == MODE_RUNS_CALLER || MODE == MODE_RUNS_ALL;
    const uint32_t sh = (g.s & 3u) * 8u;
    const uint8_t *pT = Ts + 16u * st + 32u * lane;  // this lane's two target vectors; the query vectors kTmaWin * 16 behind
    const uint32_t lt = (1u << lane) - 1u;
    do {
        const uint4 ta = lds16(pT), tb = lds16(pT + 16);
        const uint4 q0 = lds16(pT + kTmaWin * 16), q1 = lds16(pT + kTmaWin * 16 + 16), q2 = lds16(pT + kTmaWin * 16 + 32);
        uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
        realign<WS>(q0, q1, sh, a0, a1, a2, a3);
        realign<WS>(q1, q2, sh, b0, b1, b2, b3);
        const uint32_t rel = st + 2u * lane;  // chunk A; B = rel + 1
        const bool act_a = rel < wcnt, act_b = rel + 1u < wcnt;
        const uint32_t xa3 = ta.w ^ a3, xb3 = tb.w ^ b3;
        uint32_t xa = (ta.x ^ a0) | (ta.y ^ a1) | (ta.z ^ a2) | xa3;
        uint32_t xb = (tb.x ^ b0) | (tb.y ^ b1) | (tb.z ^ b2) | xb3;
        bool ne_a = xa3 > 0x00FFFFFFu, ne_b = xb3 > 0x00FFFFFFu;  // the chunk's last column differs
        if (CALLER) {  // bit 6 clear in a target byte: superset of "holds a gap" (see walk_step)
            xa |= ~(ta.x & ta.y & ta.z & ta.w) & 0x40404040u;
            xb |= ~(tb.x & tb.y & tb.z & tb.w) & 0x40404040u;
            ne_a |= (ta.w & 0x40000000u) == 0u;
            ne_b |= (tb.w & 0x40000000u) == 0u;
        }
        ne_a &= act_a;
        ne_b &= act_b;
        const uint32_t nb = __ballot_sync(0xFFFFFFFFu, ne_b);
        const bool prev_a = (((nb << 1) | P.bit) >> lane) & 1u;  // B's previous column is A's last: ne_a
        const uint32_t j = wj0 + rel;                           // chunk A's index in the item; B is j + 1 (never 0)
        const bool edge_a = j - 1u >= jlm1, edge_b = j >= jlm1;
        const bool slow_a = act_a && (edge_a || xa != 0u || prev_a);
        const bool slow_b = act_b && (edge_b || xb != 0u || ne_a);
        uint32_t ma = __ballot_sync(0xFFFFFFFFu, slow_a), mb = __ballot_sync(0xFFFFFFFFu, slow_b);
        bool again = false;  // warp-uniform: the upper half of this step is still to be pushed
        if (ma | mb) {
            uint32_t allow = half ? 0xFFFF0000u : 0xFFFFFFFFu;
            if (!half && qn + __popc(ma) + __popc(mb) > 64u) { allow = 0x0000FFFFu; again = true; }
            ma &= allow;
            mb &= allow;
            const bool push_a = slow_a && ((allow >> lane) & 1u), push_b = slow_b && ((allow >> lane) & 1u);
            // the column before chunk A (see tma_steps); the one before chunk B is A's last column, still in registers
            uint32_t pa = 0x4141u;
            if (__any_sync(0xFFFFFFFFu, push_a && prev_a)) {
                if (st == 0 && (ma & 1u)) pa = __shfl_sync(0xFFFFFFFFu, (P.tw3 >> 24) | ((P.qw3 >> 24) << 8), 31);  // warp-uniform
                if (push_a && rel != 0) pa = (uint32_t)pT[-1] | ((uint32_t)pT[kTmaWin * 16 + g.s - 1u] << 8);
            }
            const uint32_t pb = (ta.w >> 24) | ((a3 >> 24) << 8);
            uint32_t meta_a = (16u << 5) | (16u << 10), meta_b = meta_a;
            if (__any_sync(0xFFFFFFFFu, (push_a && edge_a) || (push_b && edge_b))) {
                // a <= 15, n <= kSeg: everything fits 32-bit signed arithmetic; tail: context columns only
                const int32_t lo = (int32_t)g.a - 16 * (int32_t)j, hi = (int32_t)(g.n + g.a) - 16 * (int32_t)j;
                if (edge_a) meta_a = clamp16(lo) | (clamp16(hi) << 5) | (clamp16(hi + (int32_t)g.tail) << 10);
                if (edge_b) meta_b = clamp16(lo - 16) | (clamp16(hi - 16) << 5) | (clamp16(hi - 16 + (int32_t)g.tail) << 10);
            }
            const uint32_t seg0 = (g.seg == 0 ? 1u : 0u) << 15, col = (16u * j + 16u - g.a) << 16;
            uint32_t slot = qhead + qn + __popc(ma & lt) + __popc(mb & lt);
            if (push_a) {
                const uint32_t sl = slot & 63u;
                QV.t[sl] = ta;
                QV.q[sl] = make_uint4(a0, a1, a2, a3);
                QV.m[sl] = make_uint4(j == 0 ? g.first_prev : ((1u << 16) | pa), meta_a | seg0 | col, g.item, g.rec);
                ++slot;
            }
            if (push_b) {
                const uint32_t sl = slot & 63u;
                QV.t[sl] = tb;
                QV.q[sl] = make_uint4(b0, b1, b2, b3);
                QV.m[sl] = make_uint4((1u << 16) | pb, meta_b | seg0 | (col + (16u << 16)), g.item, g.rec);
            }
            qn += __popc(ma) + __popc(mb);
        }
        half = again ? 1u : 0u;
        if (!again) {
            P.bit = nb >> 31;
            P.tw3 = tb.w;
            P.qw3 = b3;
            st += 64;
            pT += 1024;
        }
    } while (st < wcnt && qn < 32);
    return st;
}

template <int MODE>
__global__ void __launch_bounds__(kTmaThreads, kTmaCtasPerSm) k_walk_tma(WalkArgs A) {
    constexpr bool RUNS = MODE != MODE_STAT;
    extern __shared__ __align__(128) uint8_t s_raw[];
    TmaSmem &S = *reinterpret_cast<TmaSmem *>(s_raw);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    TmaQueue &QV = S.queue[wid];
    TmaWarp &W = S.w[wid];
    uint32_t qn = 0, qhead = 0;  // queued entries, ring position of the oldest
    WarpEmit E;
    E.open_item = 0xFFFFFFFFu;
    E.open_count = 0;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kTmaStages; ++s) mbar_init(&W.bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // dynamic scheduling as in k_walk: batches of kGrab consecutive items, the next batch requested one ahead
    const uint32_t kGrab = A.grab;  // 8 for short items; fewer when items are long and few (tail balance)
    const uint32_t n_items = (uint32_t)A.n_items;
    uint32_t s_it = 0, s_end = 0, nb_raw = 0;
    if (lane == 0) {
        s_it = atomicAdd(A.sched, kGrab);
        nb_raw = atomicAdd(A.sched, kGrab);
    }
    s_it = __shfl_sync(0xFFFFFFFFu, s_it, 0);
    s_end = s_it + kGrab < n_items ? s_it + kGrab : n_items;
    auto advance = [&]() -> bool {
        ++s_it;
        if (s_it < s_end) return true;
        const uint32_t nb = __shfl_sync(0xFFFFFFFFu, nb_raw, 0);
        if (nb >= n_items) { s_it = n_items; return false; }
        s_it = nb;
        s_end = nb + kGrab < n_items ? nb + kGrab : n_items;
        if (lane == 0) nb_raw = atomicAdd(A.sched, kGrab);
        return true;
    };
    if (s_it >= n_items) return;

    // ---- producer cursor: the item being issued lives in geometry slot p_gs (shared memory); only what issuing needs
    //      stays in registers.  The descriptor of the item after it is already requested, one 16-byte third per lane.
    uint32_t p_gs = 0, p_j = 0, p_nch, p_s, p_qlast;
    const uint8_t *p_tb, *p_qb;
    bool p_has = true, e_has;
    uint32_t e_item;
    uint4 ed = make_uint4(0, 0, 0, 0);
    auto fetch_desc = [&](uint32_t item) {
        if (lane < 3) ed = __ldg(reinterpret_cast<const uint4 *>(A.items + item) + lane);
    };
    auto open_item = [&](uint32_t item, uint32_t gslot) {  // descriptor thirds (lanes 0..2) -> geometry in slot gslot
        uint4 d0, d1, d2;
        d0.x = __shfl_sync(0xFFFFFFFFu, ed.x, 0); d0.y = __shfl_sync(0xFFFFFFFFu, ed.y, 0);
        d0.z = __shfl_sync(0xFFFFFFFFu, ed.z, 0); d0.w = __shfl_sync(0xFFFFFFFFu, ed.w, 0);
        d1.x = __shfl_sync(0xFFFFFFFFu, ed.x, 1); d1.y = __shfl_sync(0xFFFFFFFFu, ed.y, 1);
        d1.z = __shfl_sync(0xFFFFFFFFu, ed.z, 1); d1.w = __shfl_sync(0xFFFFFFFFu, ed.w, 1);
        d2.x = __shfl_sync(0xFFFFFFFFu, ed.x, 2); d2.y = __shfl_sync(0xFFFFFFFFu, ed.y, 2);
        d2.z = d2.w = 0;
        const Geom g = make_geom(A, d0, d1, d2, item);
        if (lane == 0) W.geom[gslot] = g;
        p_nch = g.nch; p_s = g.s; p_qlast = g.q_last; p_tb = g.tbase; p_qb = g.qbase;
        p_j = 0;
    };
    // the first kTmaPrefetch windows of the item AFTER the one being issued: its descriptor thirds sit in lanes 0..2
    auto prefetch_next_item = [&]() {
        if (kTmaPrefetch == 0) return;
        const uint32_t n = __shfl_sync(0xFFFFFFFFu, ed.z, 1);
        if (lane == 0) {
            const uint64_t t_off = ((uint64_t)ed.y << 32) | ed.x, q_off = ((uint64_t)ed.w << 32) | ed.z;
            const uint32_t cap = kTmaPrefetch * kTmaWin * 16u + 16u;
            const uint32_t want = n + 31u < cap ? n + 31u : cap;
            const uint32_t bytes = want & ~15u;
            l2_prefetch(reinterpret_cast<const void *>(reinterpret_cast<uintptr_t>(A.text + t_off) & ~(uintptr_t)15), bytes);
            l2_prefetch(reinterpret_cast<const void *>(reinterpret_cast<uintptr_t>(A.text + q_off) & ~(uintptr_t)15), bytes);
        }
    };
    fetch_desc(s_it);
    open_item(s_it, 0);
    e_has = advance();
    e_item = s_it;
    if (e_has) { fetch_desc(e_item); prefetch_next_item(); }
    uint32_t n_out = 0;  // windows issued and not yet consumed
    auto produce = [&](uint32_t stage) {
        if (!p_has) return;
        const uint32_t cnt = p_nch - p_j < kTmaWin ? p_nch - p_j : kTmaWin;
        // query vectors p_j .. min(p_j + cnt, q_last): the realign of chunk j needs vector j + 1 unless s == 0
        uint32_t qend = p_j + cnt;
        if (p_s == 0) qend -= 1;
        else if (qend > p_qlast) qend = p_qlast;
        const uint32_t qcnt = qend - p_j + 1;
        if (lane == 0) {
            W.win[stage].gslot = p_gs; W.win[stage].j0 = p_j; W.win[stage].cnt = cnt;
            mbar_expect_tx(&W.bar[stage], (cnt + qcnt) * 16u);
            bulk_g2s(W.stage[stage], p_tb + 16ull * p_j, cnt * 16u, &W.bar[stage]);
            bulk_g2s(W.stage[stage] + kTmaWin * 16, p_qb + 16ull * p_j, qcnt * 16u, &W.bar[stage]);
            if (kTmaPrefetch) {  // the window kTmaPrefetch turns ahead, while it still lies inside this item
                const uint32_t pj = p_j + kTmaPrefetch * kTmaWin;
                if (pj < p_nch) {
                    const uint32_t c2 = (p_nch - pj < kTmaWin ? p_nch - pj : kTmaWin) * 16u;
                    l2_prefetch(p_tb + 16ull * pj, c2);
                    l2_prefetch(p_qb + 16ull * pj, c2);
                }
            }
        }
        ++n_out;
        p_j += cnt;
        if (p_j >= p_nch) {  // next item
            p_has = e_has;
            if (e_has) {
                p_gs = (p_gs + 1) % kTmaGeoms;
                open_item(e_item, p_gs);
                e_has = advance();
                e_item = s_it;
                if (e_has) { fetch_desc(e_item); prefetch_next_item(); }
            }
        }
    };
#pragma unroll
    for (int s = 0; s < kTmaStages; ++s) produce(s);

    // ---- consumer
    Geom cur;
    cur.nch = 0;
    uint32_t cur_gs = 0xFFFFFFFFu, sc = 0, parity = 0, jlm1 = 0;  // parity: bit s = phase of stage s to wait for
    PrevCol carry = {0u, 0u, 0u};
    while (n_out) {
        mbar_wait(&W.bar[sc], (parity >> sc) & 1u);
        parity ^= 1u << sc;
        const uint4 wm = *reinterpret_cast<const uint4 *>(&W.win[sc]);
        const uint32_t gs = wm.x, wj0 = wm.y, wcnt = wm.z;
        if (gs != cur_gs) {
            cur = W.geom[gs];
            cur_gs = gs;
            jlm1 = ((cur.n + cur.a) >> 4) - 1u;
            carry.bit = 0; carry.tw3 = 0; carry.qw3 = 0;
        }
        const uint8_t *Ts = W.stage[sc];
        const uint32_t ws = cur.s >> 2;  // warp-uniform
        uint32_t st = 0;
#if WGA_TMA_LANE32
        uint32_t half = 0;
#define WGA_TMA_STEPS(WSV) st = tma_steps2<MODE, WSV>(QV, qhead, qn, lane, cur, jlm1, wj0, wcnt, st, half, Ts, carry)
#else
#define WGA_TMA_STEPS(WSV) st = tma_steps<MODE, WSV>(QV, qhead, qn, lane, cur, jlm1, wj0, wcnt, st, Ts, carry)
#endif
        while (st < wcnt) {
            switch (ws) {
                case 0: WGA_TMA_STEPS(0); break;
                case 1: WGA_TMA_STEPS(1); break;
                case 2: WGA_TMA_STEPS(2); break;
                default: WGA_TMA_STEPS(3); break;
            }
            while (qn >= 32) {  // the one classification site of the kernel
                __syncwarp();
                walk_batch_q<MODE>(A, QueueVec{&QV, qhead}, 32, lane, E);
                __syncwarp();
                qhead = (qhead + 32u) & 63u;
                qn -= 32;
            }
        }
#undef WGA_TMA_STEPS
        --n_out;
        // every lane has consumed its LDS results (they were compared / queued inside tma_steps), so the stage can be
        // handed back to the async proxy without a proxy fence: the fence is for generic-proxy WRITES, there are none
        __syncwarp();
        produce(sc);
        sc = sc + 1 == kTmaStages ? 0 : sc + 1;
    }
    if (qn) {
        __syncwarp();
        walk_batch_q<MODE>(A, QueueVec{&QV, qhead}, qn, lane, E);
        __syncwarp();
    }
    if (RUNS && E.open_item != 0xFFFFFFFFu && lane == 0) A.item_nheads[E.open_item] = E.open_count;
}

}  // namespace wga
