// walk_types.cuh -- work-item descriptor, kernel arguments and launch wrappers of the dense column walk.
// The kernels themselves (walk.cuh) are compiled in their own translation unit (walk_kernels.cu): ptxas'
// register allocation for k_walk sits close to the 80-register budget of 3 CTAs/SM and was observed to tip into
// spilling (-35 % throughput) when unrelated kernels in the same translation unit changed.  The Makefile checks
// that k_walk keeps a zero stack frame.
#pragma once
#include "ctx.cuh"

namespace wga {

// MODE_RUNS_ALL: caller classification (W runs kept apart) + the counters: one walk that serves maf2paf, stat AND call
enum { MODE_STAT = 0, MODE_RUNS_EXT = 1, MODE_RUNS_CALLER = 2, MODE_RUNS_ALL = 3 };
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
    uint64_t slot_base;          // first slot of the item's run-head region
    uint32_t n, rec;
    uint32_t seg, tail, pad0, pad1;  // tail: columns of the record after this item, clipped to 16
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
    uint32_t *n_w;                // MODE_RUNS_ALL: number of W (both-gap) run heads emitted
    unsigned int *sched;          // dynamic work counter (zero-initialised)
    uint32_t grab;                // items handed out per request (k_walk_tma; k_walk uses 8)
};

__device__ __host__ __forceinline__ uint64_t slot_cap(uint64_t n, uint32_t shift) { return (n >> shift) + kSlotPad; }

// host-side launch wrappers (defined in walk_kernels.cu)
void launch_walk(int mode, unsigned grid, cudaStream_t stream, const WalkArgs &A);
// TMA-staged variant (walk_tma.cuh, walk_tma_kernels.cu)
cudaError_t launch_walk_tma(int mode, unsigned grid, cudaStream_t stream, const WalkArgs &A);
unsigned walk_tma_grid(int n_sm, uint64_t n_items);
void launch_rec_slot_need(unsigned grid, unsigned block, cudaStream_t stream, uint64_t n_rec, const uint32_t *n_items,
                          const uint64_t *ncols, uint32_t shift, uint64_t *need);
void launch_fill_item_desc(unsigned grid, unsigned block, cudaStream_t stream, uint64_t n_rec, const uint32_t *n_items,
                           const uint64_t *item_base, const uint64_t *t_off, const uint64_t *q_off, const uint64_t *ncols,
                           const uint64_t *rec_slot_base, uint32_t shift, ItemDesc *items);
void launch_runs_compact(unsigned grid, unsigned block, cudaStream_t stream, uint64_t n_items, const ItemDesc *items,
                         const uint32_t *item_nheads, const uint64_t *run_base, const uint32_t *slots, uint64_t *run_pos,
                         uint32_t *run_rec, uint8_t *run_cls);

}  // namespace wga
