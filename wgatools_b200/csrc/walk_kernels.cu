// walk_kernels.cu -- translation unit of the dense column-walk kernels (see walk_types.cuh for why it is separate).
#include "walk.cuh"

namespace wga {

void launch_walk(int mode, unsigned grid, cudaStream_t stream, const WalkArgs &A) {
    switch (mode) {
        case MODE_STAT: k_walk<MODE_STAT><<<grid, kWalkThreads, 0, stream>>>(A); break;
        case MODE_RUNS_EXT: k_walk<MODE_RUNS_EXT><<<grid, kWalkThreads, 0, stream>>>(A); break;
        case MODE_RUNS_ALL: k_walk<MODE_RUNS_ALL><<<grid, kWalkThreads, 0, stream>>>(A); break;
        default: k_walk<MODE_RUNS_CALLER><<<grid, kWalkThreads, 0, stream>>>(A); break;
    }
}
void launch_rec_slot_need(unsigned grid, unsigned block, cudaStream_t stream, uint64_t n_rec, const uint32_t *n_items,
                          const uint64_t *ncols, uint32_t shift, uint64_t *need) {
    k_rec_slot_need<<<grid, block, 0, stream>>>(n_rec, n_items, ncols, shift, need);
}
void launch_fill_item_desc(unsigned grid, unsigned block, cudaStream_t stream, uint64_t n_rec, const uint32_t *n_items,
                           const uint64_t *item_base, const uint64_t *t_off, const uint64_t *q_off, const uint64_t *ncols,
                           const uint64_t *rec_slot_base, uint32_t shift, ItemDesc *items) {
    k_fill_item_desc<<<grid, block, 0, stream>>>(n_rec, n_items, item_base, t_off, q_off, ncols, rec_slot_base, shift, items);
}
void launch_runs_compact(unsigned grid, unsigned block, cudaStream_t stream, uint64_t n_items, const ItemDesc *items,
                         const uint32_t *item_nheads, const uint64_t *run_base, const uint32_t *slots, uint64_t *run_pos,
                         uint32_t *run_rec, uint8_t *run_cls) {
    k_runs_compact<<<grid, block, 0, stream>>>(n_items, items, item_nheads, run_base, slots, run_pos, run_rec, run_cls);
}

}  // namespace wga
