// walk_tma_kernels.cu -- translation unit of the TMA-staged column walk (walk_tma.cuh); separate from walk_kernels.cu
// for the same reason that one is separate from the rest (register allocation near the occupancy budget).
#define WGA_WALK_NO_SPARSE
#include "walk_tma.cuh"

namespace wga {

template <int MODE>
static cudaError_t launch_one(unsigned grid, cudaStream_t stream, const WalkArgs &A) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_walk_tma<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TmaSmem));
        if (e != cudaSuccess) return e;
        configured = true;
    }
    k_walk_tma<MODE><<<grid, kTmaThreads, sizeof(TmaSmem), stream>>>(A);
    return cudaSuccess;
}

// persistent grid of the TMA variant: kTmaCtasPerSm CTAs per SM, never more CTAs than there are warps' worth of items
unsigned walk_tma_grid(int n_sm, uint64_t n_items) {
    const uint64_t want = (n_items + kTmaWarps - 1) / kTmaWarps, cap = (uint64_t)n_sm * kTmaCtasPerSm;
    return (unsigned)(want < cap ? want : cap);
}

cudaError_t launch_walk_tma(int mode, unsigned grid, cudaStream_t stream, const WalkArgs &A) {
    switch (mode) {
        case MODE_STAT: return launch_one<MODE_STAT>(grid, stream, A);
        case MODE_RUNS_EXT: return launch_one<MODE_RUNS_EXT>(grid, stream, A);
        case MODE_RUNS_ALL: return launch_one<MODE_RUNS_ALL>(grid, stream, A);
        default: return launch_one<MODE_RUNS_CALLER>(grid, stream, A);
    }
}

}  // namespace wga
