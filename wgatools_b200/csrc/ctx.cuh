// ctx.cuh -- context, device buffer pool, error plumbing, small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/wga_b200.h"
#include "hd_core.h"

namespace wga {

#define WGA_CK(call)                                                                       \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            ctx->set_error(WGA_E_CUDA, 0, std::string("CUDA error: ") + cudaGetErrorString(e__) + \
                                              " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
            return WGA_E_CUDA;                                                             \
        }                                                                                  \
    } while (0)

#define WGA_TRY(expr)            \
    do {                         \
        int rc__ = (expr);       \
        if (rc__) return rc__;   \
    } while (0)

constexpr uint32_t kSeg = 16384;  // columns per work item of the column walk
constexpr uint64_t kNoErr = 0xFFFFFFFFFFFFFFFFull;
constexpr uint64_t kPosMask = 0x7FFFFFFFFFFFFFFFull;  // ws entry: bit 63 = newline

// named device buffers (grown on demand, reused across calls)
enum BufId {
    B_TEXT, B_WS, B_TILE_WS, B_TILE_NL, B_TILE_WS_OFF, B_TILE_NL_OFF, B_CUBTMP, B_NLIDX,
    B_LINE_FLAGS, B_LINE_HEAD, B_LINE_REC,
    B_SL_NAME_OFF, B_SL_NAME_LEN, B_SL_START, B_SL_ALI, B_SL_SRC, B_SL_NEG, B_SL_SEQ_OFF, B_SL_SEQ_LEN,
    B_REC_FIRST, B_REC_LAST, B_REC_TL, B_REC_QL, B_REC_TOFF, B_REC_QOFF, B_REC_NCOLS, B_REC_NEG,
    B_REC_SKIP, B_REC_NITEMS, B_REC_ITEMBASE, B_REC_CNT, B_ITEM_REC, B_ITEM_SEG, B_ITEM_NHEADS,
    B_RUN_BASE, B_RUN_POS, B_RUN_REC, B_RUN_CLS, B_RUN_LEN, B_RUN_DL, B_RUN_CIGOFF, B_REC_RUNFIRST,
    B_ROW_LEN, B_ROW_OFF, B_ROW_FIXED, B_OUT, B_ERR, B_QNAME, B_SCALARS,
    B_PT_KEYS, B_PT_FIRST, B_PT_VALS, B_REC_SLOT, B_PT_FLAG, B_PT_IDX, B_PAIR_FIRST, B_PAIR_VALS,
    B_NAME_LEN, B_NAME_OFF, B_NAME_BUF, B_INV_FLAG, B_INV_IDX, B_INV_PAIR, B_INV_NUM,
    B_GEN_A, B_GEN_B, B_GEN_C, B_GEN_D, B_GEN_E, B_GEN_F, B_GEN_G, B_GEN_H,
    B_TEXT_RES, B_TOK_ERR, B_CNT11, B_XRUN_POS, B_XRUN_REC, B_XRUN_CLS, B_XRUN_FLAG, B_XRUN_IDX,
    B_PP_A, B_PP_B, B_PP_C, B_PP_D, B_PP_E, B_PP_F, B_PP_G, B_PP_H, B_PP_I, B_PP_J, B_PP_K, B_PP_L, B_PP_M, B_PP_N, B_PP_O, B_PP_P,
    B_PP_Q, B_PP_R, B_PP_S, B_PP_T,
    B_PAF_A, B_PAF_B, B_PAF_C, B_PAF_D, B_PAF_E, B_PAF_F, B_PAF_G, B_PAF_H, B_PAF_I, B_PAF_J, B_CS_ARENA, B_CS_TOTAL, B_CIG_NVEC, B_CIG_PBASE, B_CIG_CHUNK, B_CIG_ACC, B_CHAIN_HDR, B_CHAIN_IRR, B_UNQ_TEXT, B_ROW_RUN,
    B_COUNT_
};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

}  // namespace wga

struct wga_ctx {
    int device = 0;
    int n_sm = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    wga::DevBuf bufs[wga::B_COUNT_];
    // pinned host staging
    void *h_out = nullptr;
    size_t h_out_cap = 0;
    uint64_t *h_scal = nullptr;  // 64 pinned u64 for small D2H reads
    // error state
    int err_code = 0;
    uint64_t err_where = 0;
    std::string err_msg;
    std::string host_text;  // library-owned host text of the last call (pafpseudo file index, validate report)
    void *tok_cache = nullptr;  // TokCache (wga_b200.cu): index of the last tokenised MAF slab
    // accounting
    uint64_t n_own = 0, n_cub = 0, n_nccl = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float walk_ms = 0.f;
    uint64_t walk_bytes = 0;
    uint32_t walk_launches = 0;
    // state of the last prepared slab
    const uint8_t *d_text = nullptr;
    const char *h_text = nullptr;  // host copy when the input was a host buffer
    uint64_t text_len = 0;
    uint64_t n_ws = 0, n_nl = 0, n_lines = 0, n_rec = 0, n_items = 0, n_runs = 0;
    bool last_counts_valid = false;
    int last_counts_kind = 0;  // 0 maf (5 partial counters/record), 1 paf (11 counters)

    void set_error(int code, uint64_t where, const std::string &msg) {
        err_code = code;
        err_where = where;
        err_msg = msg;
    }

    // returns nullptr on failure (error recorded)
    void *ensure(wga::BufId id, size_t bytes) {
        wga::DevBuf &b = bufs[id];
        if (bytes == 0) bytes = 16;
        if (b.cap >= bytes) return b.p;
        if (b.p) {
            cudaStreamSynchronize(stream);
            cudaFree(b.p);
            b.p = nullptr;
            b.cap = 0;
        }
        size_t want = bytes + bytes / 8 + 256;
        want = (want + 255) & ~(size_t)255;
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) {
            want = (bytes + 255) & ~(size_t)255;
            e = cudaMalloc(&b.p, want);
        }
        if (e != cudaSuccess) {
            b.p = nullptr;
            set_error(WGA_E_NOMEM, 0, std::string("cudaMalloc failed for ") + std::to_string(bytes) + " bytes: " +
                                          cudaGetErrorString(e));
            return nullptr;
        }
        b.cap = want;
        return b.p;
    }
    template <class T>
    T *buf(wga::BufId id, size_t count) {
        return (T *)ensure(id, count * sizeof(T));
    }
    // pinned scratch for table-sized D2H reads (pageable destinations go through a slow staging copy)
    void *h_scratch = nullptr;
    size_t h_scratch_cap = 0;
    void *ensure_host_scratch(size_t bytes) {
        if (h_scratch_cap >= bytes && h_scratch) return h_scratch;
        if (h_scratch) cudaFreeHost(h_scratch);
        h_scratch = nullptr;
        size_t want = bytes + bytes / 4 + 4096;
        if (cudaHostAlloc(&h_scratch, want, cudaHostAllocDefault) != cudaSuccess) {
            h_scratch = nullptr;
            h_scratch_cap = 0;
            set_error(WGA_E_NOMEM, 0, "cudaHostAlloc failed for host scratch");
            return nullptr;
        }
        h_scratch_cap = want;
        return h_scratch;
    }
    void *ensure_host_out(size_t bytes) {
        if (h_out_cap >= bytes && h_out) return h_out;
        if (h_out) cudaFreeHost(h_out);
        h_out = nullptr;
        size_t want = bytes + bytes / 8 + 4096;
        if (cudaHostAlloc(&h_out, want, cudaHostAllocDefault) != cudaSuccess) {
            h_out = nullptr;
            h_out_cap = 0;
            set_error(WGA_E_NOMEM, 0, "cudaHostAlloc failed for output staging");
            return nullptr;
        }
        h_out_cap = want;
        return h_out;
    }
};

namespace wga {

// ---- exclusive sums through CUB (utility scans over small index arrays; not the hot kernels)
template <class InT, class OutT>
static int exclusive_sum(wga_ctx *ctx, const InT *in, OutT *out, uint64_t n) {
    if (n == 0) return WGA_OK;
    size_t tmp = 0;
    WGA_CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int64_t)n, ctx->stream));
    void *t = ctx->ensure(B_CUBTMP, tmp);
    if (!t) return ctx->err_code;
    WGA_CK(cub::DeviceScan::ExclusiveSum(t, tmp, in, out, (int64_t)n, ctx->stream));
    ctx->n_cub++;
    return WGA_OK;
}

// running maximum (utility: turns "index j scattered at the first row of run j" into the row -> run map of call.cuh)
struct MaxU32 {
    __host__ __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};
static int inclusive_max_u32(wga_ctx *ctx, uint32_t *io, uint64_t n) {
    if (n == 0) return WGA_OK;
    size_t tmp = 0;
    WGA_CK(cub::DeviceScan::InclusiveScan(nullptr, tmp, io, io, MaxU32(), (int64_t)n, ctx->stream));
    void *t = ctx->ensure(B_CUBTMP, tmp);
    if (!t) return ctx->err_code;
    WGA_CK(cub::DeviceScan::InclusiveScan(t, tmp, io, io, MaxU32(), (int64_t)n, ctx->stream));
    ctx->n_cub++;
    return WGA_OK;
}

// stable LSD radix sort of (key, value) pairs (utility: orders the small per-record index arrays of pafpseudo)
static int sort_pairs_u64(wga_ctx *ctx, const uint64_t *k_in, uint64_t *k_out, const uint64_t *v_in, uint64_t *v_out, uint64_t n) {
    if (n == 0) return WGA_OK;
    size_t tmp = 0;
    WGA_CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, (int64_t)n, 0, 64, ctx->stream));
    void *t = ctx->ensure(B_CUBTMP, tmp);
    if (!t) return ctx->err_code;
    WGA_CK(cub::DeviceRadixSort::SortPairs(t, tmp, k_in, k_out, v_in, v_out, (int64_t)n, 0, 64, ctx->stream));
    ctx->n_cub++;
    return WGA_OK;
}

// ---- warp helpers
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t *total) {
    uint32_t lane = threadIdx.x & 31, x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= (uint32_t)d) x += y;
    }
    *total = __shfl_sync(0xFFFFFFFFu, x, 31);
    return x - v;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}

__device__ __forceinline__ uint4 ldg16(const uint8_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }

// 0x80 flags for the bytes [from, to) of a word (indices clamp to 0..4)
__device__ __forceinline__ uint32_t range_flags(int from, int to) {
    if (to <= from || to <= 0 || from >= 4) return 0u;
    uint32_t m_to = to >= 4 ? 0xFFFFFFFFu : ((1u << (8 * to)) - 1u);
    uint32_t m_from = from <= 0 ? 0u : ((1u << (8 * from)) - 1u);
    return m_to & ~m_from & kHi;
}

}  // namespace wga
