// wga_b200.cu -- C ABI entry points (include/wga_b200.h) and host orchestration of the kernels.
#include <algorithm>
#include <numeric>

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <cstring>
#include <new>
#include <string>
#include <string_view>
#include <unordered_set>
#include <vector>
#include "ctx.cuh"
#include "emit_paf.cuh"
#include "stat.cuh"
#include "stat_acc.hpp"
#include "tokenize.cuh"
#include "chain.cuh"
#include "walk_types.cuh"
#include "call.cuh"
#include "paf.cuh"
#include "paf2maf.cuh"
#include "call_paf.cuh"
#include "pafpseudo.cuh"
#include "validate.cuh"
#include "pafcov.cuh"
#include "filter.cuh"
#include "plot.cuh"
#include "chain_in.cuh"

using namespace wga;

// No C++ exception crosses the C ABI (include/wga_b200.h: entry points never abort or throw): host-side STL work runs
// under this guard.
template <class F>
static int no_throw(F &&f) {
    try {
        return f();
    } catch (const std::bad_alloc &) {
        return WGA_E_NOMEM;
    } catch (...) {
        return WGA_E_INTERNAL;
    }
}

namespace {

inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

#define LAUNCH(ctx, kernel, grid, block, ...)                          \
    do {                                                               \
        kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__);    \
        (ctx)->n_own++;                                                \
    } while (0)

int sync_check(wga_ctx *ctx) {
    WGA_CK(cudaStreamSynchronize(ctx->stream));
    WGA_CK(cudaGetLastError());
    return WGA_OK;
}

// copy `count` u64 scalars device -> pinned host and wait
int read_scalars(wga_ctx *ctx, std::initializer_list<const void *> srcs, uint64_t *dst) {
    int i = 0;
    for (const void *s : srcs) {
        WGA_CK(cudaMemcpyAsync(ctx->h_scal + i, s, 8, cudaMemcpyDeviceToHost, ctx->stream));
        ++i;
    }
    WGA_TRY(sync_check(ctx));
    for (int k = 0; k < i; ++k) dst[k] = ctx->h_scal[k];
    return WGA_OK;
}

// host restatement of parse_sline's error wording (src/parser/maf.rs:138-211, src/errors.rs)
std::string sline_error_text(const std::string &line, int code) {
    static const char *fname[7] = {"mode", "name", "start", "align_size", "strand", "size", "seq"};
    std::vector<std::string> f;
    size_t i = 0;
    while (i < line.size()) {
        while (i < line.size() && nat_ws(line[i])) ++i;
        if (i >= line.size()) break;
        size_t b = i;
        while (i < line.size() && !nat_ws(line[i])) ++i;
        f.push_back(line.substr(b, i - b));
    }
    auto bad_int = [](const std::string &s) {
        uint64_t v;
        return !parse_u64_bytes((const unsigned char *)s.data(), s.size(), &v);
    };
    for (int k = 0; k < 7; ++k) {
        if ((size_t)k >= f.size()) return std::string("Parse MAF error by: S-line Filed `") + fname[k] + "` Missing";
        if ((k == 2 || k == 3 || k == 5) && bad_int(f[k])) return "Parse `" + f[k] + "` Into Integer Error";
        if (k == 4 && f[k] != "+" && f[k] != "-") return "Parse Strand `" + f[k] + "` Error";
    }
    if (f.size() > 7) return "Parse MAF error by: Surplus Filed > 7";
    return "MAF parse error code " + std::to_string(code);
}

__global__ void k_line_bounds(LineView lv, uint64_t k, uint64_t *out) {
    out[0] = lv.start(k);
    out[1] = lv.end(k);
}

__global__ void k_init_u64(unsigned long long *p, uint64_t n, unsigned long long v) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_pt_init(unsigned long long *vals, uint64_t cap) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap * kPairVals) return;
    vals[i] = (i % kPairVals) >= 11 ? ~0ull : 0ull;
}
__global__ void k_iota(uint64_t *p, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}
__global__ void k_each_name_len(uint64_t n_rec, const uint32_t *tl, const uint32_t *ql, uint32_t *out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n_rec) { if (i == 2 * n_rec) out[i] = 0; return; }
    out[i] = (i & 1) ? ql[i >> 1] : tl[i >> 1];
}

struct Prep {
    LineView lv;
    SLines sl;
    Recs rc;
    uint64_t *item_base = nullptr;
    ItemDesc *items = nullptr;
    unsigned long long *err_key = nullptr;
    uint64_t total_cols = 0;
    uint64_t ragged = 0;  // mode 0: some block would be seen differently by `call` (k_build_records)
};

// The delimiter index, the line table and the parsed s-line fields of the last MAF slab stay valid in the ctx pool
// until another text is staged.  A command called with WGA_F_REUSE_INDEX on the SAME device text skips straight to
// record building (which depends on the command: query_name, 1-s-line semantics).
struct Runs {
    uint64_t n_runs = 0, n_w = 0;  // n_w != 0: MODE_RUNS_ALL saw both-gap ('W') runs
    uint64_t *run_base = nullptr, *run_pos = nullptr;
    uint32_t *run_rec = nullptr;
    uint8_t *run_cls = nullptr;
};

// What the last MAF command's record building and dense walk left in the ctx pool (WGA_F_SHARE_WALK): valid only together
// with the TokCache it hangs off.  Buffers involved: B_REC_*, B_ITEM_REC (records / items), B_CNT11 (counters),
// B_RUN_POS / B_RUN_REC / B_RUN_CLS / B_RUN_LEN / B_REC_RUNFIRST (run list); every command that rewrites one of them
// drops the corresponding flag first.
struct WalkCache {
    bool recs = false;          // P, n_rec, n_items describe the slab's records (built in maf2paf/stat mode, no error)
    bool call_same = false;     // ... and `call` would build exactly the same records (no ragged / single-row block)
    bool has_q = false;
    std::string qname;
    Prep P;
    uint64_t n_rec = 0, n_items = 0;
    bool counts = false;        // B_CNT11 = the 11 counters of every record
    bool runs = false;          // caller-classified run list + run_len + rec_run_first
    Runs RU;
    uint64_t *rec_run_first = nullptr, *run_len = nullptr;
    void drop() { recs = call_same = counts = runs = false; }
    bool same_q(const char *q) const { return q ? (has_q && qname == q) : !has_q; }
};

struct TokCache {
    WalkCache wc;
    bool valid = false;
    const void *src = nullptr;      // the caller's pointer
    const uint8_t *d_text = nullptr;  // where the kernels read it (== src, or the realigned copy)
    uint64_t len = 0;
    uint32_t first_slab = 0;
    LineView lv;
    SLines sl;
    uint64_t *rec_first = nullptr, *rec_last = nullptr;
    uint64_t n_lines = 0, n_rec = 0;
};
static TokCache &tok_of(wga_ctx *ctx) {
    if (!ctx->tok_cache) ctx->tok_cache = new TokCache();
    return *(TokCache *)ctx->tok_cache;
}

// stage the slab on the device (host input) or adopt the device pointer
int stage_text(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags = 0) {
    ctx->last_counts_valid = false;
    ctx->text_len = len;
    ctx->h_text = nullptr;
    TokCache &tc = tok_of(ctx);
    if ((flags & WGA_F_REUSE_INDEX) && mem == WGA_MEM_DEVICE && tc.valid && tc.src == text && tc.len == len &&
        tc.first_slab == (flags & WGA_F_FIRST_SLAB)) {
        ctx->d_text = tc.d_text;  // same bytes as last time: keep the staged copy and the index
        return WGA_OK;
    }
    tc.valid = false;
    tc.wc.drop();
    if (mem == WGA_MEM_DEVICE) {
        if (reinterpret_cast<uintptr_t>(text) & 15u) {  // e.g. a slab of a resident text: realign with one D2D copy
            uint8_t *d = (uint8_t *)ctx->ensure(B_TEXT, len + 64);
            if (!d) return ctx->err_code;
            if (len) WGA_CK(cudaMemcpyAsync(d, text, len, cudaMemcpyDeviceToDevice, ctx->stream));
            WGA_CK(cudaMemsetAsync(d + len, 0, 64, ctx->stream));
            ctx->d_text = d;
            tc.src = text;
            return WGA_OK;
        }
        ctx->d_text = (const uint8_t *)text;
        tc.src = text;
        return WGA_OK;
    }
    if (mem != WGA_MEM_HOST) {
        ctx->set_error(WGA_E_INVALID_ARG, 0, "mem must be WGA_MEM_HOST or WGA_MEM_DEVICE");
        return WGA_E_INVALID_ARG;
    }
    uint8_t *d = (uint8_t *)ctx->ensure(B_TEXT, len + 64);
    if (!d) return ctx->err_code;
    if (len) WGA_CK(cudaMemcpyAsync(d, text, len, cudaMemcpyHostToDevice, ctx->stream));
    WGA_CK(cudaMemsetAsync(d + len, 0, 64, ctx->stream));
    ctx->d_text = d;
    ctx->h_text = (const char *)text;
    tc.src = nullptr;  // host inputs are uploaded on every call: nothing to reuse
    return WGA_OK;
}

int fetch_line(wga_ctx *ctx, const LineView &lv, uint64_t k, std::string *line) {
    uint64_t *d = ctx->buf<uint64_t>(B_SCALARS, 8);
    if (!d) return ctx->err_code;
    LAUNCH(ctx, k_line_bounds, 1, 1, lv, k, d);
    uint64_t se[2];
    WGA_TRY(read_scalars(ctx, {d, d + 1}, se));
    uint64_t n = std::min<uint64_t>(se[1] - se[0], 1u << 16);
    line->resize(n);
    if (n) WGA_CK(cudaMemcpy(&(*line)[0], lv.text + se[0], n, cudaMemcpyDeviceToHost));
    while (!line->empty() && line->back() == '\r') line->pop_back();
    return WGA_OK;
}

// turn the device error key into code + reference wording
int resolve_error(wga_ctx *ctx, const Prep &P, uint64_t key, const char *qname) {
    int code = (int)(key & 0xFF);
    uint64_t where = key >> 8;
    std::string msg;
    switch (code) {
        case WGA_E_PARSE_INT: case WGA_E_PARSE_STRAND: case WGA_E_MAF_FIELD_MISSING: case WGA_E_MAF_SURPLUS_FIELD: {
            std::string line;
            if (fetch_line(ctx, P.lv, where, &line) == WGA_OK) msg = sline_error_text(line, code);
            break;
        }
        case WGA_E_QUERY_NAME_NOT_FOUND:
            msg = std::string("Query name:") + (qname ? qname : "") + " not found in MAF";
            break;
        case WGA_E_REF_PANIC:
            msg = "the reference panics on this block (s-line index out of range: a block needs a target and a query s-line)";
            break;
        default: msg = "internal error (pair-key collision or inconsistent index)"; break;
    }
    ctx->set_error(code, where, msg);
    return code;
}

// delimiter index of the slab: single-pass fast path, exact two-pass fallback when a 4 KiB tile overflows its slots.
// CLS 0 = ASCII whitespace (MAF), CLS 1 = tab/newline (PAF).  Leaves the last text byte in ctx->h_scal[8].
template <int CLS, int CAP>
int build_delim_index(wga_ctx *ctx, const uint8_t *text, uint64_t len, uint64_t *p_n_ws, uint64_t *p_n_nl, uint64_t **p_ws_ent,
                      uint64_t **p_nl_idx) {
    uint64_t n_ws = 0, n_nl = 0;
    uint64_t *ws_ent = nullptr, *nl_idx = nullptr;
    bool indexed = false;
    {
        const uint64_t n_tiles = (len + kTokTile - 1) / kTokTile;
        uint64_t *tile_ws = ctx->buf<uint64_t>(B_TILE_WS, n_tiles + 1), *tile_nl = ctx->buf<uint64_t>(B_TILE_NL, n_tiles + 1);
        uint64_t *ws_off = ctx->buf<uint64_t>(B_TILE_WS_OFF, n_tiles + 1), *nl_off = ctx->buf<uint64_t>(B_TILE_NL_OFF, n_tiles + 1);
        uint16_t *slots = ctx->buf<uint16_t>(B_PAF_A, n_tiles * CAP);
        uint32_t *ovf = ctx->buf<uint32_t>(B_SCALARS, 16);
        if (!tile_ws || !tile_nl || !ws_off || !nl_off || !slots || !ovf) return ctx->err_code;
        WGA_CK(cudaMemsetAsync(tile_ws + n_tiles, 0, 8, ctx->stream));
        WGA_CK(cudaMemsetAsync(tile_nl + n_tiles, 0, 8, ctx->stream));
        WGA_CK(cudaMemsetAsync(ovf, 0, 4, ctx->stream));
        uint64_t want = (n_tiles + kTokWarps - 1) / kTokWarps, cap = (uint64_t)ctx->n_sm * 8;
        auto k_scan = k_ws_scan<CLS, CAP>;
        LAUNCH(ctx, k_scan, (unsigned)(want < cap ? want : cap), kTokWarps * 32, text, len, n_tiles, tile_ws, tile_nl, slots, ovf);
        WGA_TRY(exclusive_sum(ctx, tile_ws, ws_off, n_tiles + 1));
        WGA_TRY(exclusive_sum(ctx, tile_nl, nl_off, n_tiles + 1));
        WGA_CK(cudaMemcpyAsync(ctx->h_scal + 8, text + len - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
        uint64_t tot[3];
        WGA_TRY(read_scalars(ctx, {ws_off + n_tiles, nl_off + n_tiles, ovf}, tot));
        if ((uint32_t)tot[2] == 0) {
            n_ws = tot[0]; n_nl = tot[1];
            ws_ent = ctx->buf<uint64_t>(B_WS, n_ws + 1); nl_idx = ctx->buf<uint64_t>(B_NLIDX, n_nl + 1);
            if (!ws_ent || !nl_idx) return ctx->err_code;
            auto k_comp = k_ws_compact<CAP>;
            LAUNCH(ctx, k_comp, grid_for(n_tiles * 32, 256), 256, n_tiles, tile_ws, ws_off, nl_off, slots, ws_ent, nl_idx);
            indexed = true;
        }
    }
    if (!indexed) {
        const uint64_t n_tiles = (len + kWsTile - 1) / kWsTile;
        uint64_t *tile_ws = ctx->buf<uint64_t>(B_TILE_WS, n_tiles + 1), *tile_nl = ctx->buf<uint64_t>(B_TILE_NL, n_tiles + 1);
        uint64_t *ws_off = ctx->buf<uint64_t>(B_TILE_WS_OFF, n_tiles + 1), *nl_off = ctx->buf<uint64_t>(B_TILE_NL_OFF, n_tiles + 1);
        if (!tile_ws || !tile_nl || !ws_off || !nl_off) return ctx->err_code;
        WGA_CK(cudaMemsetAsync(tile_ws + n_tiles, 0, 8, ctx->stream));
        WGA_CK(cudaMemsetAsync(tile_nl + n_tiles, 0, 8, ctx->stream));
        auto k_cnt = k_ws_count<CLS>;
        LAUNCH(ctx, k_cnt, (unsigned)n_tiles, kWsThreads, text, len, tile_ws, tile_nl);
        WGA_TRY(exclusive_sum(ctx, tile_ws, ws_off, n_tiles + 1));
        WGA_TRY(exclusive_sum(ctx, tile_nl, nl_off, n_tiles + 1));
        uint64_t tot[2];
        WGA_TRY(read_scalars(ctx, {ws_off + n_tiles, nl_off + n_tiles}, tot));
        n_ws = tot[0]; n_nl = tot[1];
        ws_ent = ctx->buf<uint64_t>(B_WS, n_ws + 1); nl_idx = ctx->buf<uint64_t>(B_NLIDX, n_nl + 1);
        if (!ws_ent || !nl_idx) return ctx->err_code;
        auto k_fill = k_ws_fill<CLS>;
        LAUNCH(ctx, k_fill, (unsigned)n_tiles, kWsThreads, text, len, ws_off, nl_off, ws_ent, nl_idx);
        WGA_CK(cudaMemcpyAsync(ctx->h_scal + 8, text + len - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
        WGA_TRY(sync_check(ctx));
    }
    *p_n_ws = n_ws; *p_n_nl = n_nl; *p_ws_ent = ws_ent; *p_nl_idx = nl_idx;
    return WGA_OK;
}

// tokenise the slab: whitespace index -> lines -> s-lines -> records -> work items
// qmatch: '\n'-separated names a --query-regex matched (see NameSet, tokenize.cuh); only looked at when qname is null
int prepare_maf(wga_ctx *ctx, uint32_t flags, const char *qname, int mode, Prep *P, const char *qmatch = nullptr) {
    const uint64_t len = ctx->text_len;
    const uint8_t *text = ctx->d_text;
    ctx->n_rec = ctx->n_items = ctx->n_runs = 0;
    P->err_key = ctx->buf<unsigned long long>(B_ERR, 2);  // [1]: "call would build other records" (k_build_records)
    if (!P->err_key) return ctx->err_code;
    LAUNCH(ctx, k_init_u64, 1, 1, P->err_key, 1, kNoErr);
    WGA_CK(cudaMemsetAsync(P->err_key + 1, 0, 8, ctx->stream));
    TokCache &tc = tok_of(ctx);
    tc.wc.drop();  // the record / item / run buffers are about to be rewritten
    P->ragged = 0;
    if (len == 0) return WGA_OK;
    LineView &lv = P->lv;
    SLines &sl = P->sl;
    Recs &rc = P->rc;
    uint64_t n_lines, n_rec;
    unsigned long long *tok_err = ctx->buf<unsigned long long>(B_TOK_ERR, 1);
    if (!tok_err) return ctx->err_code;
    if ((flags & WGA_F_REUSE_INDEX) && tc.valid && tc.d_text == text && tc.len == len) {
        // ---- the slab was tokenised by the previous command: take its index, restart from its error state
        lv = tc.lv; sl = tc.sl; rc.first = tc.rec_first; rc.last = tc.rec_last;
        n_lines = tc.n_lines; n_rec = tc.n_rec;
        ctx->n_ws = lv.n_ws; ctx->n_nl = lv.n_nl; ctx->n_lines = n_lines; ctx->n_rec = n_rec;
        WGA_CK(cudaMemcpyAsync(P->err_key, tok_err, 8, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
    // 1. whitespace index
    uint64_t n_ws = 0, n_nl = 0;
    uint64_t *ws_ent = nullptr, *nl_idx = nullptr;
    uint8_t last_byte = 0;
    WGA_TRY((build_delim_index<0, kTokCap>(ctx, text, len, &n_ws, &n_nl, &ws_ent, &nl_idx)));
    last_byte = *(uint8_t *)(ctx->h_scal + 8);
    n_lines = n_nl + (last_byte != '\n' ? 1 : 0);
    ctx->n_ws = n_ws; ctx->n_nl = n_nl; ctx->n_lines = n_lines;

    lv.text = text; lv.len = len; lv.ws_ent = ws_ent; lv.nl_ws_idx = nl_idx;
    lv.n_ws = n_ws; lv.n_nl = n_nl; lv.n_lines = n_lines; lv.first_slab = (flags & WGA_F_FIRST_SLAB) ? 1u : 0u;

    // 2. lines -> record heads
    uint8_t *lflags = ctx->buf<uint8_t>(B_LINE_FLAGS, n_lines + 1);
    uint32_t *lhead = ctx->buf<uint32_t>(B_LINE_HEAD, n_lines + 1), *lscan = ctx->buf<uint32_t>(B_LINE_REC, n_lines + 1);
    if (!lflags || !lhead || !lscan) return ctx->err_code;
    WGA_CK(cudaMemsetAsync(lhead + n_lines, 0, 4, ctx->stream));
    LAUNCH(ctx, k_line_flags, grid_for(n_lines, 256), 256, lv, lflags, lhead);
    WGA_TRY(exclusive_sum(ctx, lhead, lscan, n_lines + 1));
    WGA_CK(cudaMemcpyAsync(ctx->h_scal, lscan + n_lines, 4, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_TRY(sync_check(ctx));
    n_rec = *(uint32_t *)ctx->h_scal;
    ctx->n_rec = n_rec;

    // 3. s-line fields
    sl.name_off = ctx->buf<uint64_t>(B_SL_NAME_OFF, n_lines); sl.name_len = ctx->buf<uint32_t>(B_SL_NAME_LEN, n_lines);
    sl.start = ctx->buf<uint64_t>(B_SL_START, n_lines); sl.ali = ctx->buf<uint64_t>(B_SL_ALI, n_lines);
    sl.src = ctx->buf<uint64_t>(B_SL_SRC, n_lines); sl.neg = ctx->buf<uint8_t>(B_SL_NEG, n_lines);
    sl.seq_off = ctx->buf<uint64_t>(B_SL_SEQ_OFF, n_lines); sl.seq_len = ctx->buf<uint64_t>(B_SL_SEQ_LEN, n_lines);
    rc.first = ctx->buf<uint64_t>(B_REC_FIRST, n_rec + 1); rc.last = ctx->buf<uint64_t>(B_REC_LAST, n_rec + 1);
    if (!sl.name_off || !sl.name_len || !sl.start || !sl.ali || !sl.src || !sl.neg || !sl.seq_off || !sl.seq_len || !rc.first ||
        !rc.last)
        return ctx->err_code;
    LAUNCH(ctx, k_parse_slines, grid_for(n_lines, 256), 256, lv, lflags, lscan, sl, rc.first, rc.last, P->err_key);
    WGA_CK(cudaMemcpyAsync(tok_err, P->err_key, 8, cudaMemcpyDeviceToDevice, ctx->stream));
    tc.valid = true; tc.d_text = text; tc.len = len; tc.first_slab = flags & WGA_F_FIRST_SLAB;
    tc.lv = lv; tc.sl = sl; tc.rec_first = rc.first; tc.rec_last = rc.last; tc.n_lines = n_lines; tc.n_rec = n_rec;
    }
    rc.tl = ctx->buf<uint64_t>(B_REC_TL, n_rec + 1); rc.ql = ctx->buf<uint64_t>(B_REC_QL, n_rec + 1);
    rc.t_off = ctx->buf<uint64_t>(B_REC_TOFF, n_rec + 1); rc.q_off = ctx->buf<uint64_t>(B_REC_QOFF, n_rec + 1);
    rc.ncols = ctx->buf<uint64_t>(B_REC_NCOLS, n_rec + 1); rc.neg = ctx->buf<uint8_t>(B_REC_NEG, n_rec + 1);
    rc.skip = ctx->buf<uint8_t>(B_REC_SKIP, n_rec + 1); rc.n_items = ctx->buf<uint32_t>(B_REC_NITEMS, n_rec + 1);
    P->item_base = ctx->buf<uint64_t>(B_REC_ITEMBASE, n_rec + 1);
    uint64_t *colsum = ctx->buf<uint64_t>(B_GEN_H, n_rec + 1);
    if (!rc.tl || !rc.ql || !rc.t_off || !rc.q_off || !rc.ncols || !rc.neg || !rc.skip || !rc.n_items || !P->item_base || !colsum)
        return ctx->err_code;
    if (n_rec == 0) {
        uint64_t key;
        WGA_TRY(read_scalars(ctx, {P->err_key}, &key));
        if (key != kNoErr) return resolve_error(ctx, *P, key, qname);
        return WGA_OK;
    }

    // 4. records + work items
    uint8_t *d_q = nullptr;
    uint32_t qlen = qname ? (uint32_t)strlen(qname) : 0;
    NameSet qset = {nullptr, nullptr, nullptr, 0};
    std::vector<uint8_t> h_set;  // stays alive until the stream has consumed it (the read-back below synchronises)
    if (qname) {
        d_q = ctx->buf<uint8_t>(B_QNAME, qlen + 1);
        if (!d_q) return ctx->err_code;
        WGA_CK(cudaMemcpyAsync(d_q, qname, qlen, cudaMemcpyHostToDevice, ctx->stream));
    } else if (qmatch) {
        size_t off_at = 0, bytes_at = 0, n = 0;
        const int brc = no_throw([&] {
        // [hash x n | off x (n+1) | bytes], names in ascending hash order
        std::vector<std::pair<uint64_t, std::pair<const char *, uint32_t>>> nm;
        for (const char *b = qmatch; *b;) {
            const char *e = strchr(b, '\n');
            const size_t l = e ? (size_t)(e - b) : strlen(b);
            if (l) nm.push_back({fnv1a64((const uint8_t *)b, (uint32_t)l), {b, (uint32_t)l}});
            b += l + (e ? 1 : 0);
        }
        std::sort(nm.begin(), nm.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
        n = nm.size();
        size_t bytes = 0;
        for (auto &x : nm) bytes += x.second.second;
        off_at = 8 * n; bytes_at = off_at + 4 * (n + 1);
        h_set.resize(bytes_at + bytes + 8);
        uint64_t *hh = (uint64_t *)h_set.data();
        uint32_t *ho = (uint32_t *)(h_set.data() + off_at);
        uint32_t o = 0;
        for (size_t i = 0; i < n; ++i) {
            hh[i] = nm[i].first;
            ho[i] = o;
            memcpy(h_set.data() + bytes_at + o, nm[i].second.first, nm[i].second.second);
            o += nm[i].second.second;
        }
        ho[n] = o;
        return WGA_OK;
        });
        if (brc) return brc;
        d_q = (uint8_t *)ctx->ensure(B_QNAME, h_set.size());
        if (!d_q) return ctx->err_code;
        WGA_CK(cudaMemcpyAsync(d_q, h_set.data(), h_set.size(), cudaMemcpyHostToDevice, ctx->stream));
        qset.hash = (const uint64_t *)d_q; qset.off = (const uint32_t *)(d_q + off_at); qset.bytes = d_q + bytes_at; qset.n = (uint32_t)n;
    }
    LAUNCH(ctx, k_build_records, grid_for(n_rec, 256), 256, n_rec, text, sl, rc, d_q, qlen, qname ? 1 : (qmatch ? 2 : 0), qset, mode,
           P->err_key, P->err_key + 1);
    WGA_CK(cudaMemsetAsync(rc.n_items + n_rec, 0, 4, ctx->stream));
    WGA_CK(cudaMemsetAsync(rc.ncols + n_rec, 0, 8, ctx->stream));
    WGA_TRY(exclusive_sum(ctx, rc.n_items, P->item_base, n_rec + 1));
    WGA_TRY(exclusive_sum(ctx, rc.ncols, colsum, n_rec + 1));
    uint64_t sc[4];
    WGA_TRY(read_scalars(ctx, {P->item_base + n_rec, P->err_key, colsum + n_rec, P->err_key + 1}, sc));
    if (sc[1] != kNoErr) return resolve_error(ctx, *P, sc[1], qname);
    ctx->n_items = sc[0];
    P->total_cols = sc[2];
    P->ragged = sc[3];
    return WGA_OK;
}

unsigned walk_grid(wga_ctx *ctx, uint64_t n_items) {  // persistent warps: kWalkCtasPerSm CTAs per SM
    uint64_t want = (n_items + kWalkWarps - 1) / kWalkWarps;
    uint64_t cap = (uint64_t)ctx->n_sm * kWalkCtasPerSm;
    return (unsigned)(want < cap ? want : cap);
}

void walk_begin(wga_ctx *ctx) {
    ctx->walk_ms = 0.f;
    ctx->walk_bytes = 0;
    ctx->walk_launches = 0;
}

// the dense pass: one k_walk launch (timed with CUDA events on the ctx stream); for the RUNS modes the item slot
// regions are compacted into dense run arrays afterwards.  rec_cnt must be zeroed by the caller.
template <int MODE>
int run_walk(wga_ctx *ctx, Prep &P, unsigned long long *rec_cnt, Runs *R) {
    constexpr bool RUNS = MODE != MODE_STAT;
    const uint64_t n_rec = ctx->n_rec, n_items = ctx->n_items;
    P.items = ctx->buf<ItemDesc>(B_ITEM_REC, n_items);
    uint32_t *nheads = ctx->buf<uint32_t>(B_ITEM_NHEADS, n_items + 1);
    uint32_t *ovf = ctx->buf<uint32_t>(B_SCALARS, 16);
    if (!P.items || !nheads || !ovf) return ctx->err_code;
    for (uint32_t shift = 3;; shift = 0) {
        uint64_t *slot_base = nullptr;
        uint32_t *slots = nullptr;
        if (RUNS) {
            uint64_t *need = ctx->buf<uint64_t>(B_GEN_H, n_rec + 1);
            slot_base = ctx->buf<uint64_t>(B_RUN_BASE, n_rec + 1);
            if (!need || !slot_base) return ctx->err_code;
            launch_rec_slot_need(grid_for(n_rec + 1, 256), 256, ctx->stream, n_rec, P.rc.n_items, P.rc.ncols, shift, need);
            ctx->n_own++;
            WGA_TRY(exclusive_sum(ctx, need, slot_base, n_rec + 1));
            uint64_t total;
            WGA_TRY(read_scalars(ctx, {slot_base + n_rec}, &total));
            slots = ctx->buf<uint32_t>(B_PAF_B, total + 1);
            if (!slots) return ctx->err_code;
        }
        launch_fill_item_desc(grid_for(n_rec, 256), 256, ctx->stream, n_rec, P.rc.n_items, P.item_base, P.rc.t_off, P.rc.q_off,
                              P.rc.ncols, (const uint64_t *)slot_base, shift, P.items);
        ctx->n_own++;
        WGA_CK(cudaMemsetAsync(nheads, 0, (n_items + 1) * 4, ctx->stream));
        WGA_CK(cudaMemsetAsync(ovf, 0, 16, ctx->stream));
        WalkArgs A;
        A.text = ctx->d_text; A.items = P.items; A.n_items = n_items; A.rec_cnt = rec_cnt;
        A.item_nheads = nheads; A.slots = slots; A.cap_shift = shift; A.overflow = ovf; A.sched = ovf + 1; A.n_w = ovf + 2;
        WGA_CK(cudaEventRecord(ctx->ev0, ctx->stream));
        // hand out fewer items per request when there are few per warp (long items: the last request is the tail)
        {
            const uint64_t per_warp = n_items / ((uint64_t)ctx->n_sm * 20 + 1);
            A.grab = per_warp >= 256 ? 8u : per_warp >= 64 ? 4u : per_warp >= 16 ? 2u : 1u;
        }
        // default: the TMA-staged walk (walk_tma.cuh); WGA_WALK_TMA=0 selects the register-staged k_walk (walk.cuh)
        static const bool use_tma = !(getenv("WGA_WALK_TMA") != nullptr && getenv("WGA_WALK_TMA")[0] == '0');
        if (use_tma) WGA_CK(launch_walk_tma(MODE, walk_tma_grid(ctx->n_sm, n_items), ctx->stream, A));
        else launch_walk(MODE, walk_grid(ctx, n_items), ctx->stream, A);
        ctx->n_own++;
        WGA_CK(cudaEventRecord(ctx->ev1, ctx->stream));
        ctx->walk_launches++;
        ctx->walk_bytes += 2 * P.total_cols + 48 * n_items;
        if (!RUNS) return WGA_OK;  // caller collects the event time after its next sync
        R->run_base = ctx->buf<uint64_t>(B_GEN_H, n_items + 1);
        if (!R->run_base) return ctx->err_code;
        WGA_TRY(exclusive_sum(ctx, nheads, R->run_base, n_items + 1));
        uint64_t sc[3];
        WGA_TRY(read_scalars(ctx, {R->run_base + n_items, ovf, ovf + 2}, sc));
        R->n_w = (uint32_t)sc[2];
        float ms = 0.f;
        WGA_CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->walk_ms += ms;
        if ((uint32_t)sc[1]) {  // some item has more run heads than n/8+16: redo with worst-case capacity
            if (shift == 0) { ctx->set_error(WGA_E_INTERNAL, 0, "run slot overflow at full capacity"); return WGA_E_INTERNAL; }
            if (rec_cnt) WGA_CK(cudaMemsetAsync(rec_cnt, 0, n_rec * 5 * 8, ctx->stream));
            continue;
        }
        const uint64_t n_runs = sc[0];
        R->n_runs = ctx->n_runs = n_runs;
        R->run_pos = ctx->buf<uint64_t>(B_RUN_POS, n_runs + 1);
        R->run_rec = ctx->buf<uint32_t>(B_RUN_REC, n_runs + 1);
        R->run_cls = ctx->buf<uint8_t>(B_RUN_CLS, n_runs + 1);
        if (!R->run_pos || !R->run_rec || !R->run_cls) return ctx->err_code;
        launch_runs_compact(grid_for(n_items * 32, 256), 256, ctx->stream, n_items, P.items, nheads, R->run_base, slots, R->run_pos,
                            R->run_rec, R->run_cls);
        ctx->n_own++;
        return WGA_OK;
    }
}

// hand the command's output back in the caller's memory kind
int finish_output(wga_ctx *ctx, const void *d_out, uint64_t n, int mem, uint32_t flags, wga_span *out) {
    if ((mem == WGA_MEM_DEVICE && !(flags & WGA_F_OUT_HOST)) || n == 0) {
        out->data = n ? d_out : nullptr;
        out->len = n;
        return sync_check(ctx);
    }
    void *h = ctx->ensure_host_out(n);
    if (!h) return ctx->err_code;
    WGA_CK(cudaMemcpyAsync(h, d_out, n, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_TRY(sync_check(ctx));
    out->data = h;
    out->len = n;
    return WGA_OK;
}

}  // namespace

// ================================================================== C ABI
extern "C" {

int wga_ctx_create(int device, void *stream, wga_ctx **out) {
    if (!out) return WGA_E_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return WGA_E_CUDA;  // no CPU fallback
    if (device < 0 || device >= ndev) return WGA_E_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return WGA_E_CUDA;
    wga_ctx *ctx = new wga_ctx();
    ctx->device = device;
    cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (ctx->n_sm <= 0) ctx->n_sm = 148;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return WGA_E_CUDA; }
        ctx->own_stream = true;
    }
    if (cudaHostAlloc((void **)&ctx->h_scal, 64 * 8, cudaHostAllocDefault) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        delete ctx;
        return WGA_E_CUDA;
    }
    *out = ctx;
    return WGA_OK;
}

void wga_ctx_destroy(wga_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &b : ctx->bufs)
        if (b.p) cudaFree(b.p);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
    delete (TokCache *)ctx->tok_cache;
    if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *wga_last_error(const wga_ctx *ctx, uint64_t *where) {
    if (!ctx) return "no context (CUDA device unavailable)";
    if (where) *where = ctx->err_where;
    return ctx->err_msg.c_str();
}

void wga_ctx_launch_counts(const wga_ctx *ctx, uint64_t *own, uint64_t *cub_calls) {
    if (own) *own = ctx ? ctx->n_own : 0;
    if (cub_calls) *cub_calls = ctx ? ctx->n_cub : 0;
}

uint64_t wga_ctx_nccl_calls(const wga_ctx *ctx) { return ctx ? ctx->n_nccl : 0; }

void wga_ctx_last_walk(const wga_ctx *ctx, float *ms, uint64_t *alg_bytes, uint32_t *launches) {
    if (ms) *ms = ctx->walk_ms;
    if (alg_bytes) *alg_bytes = ctx->walk_bytes;
    if (launches) *launches = ctx->walk_launches;
}

void *wga_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void wga_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

uint64_t wga_maf_safe_cut(const char *text, uint64_t len) {
    // start offset of the last line (other than the first) that does not begin with 's'; 0 if none
    uint64_t i = len;
    while (i > 0) {
        uint64_t j = i - 1;  // start of the line that contains byte j
        while (j > 0 && text[j - 1] != '\n') --j;
        if (j == 0) break;
        if (text[j] != 's') return j;
        i = j;
    }
    return 0;
}

// ---------------------------------------------------------------- resident text
struct wga_text {
    wga_ctx *ctx = nullptr;
    uint8_t *d = nullptr;
    uint64_t len = 0;
    std::vector<uint64_t> off;  // n_slabs + 1
    std::vector<cudaEvent_t> ev;
    cudaStream_t copy = nullptr;
};

int wga_text_open(wga_ctx *ctx, const void *host_text, uint64_t len, uint64_t slab_bytes, wga_text **out) {
    if (!ctx || !out || (!host_text && len)) return WGA_E_INVALID_ARG;
    *out = nullptr;
    cudaSetDevice(ctx->device);
    if (slab_bytes == 0) slab_bytes = 256ull << 20;
    if (slab_bytes < 4096) slab_bytes = 4096;
    const char *h = (const char *)host_text;
    tok_of(ctx).valid = false;  // the resident buffer is about to change under any cached index
    // the previous resident text (same buffer) may still be read by work on the ctx stream
    WGA_CK(cudaStreamSynchronize(ctx->stream));
    uint8_t *d = (uint8_t *)ctx->ensure(B_TEXT_RES, len + 64);
    if (!d) return ctx->err_code;
    wga_text *t = new wga_text();
    t->ctx = ctx;
    t->d = d;
    t->len = len;
    if (cudaStreamCreateWithFlags(&t->copy, cudaStreamNonBlocking) != cudaSuccess) {
        delete t;
        ctx->set_error(WGA_E_CUDA, 0, "cudaStreamCreate failed");
        return WGA_E_CUDA;
    }
    uint64_t off = 0;
    t->off.push_back(0);
    while (off < len) {
        uint64_t want = std::min<uint64_t>(slab_bytes, len - off), cut = want;
        while (off + want < len) {  // cut before a line that is not an s-line; widen the window until one exists
            cut = wga_maf_safe_cut(h + off, want);
            if (cut) break;
            want = std::min<uint64_t>(want * 2, len - off);
            cut = want;
        }
        cudaEvent_t e = nullptr;
        if (cudaMemcpyAsync(d + off, h + off, cut, cudaMemcpyHostToDevice, t->copy) != cudaSuccess ||
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventRecord(e, t->copy) != cudaSuccess) {
            if (e) cudaEventDestroy(e);
            wga_text_close(t);
            ctx->set_error(WGA_E_CUDA, 0, "asynchronous upload failed");
            return WGA_E_CUDA;
        }
        t->ev.push_back(e);
        off += cut;
        t->off.push_back(off);
    }
    if (t->ev.empty()) {  // empty text: one empty slab
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        cudaEventRecord(e, t->copy);
        t->ev.push_back(e);
        t->off.push_back(0);
    }
    cudaMemsetAsync(d + len, 0, 64, t->copy);
    cudaEventRecord(t->ev.back(), t->copy);  // the last slab also covers the zero tail
    *out = t;
    return WGA_OK;
}

uint32_t wga_text_slabs(const wga_text *t) { return t ? (uint32_t)t->ev.size() : 0; }

int wga_text_slab(wga_text *t, uint32_t k, const void **dev, uint64_t *len, uint32_t *flags) {
    if (!t || k >= t->ev.size() || !dev || !len) return WGA_E_INVALID_ARG;
    wga_ctx *ctx = t->ctx;
    cudaSetDevice(ctx->device);
    WGA_CK(cudaStreamWaitEvent(ctx->stream, t->ev[k], 0));
    *dev = t->d + t->off[k];
    *len = t->off[k + 1] - t->off[k];
    if (flags) *flags = k == 0 ? WGA_F_FIRST_SLAB : 0u;
    return WGA_OK;
}

int wga_text_whole(wga_text *t, const void **dev, uint64_t *len) {
    if (!t || !dev || !len) return WGA_E_INVALID_ARG;
    wga_ctx *ctx = t->ctx;
    cudaSetDevice(ctx->device);
    WGA_CK(cudaStreamWaitEvent(ctx->stream, t->ev.back(), 0));
    *dev = t->d;
    *len = t->len;
    return WGA_OK;
}

void wga_text_close(wga_text *t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    if (t->copy) cudaStreamSynchronize(t->copy);
    for (cudaEvent_t e : t->ev) cudaEventDestroy(e);
    if (t->copy) cudaStreamDestroy(t->copy);
    delete t;
}

// ---------------------------------------------------------------- maf2paf
int wga_maf2paf(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, const char *query_name,
                wga_span *out) {
    if (!ctx || !out || (!text && len)) return WGA_E_INVALID_ARG;
    cudaSetDevice(ctx->device);
    out->data = nullptr;
    out->len = 0;
    walk_begin(ctx);
    WGA_TRY(stage_text(ctx, text, len, mem, flags));
    Prep P;
    WGA_TRY(prepare_maf(ctx, flags, query_name, 0, &P));
    const uint64_t n_rec = ctx->n_rec, n_items = ctx->n_items;
    if (n_rec == 0 || n_items == 0) return sync_check(ctx);

    unsigned long long *rec_cnt = ctx->buf<unsigned long long>(B_REC_CNT, n_rec * 5);
    if (!rec_cnt) return ctx->err_code;
    WGA_CK(cudaMemsetAsync(rec_cnt, 0, n_rec * 5 * 8, ctx->stream));
    const bool share = (flags & WGA_F_SHARE_WALK) != 0;
    Runs RU;
    if (share) WGA_TRY(run_walk<MODE_RUNS_ALL>(ctx, P, rec_cnt, &RU));
    else WGA_TRY(run_walk<MODE_RUNS_EXT>(ctx, P, rec_cnt, &RU));
    uint64_t n_runs = RU.n_runs;
    uint64_t *run_base = RU.run_base, *run_pos = RU.run_pos;
    uint32_t *run_rec = RU.run_rec;
    uint8_t *run_cls = RU.run_cls;
    uint64_t *run_len = ctx->buf<uint64_t>(B_RUN_LEN, n_runs + 1);
    uint32_t *run_dl = ctx->buf<uint32_t>(B_RUN_DL, n_runs + 1);
    uint64_t *rec_run_first = ctx->buf<uint64_t>(B_REC_RUNFIRST, n_rec + 1);
    uint64_t *row_len = ctx->buf<uint64_t>(B_ROW_LEN, n_rec + 1), *row_off = ctx->buf<uint64_t>(B_ROW_OFF, n_rec + 1);
    uint32_t *row_fixed = ctx->buf<uint32_t>(B_ROW_FIXED, n_rec + 1);
    if (!run_len || !run_dl || !rec_run_first || !row_len || !row_off || !row_fixed) return ctx->err_code;
    LAUNCH(ctx, k_rec_run_first, grid_for(n_rec + 1, 256), 256, n_rec, P.item_base, run_base, n_items, n_runs, rec_run_first);
    LAUNCH(ctx, k_run_len, grid_for(n_runs + 1, 256), 256, n_runs, run_pos, run_rec, P.rc.ncols, run_len, run_dl);
    WalkCache &wc = tok_of(ctx).wc;
    if (share) {  // what stat / call on the same slab can pick up (the counters join below)
        wc.P = P; wc.n_rec = n_rec; wc.n_items = n_items; wc.has_q = query_name != nullptr; wc.qname = query_name ? query_name : "";
        wc.recs = true; wc.call_same = P.ragged == 0;
        wc.RU = RU; wc.rec_run_first = rec_run_first; wc.run_len = run_len; wc.runs = true;
    }
    if (share && RU.n_w) {
        // both-gap columns: the PAF rows need the coarser EXT list (k_ext_flags); the caller list above stays for `call`
        const uint64_t n_all = n_runs;
        uint32_t *xflag = ctx->buf<uint32_t>(B_XRUN_FLAG, n_all + 1);
        uint64_t *xidx = ctx->buf<uint64_t>(B_XRUN_IDX, n_all + 1);
        if (!xflag || !xidx) return ctx->err_code;
        LAUNCH(ctx, k_ext_flags, grid_for(n_all + 1, 256), 256, n_all, (const uint32_t *)run_rec, (const uint8_t *)run_cls, xflag);
        WGA_TRY(exclusive_sum(ctx, xflag, xidx, n_all + 1));
        uint64_t n_x;
        WGA_TRY(read_scalars(ctx, {xidx + n_all}, &n_x));
        uint64_t *xpos = ctx->buf<uint64_t>(B_XRUN_POS, n_x + 1);
        uint32_t *xrec = ctx->buf<uint32_t>(B_XRUN_REC, n_x + 1);
        uint8_t *xcls = ctx->buf<uint8_t>(B_XRUN_CLS, n_x + 1);
        uint64_t *xfirst = ctx->buf<uint64_t>(B_GEN_C, n_rec + 1), *xlen = ctx->buf<uint64_t>(B_GEN_D, n_x + 1);
        uint32_t *xdl = ctx->buf<uint32_t>(B_GEN_E, n_x + 1);
        if (!xpos || !xrec || !xcls || !xfirst || !xlen || !xdl) return ctx->err_code;
        LAUNCH(ctx, k_ext_scatter, grid_for(n_all, 256), 256, n_all, (const uint32_t *)xflag, (const uint64_t *)xidx,
               (const uint64_t *)run_pos, (const uint32_t *)run_rec, (const uint8_t *)run_cls, xpos, xrec, xcls);
        LAUNCH(ctx, k_ext_rec_first, grid_for(n_rec + 1, 256), 256, n_rec, (const uint64_t *)rec_run_first, (const uint64_t *)xidx, xfirst);
        LAUNCH(ctx, k_run_len, grid_for(n_x + 1, 256), 256, n_x, xpos, xrec, P.rc.ncols, xlen, xdl);
        n_runs = n_x; run_pos = xpos; run_rec = xrec; run_cls = xcls; run_len = xlen; run_dl = xdl; rec_run_first = xfirst;
    }
    uint64_t *cig_off = ctx->buf<uint64_t>(B_RUN_CIGOFF, n_runs + 1);
    if (!cig_off) return ctx->err_code;
    WGA_TRY(exclusive_sum(ctx, run_dl, cig_off, n_runs + 1));
    PafRowArgs R;
    R.n_rec = n_rec; R.text = ctx->d_text; R.sl = P.sl; R.tl = P.rc.tl; R.ql = P.rc.ql; R.ncols = P.rc.ncols;
    R.skip = P.rc.skip; R.rec_cnt = rec_cnt; R.rec_run_first = rec_run_first; R.cig_off = cig_off;
    R.row_len = row_len; R.row_fixed = row_fixed; R.row_off = row_off; R.out = nullptr;
    R.run_len = run_len; R.run_dl = run_dl; R.run_cls = run_cls;
    uint64_t *any_long = ctx->buf<uint64_t>(B_SCALARS, 16) + 8;
    WGA_CK(cudaMemsetAsync(any_long, 0, 8, ctx->stream));
    R.any_long = (uint32_t *)any_long;
    LAUNCH(ctx, k_paf_rows<0>, grid_for(n_rec + 1, kPafThreads), kPafThreads, R);
    WGA_TRY(exclusive_sum(ctx, row_len, row_off, n_rec + 1));
    uint64_t tl2[2];
    WGA_TRY(read_scalars(ctx, {row_off + n_rec, any_long}, tl2));
    const uint64_t total = tl2[0];
    char *d_out = (char *)ctx->ensure(B_OUT, total + 16);
    if (!d_out) return ctx->err_code;
    R.out = d_out;
    LAUNCH(ctx, k_paf_rows<1>, grid_for(n_rec, kPafThreads), kPafThreads, R);
    if (tl2[1] & 1u)
        LAUNCH(ctx, k_paf_cigar, grid_for(n_runs, 256), 256, n_runs, (const uint32_t *)run_rec, (const uint8_t *)run_cls,
               (const uint64_t *)run_len, (const uint32_t *)run_dl, (const uint64_t *)cig_off, (const uint64_t *)rec_run_first,
               (const uint64_t *)row_off, (const uint32_t *)row_fixed, d_out);
    // per-record 11 counters stay available for wga_last_record_counts (and for a following wga_maf_stat)
    unsigned long long *cnt11 = ctx->buf<unsigned long long>(B_CNT11, n_rec * 11);
    if (!cnt11) return ctx->err_code;
    LAUNCH(ctx, k_maf_counts11, grid_for(n_rec, 256), 256, n_rec, rec_cnt, P.rc.ncols, P.rc.neg, P.rc.skip, cnt11);
    ctx->last_counts_valid = true;
    if (share) wc.counts = true;
    return finish_output(ctx, d_out, total, mem, flags, out);
}

// ---------------------------------------------------------------- maf2chain (converter.rs:57-92)
// kind 0: chain records (k_chain_rows); kind 1: base-level plot segments (k_plot_rows, plot.cuh) -- both are one sparse
// pass per record over the EXT-classified run list of the dense walk
static int maf_run_rows(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, const char *query_name, uint64_t id_base,
                        wga_span *out, uint64_t *n_records, int kind, uint64_t cutoff) {
    if (!ctx || !out || (!text && len)) return WGA_E_INVALID_ARG;
    cudaSetDevice(ctx->device);
    out->data = nullptr;
    out->len = 0;
    if (n_records) *n_records = 0;
    walk_begin(ctx);
    WGA_TRY(stage_text(ctx, text, len, mem, flags));
    Prep P;
    WGA_TRY(prepare_maf(ctx, flags, query_name, 0, &P));
    const uint64_t n_rec = ctx->n_rec, n_items = ctx->n_items;
    if (n_records) *n_records = n_rec;
    if (n_rec == 0 || n_items == 0) return sync_check(ctx);
    unsigned long long *rec_cnt = ctx->buf<unsigned long long>(B_REC_CNT, n_rec * 5);
    if (!rec_cnt) return ctx->err_code;
    WGA_CK(cudaMemsetAsync(rec_cnt, 0, n_rec * 5 * 8, ctx->stream));
    Runs RU;
    WGA_TRY(run_walk<MODE_RUNS_EXT>(ctx, P, rec_cnt, &RU));
    const uint64_t n_runs = RU.n_runs;
    uint64_t *run_len = ctx->buf<uint64_t>(B_RUN_LEN, n_runs + 1);
    uint64_t *rec_run_first = ctx->buf<uint64_t>(B_REC_RUNFIRST, n_rec + 1);
    uint64_t *row_len = ctx->buf<uint64_t>(B_ROW_LEN, n_rec + 1), *row_off = ctx->buf<uint64_t>(B_ROW_OFF, n_rec + 1);
    uint64_t *tno = ctx->buf<uint64_t>(B_GEN_C, n_rec), *qno = ctx->buf<uint64_t>(B_GEN_D, n_rec);
    uint32_t *tnl = ctx->buf<uint32_t>(B_GEN_E, n_rec), *qnl = ctx->buf<uint32_t>(B_GEN_F, n_rec);
    uint64_t *m6 = ctx->buf<uint64_t>(B_GEN_G, 6 * n_rec);
    if (!run_len || !rec_run_first || !row_len || !row_off || !tno || !qno || !tnl || !qnl || !m6) return ctx->err_code;
    LAUNCH(ctx, k_rec_run_first, grid_for(n_rec + 1, 256), 256, n_rec, P.item_base, RU.run_base, n_items, n_runs, rec_run_first);
    LAUNCH(ctx, k_run_len, grid_for(n_runs + 1, 256), 256, n_runs, RU.run_pos, RU.run_rec, P.rc.ncols, run_len, (uint32_t *)nullptr);
    LAUNCH(ctx, k_chain_meta_maf, grid_for(n_rec, 256), 256, n_rec, P.sl, (const uint64_t *)P.rc.tl, (const uint64_t *)P.rc.ql, tno, qno, tnl,
           qnl, m6);
    ChainView V;
    V.n_rec = n_rec; V.id_base = id_base; V.text = ctx->d_text; V.t_name_off = tno; V.q_name_off = qno; V.t_name_len = tnl;
    V.q_name_len = qnl; V.t_size = m6; V.q_size = m6 + n_rec; V.t_start = m6 + 2 * n_rec; V.t_end = m6 + 3 * n_rec;
    V.q_start = m6 + 4 * n_rec; V.q_end = m6 + 5 * n_rec; V.neg = P.rc.neg; V.skip = P.rc.skip; V.run_first = rec_run_first;
    V.run_len = run_len; V.run_type = RU.run_cls; V.maf = 1; V.row_len = row_len; V.row_off = row_off; V.out = nullptr;
    static const bool chain_seq = getenv("WGA_CHAIN_SEQ") != nullptr && getenv("WGA_CHAIN_SEQ")[0] == '1';  // one thread per record
    static const bool plot_seq = getenv("WGA_PLOT_SEQ") != nullptr && getenv("WGA_PLOT_SEQ")[0] == '1';     // one thread per record
    uint32_t *hdr_len = ctx->buf<uint32_t>(B_CHAIN_HDR, n_rec + 1);
    uint8_t *irregular = ctx->buf<uint8_t>(B_CHAIN_IRR, n_rec + 1);
    if (!hdr_len || !irregular) return ctx->err_code;
    if (kind == 0 && chain_seq) LAUNCH(ctx, k_chain_rows<0>, grid_for(n_rec + 1, 128), 128, V);
    else if (kind == 0) LAUNCH(ctx, k_chain_rows_warp<0>, grid_for((n_rec + 1) * 32, 128), 128, V, hdr_len, irregular);
    else if (plot_seq) LAUNCH(ctx, k_plot_rows<0>, grid_for(n_rec + 1, 128), 128, V, cutoff);
    else LAUNCH(ctx, k_plot_rows_warp<0>, grid_for((n_rec + 1) * 32, 128), 128, V, cutoff);
    WGA_TRY(exclusive_sum(ctx, row_len, row_off, n_rec + 1));
    uint64_t total;
    WGA_TRY(read_scalars(ctx, {row_off + n_rec}, &total));
    if (total == 0) return WGA_OK;
    char *d_out = (char *)ctx->ensure(B_OUT, total + 16);
    if (!d_out) return ctx->err_code;
    V.out = d_out;
    if (kind == 0 && chain_seq) LAUNCH(ctx, k_chain_rows<1>, grid_for(n_rec, 128), 128, V);
    else if (kind == 0) LAUNCH(ctx, k_chain_rows_warp<1>, grid_for(n_rec * 32, 128), 128, V, hdr_len, irregular);
    else if (plot_seq) LAUNCH(ctx, k_plot_rows<1>, grid_for(n_rec, 128), 128, V, cutoff);
    else LAUNCH(ctx, k_plot_rows_warp<1>, grid_for(n_rec * 32, 128), 128, V, cutoff);
    return finish_output(ctx, d_out, total, mem, flags, out);
}
int wga_maf2chain(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, const char *query_name, uint64_t id_base,
                  wga_span *out, uint64_t *n_records) {
    return maf_run_rows(ctx, text, len, mem, flags, query_name, id_base, out, n_records, 0, 0);
}
// ---------------------------------------------------------------- dotplot -m base-level, MAF (cigar.rs:955-985)
int wga_maf_plotdata(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, const char *query_name,
                     uint64_t skip_cutoff, wga_span *out) {
    return maf_run_rows(ctx, text, len, mem, flags, query_name, 0, out, nullptr, 1, skip_cutoff);
}

// ---------------------------------------------------------------- stat
wga_stat_acc *wga_stat_acc_create(int each) {
    wga_stat_acc *a = new (std::nothrow) wga_stat_acc();
    if (a) a->each = each ? 1 : 0;
    return a;
}
void wga_stat_acc_destroy(wga_stat_acc *acc) { delete acc; }
uint64_t wga_stat_acc_npairs(const wga_stat_acc *acc) { return acc ? acc->rows.size() : 0; }
int wga_stat_acc_write(wga_stat_acc *acc, wga_span *out) {
    if (!acc || !out) return WGA_E_INVALID_ARG;
    out->data = nullptr; out->len = 0;
    return no_throw([&] {
        acc->write();
        out->data = acc->text.data();
        out->len = acc->text.size();
        return (int)WGA_OK;
    });
}
int wga_stat_acc_export(wga_stat_acc *acc, wga_span *blob) {
    if (!acc || !blob) return WGA_E_INVALID_ARG;
    blob->data = nullptr; blob->len = 0;
    return no_throw([&] {
        acc->export_blob();
        blob->data = acc->blob.data();
        blob->len = acc->blob.size();
        return (int)WGA_OK;
    });
}
int wga_stat_acc_import(wga_stat_acc *acc, const void *blob, uint64_t len) {
    if (!acc || !blob) return WGA_E_INVALID_ARG;
    return no_throw([&] { return acc->import_blob((const char *)blob, len); });
}

}  // extern "C"

namespace {

// shared tail of stat_maf / stat_paf: pair reduction on the device, merge into the host accumulator
int stat_reduce(wga_ctx *ctx, uint64_t n_rec, const RecMeta &M, const unsigned long long *cnt11,
                unsigned long long *err_key, wga_stat_acc *acc) {
    static const bool dbg = getenv("WGA_DEBUG_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!dbg) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[stat_reduce] %-18s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    std::vector<uint64_t> h_first, h_vals, h_sizes, h_inv_num;
    std::vector<uint32_t> h_nlen, h_inv_pair;
    std::vector<char> h_names;
    uint64_t n_pairs = 0, n_inv = 0;
    if (!acc->each) {
        uint64_t cap = 1024;
        while (cap < 2 * n_rec) cap <<= 1;
        unsigned long long *keys = ctx->buf<unsigned long long>(B_PT_KEYS, cap);
        unsigned long long *first = ctx->buf<unsigned long long>(B_PT_FIRST, cap);
        unsigned long long *vals = ctx->buf<unsigned long long>(B_PT_VALS, cap * kPairVals);
        uint64_t *rec_slot = ctx->buf<uint64_t>(B_REC_SLOT, n_rec);
        uint32_t *flag = ctx->buf<uint32_t>(B_PT_FLAG, cap + 1), *idx = ctx->buf<uint32_t>(B_PT_IDX, cap + 1);
        if (!keys || !first || !vals || !rec_slot || !flag || !idx) return ctx->err_code;
        WGA_CK(cudaMemsetAsync(keys, 0, cap * 8, ctx->stream));
        WGA_CK(cudaMemsetAsync(first, 0xFF, cap * 8, ctx->stream));
        LAUNCH(ctx, k_pt_init, grid_for(cap * kPairVals, 256), 256, vals, cap);
        LAUNCH(ctx, k_pair_insert, grid_for(n_rec, 256), 256, n_rec, M, cnt11, cap - 1, keys, first, vals, rec_slot);
        LAUNCH(ctx, k_pair_verify, grid_for(n_rec, 256), 256, n_rec, M, first, rec_slot, err_key);
        WGA_CK(cudaMemsetAsync(flag + cap, 0, 4, ctx->stream));
        LAUNCH(ctx, k_slot_flags, grid_for(cap, 256), 256, cap, keys, flag);
        WGA_TRY(exclusive_sum(ctx, flag, idx, cap + 1));
        // '-'-strand records in order
        uint32_t *iflag = ctx->buf<uint32_t>(B_INV_FLAG, n_rec + 1);
        uint64_t *iidx = ctx->buf<uint64_t>(B_INV_IDX, n_rec + 1);
        if (!iflag || !iidx) return ctx->err_code;
        WGA_CK(cudaMemsetAsync(iflag + n_rec, 0, 4, ctx->stream));
        LAUNCH(ctx, k_inv_flags, grid_for(n_rec, 256), 256, n_rec, M.neg, M.skip, iflag);
        WGA_TRY(exclusive_sum(ctx, iflag, iidx, n_rec + 1));
        WGA_CK(cudaMemcpyAsync(ctx->h_scal, idx + cap, 4, cudaMemcpyDeviceToHost, ctx->stream));
        WGA_CK(cudaMemcpyAsync(ctx->h_scal + 1, iidx + n_rec, 8, cudaMemcpyDeviceToHost, ctx->stream));
        WGA_CK(cudaMemcpyAsync(ctx->h_scal + 2, err_key, 8, cudaMemcpyDeviceToHost, ctx->stream));
        WGA_TRY(sync_check(ctx));
        lap("table build + sync");
        n_pairs = *(uint32_t *)ctx->h_scal;
        n_inv = ctx->h_scal[1];
        if (ctx->h_scal[2] != kNoErr) {
            ctx->set_error(WGA_E_INTERNAL, ctx->h_scal[2] >> 8, "pair-key hash collision; refusing to merge distinct pairs");
            return WGA_E_INTERNAL;
        }
        uint64_t *p_first = ctx->buf<uint64_t>(B_PAIR_FIRST, n_pairs + 1);
        uint64_t *p_vals = ctx->buf<uint64_t>(B_PAIR_VALS, (n_pairs + 1) * kPairVals);
        uint64_t *p_sizes = ctx->buf<uint64_t>(B_GEN_B, (n_pairs + 1) * 3);
        uint32_t *nlen = ctx->buf<uint32_t>(B_NAME_LEN, 2 * n_pairs + 1);
        uint64_t *noff = ctx->buf<uint64_t>(B_NAME_OFF, 2 * n_pairs + 1);
        uint32_t *inv_pair = ctx->buf<uint32_t>(B_INV_PAIR, n_inv + 1);
        uint64_t *inv_num = ctx->buf<uint64_t>(B_INV_NUM, n_inv + 1);
        if (!p_first || !p_vals || !p_sizes || !nlen || !noff || !inv_pair || !inv_num) return ctx->err_code;
        WGA_CK(cudaMemsetAsync(nlen + 2 * n_pairs, 0, 4, ctx->stream));
        LAUNCH(ctx, k_pair_gather, grid_for(cap, 256), 256, cap, flag, idx, keys, first, vals, M, p_first, p_vals, p_sizes, nlen);
        // pairs in first-occurrence order (the reference's HashMap order is unspecified; natord sort is stable): sort
        // (first record, pair id) on the device and lay every array out in that order, so that the host merge reads
        // front to back
        uint64_t *ord_k = ctx->buf<uint64_t>(B_PP_A, n_pairs + 1), *ord_v = ctx->buf<uint64_t>(B_PP_B, n_pairs + 1);
        uint64_t *ord_i = ctx->buf<uint64_t>(B_PP_C, n_pairs + 1);
        uint64_t *vals2 = ctx->buf<uint64_t>(B_PP_D, (n_pairs + 1) * kPairVals), *sizes2 = ctx->buf<uint64_t>(B_PP_E, (n_pairs + 1) * 3);
        uint32_t *nlen2 = ctx->buf<uint32_t>(B_PP_F, 2 * n_pairs + 2), *invperm = ctx->buf<uint32_t>(B_PP_H, n_pairs + 1);
        if (!ord_k || !ord_v || !ord_i || !vals2 || !sizes2 || !nlen2 || !invperm) return ctx->err_code;
        if (n_pairs) {
            LAUNCH(ctx, k_iota, grid_for(n_pairs, 256), 256, ord_i, n_pairs);
            WGA_TRY(sort_pairs_u64(ctx, p_first, ord_k, ord_i, ord_v, n_pairs));
        }
        LAUNCH(ctx, k_pair_permute, grid_for(n_pairs + 1, 256), 256, n_pairs, (const uint64_t *)ord_v, (const uint64_t *)p_first,
               (const uint64_t *)p_vals, (const uint64_t *)p_sizes, (const uint32_t *)nlen, ord_i /* first2 */, vals2, sizes2, nlen2, invperm);
        uint64_t *first2 = ord_i;
        WGA_TRY(exclusive_sum(ctx, nlen2, noff, 2 * n_pairs + 1));
        uint64_t name_total;
        WGA_TRY(read_scalars(ctx, {noff + 2 * n_pairs}, &name_total));
        char *nbuf = (char *)ctx->ensure(B_NAME_BUF, name_total + 1);
        if (!nbuf) return ctx->err_code;
        LAUNCH(ctx, k_pair_names, grid_for(2 * n_pairs, 256), 256, n_pairs, first2, M, noff, nbuf);
        LAUNCH(ctx, k_inv_gather, grid_for(n_rec, 256), 256, n_rec, iflag, iidx, rec_slot, idx, (const uint32_t *)invperm, cnt11, inv_pair,
               inv_num);
        const size_t b_vals = n_pairs * kPairVals * 8, b_sizes = n_pairs * 24, b_nlen = n_pairs * 8;
        const size_t b_invp = (n_inv * 4 + 7) & ~(size_t)7, b_invn = n_inv * 8;
        const size_t o_vals = 0, o_sizes = o_vals + b_vals, o_nlen = o_sizes + b_sizes, o_invp = o_nlen + b_nlen, o_invn = o_invp + b_invp,
                     o_names = o_invn + b_invn;
        char *hs = (char *)ctx->ensure_host_scratch(o_names + name_total + 8);
        if (!hs) return ctx->err_code;
        if (n_pairs) {
            WGA_CK(cudaMemcpyAsync(hs + o_vals, vals2, b_vals, cudaMemcpyDeviceToHost, ctx->stream));
            WGA_CK(cudaMemcpyAsync(hs + o_sizes, sizes2, b_sizes, cudaMemcpyDeviceToHost, ctx->stream));
            WGA_CK(cudaMemcpyAsync(hs + o_nlen, nlen2, b_nlen, cudaMemcpyDeviceToHost, ctx->stream));
            if (name_total) WGA_CK(cudaMemcpyAsync(hs + o_names, nbuf, name_total, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (n_inv) {
            WGA_CK(cudaMemcpyAsync(hs + o_invp, inv_pair, n_inv * 4, cudaMemcpyDeviceToHost, ctx->stream));
            WGA_CK(cudaMemcpyAsync(hs + o_invn, inv_num, n_inv * 8, cudaMemcpyDeviceToHost, ctx->stream));
        }
        WGA_TRY(sync_check(ctx));
        lap("gather + sort + D2H");
        const uint64_t *h_vals = (const uint64_t *)(hs + o_vals), *h_sizes = (const uint64_t *)(hs + o_sizes);
        const uint32_t *h_nlen = (const uint32_t *)(hs + o_nlen);
        const uint32_t *h_inv_pair = (const uint32_t *)(hs + o_invp);
        const uint64_t *h_inv_num = (const uint64_t *)(hs + o_invn);
        const char *h_names = hs + o_names;
        if (acc->rows.empty() && acc->inv_pair.empty()) {
            // first table of this accumulator (the whole job when the file is one slab): the device already made the pairs
            // distinct and laid them out in first-occurrence order -- adopt them wholesale, no probing, no per-row copies
            acc->adopt_table(n_pairs, h_names, name_total, h_nlen, h_sizes, h_vals, kPairVals);
            acc->inv_pair.assign(h_inv_pair, h_inv_pair + n_inv);
            acc->inv_num.assign(h_inv_num, h_inv_num + n_inv);
            lap("host merge");
            return WGA_OK;
        }
        std::vector<uint32_t> map(n_pairs);
        acc->rows.reserve(acc->rows.size() + n_pairs);
        acc->names.reserve(acc->names.size() + name_total);
        acc->grow_table(acc->rows.size() + n_pairs);
        const char *nm = h_names;
        for (uint64_t p = 0; p < n_pairs; ++p) {  // first-occurrence order; the device's pair hash is the host table's key
            const uint32_t lr = h_nlen[2 * p], lq = h_nlen[2 * p + 1];
            map[p] = acc->find_or_add(nm, lr, h_sizes[3 * p], nm + lr, lq, h_sizes[3 * p + 1], h_sizes[3 * p + 2]);
            acc->add_pair_vals(map[p], &h_vals[p * kPairVals]);
            nm += lr + lq;
        }
        const size_t i0 = acc->inv_pair.size();
        acc->inv_pair.reserve(i0 + n_inv);
        for (uint64_t i = 0; i < n_inv; ++i) acc->inv_pair.push_back(map[h_inv_pair[i]]);
        acc->inv_num.insert(acc->inv_num.end(), h_inv_num, h_inv_num + n_inv);
        lap("host merge");
        return WGA_OK;
    }
    // --each: one row per record (stat.rs:129-164)
    uint64_t *iota = ctx->buf<uint64_t>(B_PAIR_FIRST, n_rec + 1);
    uint32_t *nlen = ctx->buf<uint32_t>(B_NAME_LEN, 2 * n_rec + 1);
    uint64_t *noff = ctx->buf<uint64_t>(B_NAME_OFF, 2 * n_rec + 1);
    if (!iota || !nlen || !noff) return ctx->err_code;
    LAUNCH(ctx, k_iota, grid_for(n_rec, 256), 256, iota, n_rec);
    LAUNCH(ctx, k_each_name_len, grid_for(2 * n_rec + 1, 256), 256, n_rec, M.t_name_len, M.q_name_len, nlen);
    WGA_TRY(exclusive_sum(ctx, nlen, noff, 2 * n_rec + 1));
    uint64_t name_total;
    WGA_TRY(read_scalars(ctx, {noff + 2 * n_rec}, &name_total));
    char *nbuf = (char *)ctx->ensure(B_NAME_BUF, name_total + 1);
    if (!nbuf) return ctx->err_code;
    LAUNCH(ctx, k_pair_names, grid_for(2 * n_rec, 256), 256, n_rec, iota, M, noff, nbuf);
    std::vector<uint64_t> c11(n_rec * 11), tsz(n_rec), qsz(n_rec), tst(n_rec), qst(n_rec);
    std::vector<uint8_t> neg(n_rec), skip(n_rec, 0);
    h_nlen.resize(2 * n_rec);
    h_names.resize(name_total + 1);
    WGA_CK(cudaMemcpyAsync(c11.data(), cnt11, n_rec * 88, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_CK(cudaMemcpyAsync(tsz.data(), M.t_size, n_rec * 8, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_CK(cudaMemcpyAsync(qsz.data(), M.q_size, n_rec * 8, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_CK(cudaMemcpyAsync(tst.data(), M.t_start, n_rec * 8, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_CK(cudaMemcpyAsync(qst.data(), M.q_start, n_rec * 8, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_CK(cudaMemcpyAsync(neg.data(), M.neg, n_rec, cudaMemcpyDeviceToHost, ctx->stream));
    if (M.skip) WGA_CK(cudaMemcpyAsync(skip.data(), M.skip, n_rec, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_CK(cudaMemcpyAsync(h_nlen.data(), nlen, n_rec * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (name_total) WGA_CK(cudaMemcpyAsync(h_names.data(), nbuf, name_total, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_TRY(sync_check(ctx));
    uint64_t pos = 0;
    acc->rows.reserve(acc->rows.size() + n_rec);
    for (uint64_t r = 0; r < n_rec; ++r) {
        const char *rn = h_names.data() + pos;
        pos += h_nlen[2 * r];
        const char *qn = h_names.data() + pos;
        pos += h_nlen[2 * r + 1];
        if (skip[r]) continue;
        wga_stat_acc::Row &row = acc->rows[acc->push_row(rn, h_nlen[2 * r], tsz[r], qn, h_nlen[2 * r + 1], qsz[r], 0)];
        const uint64_t *c = &c11[r * 11];
        for (int i = 0; i < 11; ++i) row.v[i] = c[i];
        row.v[11] = tst[r];
        row.v[12] = qst[r];
        if (c[10]) {
            uint64_t aligned = c[0] + c[1] + c[5] + c[9], qal = c[0] + c[1] + c[3] + c[7];
            row.inv_size = (float)(aligned + qal) / (float)(c[10] + 1);
        }
    }
    return WGA_OK;
}

}  // namespace

extern "C" {

int wga_maf_stat(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, const char *query_name,
                 wga_stat_acc *acc) {
    if (!ctx || !acc || (!text && len)) return WGA_E_INVALID_ARG;
    cudaSetDevice(ctx->device);
    walk_begin(ctx);
    WGA_TRY(stage_text(ctx, text, len, mem, flags));
    TokCache &tc = tok_of(ctx);
    WalkCache &wc = tc.wc;
    Prep P;
    unsigned long long *cnt11 = nullptr;
    // RecStat::from(Cigar) (common.rs:116-140) is a function of the 11 counters: when the previous command on this slab
    // left them behind (wga_maf2paf with WGA_F_SHARE_WALK), stat needs neither record building nor a walk
    const bool reuse = (flags & WGA_F_REUSE_INDEX) && tc.valid && tc.d_text == ctx->d_text && tc.len == len && wc.recs && wc.counts &&
                       wc.same_q(query_name);
    if (reuse) {
        P = wc.P;
        ctx->n_rec = wc.n_rec; ctx->n_items = wc.n_items;
        cnt11 = (unsigned long long *)ctx->bufs[B_CNT11].p;
    } else {
        WGA_TRY(prepare_maf(ctx, flags, query_name, 0, &P));
    }
    const uint64_t n_rec = ctx->n_rec, n_items = ctx->n_items;
    if (n_rec == 0) return sync_check(ctx);
    if (!reuse) {
        unsigned long long *rec_cnt = ctx->buf<unsigned long long>(B_REC_CNT, n_rec * 5);
        cnt11 = ctx->buf<unsigned long long>(B_CNT11, n_rec * 11);
        if (!rec_cnt || !cnt11) return ctx->err_code;
        WGA_CK(cudaMemsetAsync(rec_cnt, 0, n_rec * 5 * 8, ctx->stream));
        if (n_items) WGA_TRY(run_walk<MODE_STAT>(ctx, P, rec_cnt, nullptr));
        LAUNCH(ctx, k_maf_counts11, grid_for(n_rec, 256), 256, n_rec, rec_cnt, P.rc.ncols, P.rc.neg, P.rc.skip, cnt11);
    }
    RecMeta M;
    M.text = ctx->d_text;
    uint64_t *tno = ctx->buf<uint64_t>(B_GEN_C, n_rec), *qno = ctx->buf<uint64_t>(B_GEN_D, n_rec);
    uint32_t *tnl = ctx->buf<uint32_t>(B_GEN_E, n_rec), *qnl = ctx->buf<uint32_t>(B_GEN_F, n_rec);
    uint64_t *m4 = ctx->buf<uint64_t>(B_GEN_G, 4 * n_rec);
    if (!tno || !qno || !tnl || !qnl || !m4) return ctx->err_code;
    LAUNCH(ctx, k_maf_meta, grid_for(n_rec, 256), 256, n_rec, P.sl, P.rc.tl, P.rc.ql, tno, qno, tnl, qnl, m4, m4 + n_rec,
           m4 + 2 * n_rec, m4 + 3 * n_rec);
    M.t_name_off = tno; M.q_name_off = qno; M.t_name_len = tnl; M.q_name_len = qnl;
    M.t_size = m4; M.q_size = m4 + n_rec; M.t_start = m4 + 2 * n_rec; M.q_start = m4 + 3 * n_rec;
    M.neg = P.rc.neg; M.skip = P.rc.skip;
    unsigned long long *err_key = ctx->buf<unsigned long long>(B_ERR, 2);
    if (!err_key) return ctx->err_code;
    if (reuse) LAUNCH(ctx, k_init_u64, 1, 1, err_key, 1, kNoErr);
    int rc = no_throw([&] { return stat_reduce(ctx, n_rec, M, cnt11, err_key, acc); });
    if (rc) return rc;
    if (ctx->walk_launches) {
        float ms = 0.f;
        WGA_CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->walk_ms += ms;
    }
    ctx->last_counts_valid = true;
    return WGA_OK;
}

int wga_last_record_counts(wga_ctx *ctx, uint64_t *dst, uint64_t cap_records, uint64_t *n_records) {
    if (!ctx || !n_records) return WGA_E_INVALID_ARG;
    if (!ctx->last_counts_valid) {
        ctx->set_error(WGA_E_INVALID_ARG, 0, "no per-record counters available (run maf2paf/stat first)");
        return WGA_E_INVALID_ARG;
    }
    *n_records = ctx->n_rec;
    uint64_t n = std::min(cap_records, ctx->n_rec);
    if (n && dst) {
        WGA_CK(cudaMemcpyAsync(dst, ctx->bufs[B_CNT11].p, n * 88, cudaMemcpyDeviceToHost, ctx->stream));
        WGA_TRY(sync_check(ctx));
    }
    return WGA_OK;
}

// ---------------------------------------------------------------- call
int wga_vcf_header(wga_ctx *ctx, const char *sample, wga_span *out) {
    if (!ctx || !out) return WGA_E_INVALID_ARG;
    // caller.rs:304-337 build_header through noodles-vcf 0.43 (README.md:324-331 pins the text)
    static thread_local std::string hdr;
    hdr = "##fileformat=VCFv4.4\n"
          "##INFO=<ID=SVLEN,Number=A,Type=Integer,Description=\"Length of structural variant\">\n"
          "##INFO=<ID=SVTYPE,Number=1,Type=String,Description=\"Type of structural variant\">\n"
          "##INFO=<ID=END,Number=1,Type=Integer,Description=\"End position of the longest variant described in this record\">\n"
          "##INFO=<ID=INV_NEST,Number=1,Type=String,Description=\"Varations nested within inversion\">\n"
          "##FORMAT=<ID=QI,Number=1,Type=String,Description=\"Query informations\">\n"
          "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
          "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t";
    hdr += sample ? sample : "sample";
    hdr += "\n";
    out->data = hdr.data();
    out->len = hdr.size();
    return WGA_OK;
}

// noodles-vcf 0.43 contig name (Map<Contig> key, FromStr): the VCF 4.3 grammar of section 1.4.7 -- printable ASCII without
// \ , " ` ' ( ) [ ] { } < >, not starting with '*' or '='.  (third-party behaviour restated: parity unpinned)
static bool vcf_contig_name_ok(const char *s) {
    if (!*s || *s == '*' || *s == '=') return false;
    for (const unsigned char *p = (const unsigned char *)s; *p; ++p)
        if (*p < '!' || *p > '~' || strchr("\\,\"`'()[]{}<>", *p)) return false;
    return true;
}

int wga_vcf_header_contigs(wga_ctx *ctx, const char *sample, const char *const *names, const uint64_t *sizes, uint64_t n,
                           wga_span *out) {
    (void)ctx;  // host-side text only
    if (!out || (n && (!names || !sizes))) return WGA_E_INVALID_ARG;
    // caller.rs:339-357 add_header_contig: the index's isref entries, natord-sorted by name (stable), one
    // "##contig=<ID=name,length=size>" each; noodles writes them behind the FORMAT lines
    return no_throw([&]() -> int {
    std::vector<uint64_t> idx(n);
    for (uint64_t i = 0; i < n; ++i) {
        if (!names[i] || !vcf_contig_name_ok(names[i])) {
            if (ctx) ctx->set_error(WGA_E_INVALID_ARG, i, "contig name is not a valid VCF contig ID");
            return WGA_E_INVALID_ARG;
        }
        idx[i] = i;
    }
    std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return natord_cmp(names[a], names[b]) < 0; });
    static thread_local std::string hdr;
    hdr = "##fileformat=VCFv4.4\n"
          "##INFO=<ID=SVLEN,Number=A,Type=Integer,Description=\"Length of structural variant\">\n"
          "##INFO=<ID=SVTYPE,Number=1,Type=String,Description=\"Type of structural variant\">\n"
          "##INFO=<ID=END,Number=1,Type=Integer,Description=\"End position of the longest variant described in this record\">\n"
          "##INFO=<ID=INV_NEST,Number=1,Type=String,Description=\"Varations nested within inversion\">\n"
          "##FORMAT=<ID=QI,Number=1,Type=String,Description=\"Query informations\">\n"
          "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n";
    for (uint64_t i : idx) {
        hdr += "##contig=<ID=";
        hdr += names[i];
        hdr += ",length=";
        hdr += std::to_string(sizes[i]);
        hdr += ">\n";
    }
    hdr += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t";
    hdr += sample ? sample : "sample";
    hdr += "\n";
    out->data = hdr.data();
    out->len = hdr.size();
    return WGA_OK;
    });
}

int wga_maf_sline_names(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, wga_span *out) {
    if (!ctx || !out || (!text && len)) return WGA_E_INVALID_ARG;
    cudaSetDevice(ctx->device);
    out->data = nullptr;
    out->len = 0;
    walk_begin(ctx);
    WGA_TRY(stage_text(ctx, text, len, mem, flags));
    Prep P;
    WGA_TRY(prepare_maf(ctx, flags, nullptr, 1, &P));  // the records `call` builds (parse errors are call's errors)
    const uint64_t n_rec = ctx->n_rec, n_lines = ctx->n_lines;
    if (n_rec == 0) return sync_check(ctx);
    uint64_t *nlen = ctx->buf<uint64_t>(B_GEN_A, n_lines + 1), *noff = ctx->buf<uint64_t>(B_GEN_B, n_lines + 1);
    if (!nlen || !noff) return ctx->err_code;
    WGA_CK(cudaMemsetAsync(nlen, 0, (n_lines + 1) * 8, ctx->stream));
    LAUNCH(ctx, k_sline_name_len, grid_for(n_rec, 256), 256, n_rec, P.rc, P.sl, n_lines, nlen);
    WGA_TRY(exclusive_sum(ctx, nlen, noff, n_lines + 1));
    uint64_t total;
    WGA_TRY(read_scalars(ctx, {noff + n_lines}, &total));
    if (total == 0) return sync_check(ctx);
    char *d_all = (char *)ctx->ensure(B_OUT, total + 16);
    if (!d_all) return ctx->err_code;
    LAUNCH(ctx, k_sline_name_copy, grid_for(n_lines, 256), 256, n_lines, ctx->d_text, P.sl, (const uint64_t *)nlen,
           (const uint64_t *)noff, d_all);
    // distinct names in first-occurrence order: a few hundred among millions of s-lines, deduplicated on the host
    return no_throw([&]() -> int {
    std::vector<char> all(total);
    WGA_CK(cudaMemcpyAsync(all.data(), d_all, total, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_TRY(sync_check(ctx));
    static thread_local std::string names;
    names.clear();
    std::unordered_set<std::string_view> seen;
    for (uint64_t b = 0; b < total;) {
        const char *e = (const char *)memchr(all.data() + b, '\n', total - b);
        const uint64_t l = (uint64_t)(e - (all.data() + b));
        std::string_view nm(all.data() + b, l);
        if (seen.insert(nm).second) { names.append(nm); names.push_back('\n'); }
        b += l + 1;
    }
    out->data = names.data();
    out->len = names.size();
    return WGA_OK;
    });
}

int wga_maf_call(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, const wga_call_opts *opts,
                 wga_span *out) {
    if (!ctx || !out || !opts || (!text && len)) return WGA_E_INVALID_ARG;
    cudaSetDevice(ctx->device);
    out->data = nullptr;
    out->len = 0;
    if (opts->chunk_size == 0) {
        ctx->set_error(WGA_E_INVALID_ARG, 0, "chunk_size must be > 0 (the reference never terminates on 0)");
        return WGA_E_INVALID_ARG;
    }
    walk_begin(ctx);
    WGA_TRY(stage_text(ctx, text, len, mem, flags));
    TokCache &tc = tok_of(ctx);
    WalkCache &wc = tc.wc;
    Prep P;
    Runs RU;
    uint64_t *run_len = nullptr, *rec_run_first = nullptr;
    // the caller-classified run list of this slab may still be there (wga_maf2paf with WGA_F_SHARE_WALK): `call` then
    // starts at its chunk grid.  Only when it would have built exactly the same records (no single-row / ragged block).
    const bool reuse = (flags & WGA_F_REUSE_INDEX) && tc.valid && tc.d_text == ctx->d_text && tc.len == len && wc.recs && wc.runs &&
                       wc.call_same && wc.same_q(opts->query_name) && !(opts->query_match && !opts->query_name);
    if (reuse) {
        P = wc.P;
        RU = wc.RU;
        run_len = wc.run_len;
        rec_run_first = wc.rec_run_first;
        ctx->n_rec = wc.n_rec; ctx->n_items = wc.n_items; ctx->n_runs = RU.n_runs;
        P.err_key = ctx->buf<unsigned long long>(B_ERR, 2);
        if (!P.err_key) return ctx->err_code;
        LAUNCH(ctx, k_init_u64, 1, 1, P.err_key, 1, kNoErr);
    } else {
        WGA_TRY(prepare_maf(ctx, flags, opts->query_name, 1, &P, opts->query_match));
    }
    const uint64_t n_rec = ctx->n_rec, n_items = ctx->n_items;
    if (n_rec == 0 || n_items == 0) return sync_check(ctx);

    // dense pass: caller-classified run list
    if (!reuse) WGA_TRY(run_walk<MODE_RUNS_CALLER>(ctx, P, nullptr, &RU));
    const uint64_t n_runs = RU.n_runs;
    uint64_t *run_base = RU.run_base, *run_pos = RU.run_pos;
    uint32_t *run_rec = RU.run_rec;
    uint8_t *run_cls = RU.run_cls;
    if (!reuse) {
        run_len = ctx->buf<uint64_t>(B_RUN_LEN, n_runs + 1);
        rec_run_first = ctx->buf<uint64_t>(B_REC_RUNFIRST, n_rec + 1);
    }
    uint64_t *tcon = ctx->buf<uint64_t>(B_GEN_A, n_runs + 1), *qcon = ctx->buf<uint64_t>(B_GEN_B, n_runs + 1);
    uint64_t *St = ctx->buf<uint64_t>(B_GEN_C, n_runs + 1), *Sq = ctx->buf<uint64_t>(B_GEN_D, n_runs + 1);
    uint64_t *nchunks = ctx->buf<uint64_t>(B_GEN_E, n_rec + 1), *rec_chunk_first = ctx->buf<uint64_t>(B_GEN_F, n_rec + 1);
    if (!run_len || !rec_run_first || !tcon || !qcon || !St || !Sq || !nchunks || !rec_chunk_first) return ctx->err_code;
    if (!reuse) {
        LAUNCH(ctx, k_rec_run_first, grid_for(n_rec + 1, 256), 256, n_rec, P.item_base, run_base, n_items, n_runs, rec_run_first);
        LAUNCH(ctx, k_run_len, grid_for(n_runs + 1, 256), 256, n_runs, run_pos, run_rec, P.rc.ncols, run_len, (uint32_t *)nullptr);
    }
    LAUNCH(ctx, k_run_contrib, grid_for(n_runs + 1, 256), 256, n_runs, run_cls, run_len, tcon, qcon);
    WGA_TRY(exclusive_sum(ctx, tcon, St, n_runs + 1));
    WGA_TRY(exclusive_sum(ctx, qcon, Sq, n_runs + 1));
    // chunk grid: one pass into capacity-sized regions, then a dense copy
    uint64_t *ccap = ctx->buf<uint64_t>(B_PP_A, n_rec + 1), *cap_base = ctx->buf<uint64_t>(B_PP_B, n_rec + 1);
    if (!ccap || !cap_base) return ctx->err_code;
    LAUNCH(ctx, k_chunk_caps, grid_for(n_rec + 1, 256), 256, n_rec, (const uint64_t *)P.rc.ncols, (const uint8_t *)P.rc.skip,
           (uint64_t)opts->chunk_size, ccap);
    WGA_TRY(exclusive_sum(ctx, ccap, cap_base, n_rec + 1));
    const uint64_t cap_total = 2 * (P.total_cols / opts->chunk_size) + 3 * n_rec + 8;  // >= sum of the caps
    uint64_t *chunk_tmp = ctx->buf<uint64_t>(B_PP_C, cap_total);
    if (!chunk_tmp) return ctx->err_code;
    LAUNCH(ctx, k_chunks, grid_for(n_rec + 1, 128), 128, n_rec, rec_run_first, run_pos, run_cls, P.rc.ncols, P.rc.skip,
           opts->chunk_size, opts->svlen_cutoff, nchunks, (const uint64_t *)cap_base, chunk_tmp);
    WGA_TRY(exclusive_sum(ctx, nchunks, rec_chunk_first, n_rec + 1));
    uint64_t n_chunks;
    WGA_TRY(read_scalars(ctx, {rec_chunk_first + n_rec}, &n_chunks));
    uint64_t *chunk_cs = ctx->buf<uint64_t>(B_GEN_G, n_chunks + 1);
    uint32_t *chunk_rec = ctx->buf<uint32_t>(B_PT_FLAG, n_chunks + 1), *chunk_inv = ctx->buf<uint32_t>(B_PT_IDX, n_chunks + 1);
    uint64_t *inv_cum = ctx->buf<uint64_t>(B_INV_IDX, n_chunks + 1);
    uint64_t *rows = ctx->buf<uint64_t>(B_RUN_CIGOFF, n_runs + 1), *rowbase = ctx->buf<uint64_t>(B_NAME_OFF, n_runs + 1);
    if (!chunk_cs || !chunk_rec || !chunk_inv || !inv_cum || !rows || !rowbase) return ctx->err_code;
    LAUNCH(ctx, k_chunks_dense, grid_for(n_rec, 256), 256, n_rec, (const uint64_t *)cap_base, (const uint64_t *)chunk_tmp,
           (const uint64_t *)rec_chunk_first, chunk_cs, chunk_rec, P.err_key);
    CallView V;
    V.n_rec = n_rec; V.n_runs = n_runs; V.text = ctx->d_text; V.sl = P.sl; V.tl = P.rc.tl; V.ql = P.rc.ql;
    V.ncols = P.rc.ncols; V.t_off = P.rc.t_off; V.q_off = P.rc.q_off; V.neg = P.rc.neg; V.skip = P.rc.skip;
    V.rec_run_first = rec_run_first; V.run_pos = run_pos; V.run_len = run_len; V.run_rec = run_rec; V.run_cls = run_cls;
    V.St = St; V.Sq = Sq; V.chunk_size = opts->chunk_size; V.svlen = opts->svlen_cutoff; V.snp = opts->snp; V.inv = opts->inv;
    V.rec_chunk_first = rec_chunk_first; V.chunk_cs = chunk_cs; V.chunk_inv = chunk_inv; V.inv_cum = inv_cum;
    LAUNCH(ctx, k_chunk_inv, grid_for(n_chunks + 1, 256), 256, V, n_chunks, chunk_rec, chunk_inv);
    WGA_TRY(exclusive_sum(ctx, chunk_inv, inv_cum, n_chunks + 1));
    LAUNCH(ctx, k_run_rows, grid_for(n_runs + 1, 256), 256, V, rows);
    WGA_TRY(exclusive_sum(ctx, rows, rowbase, n_runs + 1));
    uint64_t n_rows;
    WGA_TRY(read_scalars(ctx, {rowbase + n_runs}, &n_rows));
    if (n_rows == 0) return sync_check(ctx);
    uint64_t *row_len = ctx->buf<uint64_t>(B_ROW_LEN, n_rows + 1), *row_off = ctx->buf<uint64_t>(B_ROW_OFF, n_rows + 1);
    if (!row_len || !row_off) return ctx->err_code;
    // row -> run map (both row passes start from it instead of searching rowbase)
    uint32_t *row_run = nullptr;
    if (n_runs < 0xFFFFFFFFull) {
        row_run = ctx->buf<uint32_t>(B_ROW_RUN, n_rows + 1);
        if (!row_run) return ctx->err_code;
        WGA_CK(cudaMemsetAsync(row_run, 0, (n_rows + 1) * 4, ctx->stream));
        LAUNCH(ctx, k_row_run_seed, grid_for(n_runs, 256), 256, n_runs, (const uint64_t *)rows, (const uint64_t *)rowbase, row_run);
        WGA_TRY(inclusive_max_u32(ctx, row_run, n_rows));
    }
    LAUNCH(ctx, k_vcf_rows<0>, grid_for(n_rows + 1, 128), 128, V, n_rows, rowbase, (const uint32_t *)row_run, row_len,
           (const uint64_t *)nullptr, (char *)nullptr, P.err_key);
    WGA_TRY(exclusive_sum(ctx, row_len, row_off, n_rows + 1));
    uint64_t sc[2];
    WGA_TRY(read_scalars(ctx, {row_off + n_rows, P.err_key}, sc));
    if (sc[1] != kNoErr) {
        ctx->set_error(WGA_E_VCF_BASE, sc[1] >> 8, "REF/ALT base outside ACGTNacgtn (noodles-vcf rejects the record)");
        return WGA_E_VCF_BASE;
    }
    char *d_out = (char *)ctx->ensure(B_OUT, sc[0] + 16);
    if (!d_out) return ctx->err_code;
    LAUNCH(ctx, k_vcf_rows<1>, grid_for(n_rows, 128), 128, V, n_rows, rowbase, (const uint32_t *)row_run, row_len, row_off, d_out,
           P.err_key);
    uint64_t key1;  // REF/ALT bases are validated while they are written
    WGA_TRY(read_scalars(ctx, {P.err_key}, &key1));
    if (key1 != kNoErr) {
        ctx->set_error(WGA_E_VCF_BASE, key1 >> 8, "REF/ALT base outside ACGTNacgtn (noodles-vcf rejects the record)");
        return WGA_E_VCF_BASE;
    }
    return finish_output(ctx, d_out, sc[0], mem, flags, out);
}

}  // extern "C"
namespace {
// host restatement of the tokenizer's error wording (cigar.rs:43-75, errors.rs) for one CIGAR string
std::string paf_cigar_error_text(const std::string &cg, int code) {
    size_t i = 0;
    while (i < cg.size()) {
        size_t d0 = i;
        while (i < cg.size() && cg[i] >= '0' && cg[i] <= '9') ++i;
        size_t o0 = i;
        while (i < cg.size() && !(cg[i] >= '0' && cg[i] <= '9')) ++i;
        std::string digits = cg.substr(d0, o0 - d0), op = cg.substr(o0, i - o0);
        if (op.size() != 1) return "CIGAR OP `" + op + "` invalid";
        uint64_t v;
        if (!parse_u64_bytes((const unsigned char *)digits.data(), digits.size(), &v)) return "Parse `" + digits + "` Into Integer Error";
        if (std::string("M=XID").find(op[0]) == std::string::npos) return "CIGAR OP `" + op + "` invalid";
    }
    return "CIGAR error code " + std::to_string(code);
}
}  // namespace
// Quoted csv fields: the front end refuses a field that starts with '"' (k_paf_parse).  The PAF entry points then
// normalise the staged text (k_unquote_lines: quoting that was not needed is stripped; anything else stays refused) and
// run once more on the normalised copy.  Inputs without quotes never get here.
static int paf_unquote(wga_ctx *ctx, const uint8_t **d_out, uint64_t *len_out) {
    const uint8_t *d_text = ctx->d_text;
    const uint64_t len = ctx->text_len;
    unsigned long long *err_key = ctx->buf<unsigned long long>(B_ERR, 1);
    if (!err_key) return ctx->err_code;
    LAUNCH(ctx, k_init_u64, 1, 1, err_key, 1, kNoErr);
    uint64_t n_ws = 0, n_nl = 0;
    uint64_t *ws_ent = nullptr, *nl_idx = nullptr;
    WGA_TRY((build_delim_index<1, kTokCapPaf>(ctx, d_text, len, &n_ws, &n_nl, &ws_ent, &nl_idx)));
    const uint8_t last_byte = *(uint8_t *)(ctx->h_scal + 8);
    const uint64_t n_lines = n_nl + (last_byte != '\n' && last_byte != '\r' ? 1 : 0);
    LineView lv;
    lv.text = d_text; lv.len = len; lv.ws_ent = ws_ent; lv.nl_ws_idx = nl_idx;
    lv.n_ws = n_ws; lv.n_nl = n_nl; lv.n_lines = n_lines; lv.first_slab = 0;
    uint8_t *lflags = ctx->buf<uint8_t>(B_LINE_FLAGS, n_lines + 1);
    uint32_t *isrec = ctx->buf<uint32_t>(B_LINE_HEAD, n_lines + 1);
    uint64_t *nlen = ctx->buf<uint64_t>(B_GEN_A, n_lines + 1), *noff = ctx->buf<uint64_t>(B_GEN_B, n_lines + 1);
    if (!lflags || !isrec || !nlen || !noff) return ctx->err_code;
    LAUNCH(ctx, k_paf_line_flags, grid_for(n_lines + 1, 256), 256, lv, lflags, isrec);
    LAUNCH(ctx, k_unquote_lines<0>, grid_for(n_lines + 1, 256), 256, lv, (const uint8_t *)lflags, nlen, (const uint64_t *)nullptr, (uint8_t *)nullptr,
           err_key);
    WGA_TRY(exclusive_sum(ctx, nlen, noff, n_lines + 1));
    uint64_t sc[2];
    WGA_TRY(read_scalars(ctx, {noff + n_lines, err_key}, sc));
    if (sc[1] != kNoErr) {
        ctx->set_error(WGA_E_UNSUPPORTED, sc[1] >> 32,
                       "unsupported PAF construct: a quoted csv field that holds a quote, a tab or a line break, or text behind its closing "
                       "quote (line " + std::to_string(sc[1] >> 32) + ")");
        return WGA_E_UNSUPPORTED;
    }
    uint8_t *d_new = ctx->buf<uint8_t>(B_UNQ_TEXT, sc[0] + 64);
    if (!d_new) return ctx->err_code;
    LAUNCH(ctx, k_unquote_lines<1>, grid_for(n_lines + 1, 256), 256, lv, (const uint8_t *)lflags, nlen, (const uint64_t *)noff, d_new, err_key);
    WGA_CK(cudaMemsetAsync(d_new + sc[0], 0, 64, ctx->stream));
    *d_out = d_new;
    *len_out = sc[0];
    return WGA_OK;
}
// rc of a PAF command -> rc, after one retry on the normalised text when the refusal was a quoted field
template <class F>
static int paf_retry_unquoted(wga_ctx *ctx, int rc, int mem, uint32_t flags, F &&again) {
    if (rc != WGA_E_UNSUPPORTED || !ctx || ctx->err_msg.find("(quoted csv field)") == std::string::npos) return rc;
    const uint8_t *d2 = nullptr;
    uint64_t n2 = 0;
    WGA_TRY(paf_unquote(ctx, &d2, &n2));
    return again((const void *)d2, n2, (int)WGA_MEM_DEVICE, flags | (mem == WGA_MEM_HOST ? (uint32_t)WGA_F_OUT_HOST : 0u));
}

// CIGAR text -> 11 counters per record: the flattened piece stream (paf.cuh k_paf_cigar_flat) and, for the records it
// flags (anything that is not `digits{1,8} one-of(M=XID)`, i.e. every error input), the exact one-thread kernel.
// WGA_PAF_CIGAR_WARP=1 selects the older warp-per-record kernel.
static int paf_cigar_counts(wga_ctx *ctx, uint64_t n_rec, uint64_t text_len, const uint8_t *d_text, const PafRecs &R,
                            unsigned long long *cnt11, uint8_t *slow, unsigned long long *err_key) {
    static const bool per_rec = getenv("WGA_PAF_CIGAR_WARP") != nullptr && getenv("WGA_PAF_CIGAR_WARP")[0] == '1';
    if (per_rec) {
        LAUNCH(ctx, k_paf_cigar_counts_warp, grid_for(n_rec * 32, 128), 128, n_rec, d_text, R, cnt11, slow);
    } else {
        // pieces <= (CIGAR bytes + 30 per record) / 16: the grid is sized for that bound, chunks past the real total exit
        const uint64_t max_pieces = text_len / 16 + 3 * n_rec + 1, n_chunks = (max_pieces + kCigChunk - 1) / kCigChunk;
        uint64_t *nvec = ctx->buf<uint64_t>(B_CIG_NVEC, n_rec + 1), *pbase = ctx->buf<uint64_t>(B_CIG_PBASE, n_rec + 1);
        uint32_t *chunk_rec = ctx->buf<uint32_t>(B_CIG_CHUNK, n_chunks + 1);
        unsigned long long *acc6 = ctx->buf<unsigned long long>(B_CIG_ACC, n_rec * 6);
        if (!nvec || !pbase || !chunk_rec || !acc6) return ctx->err_code;
        WGA_CK(cudaMemsetAsync(acc6, 0, n_rec * 48, ctx->stream));
        LAUNCH(ctx, k_cig_pieces, grid_for(n_rec + 1, 256), 256, n_rec, d_text, R, nvec, slow);
        WGA_TRY(exclusive_sum(ctx, nvec, pbase, n_rec + 1));
        LAUNCH(ctx, k_cig_chunk_rec, grid_for(n_rec, 256), 256, n_rec, (const uint64_t *)pbase, chunk_rec);
        LAUNCH(ctx, k_paf_cigar_flat, grid_for(n_chunks * 32, 128), 128, n_rec, d_text, R, (const uint64_t *)pbase,
               (const uint32_t *)chunk_rec, acc6, slow);
        LAUNCH(ctx, k_cig_finish, grid_for(n_rec, 256), 256, n_rec, R, (const unsigned long long *)acc6, (const uint8_t *)slow, cnt11);
    }
    LAUNCH(ctx, k_paf_cigar_counts, grid_for(n_rec, 128), 128, n_rec, d_text, R, cnt11, err_key, (const uint8_t *)slow);
    return WGA_OK;
}

// cs:Z: tags standing in for a missing cg:Z: (paf.rs:129-135): k_paf_parse summed their lengths into *cs_total;
// when there are any, k_cs_convert writes the converted CIGARs to an arena and repoints cig_off / cig_len.
static int paf_cs_to_cigar(wga_ctx *ctx, const uint8_t *d_text, const PafRecs &R, uint64_t n_rec, unsigned long long *cs_total,
                           unsigned long long *tag_key) {
    uint64_t total;
    WGA_TRY(read_scalars(ctx, {(const uint64_t *)cs_total}, &total));
    if (total == 0) return WGA_OK;
    uint8_t *arena = ctx->buf<uint8_t>(B_CS_ARENA, total + 16);
    if (!arena) return ctx->err_code;
    LAUNCH(ctx, k_init_u64, 1, 1, cs_total, 1, 0ull);
    LAUNCH(ctx, k_cs_convert, grid_for(n_rec, 128), 128, n_rec, d_text, R, arena, cs_total, tag_key);
    return WGA_OK;
}

extern "C" {

// ---------------------------------------------------------------- stat -f paf
static int wga_paf_stat_impl(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, wga_stat_acc *acc) {
    (void)flags;
    if (!ctx || !acc || (!text && len)) return WGA_E_INVALID_ARG;
    cudaSetDevice(ctx->device);
    walk_begin(ctx);
    WGA_TRY(stage_text(ctx, text, len, mem));
    ctx->n_rec = 0;
    if (len == 0) return WGA_OK;
    const uint8_t *d_text = ctx->d_text;
    unsigned long long *err_key = ctx->buf<unsigned long long>(B_ERR, 1);
    if (!err_key) return ctx->err_code;
    LAUNCH(ctx, k_init_u64, 1, 1, err_key, 1, kNoErr);
    uint64_t n_ws = 0, n_nl = 0;
    uint64_t *ws_ent = nullptr, *nl_idx = nullptr;
    WGA_TRY((build_delim_index<1, kTokCapPaf>(ctx, d_text, len, &n_ws, &n_nl, &ws_ent, &nl_idx)));
    const uint8_t last_byte = *(uint8_t *)(ctx->h_scal + 8);
    const uint64_t n_lines = n_nl + (last_byte != '\n' && last_byte != '\r' ? 1 : 0);  // PAF lines end at either
    LineView lv;
    lv.text = d_text; lv.len = len; lv.ws_ent = ws_ent; lv.nl_ws_idx = nl_idx;
    lv.n_ws = n_ws; lv.n_nl = n_nl; lv.n_lines = n_lines; lv.first_slab = 0;
    uint8_t *lflags = ctx->buf<uint8_t>(B_LINE_FLAGS, n_lines + 1);
    uint32_t *isrec = ctx->buf<uint32_t>(B_LINE_HEAD, n_lines + 1), *rscan = ctx->buf<uint32_t>(B_LINE_REC, n_lines + 1);
    if (!lflags || !isrec || !rscan) return ctx->err_code;
    LAUNCH(ctx, k_paf_line_flags, grid_for(n_lines + 1, 256), 256, lv, lflags, isrec);
    WGA_TRY(exclusive_sum(ctx, isrec, rscan, n_lines + 1));
    WGA_CK(cudaMemcpyAsync(ctx->h_scal, rscan + n_lines, 4, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_TRY(sync_check(ctx));
    const uint64_t n_rec = *(uint32_t *)ctx->h_scal;
    ctx->n_rec = n_rec;
    if (n_rec == 0) return WGA_OK;
    PafRecs R;
    R.q_name_off = ctx->buf<uint64_t>(B_PAF_B, n_rec); R.t_name_off = ctx->buf<uint64_t>(B_PAF_C, n_rec);
    R.cig_off = ctx->buf<uint64_t>(B_PAF_D, n_rec); R.cig_len = ctx->buf<uint64_t>(B_PAF_E, n_rec);
    R.q_name_len = ctx->buf<uint32_t>(B_PAF_F, n_rec); R.t_name_len = ctx->buf<uint32_t>(B_PAF_G, n_rec);
    uint64_t *m4 = ctx->buf<uint64_t>(B_PAF_H, 4 * n_rec);
    R.neg = ctx->buf<uint8_t>(B_PAF_I, n_rec);
    unsigned long long *cnt11 = ctx->buf<unsigned long long>(B_CNT11, n_rec * 11);
    if (!R.q_name_off || !R.t_name_off || !R.cig_off || !R.cig_len || !R.q_name_len || !R.t_name_len || !m4 || !R.neg || !cnt11)
        return ctx->err_code;
    R.q_size = m4; R.t_size = m4 + n_rec; R.q_start = m4 + 2 * n_rec; R.t_start = m4 + 3 * n_rec;
    unsigned long long *cs_total = ctx->buf<unsigned long long>(B_CS_TOTAL, 1);
    if (!cs_total) return ctx->err_code;
    LAUNCH(ctx, k_init_u64, 1, 1, cs_total, 1, 0ull);
    LAUNCH(ctx, k_paf_parse, grid_for(n_lines, 256), 256, lv, lflags, rscan, R, err_key, err_key, cs_total);
    WGA_TRY(paf_cs_to_cigar(ctx, d_text, R, n_rec, cs_total, err_key));
    WGA_CK(cudaEventRecord(ctx->ev0, ctx->stream));
    uint8_t *slow = ctx->buf<uint8_t>(B_RUN_CLS, n_rec);
    if (!slow) return ctx->err_code;
    WGA_TRY(paf_cigar_counts(ctx, n_rec, len, d_text, R, cnt11, slow, err_key));
    WGA_CK(cudaEventRecord(ctx->ev1, ctx->stream));
    uint64_t key;
    WGA_TRY(read_scalars(ctx, {err_key}, &key));
    {
        float ms = 0.f;
        WGA_CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->walk_ms = ms; ctx->walk_launches = 1; ctx->walk_bytes = len;
    }
    if (key != kNoErr) {
        const int code = (int)(key & 0xFF);
        const uint64_t rec = key >> 32;
        std::string msg;
        if (code == WGA_E_CSV) msg = "CSV deserialize error by: record " + std::to_string(rec) + " does not match the 12 typed PAF columns";
        else if (code == WGA_E_CIGAR_TAG_NOT_FOUND) msg = "CIGAR start tag not found";
        else if (code == WGA_E_UNSUPPORTED) msg = "unsupported PAF construct (quoted csv field) in record " + std::to_string(rec);
        else if (code == WGA_E_REF_PANIC) msg = "the reference panics on this record (empty CIGAR, or a cs:Z: length beyond u64)";
        else {
            uint64_t oc[2];
            WGA_TRY(read_scalars(ctx, {R.cig_off + rec, R.cig_len + rec}, oc));
            std::string cg((size_t)std::min<uint64_t>(oc[1], 1u << 20), '\0');
            if (!cg.empty()) WGA_CK(cudaMemcpy(&cg[0], d_text + oc[0], cg.size(), cudaMemcpyDeviceToHost));
            msg = paf_cigar_error_text(cg, code);
        }
        ctx->set_error(code, rec, msg);
        return code;
    }
    RecMeta M;
    M.text = d_text;
    M.t_name_off = R.t_name_off; M.q_name_off = R.q_name_off; M.t_name_len = R.t_name_len; M.q_name_len = R.q_name_len;
    M.t_size = R.t_size; M.q_size = R.q_size; M.t_start = R.t_start; M.q_start = R.q_start;
    M.neg = R.neg; M.skip = nullptr;
    int rc = no_throw([&] { return stat_reduce(ctx, n_rec, M, cnt11, err_key, acc); });
    if (rc) return rc;
    ctx->last_counts_valid = true;
    return WGA_OK;
}
int wga_paf_stat(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, wga_stat_acc *acc) {
    const int rc = wga_paf_stat_impl(ctx, text, len, mem, flags, acc);
    return paf_retry_unquoted(ctx, rc, mem, flags, [&](const void *t2, uint64_t n2, int m2, uint32_t f2) {
        return wga_paf_stat_impl(ctx, t2, n2, m2, f2, acc);
    });
}

// ---------------------------------------------------------------- paf2maf
int wga_fasta_load(wga_ctx *ctx, const void *fa, uint64_t fa_len, const void *fai, uint64_t fai_len, wga_fasta **out) {
    if (!ctx || !out || !fa || !fai) return WGA_E_INVALID_ARG;
    cudaSetDevice(ctx->device);
    *out = nullptr;
    // .fai: name \t length \t offset \t linebases \t linewidth
    struct Ent { std::string name; uint64_t len, off, lb, lw, hash; };
    std::vector<Ent> ents;
    const char *p = (const char *)fai, *end = p + fai_len;
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', end - p);
        const char *le = nl ? nl : end;
        std::vector<std::string> f;
        const char *b = p;
        for (const char *c = p; c <= le; ++c)
            if (c == le || *c == '\t') { f.emplace_back(b, c - b); b = c + 1; }
        if (f.size() >= 5) {
            Ent e;
            e.name = f[0];
            bool ok = parse_u64_bytes((const unsigned char *)f[1].data(), f[1].size(), &e.len) &&
                      parse_u64_bytes((const unsigned char *)f[2].data(), f[2].size(), &e.off) &&
                      parse_u64_bytes((const unsigned char *)f[3].data(), f[3].size(), &e.lb) &&
                      parse_u64_bytes((const unsigned char *)f[4].data(), f[4].size(), &e.lw);
            if (!ok) { ctx->set_error(WGA_E_INVALID_ARG, ents.size(), "malformed .fai line"); return WGA_E_INVALID_ARG; }
            uint64_t h = 0xCBF29CE484222325ull;
            for (unsigned char ch : e.name) h = (h ^ ch) * 0x100000001B3ull;
            e.hash = h;
            ents.push_back(std::move(e));
        }
        p = le + 1;
    }
    std::stable_sort(ents.begin(), ents.end(), [](const Ent &a, const Ent &b) { return a.hash < b.hash; });
    const uint64_t n = ents.size();
    std::vector<uint64_t> meta(n * kFaMeta + 1), unwrap(n * 5 + 1);
    std::string names;
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; ++i) {
        meta[i * kFaMeta] = ents[i].hash; meta[i * kFaMeta + 1] = names.size(); meta[i * kFaMeta + 2] = ents[i].name.size();
        meta[i * kFaMeta + 3] = total; meta[i * kFaMeta + 4] = ents[i].len;
        unwrap[i * 5] = ents[i].off; unwrap[i * 5 + 1] = ents[i].lb; unwrap[i * 5 + 2] = ents[i].lw; unwrap[i * 5 + 3] = total;
        unwrap[i * 5 + 4] = ents[i].len;
        names += ents[i].name;
        total += ents[i].len;
    }
    wga_fasta *F = new wga_fasta();
    F->device = ctx->device; F->n_seq = n; F->total = total;
    uint8_t *d_fa = nullptr;
    uint64_t *d_unwrap = nullptr;
    auto fail = [&](const char *msg) {
        if (d_fa) cudaFree(d_fa);
        if (d_unwrap) cudaFree(d_unwrap);
        wga_fasta_free(F);
        ctx->set_error(WGA_E_NOMEM, 0, msg);
        return WGA_E_NOMEM;
    };
    if (cudaMalloc(&F->d_seq, total + 64) != cudaSuccess || cudaMalloc(&F->d_names, names.size() + 16) != cudaSuccess ||
        cudaMalloc(&F->d_meta, (n * kFaMeta + 1) * 8) != cudaSuccess || cudaMalloc(&d_fa, fa_len + 64) != cudaSuccess ||
        cudaMalloc(&d_unwrap, (n * 5 + 1) * 8) != cudaSuccess)
        return fail("cudaMalloc failed while loading a FASTA");
    WGA_CK(cudaMemcpyAsync(d_fa, fa, fa_len, cudaMemcpyHostToDevice, ctx->stream));
    WGA_CK(cudaMemcpyAsync(F->d_names, names.data(), names.size(), cudaMemcpyHostToDevice, ctx->stream));
    WGA_CK(cudaMemcpyAsync(F->d_meta, meta.data(), n * kFaMeta * 8, cudaMemcpyHostToDevice, ctx->stream));
    WGA_CK(cudaMemcpyAsync(d_unwrap, unwrap.data(), n * 5 * 8, cudaMemcpyHostToDevice, ctx->stream));
    for (uint64_t s0 = 0; s0 < n; s0 += 32768) {
        const uint64_t cnt = std::min<uint64_t>(32768, n - s0);
        dim3 grid(64, (unsigned)cnt);
        k_fasta_unwrap<<<grid, 256, 0, ctx->stream>>>(cnt, d_fa, fa_len, d_unwrap + s0 * 5, F->d_seq);
        ctx->n_own++;
    }
    int rc = sync_check(ctx);
    cudaFree(d_fa);
    cudaFree(d_unwrap);
    if (rc) { wga_fasta_free(F); return rc; }
    *out = F;
    return WGA_OK;
}

void wga_fasta_free(wga_fasta *F) {
    if (!F) return;
    cudaSetDevice(F->device);
    if (F->d_seq) cudaFree(F->d_seq);
    if (F->d_names) cudaFree(F->d_names);
    if (F->d_meta) cudaFree(F->d_meta);
    delete F;
}

int wga_paf2maf_header(wga_ctx *ctx, const char *t_fa_path, const char *q_fa_path, wga_span *out) {
    if (!ctx || !out || !t_fa_path || !q_fa_path) return WGA_E_INVALID_ARG;
    static thread_local std::string hdr;
    hdr = std::string("#maf version=1.6 convert_from=paf t_seq_path=") + t_fa_path + " q_seq_path=" + q_fa_path + "\n";
    out->data = hdr.data();
    out->len = hdr.size();
    return WGA_OK;
}

// PAF text -> typed records + FASTA spans (shared by paf2maf and call -f paf)
struct PafPrep {
    LineView lv;
    PafRecs R;
    P2mRecs X;
    uint64_t n_rec = 0;
    unsigned long long *err_key = nullptr;
    uint64_t *n_ops = nullptr, *op_base = nullptr;
};
static int paf_fasta_prepare(wga_ctx *ctx, uint64_t len, const wga_fasta *TF, const wga_fasta *QF, uint64_t end_adj, PafPrep *PP,
                             bool with_spans = true, unsigned long long *tag_key = nullptr) {
    ctx->n_rec = 0;
    if (len == 0) return WGA_OK;
    const uint8_t *d_text = ctx->d_text;
    unsigned long long *err_key = ctx->buf<unsigned long long>(B_ERR, 1);
    if (!err_key) return ctx->err_code;
    LAUNCH(ctx, k_init_u64, 1, 1, err_key, 1, kNoErr);
    uint64_t n_ws = 0, n_nl = 0;
    uint64_t *ws_ent = nullptr, *nl_idx = nullptr;
    WGA_TRY((build_delim_index<1, kTokCapPaf>(ctx, d_text, len, &n_ws, &n_nl, &ws_ent, &nl_idx)));
    const uint8_t last_byte = *(uint8_t *)(ctx->h_scal + 8);
    const uint64_t n_lines = n_nl + (last_byte != '\n' && last_byte != '\r' ? 1 : 0);  // PAF lines end at either
    LineView lv;
    lv.text = d_text; lv.len = len; lv.ws_ent = ws_ent; lv.nl_ws_idx = nl_idx;
    lv.n_ws = n_ws; lv.n_nl = n_nl; lv.n_lines = n_lines; lv.first_slab = 0;
    uint8_t *lflags = ctx->buf<uint8_t>(B_LINE_FLAGS, n_lines + 1);
    uint32_t *isrec = ctx->buf<uint32_t>(B_LINE_HEAD, n_lines + 1), *rscan = ctx->buf<uint32_t>(B_LINE_REC, n_lines + 1);
    if (!lflags || !isrec || !rscan) return ctx->err_code;
    LAUNCH(ctx, k_paf_line_flags, grid_for(n_lines + 1, 256), 256, lv, lflags, isrec);
    WGA_TRY(exclusive_sum(ctx, isrec, rscan, n_lines + 1));
    WGA_CK(cudaMemcpyAsync(ctx->h_scal, rscan + n_lines, 4, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_TRY(sync_check(ctx));
    const uint64_t n_rec = *(uint32_t *)ctx->h_scal;
    ctx->n_rec = n_rec;
    if (n_rec == 0) return WGA_OK;
    PafRecs R;
    R.q_name_off = ctx->buf<uint64_t>(B_PAF_B, n_rec); R.t_name_off = ctx->buf<uint64_t>(B_PAF_C, n_rec);
    R.cig_off = ctx->buf<uint64_t>(B_PAF_D, n_rec); R.cig_len = ctx->buf<uint64_t>(B_PAF_E, n_rec);
    R.q_name_len = ctx->buf<uint32_t>(B_PAF_F, n_rec); R.t_name_len = ctx->buf<uint32_t>(B_PAF_G, n_rec);
    uint64_t *m4 = ctx->buf<uint64_t>(B_PAF_H, 4 * n_rec);
    R.neg = ctx->buf<uint8_t>(B_PAF_I, n_rec);
    uint64_t *x7 = ctx->buf<uint64_t>(B_PAF_J, 7 * n_rec);
    uint64_t *n_ops = ctx->buf<uint64_t>(B_GEN_A, n_rec + 1), *op_base = ctx->buf<uint64_t>(B_GEN_B, n_rec + 1);
    if (!R.q_name_off || !R.t_name_off || !R.cig_off || !R.cig_len || !R.q_name_len || !R.t_name_len || !m4 || !R.neg || !x7 ||
        !n_ops || !op_base)
        return ctx->err_code;
    R.q_size = m4; R.t_size = m4 + n_rec; R.q_start = m4 + 2 * n_rec; R.t_start = m4 + 3 * n_rec;
    P2mRecs X;
    X.q_end = x7; X.t_end = x7 + n_rec; X.mapq = x7 + 2 * n_rec; X.t_src = x7 + 3 * n_rec; X.q_src = x7 + 4 * n_rec;
    X.LT = x7 + 5 * n_rec; X.LQ = x7 + 6 * n_rec; X.n_ops = n_ops;
    FaView T{nullptr, nullptr, nullptr, 0}, Q{nullptr, nullptr, nullptr, 0};
    if (TF) T = FaView{TF->d_seq, TF->d_names, TF->d_meta, TF->n_seq};
    if (QF) Q = FaView{QF->d_seq, QF->d_names, QF->d_meta, QF->n_seq};
    unsigned long long *cs_total = ctx->buf<unsigned long long>(B_CS_TOTAL, 1);
    if (!cs_total) return ctx->err_code;
    LAUNCH(ctx, k_init_u64, 1, 1, cs_total, 1, 0ull);
    LAUNCH(ctx, k_paf_parse, grid_for(n_lines, 256), 256, lv, lflags, rscan, R, err_key, tag_key ? tag_key : err_key, cs_total);
    WGA_TRY(paf_cs_to_cigar(ctx, d_text, R, n_rec, cs_total, tag_key ? tag_key : err_key));
    LAUNCH(ctx, k_p2m_spans, grid_for(n_lines, 256), 256, lv, lflags, rscan, R, X, T, Q, end_adj, with_spans ? 1 : 0, err_key);
    PP->lv = lv; PP->R = R; PP->X = X; PP->n_rec = n_rec; PP->err_key = err_key; PP->n_ops = n_ops; PP->op_base = op_base;
    return WGA_OK;
}

}  // extern "C"
// row expansion of paf2maf / chain2maf: by runs between gap ops (k_p2m_runs); WGA_P2M_OPS=1 selects the per-op kernel
static int p2m_expand_rows(wga_ctx *ctx, const P2mView &V, uint64_t n_opsT) {
    static const bool per_op = getenv("WGA_P2M_OPS") != nullptr && getenv("WGA_P2M_OPS")[0] == '1';
    const uint64_t n_rec = V.n_rec;
    if (per_op) {
        LAUNCH(ctx, k_p2m_rows, grid_for((n_opsT + n_rec) * 32, 256), 256, V);
        return WGA_OK;
    }
    uint32_t *fI = ctx->buf<uint32_t>(B_XRUN_FLAG, n_opsT + 1), *fD = ctx->buf<uint32_t>(B_PP_O, n_opsT + 1);
    uint64_t *idxI = ctx->buf<uint64_t>(B_XRUN_IDX, n_opsT + 1), *idxD = ctx->buf<uint64_t>(B_PP_P, n_opsT + 1);
    if (!fI || !fD || !idxI || !idxD) return ctx->err_code;
    LAUNCH(ctx, k_p2m_gap_flags, grid_for(n_opsT + 1, 256), 256, n_opsT, V.op_type, fI, fD);
    WGA_TRY(exclusive_sum(ctx, fI, idxI, n_opsT + 1));
    WGA_TRY(exclusive_sum(ctx, fD, idxD, n_opsT + 1));
    uint64_t cnt[2];
    WGA_TRY(read_scalars(ctx, {idxI + n_opsT, idxD + n_opsT}, cnt));
    uint64_t *Ilist = ctx->buf<uint64_t>(B_XRUN_POS, cnt[0] + 1), *Dlist = ctx->buf<uint64_t>(B_PP_Q, cnt[1] + 1);
    if (!Ilist || !Dlist) return ctx->err_code;
    if (n_opsT)
        LAUNCH(ctx, k_p2m_gap_lists, grid_for(n_opsT, 256), 256, n_opsT, V.op_type, (const uint64_t *)idxI, (const uint64_t *)idxD, Ilist, Dlist);
    LAUNCH(ctx, k_p2m_runs, grid_for((cnt[0] + cnt[1] + 2 * n_rec) * 32, 256), 256, V, cnt[0], cnt[1], (const uint64_t *)idxI,
           (const uint64_t *)Ilist, (const uint64_t *)idxD, (const uint64_t *)Dlist);
    return WGA_OK;
}

extern "C" {
static int wga_paf2maf_impl(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, const wga_fasta *TF, const wga_fasta *QF,
                wga_span *out) {
    (void)flags;
    if (!ctx || !out || !TF || !QF || (!text && len)) return WGA_E_INVALID_ARG;
    cudaSetDevice(ctx->device);
    out->data = nullptr;
    out->len = 0;
    walk_begin(ctx);
    WGA_TRY(stage_text(ctx, text, len, mem));
    PafPrep PP;
    WGA_TRY(paf_fasta_prepare(ctx, len, TF, QF, 1, &PP));
    const uint64_t n_rec = PP.n_rec;
    if (n_rec == 0) return WGA_OK;
    const uint8_t *d_text = ctx->d_text;
    PafRecs R = PP.R;
    P2mRecs X = PP.X;
    unsigned long long *err_key = PP.err_key;
    uint64_t *n_ops = PP.n_ops, *op_base = PP.op_base;
    LAUNCH(ctx, k_p2m_revcomp_check, (unsigned)n_rec, 128, n_rec, R, X, QF->d_seq, err_key);
    LAUNCH(ctx, k_p2m_ops<0>, grid_for(n_rec + 1, 128), 128, n_rec, d_text, R, n_ops, (const uint64_t *)nullptr, (uint64_t *)nullptr,
           (uint8_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, err_key);
    WGA_TRY(exclusive_sum(ctx, n_ops, op_base, n_rec + 1));
    uint64_t sc[2];
    WGA_TRY(read_scalars(ctx, {op_base + n_rec, err_key}, sc));
    const uint64_t n_opsT = sc[0];
    uint64_t key = sc[1];
    uint64_t *op_len = ctx->buf<uint64_t>(B_RUN_POS, n_opsT + 1), *c_all = ctx->buf<uint64_t>(B_GEN_C, n_opsT + 1);
    uint64_t *c_i = ctx->buf<uint64_t>(B_GEN_D, n_opsT + 1), *c_d = ctx->buf<uint64_t>(B_GEN_E, n_opsT + 1);
    uint64_t *S_all = ctx->buf<uint64_t>(B_GEN_F, n_opsT + 1), *S_i = ctx->buf<uint64_t>(B_GEN_G, n_opsT + 1);
    uint64_t *S_d = ctx->buf<uint64_t>(B_GEN_H, n_opsT + 1);
    uint8_t *op_type = ctx->buf<uint8_t>(B_RUN_CLS, n_opsT + 1);
    uint32_t *op_rec = ctx->buf<uint32_t>(B_RUN_REC, n_opsT + 1), *op_pos = ctx->buf<uint32_t>(B_RUN_DL, n_opsT + 1);
    uint64_t *blk_len = ctx->buf<uint64_t>(B_ROW_LEN, n_rec + 1), *blk_off = ctx->buf<uint64_t>(B_ROW_OFF, n_rec + 1);
    uint32_t *fix_t = ctx->buf<uint32_t>(B_ROW_FIXED, n_rec + 1), *fix_q = ctx->buf<uint32_t>(B_NAME_LEN, n_rec + 1);
    if (!op_len || !c_all || !c_i || !c_d || !S_all || !S_i || !S_d || !op_type || !op_rec || !op_pos || !blk_len || !blk_off ||
        !fix_t || !fix_q)
        return ctx->err_code;
    P2mView V;
    V.n_rec = n_rec; V.n_ops = n_opsT; V.text = d_text; V.R = R; V.X = X; V.op_base = op_base; V.op_len = op_len; V.op_type = op_type;
    V.op_rec = op_rec; V.op_pos = op_pos; V.S_all = S_all; V.S_i = S_i; V.S_d = S_d; V.tseq = TF->d_seq; V.qseq = QF->d_seq;
    V.blk_len = blk_len; V.blk_off = blk_off; V.fix_t = fix_t; V.fix_q = fix_q; V.out = nullptr; V.t_neg = nullptr;
    if (key == kNoErr) {
        LAUNCH(ctx, k_p2m_ops<1>, grid_for(n_rec + 1, 128), 128, n_rec, d_text, R, n_ops, (const uint64_t *)op_base, op_len, op_type,
               op_rec, op_pos, err_key);
        LAUNCH(ctx, k_p2m_contrib, grid_for(n_opsT + 1, 256), 256, n_opsT, op_len, op_type, c_all, c_i, c_d);
        WGA_TRY(exclusive_sum(ctx, c_all, S_all, n_opsT + 1));
        WGA_TRY(exclusive_sum(ctx, c_i, S_i, n_opsT + 1));
        WGA_TRY(exclusive_sum(ctx, c_d, S_d, n_opsT + 1));
        if (n_opsT) LAUNCH(ctx, k_p2m_check, grid_for(n_opsT, 256), 256, V, err_key);
        LAUNCH(ctx, k_p2m_blocks<0>, grid_for(n_rec + 1, 256), 256, V);
        WGA_TRY(exclusive_sum(ctx, blk_len, blk_off, n_rec + 1));
        WGA_TRY(read_scalars(ctx, {blk_off + n_rec, err_key}, sc));
        key = sc[1];
    }
    if (key != kNoErr) {
        const int code = (int)(key & 0xFF);
        const uint64_t rec = key >> 32, pos = (key >> 8) & 0xFFFFFF;
        std::string msg;
        if (code == WGA_E_CSV) msg = "CSV deserialize error by: record " + std::to_string(rec) + " does not match the 12 typed PAF columns";
        else if (code == WGA_E_NOM) msg = "HTS library error by sequence of record " + std::to_string(rec) + " not found in the FASTA index";
        else if (code == WGA_E_CIGAR_TAG_NOT_FOUND) msg = "CIGAR start tag not found";
        else if (code == WGA_E_UNSUPPORTED) msg = "unsupported PAF construct (quoted csv field) in record " + std::to_string(rec);
        else if (code == WGA_E_REF_PANIC) msg = "the reference panics on this record (empty CIGAR, or an insert past the end of the fetched sequence)";
        else if (code == WGA_E_VCF_BASE) {
            uint64_t ql[2];
            WGA_TRY(read_scalars(ctx, {X.q_src + rec, X.LQ + rec}, ql));
            char c = '?';
            if (pos >= 2 && ql[1] > pos - 2) WGA_CK(cudaMemcpy(&c, QF->d_seq + ql[0] + ql[1] - 1 - (pos - 2), 1, cudaMemcpyDeviceToHost));
            msg = std::string("Invalid Base: `") + c + "`";
        } else {
            uint64_t oc[2];
            WGA_TRY(read_scalars(ctx, {R.cig_off + rec, R.cig_len + rec}, oc));
            std::string cg((size_t)std::min<uint64_t>(oc[1], 1u << 20), '\0');
            if (!cg.empty()) WGA_CK(cudaMemcpy(&cg[0], d_text + oc[0], cg.size(), cudaMemcpyDeviceToHost));
            msg = paf_cigar_error_text(cg, code);
        }
        ctx->set_error(code, rec, msg);
        return code;
    }
    const uint64_t total = sc[0];
    char *d_out = (char *)ctx->ensure(B_OUT, total + 16);
    if (!d_out) return ctx->err_code;
    V.out = d_out;
    LAUNCH(ctx, k_p2m_blocks<1>, grid_for(n_rec + 1, 256), 256, V);
    WGA_CK(cudaEventRecord(ctx->ev0, ctx->stream));
    WGA_TRY(p2m_expand_rows(ctx, V, n_opsT));
    WGA_CK(cudaEventRecord(ctx->ev1, ctx->stream));
    int rc = finish_output(ctx, d_out, total, mem, flags, out);
    if (rc) return rc;
    float ms = 0.f;
    WGA_CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->walk_ms = ms; ctx->walk_launches = 1; ctx->walk_bytes = total;  // expansion kernel: output-bound
    return WGA_OK;
}
int wga_paf2maf(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, const wga_fasta *TF, const wga_fasta *QF, wga_span *out) {
    const int rc = wga_paf2maf_impl(ctx, text, len, mem, flags, TF, QF, out);
    return paf_retry_unquoted(ctx, rc, mem, flags, [&](const void *t2, uint64_t n2, int m2, uint32_t f2) {
        return wga_paf2maf_impl(ctx, t2, n2, m2, f2, TF, QF, out);
    });
}

// ---------------------------------------------------------------- call -f paf (caller.rs:268-302,610-822)
static int wga_paf_call_impl(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, const wga_fasta *TF, const wga_fasta *QF,
                 int snp, uint64_t svlen_cutoff, wga_span *out) {
    if (!ctx || !out || !TF || !QF || (!text && len)) return WGA_E_INVALID_ARG;
    cudaSetDevice(ctx->device);
    out->data = nullptr;
    out->len = 0;
    walk_begin(ctx);
    WGA_TRY(stage_text(ctx, text, len, mem));
    PafPrep PP;
    WGA_TRY(paf_fasta_prepare(ctx, len, TF, QF, 0, &PP));
    const uint64_t n_rec = PP.n_rec;
    if (n_rec == 0) return WGA_OK;
    const uint8_t *d_text = ctx->d_text;
    PafRecs R = PP.R;
    P2mRecs X = PP.X;
    unsigned long long *err_key = PP.err_key;
    uint64_t *n_ops = PP.n_ops, *op_base = PP.op_base;
    LAUNCH(ctx, k_cp_inv_check, grid_for(n_rec, 256), 256, n_rec, R, X, (const uint8_t *)TF->d_seq, err_key);
    LAUNCH(ctx, k_cp_ops<0>, grid_for(n_rec + 1, 128), 128, n_rec, d_text, R, n_ops, (const uint64_t *)nullptr, (uint64_t *)nullptr,
           (uint8_t *)nullptr, (uint32_t *)nullptr, err_key);
    WGA_TRY(exclusive_sum(ctx, n_ops, op_base, n_rec + 1));
    uint64_t sc[2];
    WGA_TRY(read_scalars(ctx, {op_base + n_rec, err_key}, sc));
    const uint64_t n_opsT = sc[0];
    uint64_t key = sc[1];
    uint64_t total = 0, n_rows = 0;
    uint64_t *row_len = nullptr, *row_off = nullptr;
    uint32_t *rec_bad = nullptr;
    CpView V;
    if (n_opsT) {  // also when an error is already known: a lower record may still fail at row level
        uint64_t *op_len = ctx->buf<uint64_t>(B_RUN_POS, n_opsT + 1), *tcon = ctx->buf<uint64_t>(B_GEN_C, n_opsT + 1);
        uint64_t *qcon = ctx->buf<uint64_t>(B_GEN_D, n_opsT + 1), *rows = ctx->buf<uint64_t>(B_GEN_E, n_opsT + 1);
        uint64_t *St = ctx->buf<uint64_t>(B_GEN_F, n_opsT + 1), *Sq = ctx->buf<uint64_t>(B_GEN_G, n_opsT + 1);
        uint64_t *rowbase = ctx->buf<uint64_t>(B_GEN_H, n_opsT + 1);
        uint8_t *op_type = ctx->buf<uint8_t>(B_RUN_CLS, n_opsT + 1);
        uint32_t *op_rec = ctx->buf<uint32_t>(B_RUN_REC, n_opsT + 1);
        rec_bad = ctx->buf<uint32_t>(B_ROW_FIXED, 2 * n_rec);
        if (!op_len || !tcon || !qcon || !rows || !St || !Sq || !rowbase || !op_type || !op_rec || !rec_bad) return ctx->err_code;
        uint32_t *rec_panic = rec_bad + n_rec;
        WGA_CK(cudaMemsetAsync(rec_bad, 0xFF, 2 * n_rec * sizeof(uint32_t), ctx->stream));
        LAUNCH(ctx, k_cp_ops<1>, grid_for(n_rec + 1, 128), 128, n_rec, d_text, R, n_ops, (const uint64_t *)op_base, op_len, op_type,
               op_rec, err_key);
        LAUNCH(ctx, k_cp_contrib, grid_for(n_opsT + 1, 256), 256, n_opsT, op_len, op_type, op_rec, op_base, svlen_cutoff, snp, tcon,
               qcon, rows);
        WGA_TRY(exclusive_sum(ctx, tcon, St, n_opsT + 1));
        WGA_TRY(exclusive_sum(ctx, qcon, Sq, n_opsT + 1));
        WGA_TRY(exclusive_sum(ctx, rows, rowbase, n_opsT + 1));
        WGA_TRY(read_scalars(ctx, {rowbase + n_opsT}, &n_rows));
        V.n_rec = n_rec; V.n_ops = n_opsT; V.text = d_text; V.R = R; V.X = X; V.op_base = op_base; V.op_len = op_len;
        V.op_type = op_type; V.op_rec = op_rec; V.St = St; V.Sq = Sq; V.rowbase = rowbase; V.tseq = TF->d_seq; V.qseq = QF->d_seq;
        V.svlen = svlen_cutoff; V.snp = snp;
        if (n_rows) {
            row_len = ctx->buf<uint64_t>(B_ROW_LEN, n_rows + 1);
            row_off = ctx->buf<uint64_t>(B_ROW_OFF, n_rows + 1);
            if (!row_len || !row_off) return ctx->err_code;
            LAUNCH(ctx, k_cp_rows<0>, grid_for(n_rows + 1, 128), 128, V, n_rows, row_len, (const uint64_t *)nullptr, rec_bad, rec_panic,
                   (char *)nullptr, err_key);
            LAUNCH(ctx, k_cp_resolve, grid_for(n_rows, 256), 256, V, n_rows, row_len, (const uint32_t *)rec_bad,
                   (const uint32_t *)rec_panic, err_key);
            WGA_TRY(exclusive_sum(ctx, row_len, row_off, n_rows + 1));
            WGA_TRY(read_scalars(ctx, {row_off + n_rows, err_key}, sc));
            total = sc[0];
            key = sc[1];
        }
    }
    if (key != kNoErr) {
        const int code = (int)(key & 0xFF);
        const uint64_t rec = key >> 32, pos = (key >> 8) & 0xFFFFFF;
        std::string msg;
        if (code == WGA_E_CSV) msg = "CSV deserialize error by: record " + std::to_string(rec) + " does not match the 12 typed PAF columns";
        else if (code == WGA_E_NOM) msg = "HTS library error by sequence of record " + std::to_string(rec) + " not found in the FASTA index";
        else if (code == WGA_E_CIGAR_TAG_NOT_FOUND) msg = "CIGAR start tag not found";
        else if (code == WGA_E_UNSUPPORTED) msg = "unsupported PAF construct (quoted csv field) in record " + std::to_string(rec);
        else if (code == WGA_E_VCF_BASE) {
            uint64_t ts;
            WGA_TRY(read_scalars(ctx, {X.t_src + rec}, &ts));
            char c = '?';
            WGA_CK(cudaMemcpy(&c, TF->d_seq + ts, 1, cudaMemcpyDeviceToHost));
            msg = std::string("invalid VCF base `") + c + "`";
        } else msg = pos == 0x800001 ? "the reference panics on this record (empty CIGAR)"
                                     : "the reference panics on this record (slice past the fetched sequence)";
        ctx->set_error(code, rec, msg);
        return code;
    }
    if (total == 0) return sync_check(ctx);
    char *d_out = (char *)ctx->ensure(B_OUT, total + 16);
    if (!d_out) return ctx->err_code;
    LAUNCH(ctx, k_cp_rows<1>, grid_for(n_rows, 128), 128, V, n_rows, row_len, (const uint64_t *)row_off, rec_bad, rec_bad + n_rec, d_out,
           err_key);
    return finish_output(ctx, d_out, total, mem, flags, out);
}
int wga_paf_call(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, const wga_fasta *TF, const wga_fasta *QF, int snp, uint64_t svlen_cutoff, wga_span *out) {
    const int rc = wga_paf_call_impl(ctx, text, len, mem, flags, TF, QF, snp, svlen_cutoff, out);
    return paf_retry_unquoted(ctx, rc, mem, flags, [&](const void *t2, uint64_t n2, int m2, uint32_t f2) {
        return wga_paf_call_impl(ctx, t2, n2, m2, f2, TF, QF, snp, svlen_cutoff, out);
    });
}

// ---------------------------------------------------------------- paf2chain (converter.rs:148-173)
// kind 0: chain records; kind 1: base-level plot segments (ops other than M = X I D are legal there and ignored)
static int paf_op_rows(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, uint64_t id_base, wga_span *out,
                       uint64_t *n_records, int kind, uint64_t cutoff) {
    if (!ctx || !out || (!text && len)) return WGA_E_INVALID_ARG;
    cudaSetDevice(ctx->device);
    out->data = nullptr;
    out->len = 0;
    if (n_records) *n_records = 0;
    walk_begin(ctx);
    WGA_TRY(stage_text(ctx, text, len, mem));
    PafPrep PP;
    WGA_TRY(paf_fasta_prepare(ctx, len, nullptr, nullptr, 1, &PP, false, nullptr));
    const uint64_t n_rec = PP.n_rec;
    if (n_records) *n_records = n_rec;
    if (n_rec == 0) return WGA_OK;
    const uint8_t *d_text = ctx->d_text;
    PafRecs R = PP.R;
    P2mRecs X = PP.X;
    unsigned long long *err_key = PP.err_key;
    uint64_t *n_ops = PP.n_ops, *op_base = PP.op_base;
    uint8_t *slow = ctx->buf<uint8_t>(B_REC_SKIP, n_rec + 1);  // records the warp kernel leaves to the exact one
    if (!slow) return ctx->err_code;
    LAUNCH(ctx, k_paf_ops_warp<0>, grid_for(n_rec * 32, 128), 128, n_rec, d_text, R, n_ops, slow, (const uint64_t *)nullptr,
           (uint64_t *)nullptr, (uint8_t *)nullptr);
    WGA_CK(cudaMemsetAsync(n_ops + n_rec, 0, 8, ctx->stream));
    LAUNCH(ctx, k_p2m_ops<0>, grid_for(n_rec + 1, 128), 128, n_rec, d_text, R, n_ops, (const uint64_t *)nullptr, (uint64_t *)nullptr,
           (uint8_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, err_key, (const uint8_t *)slow, kind == 1);
    WGA_TRY(exclusive_sum(ctx, n_ops, op_base, n_rec + 1));
    uint64_t sc[2];
    WGA_TRY(read_scalars(ctx, {op_base + n_rec, err_key}, sc));
    if (sc[1] != kNoErr) {
        const int code = (int)(sc[1] & 0xFF);
        const uint64_t rec = sc[1] >> 32;
        std::string msg;
        if (code == WGA_E_CSV) msg = "CSV deserialize error by: record " + std::to_string(rec) + " does not match the 12 typed PAF columns";
        else if (code == WGA_E_CIGAR_TAG_NOT_FOUND) msg = "CIGAR start tag not found";
        else if (code == WGA_E_UNSUPPORTED) msg = "unsupported PAF construct (quoted csv field) in record " + std::to_string(rec);
        else if (code == WGA_E_REF_PANIC) msg = "the reference panics on this record (empty CIGAR, or a cs:Z: length beyond u64)";
        else {
            uint64_t oc[2];
            WGA_TRY(read_scalars(ctx, {R.cig_off + rec, R.cig_len + rec}, oc));
            std::string cg((size_t)std::min<uint64_t>(oc[1], 1u << 20), '\0');
            if (!cg.empty()) WGA_CK(cudaMemcpy(&cg[0], d_text + oc[0], cg.size(), cudaMemcpyDeviceToHost));
            msg = paf_cigar_error_text(cg, code);
        }
        ctx->set_error(code, rec, msg);
        return code;
    }
    const uint64_t n_opsT = sc[0];
    uint64_t *op_len = ctx->buf<uint64_t>(B_RUN_POS, n_opsT + 1);
    uint8_t *op_type = ctx->buf<uint8_t>(B_RUN_CLS, n_opsT + 1);
    uint64_t *row_len = ctx->buf<uint64_t>(B_ROW_LEN, n_rec + 1), *row_off = ctx->buf<uint64_t>(B_ROW_OFF, n_rec + 1);
    if (!op_len || !op_type || !row_len || !row_off) return ctx->err_code;
    LAUNCH(ctx, k_paf_ops_warp<1>, grid_for(n_rec * 32, 128), 128, n_rec, d_text, R, n_ops, slow, (const uint64_t *)op_base, op_len,
           op_type);
    LAUNCH(ctx, k_p2m_ops<1>, grid_for(n_rec + 1, 128), 128, n_rec, d_text, R, n_ops, (const uint64_t *)op_base, op_len, op_type,
           (uint32_t *)nullptr, (uint32_t *)nullptr, err_key, (const uint8_t *)slow, kind == 1);
    ChainView V;
    V.n_rec = n_rec; V.id_base = id_base; V.text = d_text; V.t_name_off = R.t_name_off; V.q_name_off = R.q_name_off;
    V.t_name_len = R.t_name_len; V.q_name_len = R.q_name_len; V.t_size = R.t_size; V.q_size = R.q_size; V.t_start = R.t_start;
    V.t_end = X.t_end; V.q_start = R.q_start; V.q_end = X.q_end; V.neg = R.neg; V.skip = nullptr; V.run_first = op_base;
    V.run_len = op_len; V.run_type = op_type; V.maf = 0; V.row_len = row_len; V.row_off = row_off; V.out = nullptr;
    static const bool chain_seq = getenv("WGA_CHAIN_SEQ") != nullptr && getenv("WGA_CHAIN_SEQ")[0] == '1';  // one thread per record
    static const bool plot_seq = getenv("WGA_PLOT_SEQ") != nullptr && getenv("WGA_PLOT_SEQ")[0] == '1';     // one thread per record
    uint32_t *hdr_len = ctx->buf<uint32_t>(B_CHAIN_HDR, n_rec + 1);
    uint8_t *irregular = ctx->buf<uint8_t>(B_CHAIN_IRR, n_rec + 1);
    if (!hdr_len || !irregular) return ctx->err_code;
    if (kind == 0 && chain_seq) LAUNCH(ctx, k_chain_rows<0>, grid_for(n_rec + 1, 128), 128, V);
    else if (kind == 0) LAUNCH(ctx, k_chain_rows_warp<0>, grid_for((n_rec + 1) * 32, 128), 128, V, hdr_len, irregular);
    else if (plot_seq) LAUNCH(ctx, k_plot_rows<0>, grid_for(n_rec + 1, 128), 128, V, cutoff);
    else LAUNCH(ctx, k_plot_rows_warp<0>, grid_for((n_rec + 1) * 32, 128), 128, V, cutoff);
    WGA_TRY(exclusive_sum(ctx, row_len, row_off, n_rec + 1));
    uint64_t total;
    WGA_TRY(read_scalars(ctx, {row_off + n_rec}, &total));
    if (total == 0) return WGA_OK;
    char *d_out = (char *)ctx->ensure(B_OUT, total + 16);
    if (!d_out) return ctx->err_code;
    V.out = d_out;
    if (kind == 0 && chain_seq) LAUNCH(ctx, k_chain_rows<1>, grid_for(n_rec, 128), 128, V);
    else if (kind == 0) LAUNCH(ctx, k_chain_rows_warp<1>, grid_for(n_rec * 32, 128), 128, V, hdr_len, irregular);
    else if (plot_seq) LAUNCH(ctx, k_plot_rows<1>, grid_for(n_rec, 128), 128, V, cutoff);
    else LAUNCH(ctx, k_plot_rows_warp<1>, grid_for(n_rec * 32, 128), 128, V, cutoff);
    return finish_output(ctx, d_out, total, mem, flags, out);
}
int wga_paf2chain(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, uint64_t id_base, wga_span *out,
                  uint64_t *n_records) {
    const int rc = paf_op_rows(ctx, text, len, mem, flags, id_base, out, n_records, 0, 0);
    return paf_retry_unquoted(ctx, rc, mem, flags, [&](const void *t2, uint64_t n2, int m2, uint32_t f2) {
        return paf_op_rows(ctx, t2, n2, m2, f2, id_base, out, n_records, 0, 0);
    });
}
// ---------------------------------------------------------------- dotplot -m base-level, PAF (cigar.rs:917-952)
int wga_paf_plotdata(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, uint64_t skip_cutoff, wga_span *out) {
    const int rc = paf_op_rows(ctx, text, len, mem, flags, 0, out, nullptr, 1, skip_cutoff);
    return paf_retry_unquoted(ctx, rc, mem, flags, [&](const void *t2, uint64_t n2, int m2, uint32_t f2) {
        return paf_op_rows(ctx, t2, n2, m2, f2, 0, out, nullptr, 1, skip_cutoff);
    });
}

// ---------------------------------------------------------------- chain reader front end (chain.rs:27-75,206-383)
}  // extern "C"
namespace {
struct ChainPrep {
    LineView lv;
    uint64_t n_lines = 0, n_rec = 0;
    uint32_t *hdr_scan = nullptr;
    uint64_t *hdr_line = nullptr;
    ChainRecs C;
    uint64_t *v = nullptr;
    uint32_t *dl_cig = nullptr, *dl_ops = nullptr;
    unsigned long long *err_key = nullptr;  // [0] reader errors, [1] errors of the stages after it (chain2maf)
};

// host restatement of the reader's error wording (chain.rs:206-350, errors.rs) for the line a device error points at
std::string chain_error_text(const std::string &line, int code, bool header) {
    static const char *fname[12] = {"score", "target_name", "target_size", "target_strand", "target_start", "target_end",
                                    "query_name", "query_size", "query_strand", "query_start", "query_end", "chain_id"};
    std::vector<std::string> f;
    size_t i = header ? std::min<size_t>(5, line.size()) : 0;
    while (i < line.size()) {
        while (i < line.size() && nat_ws(line[i])) ++i;
        if (i >= line.size()) break;
        size_t b = i;
        while (i < line.size() && !nat_ws(line[i])) ++i;
        f.push_back(line.substr(b, i - b));
    }
    auto bad_int = [](const std::string &t) {
        uint64_t v;
        return !parse_u64_bytes((const unsigned char *)t.data(), t.size(), &v);
    };
    if (header) {
        for (int k = 0; k < 12; ++k) {
            if ((size_t)k >= f.size()) return std::string("Parse Chain Error By: Chain Line Field `") + fname[k] + "` Missing";
            if (k == 0) { if (code == WGA_E_PARSE_FLOAT) return "Parse `" + f[0] + "` Into Float Error"; }
            else if (k == 3 || k == 8) { if (f[k] != "+" && f[k] != "-") return "Parse Strand `" + f[k] + "` Error"; }
            else if (k != 1 && k != 6 && bad_int(f[k])) return "Parse `" + f[k] + "` Into Integer Error";
        }
    } else {
        if (f.empty()) return "Parse Chain Error By: Chain Line Field `size` Missing";
        for (size_t k = 0; k < f.size() && k < 3; ++k)
            if (bad_int(f[k])) return "Parse `" + f[k] + "` Into Integer Error";
    }
    return "chain parse error code " + std::to_string(code);
}

int chain_resolve_error(wga_ctx *ctx, const ChainPrep &CP, uint64_t key) {
    const int code = (int)(key & 0xFF);
    const uint64_t rec = key >> 32, stage = (key >> 8) & 0xFFFFFF;
    std::string msg;
    if (code == WGA_E_NOM) msg = "Format error at record " + std::to_string(rec) + " Parse Error by rust::nom, please check (a chain record must start with `chain` and hold at least one data line)";
    else if (code == WGA_E_REF_PANIC) msg = "the reference panics on this input (nom error with fewer than 10 bytes left, errors.rs:92)";
    else if (code == WGA_E_UNSUPPORTED) msg = "unsupported chain construct: `chain` in the middle of a line";
    else if (rec < CP.n_rec && (stage == 1 || stage >= 3)) {
        uint64_t h;
        WGA_TRY(read_scalars(ctx, {CP.hdr_line + rec}, &h));
        std::string line;
        if (fetch_line(ctx, CP.lv, stage == 1 ? h : h + (stage - 3), &line) == WGA_OK) msg = chain_error_text(line, code, stage == 1);
    }
    if (msg.empty()) msg = "chain parse error code " + std::to_string(code);
    ctx->set_error(code, rec, msg);
    return code;
}

// text -> line classes -> records (headers) + per-line data values; leaves the reader's error key unread in err_key[0]
int chain_prepare(wga_ctx *ctx, uint64_t len, ChainPrep *CP) {
    ctx->n_rec = 0;
    const uint8_t *d_text = ctx->d_text;
    unsigned long long *err_key = ctx->buf<unsigned long long>(B_ERR, 2);
    if (!err_key) return ctx->err_code;
    LAUNCH(ctx, k_init_u64, 1, 2, err_key, 2, kNoErr);
    CP->err_key = err_key;
    if (len == 0) return WGA_OK;
    uint64_t n_ws = 0, n_nl = 0;
    uint64_t *ws_ent = nullptr, *nl_idx = nullptr;
    WGA_TRY((build_delim_index<0, kTokCap>(ctx, d_text, len, &n_ws, &n_nl, &ws_ent, &nl_idx)));
    const uint8_t last_byte = *(uint8_t *)(ctx->h_scal + 8);
    const uint64_t n_lines = n_nl + (last_byte != '\n' ? 1 : 0);
    LineView lv;
    lv.text = d_text; lv.len = len; lv.ws_ent = ws_ent; lv.nl_ws_idx = nl_idx;
    lv.n_ws = n_ws; lv.n_nl = n_nl; lv.n_lines = n_lines; lv.first_slab = 0;
    uint8_t *cls = ctx->buf<uint8_t>(B_LINE_FLAGS, n_lines + 1);
    uint32_t *is_hdr = ctx->buf<uint32_t>(B_LINE_HEAD, n_lines + 1), *hdr_scan = ctx->buf<uint32_t>(B_LINE_REC, n_lines + 1);
    uint32_t *nonvalid = ctx->buf<uint32_t>(B_PT_FLAG, n_lines + 1), *nv_scan = ctx->buf<uint32_t>(B_PT_IDX, n_lines + 1);
    uint64_t *first_c = ctx->buf<uint64_t>(B_PP_J, n_lines + 1);
    if (!cls || !is_hdr || !hdr_scan || !nonvalid || !nv_scan || !first_c) return ctx->err_code;
    LAUNCH(ctx, k_chain_classify, grid_for(n_lines + 1, 256), 256, lv, cls, is_hdr, nonvalid, first_c);
    WGA_TRY(exclusive_sum(ctx, is_hdr, hdr_scan, n_lines + 1));
    WGA_TRY(exclusive_sum(ctx, nonvalid, nv_scan, n_lines + 1));
    WGA_CK(cudaMemcpyAsync(ctx->h_scal, hdr_scan + n_lines, 4, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_CK(cudaMemcpyAsync(ctx->h_scal + 1, nv_scan + n_lines, 4, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_TRY(sync_check(ctx));
    const uint64_t n_rec = *(uint32_t *)ctx->h_scal, n_nv = *(uint32_t *)(ctx->h_scal + 1);
    ctx->n_rec = n_rec;
    uint64_t *hdr_line = ctx->buf<uint64_t>(B_REC_FIRST, n_rec + 1), *nv_line = ctx->buf<uint64_t>(B_REC_LAST, n_nv + 1);
    ChainRecs C;
    PafRecs &R = C.R;
    R.q_name_off = ctx->buf<uint64_t>(B_PAF_B, n_rec + 1); R.t_name_off = ctx->buf<uint64_t>(B_PAF_C, n_rec + 1);
    R.cig_off = ctx->buf<uint64_t>(B_PAF_D, n_rec + 1); R.cig_len = ctx->buf<uint64_t>(B_PAF_E, n_rec + 1);
    R.q_name_len = ctx->buf<uint32_t>(B_PAF_F, n_rec + 1); R.t_name_len = ctx->buf<uint32_t>(B_PAF_G, n_rec + 1);
    uint64_t *m4 = ctx->buf<uint64_t>(B_PAF_H, 4 * (n_rec + 1));
    R.neg = ctx->buf<uint8_t>(B_PAF_I, n_rec + 1);
    uint64_t *e2 = ctx->buf<uint64_t>(B_PP_K, 2 * (n_rec + 1));
    C.first_data = ctx->buf<uint64_t>(B_PP_L, n_rec + 1); C.n_data = ctx->buf<uint64_t>(B_PP_M, n_rec + 1);
    C.t_neg = ctx->buf<uint8_t>(B_PP_N, n_rec + 1);
    uint64_t *v = ctx->buf<uint64_t>(B_PP_A, 3 * (n_lines + 1));
    uint32_t *dl_cig = ctx->buf<uint32_t>(B_PP_B, n_lines + 1), *dl_ops = ctx->buf<uint32_t>(B_PP_C, n_lines + 1);
    if (!hdr_line || !nv_line || !R.q_name_off || !R.t_name_off || !R.cig_off || !R.cig_len || !R.q_name_len || !R.t_name_len || !m4 ||
        !R.neg || !e2 || !C.first_data || !C.n_data || !C.t_neg || !v || !dl_cig || !dl_ops)
        return ctx->err_code;
    R.q_size = m4; R.t_size = m4 + (n_rec + 1); R.q_start = m4 + 2 * (n_rec + 1); R.t_start = m4 + 3 * (n_rec + 1);
    C.t_end = e2; C.q_end = e2 + (n_rec + 1);
    LAUNCH(ctx, k_chain_lists, grid_for(n_lines, 256), 256, n_lines, (const uint8_t *)cls, (const uint32_t *)hdr_scan, (const uint32_t *)nv_scan,
           hdr_line, nv_line);
    if (n_rec)
        LAUNCH(ctx, k_chain_headers, grid_for(n_rec, 128), 128, lv, n_rec, (const uint64_t *)hdr_line, (const uint32_t *)nv_scan,
               (const uint64_t *)nv_line, n_nv, C, err_key);
    LAUNCH(ctx, k_chain_data, grid_for(n_lines + 1, 256), 256, lv, (const uint8_t *)cls, (const uint32_t *)hdr_scan, (const uint32_t *)nv_scan,
           (const uint64_t *)hdr_line, (const uint64_t *)first_c, v, dl_cig, dl_ops, err_key);
    CP->lv = lv; CP->n_lines = n_lines; CP->n_rec = n_rec; CP->hdr_scan = hdr_scan; CP->hdr_line = hdr_line; CP->C = C; CP->v = v;
    CP->dl_cig = dl_cig; CP->dl_ops = dl_ops;
    return WGA_OK;
}
}  // namespace
extern "C" {

// ---------------------------------------------------------------- chain2paf (converter.rs:391-416, chain.rs:430-451)
int wga_chain2paf(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, wga_span *out) {
    if (!ctx || !out || (!text && len)) return WGA_E_INVALID_ARG;
    cudaSetDevice(ctx->device);
    out->data = nullptr;
    out->len = 0;
    walk_begin(ctx);
    WGA_TRY(stage_text(ctx, text, len, mem));
    ChainPrep CP;
    WGA_TRY(chain_prepare(ctx, len, &CP));
    if (len == 0) return WGA_OK;
    const uint64_t n_rec = CP.n_rec, n_lines = CP.n_lines;
    uint64_t *a = ctx->buf<uint64_t>(B_PP_H, n_lines + 1), *b = ctx->buf<uint64_t>(B_PP_I, n_lines + 1);
    uint64_t *S_size = ctx->buf<uint64_t>(B_PP_D, n_lines + 1), *S_qd = ctx->buf<uint64_t>(B_PP_E, n_lines + 1);
    uint64_t *S_cig = ctx->buf<uint64_t>(B_PP_F, n_lines + 1);
    uint64_t *row_len = ctx->buf<uint64_t>(B_ROW_LEN, n_rec + 1), *row_off = ctx->buf<uint64_t>(B_ROW_OFF, n_rec + 1);
    uint32_t *row_fixed = ctx->buf<uint32_t>(B_ROW_FIXED, n_rec + 1);
    if (!a || !b || !S_size || !S_qd || !S_cig || !row_len || !row_off || !row_fixed) return ctx->err_code;
    LAUNCH(ctx, k_chain_split3, grid_for(n_lines + 1, 256), 256, n_lines, (const uint64_t *)CP.v, a, b);
    WGA_TRY(exclusive_sum(ctx, a, S_size, n_lines + 1));
    WGA_TRY(exclusive_sum(ctx, b, S_qd, n_lines + 1));
    WGA_TRY(exclusive_sum(ctx, CP.dl_cig, S_cig, n_lines + 1));
    uint64_t key;
    WGA_TRY(read_scalars(ctx, {CP.err_key}, &key));
    if (key != kNoErr) return chain_resolve_error(ctx, CP, key);
    if (n_rec == 0) return WGA_OK;
    C2pView V;
    V.n_rec = n_rec; V.text = ctx->d_text; V.C = CP.C; V.S_size = S_size; V.S_qd = S_qd; V.S_cig = S_cig; V.v = CP.v;
    V.row_len = row_len; V.row_fixed = row_fixed; V.row_off = row_off; V.out = nullptr;
    LAUNCH(ctx, k_c2p_rows<0>, grid_for(n_rec + 1, 128), 128, V);
    WGA_TRY(exclusive_sum(ctx, row_len, row_off, n_rec + 1));
    uint64_t total;
    WGA_TRY(read_scalars(ctx, {row_off + n_rec}, &total));
    char *d_out = (char *)ctx->ensure(B_OUT, total + 16);
    if (!d_out) return ctx->err_code;
    V.out = d_out;
    LAUNCH(ctx, k_c2p_rows<1>, grid_for(n_rec, 128), 128, V);
    LAUNCH(ctx, k_c2p_cigar, grid_for(n_lines, 256), 256, n_lines, (const uint32_t *)CP.dl_cig, (const uint32_t *)CP.hdr_scan, V);
    return finish_output(ctx, d_out, total, mem, flags, out);
}

// ---------------------------------------------------------------- chain2maf (converter.rs:268-388)
int wga_chain2maf_header(wga_ctx *ctx, const char *t_fa_path, const char *q_fa_path, wga_span *out) {
    if (!ctx || !out || !t_fa_path || !q_fa_path) return WGA_E_INVALID_ARG;
    static thread_local std::string hdr;
    hdr = std::string("#maf version=1.6 convert_from=chain t_seq_path=") + t_fa_path + " q_seq_path=" + q_fa_path + "\n";
    out->data = hdr.data();
    out->len = hdr.size();
    return WGA_OK;
}

int wga_chain2maf(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, const wga_fasta *TF, const wga_fasta *QF,
                  wga_span *out) {
    if (!ctx || !out || !TF || !QF || (!text && len)) return WGA_E_INVALID_ARG;
    cudaSetDevice(ctx->device);
    out->data = nullptr;
    out->len = 0;
    walk_begin(ctx);
    WGA_TRY(stage_text(ctx, text, len, mem));
    ChainPrep CP;
    WGA_TRY(chain_prepare(ctx, len, &CP));
    if (len == 0) return WGA_OK;
    const uint64_t n_rec = CP.n_rec, n_lines = CP.n_lines;
    const uint8_t *d_text = ctx->d_text;
    unsigned long long *fe_key = CP.err_key, *be_key = CP.err_key + 1;
    uint64_t *S_ops = ctx->buf<uint64_t>(B_PP_G, n_lines + 1);
    if (!S_ops) return ctx->err_code;
    WGA_TRY(exclusive_sum(ctx, CP.dl_ops, S_ops, n_lines + 1));
    uint64_t sc[3];
    WGA_TRY(read_scalars(ctx, {S_ops + n_lines, fe_key}, sc));
    if (n_rec == 0) return sc[1] != kNoErr ? chain_resolve_error(ctx, CP, sc[1]) : WGA_OK;
    // every data line of a record the reader accepted has >= 1 op, so the ops of record r are S_ops[first_data .. +n_data)
    const uint64_t n_opsT = sc[0];
    const uint64_t fe = sc[1];
    PafRecs R = CP.C.R;
    uint64_t *x7 = ctx->buf<uint64_t>(B_PAF_J, 7 * (n_rec + 1));
    uint64_t *n_ops = ctx->buf<uint64_t>(B_GEN_A, n_rec + 1), *op_base = ctx->buf<uint64_t>(B_GEN_B, n_rec + 1);
    uint64_t *op_len = ctx->buf<uint64_t>(B_RUN_POS, n_opsT + 1), *c_all = ctx->buf<uint64_t>(B_GEN_C, n_opsT + 1);
    uint64_t *c_i = ctx->buf<uint64_t>(B_GEN_D, n_opsT + 1), *c_d = ctx->buf<uint64_t>(B_GEN_E, n_opsT + 1);
    uint64_t *S_all = ctx->buf<uint64_t>(B_GEN_F, n_opsT + 1), *S_i = ctx->buf<uint64_t>(B_GEN_G, n_opsT + 1);
    uint64_t *S_d = ctx->buf<uint64_t>(B_GEN_H, n_opsT + 1);
    uint8_t *op_type = ctx->buf<uint8_t>(B_RUN_CLS, n_opsT + 1);
    uint32_t *op_rec = ctx->buf<uint32_t>(B_RUN_REC, n_opsT + 1), *op_pos = ctx->buf<uint32_t>(B_RUN_DL, n_opsT + 1);
    uint64_t *blk_len = ctx->buf<uint64_t>(B_ROW_LEN, n_rec + 1), *blk_off = ctx->buf<uint64_t>(B_ROW_OFF, n_rec + 1);
    uint32_t *fix_t = ctx->buf<uint32_t>(B_ROW_FIXED, n_rec + 1), *fix_q = ctx->buf<uint32_t>(B_NAME_LEN, n_rec + 1);
    if (!x7 || !n_ops || !op_base || !op_len || !c_all || !c_i || !c_d || !S_all || !S_i || !S_d || !op_type || !op_rec || !op_pos ||
        !blk_len || !blk_off || !fix_t || !fix_q)
        return ctx->err_code;
    P2mRecs X;
    X.q_end = x7; X.t_end = x7 + (n_rec + 1); X.mapq = x7 + 2 * (n_rec + 1); X.t_src = x7 + 3 * (n_rec + 1);
    X.q_src = x7 + 4 * (n_rec + 1); X.LT = x7 + 5 * (n_rec + 1); X.LQ = x7 + 6 * (n_rec + 1); X.n_ops = n_ops;
    FaView T{TF->d_seq, TF->d_names, TF->d_meta, TF->n_seq}, Q{QF->d_seq, QF->d_names, QF->d_meta, QF->n_seq};
    uint64_t be = kNoErr, total = 0;
    P2mView V;
    if (fe == kNoErr || (fe >> 32) > 0) {  // a reader error in record 0 precedes everything else
        LAUNCH(ctx, k_c2m_spans, grid_for(n_rec + 1, 256), 256, n_rec, d_text, CP.C, X, T, Q, (const uint64_t *)S_ops, op_base, be_key);
        LAUNCH(ctx, k_p2m_revcomp_check, (unsigned)n_rec, 128, n_rec, R, X, QF->d_seq, be_key);
        LAUNCH(ctx, k_c2m_ops, grid_for(n_lines, 256), 256, n_lines, (const uint32_t *)CP.dl_ops, (const uint32_t *)CP.hdr_scan,
               (const uint64_t *)S_ops, (const uint64_t *)CP.v, op_len, op_type, op_rec, op_pos);
        V.n_rec = n_rec; V.n_ops = n_opsT; V.text = d_text; V.R = R; V.X = X; V.op_base = op_base; V.op_len = op_len; V.op_type = op_type;
        V.op_rec = op_rec; V.op_pos = op_pos; V.S_all = S_all; V.S_i = S_i; V.S_d = S_d; V.tseq = TF->d_seq; V.qseq = QF->d_seq;
        V.blk_len = blk_len; V.blk_off = blk_off; V.fix_t = fix_t; V.fix_q = fix_q; V.out = nullptr; V.t_neg = CP.C.t_neg;
        LAUNCH(ctx, k_p2m_contrib, grid_for(n_opsT + 1, 256), 256, n_opsT, op_len, op_type, c_all, c_i, c_d);
        WGA_TRY(exclusive_sum(ctx, c_all, S_all, n_opsT + 1));
        WGA_TRY(exclusive_sum(ctx, c_i, S_i, n_opsT + 1));
        WGA_TRY(exclusive_sum(ctx, c_d, S_d, n_opsT + 1));
        if (n_opsT) LAUNCH(ctx, k_p2m_check, grid_for(n_opsT, 256), 256, V, be_key);
        LAUNCH(ctx, k_p2m_blocks<0>, grid_for(n_rec + 1, 256), 256, V);
        WGA_TRY(exclusive_sum(ctx, blk_len, blk_off, n_rec + 1));
        WGA_TRY(read_scalars(ctx, {blk_off + n_rec, be_key}, sc));
        total = sc[0];
        be = sc[1];
    }
    // the reference reads record r, then fetches / reverse-complements / expands it, then reads record r+1: the reader's
    // error in record a comes before a later stage's error in record b iff a <= b
    if (fe != kNoErr && (be == kNoErr || (fe >> 32) <= (be >> 32))) return chain_resolve_error(ctx, CP, fe);
    if (be != kNoErr) {
        const int code = (int)(be & 0xFF);
        const uint64_t rec = be >> 32, pos = (be >> 8) & 0xFFFFFF;
        std::string msg;
        if (code == WGA_E_NOM) msg = "HTS library error by sequence of record " + std::to_string(rec) + " not found in the FASTA index";
        else if (code == WGA_E_VCF_BASE) {
            uint64_t ql[2];
            WGA_TRY(read_scalars(ctx, {X.q_src + rec, X.LQ + rec}, ql));
            char c = '?';
            if (pos >= 2 && ql[1] > pos - 2) WGA_CK(cudaMemcpy(&c, QF->d_seq + ql[0] + ql[1] - 1 - (pos - 2), 1, cudaMemcpyDeviceToHost));
            msg = std::string("Invalid Base: `") + c + "`";
        } else msg = "the reference panics on this record (an insert past the end of the fetched sequence)";
        ctx->set_error(code, rec, msg);
        return code;
    }
    char *d_out = (char *)ctx->ensure(B_OUT, total + 16);
    if (!d_out) return ctx->err_code;
    V.out = d_out;
    LAUNCH(ctx, k_p2m_blocks<1>, grid_for(n_rec + 1, 256), 256, V);
    WGA_TRY(p2m_expand_rows(ctx, V, n_opsT));
    return finish_output(ctx, d_out, total, mem, flags, out);
}

// ---------------------------------------------------------------- pafcov (pafcov.rs:13-60, cigar.rs:710-741)
static int wga_paf_cov_impl(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, wga_span *out) {
    if (!ctx || !out || (!text && len)) return WGA_E_INVALID_ARG;
    cudaSetDevice(ctx->device);
    out->data = nullptr;
    out->len = 0;
    walk_begin(ctx);
    WGA_TRY(stage_text(ctx, text, len, mem));
    PafPrep PP;
    WGA_TRY(paf_fasta_prepare(ctx, len, nullptr, nullptr, 1, &PP, false, nullptr));
    const uint64_t n_rec = PP.n_rec;
    if (n_rec == 0) return WGA_OK;
    const uint8_t *d_text = ctx->d_text;
    PafRecs R = PP.R;
    unsigned long long *err_key = PP.err_key;
    // ---- targets in first-occurrence order
    uint64_t *kt = ctx->buf<uint64_t>(B_PP_A, n_rec), *kp = ctx->buf<uint64_t>(B_PP_B, n_rec), *ft = ctx->buf<uint64_t>(B_PP_C, n_rec);
    uint64_t *slot_of = ctx->buf<uint64_t>(B_PP_E, n_rec);
    uint32_t *head = ctx->buf<uint32_t>(B_PP_J, n_rec + 1), *tidx = ctx->buf<uint32_t>(B_PP_K, n_rec + 1);
    uint64_t cap = 64;
    while (cap < 2 * n_rec) cap <<= 1;
    unsigned long long *tkeys = ctx->buf<unsigned long long>(B_PT_KEYS, cap), *tfirst = ctx->buf<unsigned long long>(B_PT_FIRST, cap);
    if (!kt || !kp || !ft || !slot_of || !head || !tidx || !tkeys || !tfirst) return ctx->err_code;
    LAUNCH(ctx, k_pp_keys, grid_for(n_rec, 256), 256, n_rec, d_text, R, (const uint8_t *)nullptr, 0u, 0, kt, kp);
    WGA_CK(cudaMemsetAsync(tkeys, 0, cap * 8, ctx->stream));
    WGA_CK(cudaMemsetAsync(tfirst, 0xFF, cap * 8, ctx->stream));
    LAUNCH(ctx, k_pp_first_insert, grid_for(n_rec, 256), 256, n_rec, (const uint64_t *)kt, cap - 1, tkeys, tfirst, slot_of);
    LAUNCH(ctx, k_pp_first_read, grid_for(n_rec, 256), 256, n_rec, (const uint64_t *)slot_of, (const unsigned long long *)tfirst, ft);
    LAUNCH(ctx, k_cov_verify, grid_for(n_rec, 256), 256, n_rec, d_text, R, (const uint64_t *)ft, err_key);
    LAUNCH(ctx, k_cov_heads, grid_for(n_rec + 1, 256), 256, n_rec, (const uint64_t *)ft, head);
    WGA_TRY(exclusive_sum(ctx, head, tidx, n_rec + 1));
    WGA_CK(cudaMemcpyAsync(ctx->h_scal, tidx + n_rec, 4, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_TRY(sync_check(ctx));
    const uint64_t n_targets = *(uint32_t *)ctx->h_scal;
    uint64_t *t_rec = ctx->buf<uint64_t>(B_PP_F, n_targets + 1), *t_len1 = ctx->buf<uint64_t>(B_PP_G, n_targets + 1);
    uint64_t *t_base = ctx->buf<uint64_t>(B_PP_H, n_targets + 1), *t_ntile = ctx->buf<uint64_t>(B_PP_I, n_targets + 1);
    uint64_t *t_tile0 = ctx->buf<uint64_t>(B_PP_L, n_targets + 1);
    if (!t_rec || !t_len1 || !t_base || !t_ntile || !t_tile0) return ctx->err_code;
    WGA_CK(cudaMemsetAsync(t_len1 + n_targets, 0, 8, ctx->stream));
    LAUNCH(ctx, k_cov_targets, grid_for(n_rec, 256), 256, n_rec, (const uint32_t *)head, (const uint32_t *)tidx, R, t_rec, t_len1);
    WGA_TRY(exclusive_sum(ctx, t_len1, t_base, n_targets + 1));
    LAUNCH(ctx, k_cov_tiles, grid_for(n_targets + 1, 256), 256, n_targets, (const uint64_t *)t_len1, t_ntile);
    WGA_TRY(exclusive_sum(ctx, t_ntile, t_tile0, n_targets + 1));
    uint64_t sc[3];
    WGA_TRY(read_scalars(ctx, {t_base + n_targets, t_tile0 + n_targets, err_key}, sc));
    const uint64_t n_slots = sc[0], n_tiles = sc[1];
    // ---- difference array -> coverage
    int *diff = ctx->buf<int>(B_PP_M, n_slots + 1), *cov = ctx->buf<int>(B_PP_N, n_slots + 1);
    if (!diff || !cov) return ctx->err_code;
    WGA_CK(cudaMemsetAsync(diff, 0, (n_slots + 1) * 4, ctx->stream));
    LAUNCH(ctx, k_cov_ops, grid_for(n_rec, 128), 128, n_rec, d_text, R, (const uint64_t *)ft, (const uint32_t *)tidx, (const uint64_t *)t_base,
           (const uint64_t *)t_len1, diff, err_key);
    {
        size_t tmp = 0;
        WGA_CK(cub::DeviceScan::InclusiveSum(nullptr, tmp, diff, cov, (int64_t)n_slots, ctx->stream));
        void *t = ctx->ensure(B_CUBTMP, tmp);
        if (!t) return ctx->err_code;
        WGA_CK(cub::DeviceScan::InclusiveSum(t, tmp, diff, cov, (int64_t)n_slots, ctx->stream));
        ctx->n_cub++;
    }
    // ---- BED text, tile by tile
    uint64_t *tile_len = ctx->buf<uint64_t>(B_PP_O, n_tiles + 1), *tile_off = ctx->buf<uint64_t>(B_PP_P, n_tiles + 1);
    if (!tile_len || !tile_off) return ctx->err_code;
    CovView V;
    V.n_targets = n_targets; V.n_tiles = n_tiles; V.text = d_text; V.R = R; V.t_rec = t_rec; V.t_base = t_base; V.t_len1 = t_len1;
    V.t_tile0 = t_tile0; V.cov = cov; V.tile_len = tile_len; V.tile_off = tile_off; V.out = nullptr;
    LAUNCH(ctx, k_cov_rows<0>, grid_for(n_tiles + 1, 128), 128, V);
    WGA_TRY(exclusive_sum(ctx, tile_len, tile_off, n_tiles + 1));
    WGA_TRY(read_scalars(ctx, {tile_off + n_tiles, err_key}, sc));
    if (sc[1] != kNoErr) {
        const int code = (int)(sc[1] & 0xFF);
        const uint64_t rec = sc[1] >> 32;
        std::string msg;
        if (code == WGA_E_CSV) msg = "CSV deserialize error by: record " + std::to_string(rec) + " does not match the 12 typed PAF columns";
        else if (code == WGA_E_CIGAR_TAG_NOT_FOUND) msg = "CIGAR start tag not found";
        else if (code == WGA_E_UNSUPPORTED) msg = "unsupported PAF construct (quoted csv field) in record " + std::to_string(rec);
        else if (code == WGA_E_REF_PANIC) msg = "the reference panics on this record (empty CIGAR, or a cs:Z: length beyond u64)";
        else if (code == WGA_E_INTERNAL) msg = "name hash collision";
        else {
            uint64_t oc[2];
            WGA_TRY(read_scalars(ctx, {R.cig_off + rec, R.cig_len + rec}, oc));
            std::string cg((size_t)std::min<uint64_t>(oc[1], 1u << 20), '\0');
            if (!cg.empty()) WGA_CK(cudaMemcpy(&cg[0], d_text + oc[0], cg.size(), cudaMemcpyDeviceToHost));
            msg = paf_cigar_error_text(cg, code);
        }
        ctx->set_error(code, rec, msg);
        return code;
    }
    const uint64_t total = sc[0];
    if (total == 0) return WGA_OK;
    char *d_out = (char *)ctx->ensure(B_OUT, total + 16);
    if (!d_out) return ctx->err_code;
    V.out = d_out;
    LAUNCH(ctx, k_cov_rows<1>, grid_for(n_tiles, 128), 128, V);
    return finish_output(ctx, d_out, total, mem, flags, out);
}
int wga_paf_cov(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, wga_span *out) {
    const int rc = wga_paf_cov_impl(ctx, text, len, mem, flags, out);
    return paf_retry_unquoted(ctx, rc, mem, flags, [&](const void *t2, uint64_t n2, int m2, uint32_t f2) {
        return wga_paf_cov_impl(ctx, t2, n2, m2, f2, out);
    });
}

// ---------------------------------------------------------------- validate (validate.rs:44-147)
static int wga_paf_validate_impl(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, int fix, wga_span *report,
                     wga_span *fixed) {
    if (!ctx || !report || !fixed || (!text && len)) return WGA_E_INVALID_ARG;
    cudaSetDevice(ctx->device);
    report->data = nullptr; report->len = 0;
    fixed->data = nullptr; fixed->len = 0;
    walk_begin(ctx);
    WGA_TRY(stage_text(ctx, text, len, mem));
    PafPrep PP;
    WGA_TRY(paf_fasta_prepare(ctx, len, nullptr, nullptr, 1, &PP, false, nullptr));
    const uint64_t n_rec = PP.n_rec;
    std::string &rep = ctx->host_text;  // host text owned by the ctx
    rep.clear();
    uint64_t nbad[2] = {0, 0};
    std::string qlist, tlist;
    uint64_t total = 0;
    char *d_rows = nullptr;
    if (n_rec) {
        const uint8_t *d_text = ctx->d_text;
        PafRecs R = PP.R;
        unsigned long long *err_key = PP.err_key;
        unsigned long long *cnt11 = ctx->buf<unsigned long long>(B_GEN_C, n_rec * 11);
        uint8_t *slow = ctx->buf<uint8_t>(B_RUN_CLS, n_rec);
        uint64_t *c7 = ctx->buf<uint64_t>(B_GEN_D, 7 * n_rec);
        uint64_t *l3 = ctx->buf<uint64_t>(B_GEN_E, 6 * (n_rec + 1));
        unsigned long long *d_nbad = ctx->buf<unsigned long long>(B_TOK_ERR, 4) + 1;
        if (!cnt11 || !slow || !c7 || !l3 || !d_nbad) return ctx->err_code;
        ValCols C;
        C.q_end = c7; C.t_end = c7 + n_rec; C.matches = c7 + 2 * n_rec; C.block_len = c7 + 3 * n_rec; C.mapq = c7 + 4 * n_rec;
        C.tags_off = c7 + 5 * n_rec; C.line_end = c7 + 6 * n_rec;
        // the line table of paf_fasta_prepare is still in the pool
        uint8_t *lflags = ctx->buf<uint8_t>(B_LINE_FLAGS, PP.lv.n_lines + 1);
        uint32_t *rscan = ctx->buf<uint32_t>(B_LINE_REC, PP.lv.n_lines + 1);
        LAUNCH(ctx, k_val_cols, grid_for(PP.lv.n_lines, 256), 256, PP.lv, (const uint8_t *)lflags, (const uint32_t *)rscan, C);
        WGA_TRY(paf_cigar_counts(ctx, n_rec, len, d_text, R, cnt11, slow, err_key));
        WGA_CK(cudaMemsetAsync(d_nbad, 0, 16, ctx->stream));
        ValView V;
        V.n_rec = n_rec; V.text = d_text; V.R = R; V.C = C; V.cnt11 = cnt11; V.fix = fix ? 1 : 0;
        V.q_len = l3; V.t_len = l3 + (n_rec + 1); V.row_len = l3 + 2 * (n_rec + 1);
        uint64_t *q_off = l3 + 3 * (n_rec + 1), *t_off = l3 + 4 * (n_rec + 1), *row_off = l3 + 5 * (n_rec + 1);
        V.q_off = q_off; V.t_off = t_off; V.row_off = row_off; V.q_out = V.t_out = V.row_out = nullptr; V.n_bad = d_nbad;
        LAUNCH(ctx, k_val_rows_warp<0>, grid_for((n_rec + 1) * 32, 128), 128, V);
        WGA_TRY(exclusive_sum(ctx, V.q_len, q_off, n_rec + 1));
        WGA_TRY(exclusive_sum(ctx, V.t_len, t_off, n_rec + 1));
        WGA_TRY(exclusive_sum(ctx, V.row_len, row_off, n_rec + 1));
        uint64_t sc[6];
        WGA_TRY(read_scalars(ctx, {q_off + n_rec, t_off + n_rec, row_off + n_rec, err_key, d_nbad, d_nbad + 1}, sc));
        if (sc[3] != kNoErr) {
            int code = (int)(sc[3] & 0xFF);
            const uint64_t rec = sc[3] >> 32;
            std::string msg;
            if (code == WGA_E_CSV) msg = "CSV deserialize error by: record " + std::to_string(rec) + " does not match the 12 typed PAF columns";
            else if (code == WGA_E_UNSUPPORTED) msg = "unsupported PAF construct (quoted csv field) in record " + std::to_string(rec);
            else { code = WGA_E_REF_PANIC; msg = "the reference panics on this record (get_stat().unwrap() on a CIGAR error, validate.rs:79)"; }
            ctx->set_error(code, rec, msg);
            return code;
        }
        nbad[0] = sc[4]; nbad[1] = sc[5];
        total = sc[2];
        char *d_q = (char *)ctx->ensure(B_NAME_BUF, sc[0] + sc[1] + 16);
        d_rows = (char *)ctx->ensure(B_OUT, total + 16);
        if (!d_q || !d_rows) return ctx->err_code;
        V.q_out = d_q; V.t_out = d_q + sc[0]; V.row_out = d_rows;
        LAUNCH(ctx, k_val_rows_warp<1>, grid_for((n_rec + 1) * 32, 128), 128, V);
        qlist.resize(sc[0]); tlist.resize(sc[1]);
        if (sc[0]) WGA_CK(cudaMemcpyAsync(&qlist[0], d_q, sc[0], cudaMemcpyDeviceToHost, ctx->stream));
        if (sc[1]) WGA_CK(cudaMemcpyAsync(&tlist[0], d_q + sc[0], sc[1], cudaMemcpyDeviceToHost, ctx->stream));
        WGA_TRY(sync_check(ctx));
    }
    // impl Display for Validations (validate.rs:25-41) + writeln!(writer, "{}", validations)
    rep += "Total records: " + std::to_string(n_rec) + "\nQuery invalid records: " + std::to_string(nbad[0]) +
           "\nTarget invalid records: " + std::to_string(nbad[1]) + "\nQuery invalid list:\n" + qlist + "Target invalid list:\n" + tlist + "\n";
    report->data = rep.data();
    report->len = rep.size();
    if (fix && total) return finish_output(ctx, d_rows, total, mem, flags, fixed);
    return WGA_OK;
}
int wga_paf_validate(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, int fix, wga_span *report, wga_span *fixed) {
    const int rc = wga_paf_validate_impl(ctx, text, len, mem, flags, fix, report, fixed);
    return paf_retry_unquoted(ctx, rc, mem, flags, [&](const void *t2, uint64_t n2, int m2, uint32_t f2) {
        return wga_paf_validate_impl(ctx, t2, n2, m2, f2, fix, report, fixed);
    });
}

// ---------------------------------------------------------------- filter -f paf (filter.rs:41-63,92-160)
static int wga_paf_filter_impl(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, uint64_t min_block_size,
                   uint64_t min_query_size, const uint64_t *min_align_size, wga_span *out) {
    if (!ctx || !out || (!text && len)) return WGA_E_INVALID_ARG;
    cudaSetDevice(ctx->device);
    out->data = nullptr; out->len = 0;
    walk_begin(ctx);
    WGA_TRY(stage_text(ctx, text, len, mem));
    unsigned long long *tag_key = ctx->buf<unsigned long long>(B_PP_T, 1);  // filter never looks at the CIGAR: not an error here
    if (!tag_key) return ctx->err_code;
    PafPrep PP;
    WGA_TRY(paf_fasta_prepare(ctx, len, nullptr, nullptr, 1, &PP, false, tag_key));
    const uint64_t n_rec = PP.n_rec;
    if (n_rec == 0) return WGA_OK;
    const uint8_t *d_text = ctx->d_text;
    PafRecs R = PP.R;
    unsigned long long *err_key = PP.err_key;
    uint64_t *c7 = ctx->buf<uint64_t>(B_GEN_D, 7 * n_rec);
    uint64_t *l2 = ctx->buf<uint64_t>(B_GEN_E, 2 * (n_rec + 1));
    uint8_t *keep = ctx->buf<uint8_t>(B_RUN_CLS, n_rec);
    if (!c7 || !l2 || !keep) return ctx->err_code;
    ValCols C;
    C.q_end = c7; C.t_end = c7 + n_rec; C.matches = c7 + 2 * n_rec; C.block_len = c7 + 3 * n_rec; C.mapq = c7 + 4 * n_rec;
    C.tags_off = c7 + 5 * n_rec; C.line_end = c7 + 6 * n_rec;
    uint8_t *lflags = ctx->buf<uint8_t>(B_LINE_FLAGS, PP.lv.n_lines + 1);  // the line table of paf_fasta_prepare
    uint32_t *rscan = ctx->buf<uint32_t>(B_LINE_REC, PP.lv.n_lines + 1);
    LAUNCH(ctx, k_val_cols, grid_for(PP.lv.n_lines, 256), 256, PP.lv, (const uint8_t *)lflags, (const uint32_t *)rscan, C);
    if (min_align_size) {
        unsigned long long *cnt11 = ctx->buf<unsigned long long>(B_GEN_C, n_rec * 11);
        uint64_t *zeros = ctx->buf<uint64_t>(B_GEN_F, n_rec);
        uint64_t cap = 1024;
        while (cap < 2 * n_rec) cap <<= 1;
        unsigned long long *keys = ctx->buf<unsigned long long>(B_PT_KEYS, cap);
        unsigned long long *first = ctx->buf<unsigned long long>(B_PT_FIRST, cap);
        unsigned long long *vals = ctx->buf<unsigned long long>(B_PT_VALS, cap * kPairVals);
        uint64_t *rec_slot = ctx->buf<uint64_t>(B_REC_SLOT, n_rec);
        if (!cnt11 || !zeros || !keys || !first || !vals || !rec_slot) return ctx->err_code;
        WGA_CK(cudaMemsetAsync(zeros, 0, n_rec * 8, ctx->stream));
        WGA_CK(cudaMemsetAsync(keys, 0, cap * 8, ctx->stream));
        WGA_CK(cudaMemsetAsync(first, 0xFF, cap * 8, ctx->stream));
        LAUNCH(ctx, k_pt_init, grid_for(cap * kPairVals, 256), 256, vals, cap);
        LAUNCH(ctx, k_flt_contrib, grid_for(n_rec, 256), 256, n_rec, R, (const uint64_t *)C.t_end, cnt11);
        // the key of filter.rs:125 is (query_name, target_name): the sizes of stat's Pair key are left out (all zero)
        RecMeta M;
        M.text = d_text; M.t_name_off = R.t_name_off; M.q_name_off = R.q_name_off; M.t_name_len = R.t_name_len; M.q_name_len = R.q_name_len;
        M.t_size = zeros; M.q_size = zeros; M.t_start = zeros; M.q_start = zeros; M.neg = R.neg; M.skip = nullptr;
        LAUNCH(ctx, k_pair_insert, grid_for(n_rec, 256), 256, n_rec, M, (const unsigned long long *)cnt11, cap - 1, keys, first, vals, rec_slot);
        LAUNCH(ctx, k_pair_verify, grid_for(n_rec, 256), 256, n_rec, M, (const unsigned long long *)first, (const uint64_t *)rec_slot, err_key);
        LAUNCH(ctx, k_flt_keep_pair, grid_for(n_rec, 256), 256, n_rec, (const uint64_t *)rec_slot, (const unsigned long long *)vals,
               *min_align_size, keep);
    } else {
        LAUNCH(ctx, k_flt_keep, grid_for(n_rec, 256), 256, n_rec, R, (const uint64_t *)C.t_end, min_block_size, min_query_size, keep);
    }
    FltView V;
    V.n_rec = n_rec; V.text = d_text; V.R = R; V.C = C; V.keep = keep; V.row_len = l2; V.row_off = l2 + (n_rec + 1); V.out = nullptr;
    LAUNCH(ctx, k_flt_rows_warp<0>, grid_for((n_rec + 1) * 32, 128), 128, V);
    WGA_TRY(exclusive_sum(ctx, V.row_len, l2 + (n_rec + 1), n_rec + 1));
    uint64_t sc[2];
    WGA_TRY(read_scalars(ctx, {l2 + (n_rec + 1) + n_rec, err_key}, sc));
    if (sc[1] != kNoErr) {
        const int code = (int)(sc[1] & 0xFF);
        const uint64_t rec = sc[1] >> 32;
        std::string msg;
        if (code == WGA_E_CSV) msg = "CSV deserialize error by: record " + std::to_string(rec) + " does not match the 12 typed PAF columns";
        else if (code == WGA_E_UNSUPPORTED) msg = "unsupported PAF construct (quoted csv field) in record " + std::to_string(rec);
        else msg = "internal: pair key collision";
        ctx->set_error(code, rec, msg);
        return code;
    }
    const uint64_t total = sc[0];
    if (total == 0) return WGA_OK;
    char *d_rows = (char *)ctx->ensure(B_OUT, total + 16);
    if (!d_rows) return ctx->err_code;
    V.out = d_rows;
    LAUNCH(ctx, k_flt_rows_warp<1>, grid_for((n_rec + 1) * 32, 128), 128, V);
    return finish_output(ctx, d_rows, total, mem, flags, out);
}
int wga_paf_filter(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, uint64_t min_block_size, uint64_t min_query_size, const uint64_t *min_align_size, wga_span *out) {
    const int rc = wga_paf_filter_impl(ctx, text, len, mem, flags, min_block_size, min_query_size, min_align_size, out);
    return paf_retry_unquoted(ctx, rc, mem, flags, [&](const void *t2, uint64_t n2, int m2, uint32_t f2) {
        return wga_paf_filter_impl(ctx, t2, n2, m2, f2, min_block_size, min_query_size, min_align_size, out);
    });
}

// ---------------------------------------------------------------- pafpseudo (pseudomaf.rs:18-237, cigar.rs:744-804)
static int wga_paf_pseudo_impl(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, const wga_fasta *FA, const char *target,
                   wga_span *out, wga_span *index) {
    if (!ctx || !out || !index || (!text && len)) return WGA_E_INVALID_ARG;
    cudaSetDevice(ctx->device);
    out->data = nullptr; out->len = 0;
    index->data = nullptr; index->len = 0;
    ctx->host_text.clear();
    walk_begin(ctx);
    WGA_TRY(stage_text(ctx, text, len, mem));
    unsigned long long *tag_key = ctx->buf<unsigned long long>(B_PP_T, 1);
    if (!tag_key) return ctx->err_code;
    PafPrep PP;
    WGA_TRY(paf_fasta_prepare(ctx, len, nullptr, nullptr, 1, &PP, false, tag_key));
    const uint64_t n_rec = PP.n_rec;
    if (n_rec == 0) return WGA_OK;
    const uint8_t *d_text = ctx->d_text;
    PafRecs R = PP.R;
    P2mRecs X = PP.X;
    unsigned long long *err_key = PP.err_key;
    const int base = FA != nullptr;
    uint64_t key;
    WGA_TRY(read_scalars(ctx, {err_key}, &key));  // csv-level errors come first: the reference reads every record up front
    auto fail = [&](uint64_t k) {
        const int code = (int)(k & 0xFF);
        const uint64_t rec = k >> 32;
        std::string msg;
        if (code == WGA_E_CSV) msg = "CSV deserialize error by: record " + std::to_string(rec) + " does not match the 12 typed PAF columns";
        else if (code == WGA_E_NOM) msg = "HTS library error by sequence not found in the FASTA index";
        else if (code == WGA_E_CIGAR_TAG_NOT_FOUND) msg = "CIGAR start tag not found";
        else if (code == WGA_E_UNSUPPORTED) msg = "unsupported PAF construct (quoted csv field)";
        else if (code == WGA_E_VCF_BASE) msg = "Invalid Base in a '-' strand query";
        else if (code == WGA_E_CIGAR_OP_INVALID) msg = "CIGAR OP invalid";
        else if (code == WGA_E_PARSE_INT) msg = "Parse Into Integer Error";
        else if (code == WGA_E_INTERNAL) msg = "name hash collision";
        else msg = "the reference panics on this input (drain / insert_str / slice past the end, or an underflowing subtraction)";
        ctx->set_error(code, rec, msg);
        return code;
    };
    if (key != kNoErr) return fail(key);

    // ---- group keys, first-occurrence ranks, sorted order
    uint64_t *kt = ctx->buf<uint64_t>(B_PP_A, n_rec), *kp = ctx->buf<uint64_t>(B_PP_B, n_rec);
    uint64_t *ft = ctx->buf<uint64_t>(B_PP_C, n_rec), *fp = ctx->buf<uint64_t>(B_PP_D, n_rec);
    uint64_t *slot_of = ctx->buf<uint64_t>(B_PP_E, n_rec);
    uint64_t cap = 64;
    while (cap < 2 * n_rec) cap <<= 1;
    unsigned long long *tkeys = ctx->buf<unsigned long long>(B_PT_KEYS, cap), *tfirst = ctx->buf<unsigned long long>(B_PT_FIRST, cap);
    uint8_t *d_filter = nullptr;
    const uint32_t flen = target ? (uint32_t)strlen(target) : 0;
    if (target) {
        d_filter = (uint8_t *)ctx->ensure(B_QNAME, flen + 1);
        if (!d_filter) return ctx->err_code;
        WGA_CK(cudaMemcpyAsync(d_filter, target, flen, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!kt || !kp || !ft || !fp || !slot_of || !tkeys || !tfirst) return ctx->err_code;
    LAUNCH(ctx, k_pp_keys, grid_for(n_rec, 256), 256, n_rec, d_text, R, (const uint8_t *)d_filter, flen, target ? 1 : 0, kt, kp);
    for (int pass = 0; pass < 2; ++pass) {
        WGA_CK(cudaMemsetAsync(tkeys, 0, cap * 8, ctx->stream));
        WGA_CK(cudaMemsetAsync(tfirst, 0xFF, cap * 8, ctx->stream));
        LAUNCH(ctx, k_pp_first_insert, grid_for(n_rec, 256), 256, n_rec, (const uint64_t *)(pass ? kp : kt), cap - 1, tkeys, tfirst, slot_of);
        LAUNCH(ctx, k_pp_first_read, grid_for(n_rec, 256), 256, n_rec, (const uint64_t *)slot_of, (const unsigned long long *)tfirst,
               pass ? fp : ft);
    }
    LAUNCH(ctx, k_pp_verify, grid_for(n_rec, 256), 256, n_rec, d_text, R, (const uint64_t *)ft, (const uint64_t *)fp, err_key);
    uint64_t *ka = ctx->buf<uint64_t>(B_PP_F, n_rec), *kb = ctx->buf<uint64_t>(B_PP_G, n_rec);
    uint64_t *va = ctx->buf<uint64_t>(B_PP_H, n_rec), *vb = ctx->buf<uint64_t>(B_PP_I, n_rec);
    if (!ka || !kb || !va || !vb) return ctx->err_code;
    LAUNCH(ctx, k_pp_iota, grid_for(n_rec, 256), 256, n_rec, va);
    WGA_TRY(sort_pairs_u64(ctx, R.t_start, kb, va, vb, n_rec));                       // by target_start
    LAUNCH(ctx, k_pp_gather_u64, grid_for(n_rec, 256), 256, n_rec, (const uint64_t *)fp, (const uint64_t *)vb, ka);
    WGA_TRY(sort_pairs_u64(ctx, ka, kb, vb, va, n_rec));                              // then by query group
    LAUNCH(ctx, k_pp_gather_u64, grid_for(n_rec, 256), 256, n_rec, (const uint64_t *)ft, (const uint64_t *)va, ka);
    WGA_TRY(sort_pairs_u64(ctx, ka, kb, va, vb, n_rec));                              // then by target
    const uint64_t *ord = vb;   // sorted position -> file-order record; dropped records (key ~0) are last
    // number of used records = first sorted position whose target key is "dropped"
    uint64_t n = n_rec;
    if (target) {
        uint32_t *usedf = ctx->buf<uint32_t>(B_PP_J, n_rec + 1), *useds = ctx->buf<uint32_t>(B_PP_K, n_rec + 1);
        if (!usedf || !useds) return ctx->err_code;
        LAUNCH(ctx, k_pp_used, grid_for(n_rec + 1, 256), 256, n_rec, (const uint64_t *)ft, usedf);
        WGA_TRY(exclusive_sum(ctx, usedf, useds, n_rec + 1));
        WGA_CK(cudaMemcpyAsync(ctx->h_scal, useds + n_rec, 4, cudaMemcpyDeviceToHost, ctx->stream));
        WGA_TRY(sync_check(ctx));
        n = *(uint32_t *)ctx->h_scal;
    }
    if (n == 0) return WGA_OK;

    // ---- group structure + the sequential last_target_end state
    PpRecs S;
    S.rec = const_cast<uint64_t *>(ord);
    S.flags = ctx->buf<uint32_t>(B_PP_J, n + 1);
    uint32_t *ghead = ctx->buf<uint32_t>(B_PP_K, n + 1), *gid = ctx->buf<uint32_t>(B_PP_L, n + 1);
    uint64_t *p8 = ctx->buf<uint64_t>(B_PP_M, 10 * (n + 1));
    if (!S.flags || !ghead || !gid || !p8) return ctx->err_code;
    S.gap = p8; S.overlap = p8 + (n + 1); S.tail = p8 + 2 * (n + 1); S.q_src = p8 + 3 * (n + 1); S.LQ = p8 + 4 * (n + 1);
    S.P = p8 + 5 * (n + 1); S.seg_len = p8 + 6 * (n + 1); S.seg_off = p8 + 7 * (n + 1);
    uint64_t *t_src = p8 + 8 * (n + 1), *t_cnt = p8 + 9 * (n + 1);
    unsigned long long *d_nties = ctx->buf<unsigned long long>(B_CS_TOTAL, 1);
    if (!d_nties) return ctx->err_code;
    LAUNCH(ctx, k_init_u64, 1, 1, d_nties, 1, 0ull);
    LAUNCH(ctx, k_pp_flags, grid_for(n + 1, 256), 256, n, ord, (const uint64_t *)ft, (const uint64_t *)fp, (const uint64_t *)R.t_start, S.flags,
           ghead, d_nties);
    WGA_TRY(exclusive_sum(ctx, ghead, gid, n + 1));
    WGA_CK(cudaMemcpyAsync(ctx->h_scal, gid + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_CK(cudaMemcpyAsync(ctx->h_scal + 1, d_nties, 8, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_TRY(sync_check(ctx));
    const uint64_t n_groups = *(uint32_t *)ctx->h_scal, n_ties = ctx->h_scal[1];
    uint64_t *gstart = ctx->buf<uint64_t>(B_PP_N, n_groups + 1);
    if (!gstart) return ctx->err_code;
    LAUNCH(ctx, k_pp_group_starts, grid_for(n + 1, 256), 256, n, (const uint32_t *)ghead, (const uint32_t *)gid, gstart, n_groups);
    if (n_ties) {  // equal target_start inside a (target, query) group: the order of pseudomaf.rs:91-94
        LAUNCH(ctx, k_pp_ties, grid_for(n, 128), 128, n, ord, (const uint64_t *)fp, (const uint64_t *)R.t_start, (const uint32_t *)gid,
               (const uint64_t *)gstart, va);
        ord = va;
        S.rec = va;
    }
    LAUNCH(ctx, k_pp_groups, grid_for(n_groups, 128), 128, n_groups, (const uint64_t *)gstart, ord, R, X, S, err_key);
    // target heads
    uint32_t *tflag = ghead, *tid = gid;  // reuse
    LAUNCH(ctx, k_pp_thead_flag, grid_for(n + 1, 256), 256, n, (const uint32_t *)S.flags, tflag);
    WGA_TRY(exclusive_sum(ctx, tflag, tid, n + 1));
    WGA_CK(cudaMemcpyAsync(ctx->h_scal, tid + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_TRY(sync_check(ctx));
    const uint64_t n_targets = *(uint32_t *)ctx->h_scal;
    uint64_t *thead_pos = ctx->buf<uint64_t>(B_PP_O, n_targets + 1), *finfo = ctx->buf<uint64_t>(B_PP_P, 4 * n_targets + 4);
    if (!thead_pos || !finfo) return ctx->err_code;
    LAUNCH(ctx, k_pp_thead_pos, grid_for(n, 256), 256, n, (const uint32_t *)S.flags, (const uint32_t *)tid, thead_pos);

    // ---- FASTA spans, CIGAR ops, layout
    FaView F{nullptr, nullptr, nullptr, 0};
    if (FA) F = FaView{FA->d_seq, FA->d_names, FA->d_meta, FA->n_seq};
    LAUNCH(ctx, k_pp_spans, grid_for(n, 256), 256, n, d_text, ord, R, X, S, F, base, t_src, t_cnt, err_key);
    if (base) LAUNCH(ctx, k_pp_revcomp_check, (unsigned)std::min<uint64_t>(n, 0x7FFFFFFFull), 128, n, ord, R, S, (const uint8_t *)FA->d_seq, err_key);
    uint64_t *n_ops = ctx->buf<uint64_t>(B_PP_Q, n + 1), *op_base = ctx->buf<uint64_t>(B_PP_R, n + 1);
    if (!n_ops || !op_base) return ctx->err_code;
    LAUNCH(ctx, k_pp_ops<0>, grid_for(n + 1, 128), 128, n, d_text, ord, R, S, n_ops, (const uint64_t *)nullptr, (uint64_t *)nullptr,
           (uint8_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, err_key);
    WGA_TRY(exclusive_sum(ctx, n_ops, op_base, n + 1));
    uint64_t sc[2];
    WGA_TRY(read_scalars(ctx, {op_base + n, err_key}, sc));
    const uint64_t n_opsT = sc[0];
    uint64_t *op_len = ctx->buf<uint64_t>(B_RUN_POS, n_opsT + 1), *c_out = ctx->buf<uint64_t>(B_GEN_C, n_opsT + 1);
    uint64_t *c_src = ctx->buf<uint64_t>(B_GEN_D, n_opsT + 1), *S_out = ctx->buf<uint64_t>(B_GEN_E, n_opsT + 1);
    uint64_t *S_src = ctx->buf<uint64_t>(B_GEN_F, n_opsT + 1);
    uint8_t *op_type = ctx->buf<uint8_t>(B_RUN_CLS, n_opsT + 1);
    uint32_t *op_rec = ctx->buf<uint32_t>(B_RUN_REC, n_opsT + 1), *op_pos = ctx->buf<uint32_t>(B_RUN_DL, n_opsT + 1);
    if (!op_len || !c_out || !c_src || !S_out || !S_src || !op_type || !op_rec || !op_pos) return ctx->err_code;
    LAUNCH(ctx, k_pp_ops<1>, grid_for(n + 1, 128), 128, n, d_text, ord, R, S, n_ops, (const uint64_t *)op_base, op_len, op_type, op_rec, op_pos,
           err_key);
    LAUNCH(ctx, k_pp_contrib, grid_for(n_opsT + 1, 256), 256, n_opsT, (const uint64_t *)op_len, (const uint8_t *)op_type, c_out, c_src);
    WGA_TRY(exclusive_sum(ctx, c_out, S_out, n_opsT + 1));
    WGA_TRY(exclusive_sum(ctx, c_src, S_src, n_opsT + 1));
    PpView V;
    V.n = n; V.n_ops = n_opsT; V.text = d_text; V.ord = ord; V.R = R; V.X = X; V.S = S; V.op_base = op_base; V.op_len = op_len;
    V.op_type = op_type; V.op_rec = op_rec; V.op_pos = op_pos; V.S_out = S_out; V.S_src = S_src; V.seq = FA ? FA->d_seq : nullptr;
    V.t_src = t_src; V.t_cnt = t_cnt; V.base = base; V.out = nullptr;
    if (n_opsT) LAUNCH(ctx, k_pp_check, grid_for(n_opsT, 256), 256, V, err_key);
    LAUNCH(ctx, k_pp_layout, grid_for(n + 1, 256), 256, V, err_key);
    WGA_TRY(exclusive_sum(ctx, S.seg_len, S.seg_off, n + 1));
    WGA_TRY(read_scalars(ctx, {S.seg_off + n, err_key}, sc));
    if (sc[1] != kNoErr) return fail(sc[1]);
    const uint64_t total = sc[0];
    char *d_out = (char *)ctx->ensure(B_OUT, total + 16);
    if (!d_out) return ctx->err_code;
    V.out = d_out;
    WGA_CK(cudaEventRecord(ctx->ev0, ctx->stream));
    WGA_CK(cudaMemsetAsync(d_out, '-', total, ctx->stream));
    LAUNCH(ctx, k_pp_fixed, grid_for(n, 256), 256, V);
    {
        const uint64_t per = 256ull * 64;
        uint64_t gx = 1;
        // enough blocks per target for the longest sequence; grid-stride covers the rest
        gx = std::min<uint64_t>(2048, std::max<uint64_t>(1, (total / std::max<uint64_t>(n_targets, 1) + per - 1) / per));
        dim3 grid((unsigned)gx, (unsigned)std::min<uint64_t>(n_targets, 65535));
        if (n_targets > 65535) { ctx->set_error(WGA_E_UNSUPPORTED, 0, "more than 65535 targets in one call"); return WGA_E_UNSUPPORTED; }
        k_pp_target_rows<<<grid, 256, 0, ctx->stream>>>(V, (const uint64_t *)thead_pos, n_targets);
        ctx->n_own++;
    }
    LAUNCH(ctx, k_pp_rows, grid_for((n_opsT + n) * 32, 256), 256, V);
    WGA_CK(cudaEventRecord(ctx->ev1, ctx->stream));
    LAUNCH(ctx, k_pp_files, grid_for(n_targets, 128), 128, V, (const uint64_t *)thead_pos, n_targets, total, finfo);
    // index: "{target}\t{bytes}\n" per file, built on the host from the (tiny) per-file table
    std::vector<uint64_t> info(4 * n_targets);
    WGA_CK(cudaMemcpyAsync(info.data(), finfo, info.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    WGA_TRY(sync_check(ctx));
    std::string &idx = ctx->host_text;
    for (uint64_t t = 0; t < n_targets; ++t) {
        std::string name((size_t)info[4 * t + 1], '\0');
        if (!name.empty()) WGA_CK(cudaMemcpy(&name[0], d_text + info[4 * t], name.size(), cudaMemcpyDeviceToHost));
        idx += name;
        idx += '\t';
        idx += std::to_string(info[4 * t + 3]);
        idx += '\n';
    }
    index->data = idx.data();
    index->len = idx.size();
    int rc = finish_output(ctx, d_out, total, mem, flags, out);
    if (rc) return rc;
    float ms = 0.f;
    WGA_CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->walk_ms = ms; ctx->walk_launches = 1; ctx->walk_bytes = total;  // expansion: output-bound
    return WGA_OK;
}
int wga_paf_pseudo(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, const wga_fasta *FA, const char *target, wga_span *out, wga_span *index) {
    const int rc = wga_paf_pseudo_impl(ctx, text, len, mem, flags, FA, target, out, index);
    return paf_retry_unquoted(ctx, rc, mem, flags, [&](const void *t2, uint64_t n2, int m2, uint32_t f2) {
        return wga_paf_pseudo_impl(ctx, t2, n2, m2, f2, FA, target, out, index);
    });
}


}  // extern "C"
