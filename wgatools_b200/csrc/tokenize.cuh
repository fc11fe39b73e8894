// tokenize.cuh -- on-device tokenisation of raw MAF text.
//
// Replaces the reference's serial reader (src/parser/maf.rs:25-36 MAFReader::new, :371-421
// MAFRecords::next, :138-211 parse_sline) with index-building kernels:
//   1. whitespace index: position of every ASCII whitespace byte (newlines flagged)  -- the only pass
//      that touches every byte of the text besides the column walk itself;
//   2. lines, s-line flags, record heads (maximal runs of consecutive lines starting with 's');
//   3. per-s-line field parse from the whitespace index (split_whitespace semantics);
//   4. per-record (target line, query line) selection and the work-item table of the column walk.
#pragma once
#include "ctx.cuh"

namespace wga {

constexpr int kWsTile = 32768;    // bytes per CTA in the whitespace-index kernels
constexpr int kWsThreads = 256;

// ---------------------------------------------------------------- 1. whitespace index
// CLS 0: ASCII whitespace (MAF, split_whitespace), lines end at '\n'; CLS 1: PAF delimiters (csv reader, paf.rs:22-31):
// '\t' between fields, and the csv crate's default Terminator::CRLF ends a record at '\r', '\n' or "\r\n" -- both bytes
// are line ends here ("\r\n" = a line end plus an empty line, and empty lines are skipped)
template <int CLS>
__device__ __forceinline__ uint32_t line_end_bytes(uint32_t w) {
    return CLS == 0 ? nl_bytes(w) : (zero_bytes(w ^ 0x0A0A0A0Au) | zero_bytes(w ^ 0x0D0D0D0Du));
}
template <int CLS>
__device__ __forceinline__ uint32_t delim_bytes(uint32_t w) {
    return CLS == 0 ? ws_bytes(w) : (zero_bytes(w ^ 0x09090909u) | line_end_bytes<1>(w));
}
template <int CLS>
__device__ __forceinline__ void ws_flags16(const uint8_t *text, uint64_t off, uint64_t len, uint32_t ws[4],
                                           uint32_t nl[4]) {
    uint4 v = ldg16(text + off);
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    int64_t rem = (int64_t)(len - off);  // > 0
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t valid = rem >= 4 * (k + 1) ? kHi : range_flags(0, (int)(rem - 4 * k));
        ws[k] = delim_bytes<CLS>(w[k]) & valid;
        nl[k] = line_end_bytes<CLS>(w[k]) & valid;
    }
}

template <int CLS>
__global__ void __launch_bounds__(kWsThreads) k_ws_count(const uint8_t *__restrict__ text, uint64_t len,
                                                         uint64_t *__restrict__ tile_ws,
                                                         uint64_t *__restrict__ tile_nl) {
    uint64_t base = (uint64_t)blockIdx.x * kWsTile;
    uint32_t cw = 0, cn = 0;
    for (int i = threadIdx.x; i < kWsTile / 16; i += kWsThreads) {
        uint64_t off = base + (uint64_t)i * 16;
        if (off >= len) break;
        uint32_t ws[4], nl[4];
        ws_flags16<CLS>(text, off, len, ws, nl);
        cw += __popc(ws[0]) + __popc(ws[1]) + __popc(ws[2]) + __popc(ws[3]);
        cn += __popc(nl[0]) + __popc(nl[1]) + __popc(nl[2]) + __popc(nl[3]);
    }
    __shared__ uint32_t s_w[kWsThreads / 32], s_n[kWsThreads / 32];
    cw = warp_sum(cw);
    cn = warp_sum(cn);
    if ((threadIdx.x & 31) == 0) {
        s_w[threadIdx.x >> 5] = cw;
        s_n[threadIdx.x >> 5] = cn;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t a = 0, b = 0;
        for (int i = 0; i < kWsThreads / 32; ++i) { a += s_w[i]; b += s_n[i]; }
        tile_ws[blockIdx.x] = a;
        tile_nl[blockIdx.x] = b;
    }
}

// block-wide exclusive scan of one value per thread (kWsThreads threads); returns offset, *total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *s_warp /*[8]*/, uint32_t *total) {
    uint32_t wtot;
    uint32_t off = warp_excl_scan(v, &wtot);
    uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();  // protect s_warp reuse
    if (lane == 0) s_warp[wid] = wtot;
    __syncthreads();
    uint32_t pre = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < kWsThreads / 32; ++i) {
        uint32_t x = s_warp[i];
        if ((uint32_t)i < wid) pre += x;
        tot += x;
    }
    *total = tot;
    return off + pre;
}

template <int CLS>
__global__ void __launch_bounds__(kWsThreads) k_ws_fill(const uint8_t *__restrict__ text, uint64_t len,
                                                        const uint64_t *__restrict__ tile_ws_off,
                                                        const uint64_t *__restrict__ tile_nl_off,
                                                        uint64_t *__restrict__ ws_ent,
                                                        uint64_t *__restrict__ nl_ws_idx) {
    __shared__ uint32_t s_a[kWsThreads / 32], s_b[kWsThreads / 32];
    uint64_t base = (uint64_t)blockIdx.x * kWsTile;
    uint64_t wbase = tile_ws_off[blockIdx.x], nbase = tile_nl_off[blockIdx.x];
    for (int it = 0; it < kWsTile / 16 / kWsThreads; ++it) {
        uint64_t off = base + ((uint64_t)it * kWsThreads + threadIdx.x) * 16;
        uint32_t ws[4] = {0, 0, 0, 0}, nl[4] = {0, 0, 0, 0};
        if (off < len) ws_flags16<CLS>(text, off, len, ws, nl);
        uint32_t cw = __popc(ws[0]) + __popc(ws[1]) + __popc(ws[2]) + __popc(ws[3]);
        uint32_t cn = __popc(nl[0]) + __popc(nl[1]) + __popc(nl[2]) + __popc(nl[3]);
        uint32_t tw, tn;
        uint32_t ow = block_excl_scan(cw, s_a, &tw);
        uint32_t on = block_excl_scan(cn, s_b, &tn);
        uint64_t wi = wbase + ow, ni = nbase + on;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t m = ws[k];
            while (m) {
                int bit = __ffs(m) - 1;
                m &= m - 1;
                uint64_t pos = off + 4 * k + (bit >> 3);
                bool isnl = (nl[k] >> bit) & 1u;
                ws_ent[wi] = pos | (isnl ? ~kPosMask : 0ull);
                if (isnl) nl_ws_idx[ni++] = wi;
                ++wi;
            }
        }
        wbase += tw;
        nbase += tn;
        if (base + ((uint64_t)(it + 1) * kWsThreads) * 16 >= len) break;
    }
}

// ---------------------------------------------------------------- 1b. single-pass whitespace index
// The exact two-pass index above reads the slab twice.  Whitespace is sparse in MAF (~7 bytes per s-line of
// several kB), so the fast path reads the slab ONCE: a warp owns a 4 KiB tile, keeps its 8 x 512 B of loads
// in flight, and writes the (tile-relative) whitespace offsets into a fixed slot region of kTokCap entries
// per tile; a sparse pass then compacts the slots into ws_ent / nl_ws_idx.  A tile with more than kTokCap
// whitespace bytes raises the overflow flag and the caller falls back to the exact two-pass kernels.
constexpr int kTokTile = 4096;
constexpr int kTokCap = 128;     // MAF: whitespace slots per 4 KiB tile
constexpr int kTokCapPaf = 512;  // PAF: tab/newline slots per 4 KiB tile
constexpr int kTokWarps = 8;

// 0x80 in every byte below 0x21 (control characters and space): the superset of all delimiter classes
__device__ __forceinline__ uint32_t lt21_bytes(uint32_t w) { return ~(((w & 0x7F7F7F7Fu) + 0x5F5F5F5Fu) | w) & kHi; }
template <int CLS>
__device__ __forceinline__ bool is_delim(uint32_t ch) {
    return CLS == 0 ? (ch == 0x20u || (ch - 9u) < 5u) : (ch == 9u || ch == 10u || ch == 13u);
}

template <int CLS, int CAP>
__global__ void __launch_bounds__(kTokWarps * 32) k_ws_scan(const uint8_t *__restrict__ text, uint64_t len, uint64_t n_tiles,
                                                            uint64_t *__restrict__ tile_ws, uint64_t *__restrict__ tile_nl,
                                                            uint16_t *__restrict__ slots, uint32_t *__restrict__ overflow) {
    __shared__ uint32_t s_cnt[kTokWarps][2];  // per warp: delimiters, newlines of the current tile
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t n_warps = (uint64_t)gridDim.x * kTokWarps;
    for (uint64_t tile = (uint64_t)blockIdx.x * kTokWarps + wid; tile < n_tiles; tile += n_warps) {
        const uint64_t base = tile * kTokTile;
        uint4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint64_t off = base + (uint64_t)k * 512 + lane * 16;
            v[k] = off < len ? ldg16(text + off) : make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
        }
        if (lane == 0) { s_cnt[wid][0] = 0; s_cnt[wid][1] = 0; }
        __syncwarp();
        uint16_t *slot = slots + tile * CAP;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            // dense part: 3 ops per word.  Delimiters are sparse, the exact test runs per candidate byte only;
            // entries of one step land in arbitrary order (shared atomic), k_ws_compact sorts the tile's few entries
            const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
            const uint32_t c0 = lt21_bytes(w[0]), c1 = lt21_bytes(w[1]), c2 = lt21_bytes(w[2]), c3 = lt21_bytes(w[3]);
            if (c0 | c1 | c2 | c3) {
                const uint32_t cand[4] = {c0, c1, c2, c3};
                const uint32_t rel0 = (uint32_t)k * 512u + lane * 16u;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t m = cand[i];
                    while (m) {
                        const int bit = __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t byte = (uint32_t)(bit >> 3);
                        const uint32_t ch = (w[i] >> (8 * byte)) & 0xFFu;
                        const uint32_t rel = rel0 + 4u * i + byte;
                        if (is_delim<CLS>(ch) && base + rel < len) {
                            const uint32_t isnl = ch == 0x0Au || (CLS == 1 && ch == 0x0Du);
                            const uint32_t idx = atomicAdd(&s_cnt[wid][0], 1u);
                            if (isnl) atomicAdd(&s_cnt[wid][1], 1u);
                            if (idx < (uint32_t)CAP) slot[idx] = (uint16_t)(rel | (isnl << 15));
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            const uint32_t cw = s_cnt[wid][0];
            tile_ws[tile] = cw;
            tile_nl[tile] = s_cnt[wid][1];
            if (cw > (uint32_t)CAP) atomicOr(overflow, 1u);
        }
        __syncwarp();
    }
}

// one warp per tile: the slots of a tile arrive ordered step by step but unordered inside a step (shared atomics), so
// every entry computes its rank among the tile's (few) entries with shuffles and is written straight to its place
template <int CAP>
__global__ void k_ws_compact(uint64_t n_tiles, const uint64_t *__restrict__ tile_ws, const uint64_t *__restrict__ ws_off,
                             const uint64_t *__restrict__ nl_off, const uint16_t *__restrict__ slots,
                             uint64_t *__restrict__ ws_ent, uint64_t *__restrict__ nl_ws_idx) {
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (t >= n_tiles) return;
    uint32_t c = (uint32_t)tile_ws[t];
    if (c > (uint32_t)CAP) c = CAP;  // overflow: the caller discards this index
    if (c == 0) return;
    const uint64_t wo = ws_off[t], no = nl_off[t], base = t * kTokTile;
    const uint16_t *slot = slots + t * CAP;
    if (c <= 32) {
        const uint32_t e = lane < c ? slot[lane] : 0xFFFFu;
        const uint32_t rel = e & 0x7FFFu, isnl = lane < c ? (e >> 15) : 0u;
        uint32_t rank = 0, nlrank = 0;
        for (uint32_t j = 0; j < c; ++j) {
            const uint32_t v = __shfl_sync(0xFFFFFFFFu, e, j);
            const bool before = (v & 0x7FFFu) < rel;
            rank += before;
            nlrank += before && (v >> 15);
        }
        if (lane < c) {
            ws_ent[wo + rank] = (base + rel) | (isnl ? ~kPosMask : 0ull);
            if (isnl) nl_ws_idx[no + nlrank] = wo + rank;
        }
        return;
    }
    // many delimiters in one tile (rare): ranks by a cooperative O(c^2) count over the slot array
    for (uint32_t i = lane; i < c; i += 32) {
        const uint32_t e = slot[i], rel = e & 0x7FFFu;
        uint32_t rank = 0, nlrank = 0;
        for (uint32_t j = 0; j < c; ++j) {
            const uint32_t v = slot[j];
            const bool before = (v & 0x7FFFu) < rel;
            rank += before;
            nlrank += before && (v >> 15);
        }
        ws_ent[wo + rank] = (base + rel) | ((e >> 15) ? ~kPosMask : 0ull);
        if (e >> 15) nl_ws_idx[no + nlrank] = wo + rank;
    }
}

// ---------------------------------------------------------------- 2. lines and records
struct LineView {
    const uint8_t *text;
    uint64_t len;
    const uint64_t *ws_ent;
    const uint64_t *nl_ws_idx;
    uint64_t n_ws, n_nl, n_lines;
    uint32_t first_slab;
    __device__ __forceinline__ uint64_t start(uint64_t k) const {
        return k == 0 ? 0 : (ws_ent[nl_ws_idx[k - 1]] & kPosMask) + 1;
    }
    __device__ __forceinline__ uint64_t end(uint64_t k) const {
        return k < n_nl ? (ws_ent[nl_ws_idx[k]] & kPosMask) : len;
    }
    // lines starting with 's' are s-lines (maf.rs:378, :398); line 0 of the first slab is the header
    __device__ __forceinline__ bool is_s(uint64_t k) const {
        if (k == 0 && first_slab) return false;
        uint64_t s = start(k), e = end(k);
        return e > s && text[s] == 's';
    }
};

// flags: bit0 = s-line, bit1 = first s-line of a record
__global__ void k_line_flags(LineView lv, uint8_t *__restrict__ flags, uint32_t *__restrict__ head) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= lv.n_lines) return;
    bool s = lv.is_s(k);
    bool h = s && !(k > 0 && lv.is_s(k - 1));
    flags[k] = (uint8_t)((s ? 1 : 0) | (h ? 2 : 0));
    head[k] = h ? 1u : 0u;
}

struct SLines {  // per-line SoA (meaningful for s-lines only)
    uint64_t *name_off;
    uint32_t *name_len;
    uint64_t *start, *ali, *src;
    uint8_t *neg;
    uint64_t *seq_off, *seq_len;
};

// err_key: atomicMin of (line_index << 8 | code); parse codes are smaller than record-level codes
__device__ __forceinline__ void report(unsigned long long *err_key, uint64_t where, int code) {
    atomicMin(err_key, (unsigned long long)((where << 8) | (uint64_t)code));
}

// ---------------------------------------------------------------- 3. s-line field parse
// maf.rs:138-211 parse_sline: split_whitespace into exactly 7 fields (mode name start size strand
// srcSize text); errors in field order.  Token boundaries come from the whitespace index.
__global__ void k_parse_slines(LineView lv, const uint8_t *__restrict__ flags,
                               const uint32_t *__restrict__ head_scan, SLines sl,
                               uint64_t *__restrict__ rec_first, uint64_t *__restrict__ rec_last,
                               unsigned long long *err_key) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= lv.n_lines) return;
    uint8_t f = flags[k];
    if (!(f & 1)) return;
    uint32_t rec = head_scan[k] + ((f & 2) ? 1u : 0u) - 1u;
    if (f & 2) rec_first[rec] = k;
    bool tail = !(k + 1 < lv.n_lines && (flags[k + 1] & 1));
    if (tail) rec_last[rec] = k;

    uint64_t S = lv.start(k), E = lv.end(k);
    uint64_t wlo = k == 0 ? 0 : lv.nl_ws_idx[k - 1] + 1;
    uint64_t whi = k < lv.n_nl ? lv.nl_ws_idx[k] : lv.n_ws;
    uint64_t tb[8], te[8];
    int nt = 0;
    uint64_t p = S;
    for (uint64_t w = wlo; w <= whi && nt < 8; ++w) {
        uint64_t x = w < whi ? (lv.ws_ent[w] & kPosMask) : E;
        if (x > p) {
            tb[nt] = p;
            te[nt] = x;
            ++nt;
        }
        p = x + 1;
    }
    int code = 0;
    uint64_t v_start = 0, v_ali = 0, v_src = 0;
    uint8_t neg = 0;
    if (nt < 2) code = WGA_E_MAF_FIELD_MISSING;  // "name" (mode always exists: the line starts with 's')
    if (!code) {
        if (nt < 3) code = WGA_E_MAF_FIELD_MISSING;
        else if (!parse_u64_bytes(lv.text + tb[2], te[2] - tb[2], &v_start)) code = WGA_E_PARSE_INT;
    }
    if (!code) {
        if (nt < 4) code = WGA_E_MAF_FIELD_MISSING;
        else if (!parse_u64_bytes(lv.text + tb[3], te[3] - tb[3], &v_ali)) code = WGA_E_PARSE_INT;
    }
    if (!code) {
        if (nt < 5) code = WGA_E_MAF_FIELD_MISSING;
        else if (te[4] - tb[4] == 1 && lv.text[tb[4]] == '+') neg = 0;
        else if (te[4] - tb[4] == 1 && lv.text[tb[4]] == '-') neg = 1;
        else code = WGA_E_PARSE_STRAND;
    }
    if (!code) {
        if (nt < 6) code = WGA_E_MAF_FIELD_MISSING;
        else if (!parse_u64_bytes(lv.text + tb[5], te[5] - tb[5], &v_src)) code = WGA_E_PARSE_INT;
    }
    if (!code) {
        if (nt < 7) code = WGA_E_MAF_FIELD_MISSING;
        else if (nt > 7) code = WGA_E_MAF_SURPLUS_FIELD;
    }
    if (code) {
        report(err_key, k, code);
        sl.seq_len[k] = 0;
        sl.name_len[k] = 0;
        return;
    }
    sl.name_off[k] = tb[1];
    sl.name_len[k] = (uint32_t)(te[1] - tb[1]);
    sl.start[k] = v_start;
    sl.ali[k] = v_ali;
    sl.src[k] = v_src;
    sl.neg[k] = neg;
    sl.seq_off[k] = tb[6];
    sl.seq_len[k] = te[6] - tb[6];
}

// ---------------------------------------------------------------- 4. records and work items
struct Recs {
    uint64_t *first, *last;  // line range
    uint64_t *tl, *ql;       // target / query line index
    uint64_t *t_off, *q_off, *ncols;
    uint8_t *neg, *skip;
    uint32_t *n_items;
};

// mode 0: maf2paf/stat (missing query => error like the reference); mode 1: call (skip the block)
// target = slines[0]; query = slines[1] or the first s-line named `qname`
// (maf.rs:263-285 set_query_idx_byname, :424-540 AlignRecord; caller.rs:62-108)
// --query-regex (maf.rs:267-271, caller.rs:91-104): the host evaluates its Regex on the slab's distinct s-line names
// (wga_maf_sline_names) and hands the matching names back; "first s-line whose name matches" is then "first s-line whose
// name is in the set".  The set travels as FNV-1a hashes in ascending order + the names themselves (exact compare).
struct NameSet {
    const uint64_t *hash;  // [n] ascending
    const uint32_t *off;   // [n+1] into bytes, same order
    const uint8_t *bytes;
    uint32_t n;
};
__host__ __device__ __forceinline__ uint64_t fnv1a64(const uint8_t *p, uint32_t n) {
    uint64_t h = 0xCBF29CE484222325ull;
    for (uint32_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001B3ull; }
    return h;
}
__device__ __forceinline__ bool name_in_set(const NameSet &S, const uint8_t *nm, uint32_t len) {
    const uint64_t h = fnv1a64(nm, len);
    uint32_t lo = 0, hi = S.n;
    while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (S.hash[m] < h) lo = m + 1; else hi = m; }
    for (; lo < S.n && S.hash[lo] == h; ++lo) {
        if (S.off[lo + 1] - S.off[lo] != len) continue;
        const uint8_t *c = S.bytes + S.off[lo];
        bool eq = true;
        for (uint32_t j = 0; j < len && eq; ++j) eq = nm[j] == c[j];
        if (eq) return true;
    }
    return false;
}

// has_qname: 0 = slines[1], 1 = the first s-line named `qname`, 2 = the first s-line whose name is in `qset`
__global__ void k_build_records(uint64_t n_rec, const uint8_t *__restrict__ text, SLines sl, Recs rc,
                                const uint8_t *__restrict__ qname, uint32_t qname_len, int has_qname, NameSet qset,
                                int mode, unsigned long long *err_key, unsigned long long *ragged) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    uint64_t first = rc.first[r], last = rc.last[r], nl = last - first + 1;
    uint64_t qi = 1;
    bool skip = false;
    if (mode == 1 && nl == 1) skip = true;  // caller.rs:64-74
    if (!skip && has_qname) {
        bool found = false;
        for (uint64_t i = 0; i < nl && !found; ++i) {
            uint64_t k = first + i;
            if (has_qname == 2) {
                if (name_in_set(qset, text + sl.name_off[k], sl.name_len[k])) { found = true; qi = i; }
                continue;
            }
            if (sl.name_len[k] != qname_len) continue;
            const uint8_t *nm = text + sl.name_off[k];
            bool eq = true;
            for (uint32_t j = 0; j < qname_len; ++j)
                if (nm[j] != qname[j]) { eq = false; break; }
            if (eq) { found = true; qi = i; }
        }
        if (!found) {
            if (mode == 1) skip = true;
            else { report(err_key, last, WGA_E_QUERY_NAME_NOT_FOUND); skip = true; }
        }
    }
    if (!skip && qi >= nl) {  // slines[1] on a 1-s-line block: the reference panics (maf.rs:426-453)
        report(err_key, last, WGA_E_REF_PANIC);
        skip = true;
    }
    uint64_t tl = first, ql = first + (skip ? 0 : qi);
    uint64_t nt = sl.seq_len[tl], nq = sl.seq_len[ql];
    uint64_t n = nt < nq ? nt : nq;  // zip() stops at the shorter sequence (cigar.rs:359-361)
    if (mode == 1 && !skip) {
        // query_seq()[start..end] slicing panics (caller.rs:173-175), and create_chunk_record slices EVERY s-line of the
        // block with the chunk bounds (caller.rs:233-234): any row shorter than the target row panics
        bool bad = nq < nt;
        for (uint64_t i = 0; i < nl && !bad; ++i) bad = sl.seq_len[first + i] < nt;
        if (bad) {
            report(err_key, last, WGA_E_REF_PANIC);
            skip = true;
        }
    }
    if (mode == 0 && !skip) {  // would `call` see another record here?  (rows shorter than the target, single s-line blocks)
        bool bad = nl == 1;
        for (uint64_t i = 0; i < nl && !bad; ++i) bad = sl.seq_len[first + i] < nt;
        if (bad) *ragged = 1ull;
    }
    if (mode == 1) n = nt;
    rc.tl[r] = tl;
    rc.ql[r] = ql;
    rc.t_off[r] = sl.seq_off[tl];
    rc.q_off[r] = sl.seq_off[ql];
    rc.ncols[r] = skip ? 0 : n;
    rc.neg[r] = sl.neg[ql];
    rc.skip[r] = skip ? 1 : 0;
    rc.n_items[r] = skip ? 0u : (uint32_t)((n + kSeg - 1) / kSeg);
}

// wga_maf_sline_names: byte count (name + '\n') of every s-line that belongs to a record, then the names themselves
__global__ void k_sline_name_len(uint64_t n_rec, Recs rc, SLines sl, uint64_t n_lines, uint64_t *__restrict__ nlen) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    for (uint64_t k = rc.first[r]; k <= rc.last[r] && k < n_lines; ++k) nlen[k] = (uint64_t)sl.name_len[k] + 1;
}
__global__ void k_sline_name_copy(uint64_t n_lines, const uint8_t *__restrict__ text, SLines sl, const uint64_t *__restrict__ nlen,
                                  const uint64_t *__restrict__ noff, char *__restrict__ out) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_lines || nlen[k] == 0) return;
    char *o = out + noff[k];
    const uint8_t *nm = text + sl.name_off[k];
    const uint32_t n = sl.name_len[k];
    for (uint32_t j = 0; j < n; ++j) o[j] = (char)nm[j];
    o[n] = '\n';
}

__global__ void k_fill_items(uint64_t n_rec, const uint32_t *__restrict__ n_items,
                             const uint64_t *__restrict__ item_base, uint32_t *__restrict__ item_rec,
                             uint32_t *__restrict__ item_seg) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    uint64_t b = item_base[r];
    uint32_t n = n_items[r];
    for (uint32_t s = 0; s < n; ++s) {
        item_rec[b + s] = (uint32_t)r;
        item_seg[b + s] = s;
    }
}

}  // namespace wga
