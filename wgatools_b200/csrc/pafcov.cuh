// pafcov.cuh -- per-base target coverage of a PAF in BED form.
//
// Replaces src/tools/pafcov.rs:13-60 and src/parser/cigar.rs:710-741 update_cov_vec: per target a vector of the first
// record's target_length; M and = add 1 over [pos, pos+len) clipped to the vector, I and S do not move, every other op
// (X, D, N, ...) advances pos; one line "{target}\t{pos}\t{pos+1}\t{count}\n" per base.  The reference increments the
// vector base by base and folds per-worker HashMaps; here every M/= op is two atomics on a difference array, one
// prefix sum turns differences into counts, and the BED text is sized and written per tile of 256 positions.
// Targets come in first-occurrence order (the reference iterates a HashMap).
#pragma once
#include "ctx.cuh"
#include "paf.cuh"
#include "pafpseudo.cuh"

namespace wga {

// target table: head records (first record of each target) -> index, vector length, start of its difference segment
__global__ void k_cov_heads(uint64_t n_rec, const uint64_t *__restrict__ ft, uint32_t *__restrict__ head) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rec) return;
    head[r] = r < n_rec && ft[r] == r ? 1u : 0u;
}
__global__ void k_cov_targets(uint64_t n_rec, const uint32_t *__restrict__ head, const uint32_t *__restrict__ tidx, PafRecs R,
                              uint64_t *__restrict__ t_rec, uint64_t *__restrict__ t_len1) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec || !head[r]) return;
    t_rec[tidx[r]] = r;
    t_len1[tidx[r]] = R.t_size[r] + 1;  // one extra slot takes the closing -1 of spans that reach the end
}
__global__ void k_cov_verify(uint64_t n_rec, const uint8_t *__restrict__ text, PafRecs R, const uint64_t *__restrict__ ft,
                             unsigned long long *err_key) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint64_t a = ft[r];
    bool ok = R.t_name_len[r] == R.t_name_len[a];
    if (ok)
        for (uint32_t i = 0; i < R.t_name_len[r]; ++i) ok &= text[R.t_name_off[r] + i] == text[R.t_name_off[a] + i];
    if (!ok) paf_report(err_key, 0, 0, WGA_E_INTERNAL);
}

// cigar.rs:710-741, one record per thread: the tokenizer's errors are raised (update_cov_vec returns `res`)
__global__ void k_cov_ops(uint64_t n_rec, const uint8_t *__restrict__ text, PafRecs R, const uint64_t *__restrict__ ft,
                          const uint32_t *__restrict__ tidx, const uint64_t *__restrict__ t_base, const uint64_t *__restrict__ t_len1,
                          int *__restrict__ diff, unsigned long long *err_key) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint8_t negf = R.neg[r];
    const uint64_t len = R.cig_len[r];
    if (negf & 2) return;  // no cg:Z: tag: already reported by k_paf_parse
    if (len == 0) { paf_report(err_key, r, 0x800001, WGA_E_REF_PANIC); return; }
    const uint32_t t = tidx[ft[r]];
    const uint64_t base = t_base[t], L = t_len1[t] - 1;
    const uint8_t *p = text + R.cig_off[r];
    uint64_t pos = R.t_start[r], val = 0;
    uint32_t nd = 0, op_run = 0;
    bool ovf = false, in_op = false, failed = false;
    uint8_t op = 0;
    auto close_token = [&](uint64_t at) {
        if (failed) return;
        int code = 0;
        if (op_run != 1) code = WGA_E_CIGAR_OP_INVALID;
        else if (nd == 0 || ovf) code = WGA_E_PARSE_INT;
        if (code) { failed = true; paf_report(err_key, r, 0x800002 + (at > 0x7FFFF0ull ? 0x7FFFF0ull : at), code); return; }
        if (op == 'M' || op == '=') {
            const uint64_t lo = pos < L ? pos : L, hi = (val > L || pos + val > L) ? L : pos + val;
            if (hi > lo) { atomicAdd(diff + base + lo, 1); atomicAdd(diff + base + hi, -1); }
            pos += val;
        } else if (op != 'I' && op != 'S') pos += val;
    };
    for (uint64_t i = 0; i < len && !failed; ++i) {
        const uint8_t ch = p[i];
        if (ch >= '0' && ch <= '9') {
            if (in_op) { close_token(i); val = 0; nd = 0; ovf = false; op_run = 0; in_op = false; }
            const uint32_t d = ch - '0';
            if (val > (0xFFFFFFFFFFFFFFFFull - d) / 10) ovf = true;
            val = val * 10 + d;
            ++nd;
        } else { in_op = true; ++op_run; op = ch; }
    }
    if (!failed) close_token(len);
}

constexpr uint32_t kCovTile = 256;  // positions per formatting tile (tiles never cross targets)

struct CovView {
    uint64_t n_targets, n_tiles;
    const uint8_t *text;
    PafRecs R;
    const uint64_t *t_rec, *t_base, *t_len1;
    const uint64_t *t_tile0;   // [n_targets+1] first tile of every target
    const int *cov;            // inclusive prefix sums of the difference array
    uint64_t *tile_len;        // [n_tiles+1]
    const uint64_t *tile_off;
    char *out;
};
__global__ void k_cov_tiles(uint64_t n_targets, const uint64_t *__restrict__ t_len1, uint64_t *__restrict__ n_tiles) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_targets) return;
    n_tiles[t] = t < n_targets ? (t_len1[t] - 1 + kCovTile - 1) / kCovTile : 0;
}
template <int PASS>
__global__ void k_cov_rows(CovView V) {
    const uint64_t tile = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (PASS == 0) {
        if (tile > V.n_tiles) return;
        if (tile == V.n_tiles) { V.tile_len[tile] = 0; return; }
    } else if (tile >= V.n_tiles) return;
    uint64_t lo = 0, hi = V.n_targets;  // target owning this tile: last t with t_tile0[t] <= tile
    while (hi - lo > 1) { const uint64_t m = (lo + hi) >> 1; if (V.t_tile0[m] <= tile) lo = m; else hi = m; }
    const uint64_t t = lo, r = V.t_rec[t], L = V.t_len1[t] - 1;
    const uint64_t p0 = (tile - V.t_tile0[t]) * kCovTile, p1 = p0 + kCovTile < L ? p0 + kCovTile : L;
    const uint8_t *nm = V.text + V.R.t_name_off[r];
    const uint32_t nl = V.R.t_name_len[r];
    const int *c = V.cov + V.t_base[t];
    if (PASS == 0) {
        uint64_t n = 0;
        for (uint64_t p = p0; p < p1; ++p) n += nl + 1 + ndigits_u64(p) + 1 + ndigits_u64(p + 1) + 1 + ndigits_u64((uint64_t)(uint32_t)c[p]) + 1;
        V.tile_len[tile] = n;
        return;
    }
    char *o = V.out + V.tile_off[tile];
    for (uint64_t p = p0; p < p1; ++p) {
        for (uint32_t i = 0; i < nl; ++i) *o++ = (char)nm[i];
        *o++ = '\t'; o = put_u64(o, p); *o++ = '\t'; o = put_u64(o, p + 1); *o++ = '\t'; o = put_u64(o, (uint64_t)(uint32_t)c[p]); *o++ = '\n';
    }
}

}  // namespace wga
