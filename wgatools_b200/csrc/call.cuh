// call.cuh -- sparse passes of `call` on top of the dense caller-classified run list.
//
// Replaces src/tools/caller.rs:109-149 (chunk loop), :159-219 find_safe_chunk_boundary, :221-265
// create_chunk_record and :388-608 call_within_var + :359-386 get_variant_rec (noodles-vcf record text).
// Nothing here touches columns: coordinates come from segmented prefix sums over run lengths, the chunk
// grid is recomputed per record from its runs, and VCF rows are sized, scanned and
// written one thread per row.
//
// Row order inside a record = column order; a chunk's <INV> row precedes everything at its first column.
#pragma once
#include "ctx.cuh"
#include "emit_paf.cuh"
#include "tokenize.cuh"

namespace wga {

__device__ __forceinline__ bool cls_gap(uint8_t c) { return c >= CLS_I; }                 // I, D, W
__device__ __forceinline__ bool cls_has_t(uint8_t c) { return c == CLS_EQ || c == CLS_X || c == CLS_D; }
__device__ __forceinline__ bool cls_has_q(uint8_t c) { return c == CLS_EQ || c == CLS_X || c == CLS_I; }

// contributions to the non-gap column counts of target / query
__global__ void k_run_contrib(uint64_t n_runs, const uint8_t *__restrict__ cls, const uint64_t *__restrict__ len,
                              uint64_t *__restrict__ tcon, uint64_t *__restrict__ qcon) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j > n_runs) return;
    if (j == n_runs) { tcon[j] = 0; qcon[j] = 0; return; }
    uint8_t c = cls[j];
    tcon[j] = cls_has_t(c) ? len[j] : 0;
    qcon[j] = cls_has_q(c) ? len[j] : 0;
}

struct CallView {
    uint64_t n_rec, n_runs;
    const uint8_t *text;
    SLines sl;
    const uint64_t *tl, *ql, *ncols, *t_off, *q_off;
    const uint8_t *neg, *skip;
    const uint64_t *rec_run_first;              // [n_rec+1]
    const uint64_t *run_pos, *run_len;
    const uint32_t *run_rec;
    const uint8_t *run_cls;
    const uint64_t *St, *Sq;                    // exclusive scans of tcon / qcon over all runs [n_runs+1]
    uint64_t chunk_size, svlen;
    int snp, inv;
    // chunk grid
    const uint64_t *rec_chunk_first;            // [n_rec+1]
    const uint64_t *chunk_cs;                   // chunk start columns, grouped by record
    const uint32_t *chunk_inv;                  // 1 if the chunk emits an <INV> row
    const uint64_t *inv_cum;                    // exclusive scan of chunk_inv [n_chunks+1]
};

// caller.rs:119-149 + 159-219 over runs instead of columns.  emit(k, s, e) per chunk.
template <class F>
__device__ __forceinline__ uint64_t chunk_walk(const uint64_t *__restrict__ run_pos, const uint8_t *__restrict__ run_cls,
                                               uint64_t j0, uint64_t j1, uint64_t L, uint64_t C, uint64_t K, F emit) {
    uint64_t s = 0, j = j0, k = 0;
    while (s < L) {
        uint64_t p = (L - s > C) ? s + C : L, e = p;
        bool in_gap = false;
        uint64_t gap_from = 0, jj = j;
        while (jj < j1 && run_pos[jj] < p) {
            uint64_t a = run_pos[jj] > s ? run_pos[jj] : s;
            if (cls_gap(run_cls[jj])) {
                if (!in_gap) { in_gap = true; gap_from = a; }
            } else if (in_gap) {
                if (a - gap_from >= K) e = a;  // caller.rs:187-194, last qualifying closed gap run wins
                in_gap = false;
            }
            ++jj;
        }
        if (in_gap && p - gap_from >= K) {  // caller.rs:201-215: extend to the first both-non-gap column
            uint64_t k2 = jj;
            while (k2 < j1 && cls_gap(run_cls[k2])) ++k2;
            e = k2 < j1 ? run_pos[k2] : p;
        }
        emit(k, s, e);
        ++k;
        s = e;
        while (j + 1 < j1 && run_pos[j + 1] <= s) ++j;
    }
    return k;
}

// Two consecutive chunks always span at least C columns (the first ends at the LAST qualifying gap end inside its
// window, so the second finds none before the first window's end), hence a record has at most 2*ncols/C + 2 chunks:
// one pass writes the chunk starts into a region of that capacity and counts them, a tiny copy densifies.
__device__ __forceinline__ uint64_t chunk_cap(uint64_t ncols, uint64_t C) { return 2 * (ncols / C) + 3; }
__global__ void k_chunk_caps(uint64_t n_rec, const uint64_t *__restrict__ ncols, const uint8_t *__restrict__ skip, uint64_t C,
                             uint64_t *__restrict__ cap) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rec) return;
    cap[r] = (r == n_rec || skip[r] || ncols[r] == 0) ? 0 : chunk_cap(ncols[r], C);
}
__global__ void k_chunks(uint64_t n_rec, const uint64_t *__restrict__ rec_run_first, const uint64_t *__restrict__ run_pos,
                         const uint8_t *__restrict__ run_cls, const uint64_t *__restrict__ ncols,
                         const uint8_t *__restrict__ skip, uint64_t C, uint64_t K, uint64_t *__restrict__ n_chunks,
                         const uint64_t *__restrict__ cap_base, uint64_t *__restrict__ chunk_tmp) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rec) return;
    if (r == n_rec) { n_chunks[r] = 0; return; }
    if (skip[r] || ncols[r] == 0) { n_chunks[r] = 0; return; }
    const uint64_t j0 = rec_run_first[r], j1 = rec_run_first[r + 1], base = cap_base[r], cap = cap_base[r + 1] - base;
    n_chunks[r] = chunk_walk(run_pos, run_cls, j0, j1, ncols[r], C, K,
                             [&](uint64_t k, uint64_t s, uint64_t) { if (k < cap) chunk_tmp[base + k] = s; });
}
__global__ void k_chunks_dense(uint64_t n_rec, const uint64_t *__restrict__ cap_base, const uint64_t *__restrict__ chunk_tmp,
                               const uint64_t *__restrict__ rec_chunk_first, uint64_t *__restrict__ chunk_cs,
                               uint32_t *__restrict__ chunk_rec, unsigned long long *err_key) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint64_t c0 = rec_chunk_first[r], n = rec_chunk_first[r + 1] - c0, base = cap_base[r];
    if (n > cap_base[r + 1] - base) { atomicMin(err_key, (unsigned long long)((r << 8) | WGA_E_INTERNAL)); return; }
    for (uint64_t k = 0; k < n; ++k) { chunk_cs[c0 + k] = chunk_tmp[base + k]; chunk_rec[c0 + k] = (uint32_t)r; }
}

// index of the run of record r that contains column col (runs [j0,j1) sorted by pos)
__device__ __forceinline__ uint64_t run_at(const uint64_t *__restrict__ run_pos, uint64_t j0, uint64_t j1, uint64_t col) {
    uint64_t lo = j0, hi = j1;  // last run with pos <= col
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (run_pos[mid] <= col) lo = mid; else hi = mid;
    }
    return lo;
}
// non-gap target / query columns of record before column col (col may be inside run j)
__device__ __forceinline__ uint64_t nongap_before(const CallView &V, uint64_t rec, uint64_t col, bool target) {
    uint64_t j0 = V.rec_run_first[rec], j1 = V.rec_run_first[rec + 1];
    const uint64_t *S = target ? V.St : V.Sq;
    if (col >= V.ncols[rec]) return S[j1] - S[j0];
    uint64_t j = run_at(V.run_pos, j0, j1, col);
    uint8_t c = V.run_cls[j];
    bool has = target ? cls_has_t(c) : cls_has_q(c);
    return S[j] - S[j0] + (has ? col - V.run_pos[j] : 0);
}
// chunk index (global) of the chunk of record rec containing column col
__device__ __forceinline__ uint64_t chunk_at(const CallView &V, uint64_t rec, uint64_t col) {
    uint64_t lo = V.rec_chunk_first[rec], hi = V.rec_chunk_first[rec + 1];
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (V.chunk_cs[mid] <= col) lo = mid; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ uint64_t chunk_end(const CallView &V, uint64_t rec, uint64_t c) {
    return c + 1 < V.rec_chunk_first[rec + 1] ? V.chunk_cs[c + 1] : V.ncols[rec];
}

// caller.rs:423-440: a chunk of a '-' block with at least one target base gets an <INV> row (--inv)
__global__ void k_chunk_inv(CallView V, uint64_t n_chunks, const uint32_t *__restrict__ chunk_rec, uint32_t *__restrict__ flag) {
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_chunks) return;
    if (c == n_chunks) { flag[c] = 0; return; }
    uint32_t rec = chunk_rec[c];
    uint32_t f = 0;
    if (V.inv && V.neg[rec]) {
        uint64_t cs = V.chunk_cs[c], ce = chunk_end(V, rec, c);
        f = nongap_before(V, rec, ce, true) > nongap_before(V, rec, cs, true);
    }
    flag[c] = f;
}
__global__ void k_chunk_rec(uint64_t n_rec, const uint64_t *__restrict__ rec_chunk_first, uint32_t *__restrict__ chunk_rec) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    for (uint64_t c = rec_chunk_first[r]; c < rec_chunk_first[r + 1]; ++c) chunk_rec[c] = (uint32_t)r;
}

struct RunPlan {  // per-run emission plan
    uint64_t rows;      // class rows + owned <INV> rows
    uint64_t eff_len;   // I/D: length of the part that is emitted
    uint64_t anchor;    // I/D: column of the base preceding the run (last column of the previous =/X run)
};

// decides what each run emits (caller.rs:453-605) -- state machine `after_m` resolved locally:
// after_m is true iff the nearest previous non-W run of the record is =/X and at least one of its columns
// lies in the current chunk.
__device__ __forceinline__ void plan_run(const CallView &V, uint64_t j, uint64_t *class_rows, uint64_t *eff_len,
                                         uint64_t *anchor, uint64_t *oc0, uint64_t *oc1) {
    uint32_t rec = V.run_rec[j];
    uint64_t pos = V.run_pos[j], len = V.run_len[j];
    uint8_t c = V.run_cls[j];
    uint64_t cf = V.rec_chunk_first[rec], cl = V.rec_chunk_first[rec + 1];
    // owned chunk starts: pos <= cs < pos + len
    uint64_t lo = cf, hi = cl;
    while (lo < hi) { uint64_t m = (lo + hi) >> 1; if (V.chunk_cs[m] < pos) lo = m + 1; else hi = m; }
    *oc0 = lo;
    hi = cl;
    while (lo < hi) { uint64_t m = (lo + hi) >> 1; if (V.chunk_cs[m] < pos + len) lo = m + 1; else hi = m; }
    *oc1 = lo;
    *class_rows = 0;
    *eff_len = len;
    *anchor = 0;
    if (c == CLS_X) {
        if (V.snp) *class_rows = len;
    } else if (c == CLS_I || c == CLS_D) {
        bool at_chunk_start = *oc0 < *oc1 && V.chunk_cs[*oc0] == pos;
        uint64_t j0 = V.rec_run_first[rec];
        if (!at_chunk_start && j > j0) {
            uint64_t first_inner = (*oc0 < *oc1) ? V.chunk_cs[*oc0] : pos + len;  // the run is cut there
            uint64_t l1 = first_inner - pos;
            uint64_t jp = j - 1;
            uint64_t e_prev = pos;
            bool ok = true;
            if (V.run_cls[jp] == CLS_W) {
                e_prev = V.run_pos[jp];
                if (jp > j0) --jp; else ok = false;
            }
            if (ok && (V.run_cls[jp] == CLS_EQ || V.run_cls[jp] == CLS_X)) {
                uint64_t cs = V.chunk_cs[*oc0 - 1];  // chunk containing pos (pos is not a chunk start)
                if (cs < e_prev && l1 > V.svlen) {
                    *class_rows = 1;
                    *eff_len = l1;
                    *anchor = e_prev - 1;
                }
            }
        }
    }
}

__global__ void k_run_rows(CallView V, uint64_t *__restrict__ rows) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j > V.n_runs) return;
    if (j == V.n_runs) { rows[j] = 0; return; }
    uint64_t cr, el, an, oc0, oc1;
    plan_run(V, j, &cr, &el, &an, &oc0, &oc1);
    rows[j] = cr + (V.inv_cum[oc1] - V.inv_cum[oc0]);
}

__device__ __forceinline__ bool vcf_base(uint8_t b, char *out) {  // noodles Base: upper-case, ACGTN only
    uint8_t u = (b >= 'a' && b <= 'z') ? b - 32 : b;
    *out = (char)u;
    return u == 'A' || u == 'C' || u == 'G' || u == 'T' || u == 'N';
}

struct RowOut {
    char *o;      // nullptr in the sizing pass
    uint64_t n;
    bool bad_base;
    __device__ __forceinline__ void ch(char c) { if (o) o[n] = c; ++n; }
    template <int N>
    __device__ __forceinline__ void lit(const char (&s)[N]) { for (int i = 0; i < N - 1; ++i) ch(s[i]); }
    __device__ __forceinline__ void bytes(const uint8_t *s, uint64_t k) { for (uint64_t i = 0; i < k; ++i) ch((char)s[i]); }
    __device__ __forceinline__ void u64(uint64_t v) {
        int nd = ndigits_u64(v);
        if (o) write_u64(o + n, v, nd);
        n += nd;
    }
    __device__ __forceinline__ void bases(const uint8_t *s, uint64_t k) {
        for (uint64_t i = 0; i < k; ++i) { char c; if (!vcf_base(s[i], &c)) bad_base = true; ch(c); }
    }
    // k_vcf_rows: sizing never looks at the bases (lengths are arithmetic); writing validates them, and a long REF/ALT
    // (an SV of up to 100 kb) is left to the whole warp (def_*: at most one such payload per row)
    const uint8_t *def_src;
    uint64_t def_off, def_len;
    __device__ __forceinline__ void bases_sized(uint64_t k) { n += k; }
    __device__ __forceinline__ void bases_or_defer(const uint8_t *s, uint64_t k) {
        if (k > 96 && def_len == 0) { def_src = s; def_off = n; def_len = k; n += k; return; }
        bases(s, k);
    }
};

constexpr int kVcfThreads = 128;
constexpr int kVcfSpan = 6144;  // bytes of output one warp (32 consecutive rows) may stage in shared memory

// one VCF body row per thread.  PASS 0: row_len[row] = bytes.  PASS 1: the 32 rows of a warp are consecutive in
// the output, so they are formatted into one shared-memory span and copied out with aligned word stores
// (a span longer than kVcfSpan -- long SV rows -- is written directly).
// row -> run map: run j is written at its first row (runs without rows write nothing), a running maximum fills the rest
__global__ void k_row_run_seed(uint64_t n_runs, const uint64_t *__restrict__ rows, const uint64_t *__restrict__ rowbase,
                               uint32_t *__restrict__ row_run) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_runs && rows[j] != 0) row_run[rowbase[j]] = (uint32_t)j;
}

template <int PASS>
__global__ void __launch_bounds__(kVcfThreads) k_vcf_rows(CallView V, uint64_t n_rows, const uint64_t *__restrict__ rowbase,
                                                          const uint32_t *__restrict__ row_run,
                                                          uint64_t *__restrict__ row_len, const uint64_t *__restrict__ row_off,
                                                          char *__restrict__ out, unsigned long long *err_key) {
    __shared__ __align__(16) char s_span[PASS == 1 ? kVcfThreads / 32 : 1][PASS == 1 ? kVcfSpan + 16 : 16];
    uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    uint64_t span_lo = 0, span_hi = 0;
    bool staged = false;
    if (PASS == 1) {
        const uint64_t first = row - lane;
        if (first < n_rows) {
            const uint64_t lastp1 = first + 32 < n_rows ? first + 32 : n_rows;
            span_lo = row_off[first];
            span_hi = row_off[lastp1];
            staged = span_hi - span_lo <= (uint64_t)kVcfSpan;
        }
    }
    if (PASS == 0 && row == n_rows) row_len[row] = 0;  // (no early exit: the warp shares a search below)
    const bool have_row = row < n_rows;
    const uint8_t *d_src = nullptr;
    char *d_dst = nullptr;
    uint64_t d_len = 0;
    unsigned long long ekey = 0;
    // run owning a row: last j with rowbase[j] <= row -- read from the row -> run map (k_row_run_seed + running maximum);
    // without the map (more than 2^32 runs) one full binary search per warp and a bounded one per lane
    uint64_t j_warp = 0;
    if (!row_run) {
        const uint64_t row0 = row - lane;
        if (lane == 0 && row0 < n_rows) {
            uint64_t lo = 0, hi = V.n_runs;
            while (hi - lo > 1) { uint64_t m = (lo + hi) >> 1; if (rowbase[m] <= row0) lo = m; else hi = m; }
            j_warp = lo;
        }
        j_warp = __shfl_sync(0xFFFFFFFFu, j_warp, 0);
    }
    if (have_row) {
    uint64_t lo;
    if (row_run) lo = row_run[row];
    else {
        lo = j_warp;
        uint64_t hi = j_warp + 256 < V.n_runs ? j_warp + 256 : V.n_runs;
        if (hi < V.n_runs && rowbase[hi] <= row) { lo = hi; hi = V.n_runs; }  // a long stretch of runs without rows
        while (hi - lo > 1) { uint64_t m = (lo + hi) >> 1; if (rowbase[m] <= row) lo = m; else hi = m; }
    }
    const uint64_t j = lo;
    uint64_t k = row - rowbase[j];
    const uint32_t rec = V.run_rec[j];
    const uint64_t pos = V.run_pos[j];
    const uint8_t c = V.run_cls[j];
    uint64_t cr, el, an, oc0, oc1;
    plan_run(V, j, &cr, &el, &an, &oc0, &oc1);
    // decode row k of this run: kind 0 = class row at column `col`, 1 = <INV> row of chunk `ic`
    int kind = 0;
    uint64_t col = pos, ic = 0;
    if (c == CLS_X && V.snp) {
        uint64_t done_cols = 0;  // SNP rows already accounted for
        bool found = false;
        for (uint64_t oc = oc0; oc < oc1 && !found; ++oc) {
            uint64_t before = V.chunk_cs[oc] - pos - done_cols;  // SNP rows before this chunk start
            if (k < before) { col = pos + done_cols + k; found = true; break; }
            k -= before;
            done_cols += before;
            if (V.chunk_inv[oc]) {
                if (k == 0) { kind = 1; ic = oc; found = true; break; }
                --k;
            }
        }
        if (!found) col = pos + done_cols + k;
    } else {
        if (cr && k == 0) { kind = 0; }
        else {
            k -= cr;
            kind = 1;
            for (uint64_t oc = oc0; oc < oc1; ++oc)
                if (V.chunk_inv[oc]) { if (k == 0) { ic = oc; break; } --k; }
        }
    }
    const uint64_t tl = V.tl[rec], ql = V.ql[rec];
    const uint8_t *tseq = V.text + V.t_off[rec], *qseq = V.text + V.q_off[rec];
    const bool neg = V.neg[rec];
    const uint64_t t_start = V.sl.start[tl], q_raw = V.sl.start[ql], q_size = V.sl.src[ql];
    RowOut R;
    R.o = nullptr;
    if (PASS == 1) R.o = staged ? s_span[threadIdx.x >> 5] + (span_lo & 3u) + (row_off[row] - span_lo) : out + row_off[row];
    R.n = 0;
    R.bad_base = false;
    R.def_src = nullptr; R.def_off = 0; R.def_len = 0;
    R.bytes(V.text + V.sl.name_off[tl], V.sl.name_len[tl]);
    R.ch('\t');
    const uint8_t *qn = V.text + V.sl.name_off[ql];
    const uint32_t qnl = V.sl.name_len[ql];
    if (kind == 1) {
        uint64_t cs = V.chunk_cs[ic], ce = chunk_end(V, rec, ic);
        uint64_t tb0 = nongap_before(V, rec, cs, true), tb1 = nongap_before(V, rec, ce, true);
        uint64_t qb0 = nongap_before(V, rec, cs, false), qb1 = nongap_before(V, rec, ce, false);
        // first target base of the chunk
        uint64_t j0 = V.rec_run_first[rec], j1 = V.rec_run_first[rec + 1];
        uint64_t jr = run_at(V.run_pos, j0, j1, cs), fc = cs;
        while (jr < j1 && !cls_has_t(V.run_cls[jr])) { ++jr; if (jr < j1) fc = V.run_pos[jr]; }
        R.u64(t_start + tb0 + 1);
        R.lit("\t.\t");
        if (PASS == 0) R.bases_sized(1); else R.bases_or_defer(tseq + fc, 1);
        R.lit("\t<INV>\t.\t.\tSVTYPE=INV;END=");
        R.u64(t_start + tb1);
        R.lit("\tGT:QI\t1|1:");
        R.bytes(qn, qnl);
        R.ch('@');
        R.u64(q_size - q_raw - qb1);  // chunk query_start() for '-' (maf.rs:433-441)
        R.ch('@');
        R.u64(q_size - q_raw - qb0);  // chunk query_end()
        R.lit("@N\n");
    } else {
        // running offsets before column `col` (caller.rs:399-400,457-458...)
        uint64_t cidx = chunk_at(V, rec, col);
        uint64_t j0 = V.rec_run_first[rec];
        uint64_t tb = V.St[j] - V.St[j0] + (cls_has_t(c) ? col - pos : 0);
        uint64_t qb = V.Sq[j] - V.Sq[j0] + (cls_has_q(c) ? col - pos : 0);
        uint64_t t_offv = t_start + tb, q_offv;
        if (!neg) q_offv = q_raw + qb;
        else {
            uint64_t ce = chunk_end(V, rec, cidx);
            q_offv = q_size - q_raw - nongap_before(V, rec, ce, false) + (qb - nongap_before(V, rec, V.chunk_cs[cidx], false));
        }
        const char sfx = neg ? 'N' : 'P';
        if (c == CLS_X) {
            R.u64(t_offv + 1);
            R.lit("\t.\t");
            if (PASS == 0) R.bases_sized(1); else R.bases_or_defer(tseq + col, 1);
            R.ch('\t');
            if (PASS == 0) R.bases_sized(1); else R.bases_or_defer(qseq + col, 1);
            R.lit("\t.\t.\t.\tGT:QI\t1|1:");
            R.bytes(qn, qnl);
            R.ch('@');
            R.u64(q_offv);
            R.ch('@'); R.ch(sfx); R.ch('\n');
        } else {
            R.u64(t_offv);
            R.lit("\t.\t");
            if (c == CLS_I) {
                if (PASS == 0) R.bases_sized(1); else R.bases_or_defer(tseq + an, 1);
                R.ch('\t');
                if (PASS == 0) R.bases_sized(1); else R.bases_or_defer(qseq + an, 1);
                if (PASS == 0) R.bases_sized(el); else R.bases_or_defer(qseq + pos, el);
            } else {
                if (PASS == 0) R.bases_sized(1); else R.bases_or_defer(tseq + an, 1);
                if (PASS == 0) R.bases_sized(el); else R.bases_or_defer(tseq + pos, el);
                R.ch('\t');
                if (PASS == 0) R.bases_sized(1); else R.bases_or_defer(qseq + an, 1);
            }
            R.lit("\t.\t.\t");
            if (neg) R.lit("INV_NEST=TRUE;");
            if (c == CLS_I) R.lit("SVTYPE=INS;SVLEN="); else R.lit("SVTYPE=DEL;SVLEN=");
            R.u64(el);
            R.lit(";END=");
            R.u64(c == CLS_I ? t_offv : t_offv + el);
            R.lit("\tGT:QI\t1|1:");
            R.bytes(qn, qnl);
            R.ch('@');
            R.u64(q_offv);
            R.ch('@');
            R.u64(c == CLS_I ? q_offv + el : q_offv);
            R.ch('@'); R.ch(sfx); R.ch('\n');
        }
    }
    if (PASS == 0) row_len[row] = R.n;
    else {
        ekey = (unsigned long long)((V.rec_run_first[rec] /*stable per record*/ << 8) | WGA_E_VCF_BASE);
        if (R.bad_base) atomicMin(err_key, ekey);
        d_src = R.def_src; d_len = R.def_len; d_dst = R.o + R.def_off;
    }
    }  // have_row
    if (PASS == 1) {
        // long REF/ALT payloads: the whole warp copies (upper-cases, validates) one payload at a time
        uint32_t pend = __ballot_sync(0xFFFFFFFFu, d_len != 0);
        while (pend) {
            const int sl = __ffs(pend) - 1;
            pend &= pend - 1;
            const uint8_t *src = (const uint8_t *)__shfl_sync(0xFFFFFFFFu, (unsigned long long)d_src, sl);
            char *dst = (char *)__shfl_sync(0xFFFFFFFFu, (unsigned long long)d_dst, sl);
            const uint64_t len = __shfl_sync(0xFFFFFFFFu, (unsigned long long)d_len, sl);
            const unsigned long long key = __shfl_sync(0xFFFFFFFFu, ekey, sl);
            bool bad = false;
            for (uint64_t i = lane; i < len; i += 32) { char c; if (!vcf_base(src[i], &c)) bad = true; dst[i] = c; }
            if (bad) atomicMin(err_key, key);
        }
        __syncwarp();
        if (staged) {
            const char *sp = s_span[threadIdx.x >> 5];
            const uint32_t al = (uint32_t)(span_lo & 3u);
            const uint64_t nbytes = span_hi - span_lo;
            char *gbase = out + (span_lo - al);            // 4-byte aligned (out comes from cudaMalloc)
            const uint64_t nwords = (al + nbytes + 3) >> 2;
            for (uint64_t wi = lane; wi < nwords; wi += 32) {
                const uint64_t b0 = wi * 4;
                if (b0 >= al && b0 + 4 <= al + nbytes) {
                    *reinterpret_cast<uint32_t *>(gbase + b0) = *reinterpret_cast<const uint32_t *>(sp + b0);
                } else {
                    for (uint64_t b = b0; b < b0 + 4; ++b)
                        if (b >= al && b < al + nbytes) gbase[b] = sp[b];
                }
            }
        }
    }
}

}  // namespace wga
