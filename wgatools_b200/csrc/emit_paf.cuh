// emit_paf.cuh -- sparse passes that turn the run list into PAF rows.
// Replaces the string building of parse_maf_seq_to_cigar (src/parser/cigar.rs:409-410), MAFRecord::
// convert2paf (src/parser/maf.rs:484-520) and the csv serialisation of src/converter.rs:35-52.
// Everything here works on runs (~1 per several hundred columns) or on records, never on columns.
#pragma once
#include "ctx.cuh"
#include "tokenize.cuh"

namespace wga {

__constant__ char kClsChar[5] = {'=', 'X', 'I', 'D', 'W'};

// first run index of every record (+ sentinel): run_base of the record's first work item
__global__ void k_rec_run_first(uint64_t n_rec, const uint64_t *__restrict__ item_base,
                                const uint64_t *__restrict__ run_base, uint64_t n_items, uint64_t n_runs,
                                uint64_t *__restrict__ rec_run_first) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rec) return;
    if (r == n_rec) { rec_run_first[r] = n_runs; return; }
    uint64_t ib = item_base[r];
    rec_run_first[r] = ib < n_items ? run_base[ib] : n_runs;
}

// run length = next head of the same record - this head, or the record end; dl = bytes of "<len><op>"
__global__ void k_run_len(uint64_t n_runs, const uint64_t *__restrict__ run_pos,
                          const uint32_t *__restrict__ run_rec, const uint64_t *__restrict__ ncols,
                          uint64_t *__restrict__ run_len, uint32_t *__restrict__ run_dl) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j > n_runs) return;
    if (j == n_runs) { if (run_dl) run_dl[j] = 0; return; }
    uint32_t rec = run_rec[j];
    uint64_t pos = run_pos[j];
    uint64_t nxt = (j + 1 < n_runs && run_rec[j + 1] == rec) ? run_pos[j + 1] : ncols[rec];
    uint64_t len = nxt - pos;
    run_len[j] = len;
    if (run_dl) run_dl[j] = (uint32_t)ndigits_u64(len) + 1u;
}

// ---- EXT run list out of the caller-classified one (WGA_F_SHARE_WALK, only when the slab has both-gap columns):
// cigar_cat_ext (cigar.rs:298-308) calls a both-gap column '=', cigar_cat_ext_caller (cigar.rs:314-328) calls it 'W';
// I / D / X columns are the same in both.  So the EXT list is the caller list with W relabelled '=' and every run
// dropped whose relabelled class equals that of the run before it in the same record.
__global__ void k_ext_flags(uint64_t n_runs, const uint32_t *__restrict__ run_rec, const uint8_t *__restrict__ run_cls,
                            uint32_t *__restrict__ flag) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j > n_runs) return;
    if (j == n_runs) { flag[j] = 0; return; }
    uint32_t keep = 1;
    if (j > 0 && run_rec[j - 1] == run_rec[j]) {
        const uint8_t a = run_cls[j - 1] == CLS_W ? (uint8_t)CLS_EQ : run_cls[j - 1];
        const uint8_t b = run_cls[j] == CLS_W ? (uint8_t)CLS_EQ : run_cls[j];
        keep = a != b;
    }
    flag[j] = keep;
}
__global__ void k_ext_scatter(uint64_t n_runs, const uint32_t *__restrict__ flag, const uint64_t *__restrict__ idx,
                              const uint64_t *__restrict__ run_pos, const uint32_t *__restrict__ run_rec,
                              const uint8_t *__restrict__ run_cls, uint64_t *__restrict__ xpos, uint32_t *__restrict__ xrec,
                              uint8_t *__restrict__ xcls) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_runs || !flag[j]) return;
    const uint64_t o = idx[j];
    xpos[o] = run_pos[j];
    xrec[o] = run_rec[j];
    xcls[o] = run_cls[j] == CLS_W ? (uint8_t)CLS_EQ : run_cls[j];
}
// the first run of a record always survives: its new index is the scan value at its old index
__global__ void k_ext_rec_first(uint64_t n_rec, const uint64_t *__restrict__ rec_run_first, const uint64_t *__restrict__ idx,
                                uint64_t *__restrict__ xfirst) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= n_rec) xfirst[r] = idx[rec_run_first[r]];
}

__device__ __forceinline__ uint32_t csv_name_len(const uint8_t *p, uint32_t n) {
    uint32_t q = 0;
    for (uint32_t i = 0; i < n; ++i) q += p[i] == '"';
    return q ? n + q + 2 : n;  // csv QuoteStyle::Necessary: quote + double the quotes
}
__device__ __forceinline__ char *csv_name_put(char *o, const uint8_t *p, uint32_t n) {
    bool q = false;
    for (uint32_t i = 0; i < n; ++i) q |= p[i] == '"';
    if (q) *o++ = '"';
    for (uint32_t i = 0; i < n; ++i) {
        if (p[i] == '"') *o++ = '"';
        *o++ = (char)p[i];
    }
    if (q) *o++ = '"';
    return o;
}
__device__ __forceinline__ char *put_u64(char *o, uint64_t v) { return o + write_u64(o, v, ndigits_u64(v)); }

struct PafRowArgs {
    uint64_t n_rec;
    const uint8_t *text;
    SLines sl;
    const uint64_t *tl, *ql, *ncols;
    const uint8_t *skip;
    const unsigned long long *rec_cnt;
    const uint64_t *rec_run_first, *cig_off;
    uint64_t *row_len;   // [n_rec+1]
    uint32_t *row_fixed; // bytes before the CIGAR text
    const uint64_t *row_off;
    char *out;
    const uint64_t *run_len;   // PASS 1: the CIGAR tokens are written by the row's own thread ...
    const uint32_t *run_dl;
    const uint8_t *run_cls;
    uint32_t *any_long;        // ... unless the record has more than kPafLongRuns runs: those are left to k_paf_cigar
};
constexpr uint64_t kPafLongRuns = 64;

// forward-strand query coordinates (maf.rs:433-450), wrapping like release Rust
__device__ __forceinline__ void query_fwd(const SLines &sl, uint64_t ql, uint64_t *qs, uint64_t *qe) {
    uint64_t st = sl.start[ql], al = sl.ali[ql], sz = sl.src[ql];
    if (sl.neg[ql]) { *qs = sz - st - al; *qe = sz - st; }
    else { *qs = st; *qe = st + al; }
}

constexpr int kPafThreads = 128;
constexpr int kPafSpan = 8192;  // bytes of output one warp (32 consecutive rows) may stage in shared memory

// pass 0: row lengths.  pass 1: one thread per record formats its whole row (fixed fields + CIGAR tokens); the 32 rows
// of a warp are consecutive in the output, so they are built in one shared-memory span and copied out with aligned
// word stores (a span longer than kPafSpan -- records with thousands of runs -- is written directly).
template <int PASS>
__global__ void __launch_bounds__(kPafThreads) k_paf_rows(PafRowArgs A) {
    __shared__ __align__(16) char s_span[PASS == 1 ? kPafThreads / 32 : 1][PASS == 1 ? kPafSpan + 16 : 16];
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    uint64_t span_lo = 0, span_hi = 0;
    bool staged = false;
    if (PASS == 1) {
        const uint64_t first = r - lane;
        if (first < A.n_rec) {
            const uint64_t lastp1 = first + 32 < A.n_rec ? first + 32 : A.n_rec;
            span_lo = A.row_off[first];
            span_hi = A.row_off[lastp1];
            staged = span_hi - span_lo <= (uint64_t)kPafSpan;
        }
    }
    if (PASS == 0) {
        if (r > A.n_rec) return;
        if (r == A.n_rec) { A.row_len[r] = 0; return; }
        if (A.skip[r]) { A.row_len[r] = 0; A.row_fixed[r] = 0; return; }
    }
    if (r < A.n_rec && !A.skip[r]) {
    uint64_t tl = A.tl[r], ql = A.ql[r], n = A.ncols[r];
    const unsigned long long *c = A.rec_cnt + r * 5;
    uint64_t match = n - c[0] - c[2] - c[4];
    uint64_t qs, qe;
    query_fwd(A.sl, ql, &qs, &qe);
    uint64_t ts = A.sl.start[tl], te = ts + A.sl.ali[tl];
    const uint8_t *qn = A.text + A.sl.name_off[ql], *tn = A.text + A.sl.name_off[tl];
    uint32_t qnl = A.sl.name_len[ql], tnl = A.sl.name_len[tl];
    const uint64_t j0 = A.rec_run_first[r], j1 = A.rec_run_first[r + 1];
    uint64_t cig = A.cig_off[j1] - A.cig_off[j0];
    if (PASS == 0) {
        // q_name q_size q_start q_end strand t_name t_size t_start t_end matches block_len 255 NM:i:x cg:Z:...
        uint32_t fixed = csv_name_len(qn, qnl) + 1 + ndigits_u64(A.sl.src[ql]) + 1 + ndigits_u64(qs) + 1 +
                         ndigits_u64(qe) + 1 + 1 + 1 + csv_name_len(tn, tnl) + 1 + ndigits_u64(A.sl.src[tl]) + 1 +
                         ndigits_u64(ts) + 1 + ndigits_u64(te) + 1 + ndigits_u64(match) + 1 + ndigits_u64(n) + 1 +
                         3 + 1 + 5 + ndigits_u64(n - match) + 1 + 5;
        A.row_fixed[r] = fixed;
        A.row_len[r] = (uint64_t)fixed + cig + 1;
        if (j1 - j0 > kPafLongRuns) atomicOr(A.any_long, 1u);
    } else {
        char *o = staged ? s_span[threadIdx.x >> 5] + (span_lo & 3u) + (A.row_off[r] - span_lo) : A.out + A.row_off[r];
        o = csv_name_put(o, qn, qnl); *o++ = '\t';
        o = put_u64(o, A.sl.src[ql]); *o++ = '\t';
        o = put_u64(o, qs); *o++ = '\t';
        o = put_u64(o, qe); *o++ = '\t';
        *o++ = A.sl.neg[ql] ? '-' : '+'; *o++ = '\t';
        o = csv_name_put(o, tn, tnl); *o++ = '\t';
        o = put_u64(o, A.sl.src[tl]); *o++ = '\t';
        o = put_u64(o, ts); *o++ = '\t';
        o = put_u64(o, te); *o++ = '\t';
        o = put_u64(o, match); *o++ = '\t';
        o = put_u64(o, n); *o++ = '\t';
        *o++ = '2'; *o++ = '5'; *o++ = '5'; *o++ = '\t';
        *o++ = 'N'; *o++ = 'M'; *o++ = ':'; *o++ = 'i'; *o++ = ':';
        o = put_u64(o, n - match); *o++ = '\t';
        *o++ = 'c'; *o++ = 'g'; *o++ = ':'; *o++ = 'Z'; *o++ = ':';
        if (j1 - j0 > kPafLongRuns) o += cig;  // one thread per run (k_paf_cigar) fills this in afterwards
        else
            for (uint64_t j = j0; j < j1; ++j) {  // "<len><op>" per run (cigar.rs:409-410)
                const int nd = (int)A.run_dl[j] - 1;
                write_u64(o, A.run_len[j], nd);
                o[nd] = kClsChar[A.run_cls[j]];
                o += nd + 1;
            }
        *o = '\n';
    }
    }
    if (PASS == 1) {
        __syncwarp();
        if (staged) {
            const char *sp = s_span[threadIdx.x >> 5];
            const uint32_t al = (uint32_t)(span_lo & 3u);
            const uint64_t nbytes = span_hi - span_lo;
            char *gbase = A.out + (span_lo - al);  // 4-byte aligned (out comes from cudaMalloc)
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

// one CIGAR token per run, for the records with more than kPafLongRuns runs (a Mbp-long block has 10^4 of them)
__global__ void k_paf_cigar(uint64_t n_runs, const uint32_t *__restrict__ run_rec,
                            const uint8_t *__restrict__ run_cls, const uint64_t *__restrict__ run_len,
                            const uint32_t *__restrict__ run_dl, const uint64_t *__restrict__ cig_off,
                            const uint64_t *__restrict__ rec_run_first, const uint64_t *__restrict__ row_off,
                            const uint32_t *__restrict__ row_fixed, char *__restrict__ out) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_runs) return;
    uint32_t rec = run_rec[j];
    if (rec_run_first[rec + 1] - rec_run_first[rec] <= kPafLongRuns) return;
    char *o = out + row_off[rec] + row_fixed[rec] + (cig_off[j] - cig_off[rec_run_first[rec]]);
    int nd = (int)run_dl[j] - 1;
    write_u64(o, run_len[j], nd);
    o[nd] = kClsChar[run_cls[j]];
}

}  // namespace wga
