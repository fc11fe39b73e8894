// pafpseudo.cuh -- PAF (+ optional FASTA) -> one pseudo-MAF per target.
//
// Replaces src/tools/pseudomaf.rs:18-237 (generate_pesudo_maf, write_pmaf, get_sline_seq) and
// src/parser/cigar.rs:744-804 gen_pesudo_maf_by_cigar.  Per target file:
//   "a score=0\n", the target s-line (whole sequence, or 'N' x size without a FASTA), then one s-line per query whose
//   sequence is the query projected onto target coordinates: records of the (target, query) group in target_start
//   order, '-' between records, a record contained in what was already written is skipped, a head overlap is
//   trimmed, '-' up to target_size at the end; a blank line closes the file.
// Projection of one record (cigar.rs:758-799): with FASTA M/=/X keep the (reverse-complemented if '-') query base,
// I/S delete query bases, D inserts '-'; without FASTA M/= -> '1', X -> '0', D -> '-'; other ops are ignored.
//
// The reference's two HashMaps iterate in arbitrary order; here targets, and queries inside a target, come in
// first-occurrence order (the convention of the stat rows).  The reference edits Strings (drain / insert_str, O(L) per
// indel); here the whole output is laid out by prefix sums, filled with '-' once, and every CIGAR op writes its own
// slice.  Records of one group with EQUAL target_start sit where `binary_search_by(..).unwrap_or_else(|e| e)` +
// `insert(idx, rec)` (pseudomaf.rs:91-94) puts them: k_pp_ties replays that search for every record of a tie class
// (the answer is a function of how many smaller / equal / greater records of the group came earlier in the file).
// The search is the three-way loop of the std the reference's flake.lock pins (nixpkgs 2024-01: rustc 1.75).
#pragma once
#include "ctx.cuh"
#include "paf.cuh"
#include "paf2maf.cuh"

namespace wga {

constexpr uint64_t kPpDropped = ~0ull;

// hash keys of the target name and of the (target, query) pair; dropped = filtered out by --target
__global__ void k_pp_keys(uint64_t n_rec, const uint8_t *__restrict__ text, PafRecs R, const uint8_t *__restrict__ filter,
                          uint32_t filter_len, int have_filter, uint64_t *__restrict__ kt, uint64_t *__restrict__ kp) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint8_t *tn = text + R.t_name_off[r];
    const uint32_t tl = R.t_name_len[r];
    if (have_filter) {
        bool eq = tl == filter_len;
        for (uint32_t i = 0; eq && i < tl; ++i) eq = tn[i] == filter[i];
        if (!eq) { kt[r] = kPpDropped; kp[r] = kPpDropped; return; }
    }
    uint64_t h = 0xCBF29CE484222325ull;
    for (uint32_t i = 0; i < tl; ++i) h = (h ^ tn[i]) * 0x100000001B3ull;
    uint64_t t = mix64(h);
    const uint8_t *qn = text + R.q_name_off[r];
    h ^= 0x9E3779B97F4A7C15ull;
    for (uint32_t i = 0; i < R.q_name_len[r]; ++i) h = (h ^ qn[i]) * 0x100000001B3ull;
    uint64_t p = mix64(h ^ tl);
    kt[r] = t >= kPpDropped - 1 ? 1 : (t ? t : 1);
    kp[r] = p >= kPpDropped - 1 ? 1 : (p ? p : 1);
}

// first record index of every distinct key (open addressing; keys != 0, != kPpDropped)
__global__ void k_pp_first_insert(uint64_t n_rec, const uint64_t *__restrict__ key, uint64_t cap_mask,
                                  unsigned long long *__restrict__ tkeys, unsigned long long *__restrict__ tfirst,
                                  uint64_t *__restrict__ slot_of) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const unsigned long long k = key[r];
    if (k == kPpDropped) { slot_of[r] = ~0ull; return; }
    uint64_t slot = mix64(k) & cap_mask;
    for (;;) {
        unsigned long long cur = tkeys[slot];
        if (cur == k) break;
        if (cur == 0ull) {
            unsigned long long old = atomicCAS(&tkeys[slot], 0ull, k);
            if (old == 0ull || old == k) break;
        }
        slot = (slot + 1) & cap_mask;
    }
    slot_of[r] = slot;
    atomicMin(&tfirst[slot], (unsigned long long)r);
}
__global__ void k_pp_first_read(uint64_t n_rec, const uint64_t *__restrict__ slot_of, const unsigned long long *__restrict__ tfirst,
                                uint64_t *__restrict__ first_of) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    first_of[r] = slot_of[r] == ~0ull ? kPpDropped : tfirst[slot_of[r]];
}
// a 64-bit key collision would merge two groups: compare the names with the group's first record
__global__ void k_pp_verify(uint64_t n_rec, const uint8_t *__restrict__ text, PafRecs R, const uint64_t *__restrict__ ft,
                            const uint64_t *__restrict__ fp, unsigned long long *err_key) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec || ft[r] == kPpDropped) return;
    const uint64_t a = ft[r], b = fp[r];
    bool ok = R.t_name_len[r] == R.t_name_len[a] && R.t_name_len[r] == R.t_name_len[b] && R.q_name_len[r] == R.q_name_len[b];
    if (ok) {
        for (uint32_t i = 0; i < R.t_name_len[r]; ++i)
            ok &= text[R.t_name_off[r] + i] == text[R.t_name_off[a] + i] && text[R.t_name_off[r] + i] == text[R.t_name_off[b] + i];
        for (uint32_t i = 0; i < R.q_name_len[r]; ++i) ok &= text[R.q_name_off[r] + i] == text[R.q_name_off[b] + i];
    }
    if (!ok) paf_report(err_key, 0, 0, WGA_E_INTERNAL);
}

__global__ void k_pp_used(uint64_t n_rec, const uint64_t *__restrict__ ft, uint32_t *__restrict__ used) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rec) return;
    used[r] = r < n_rec && ft[r] != kPpDropped ? 1u : 0u;
}
__global__ void k_pp_iota(uint64_t n, uint64_t *__restrict__ v) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}
__global__ void k_pp_gather_u64(uint64_t n, const uint64_t *__restrict__ src, const uint64_t *__restrict__ idx, uint64_t *__restrict__ dst) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

struct PpRecs {  // per SORTED record (targets by first occurrence, then query groups by first occurrence, then target_start)
    uint64_t *rec;                // file-order record index
    uint32_t *flags;              // bit0 group head, bit1 target head, bit2 group last, bit3 target last, bit4 kept
    uint64_t *gap, *overlap;      // '-' written before the record; projected chars trimmed from its head
    uint64_t *tail;               // group last: '-' after the last record
    uint64_t *q_src, *LQ;         // FASTA span of the query (base mode)
    uint64_t *P;                  // projected length
    uint64_t *seg_len, *seg_off;  // bytes this record contributes to the output / where they start
};
enum : uint32_t { PP_GHEAD = 1, PP_THEAD = 2, PP_GLAST = 4, PP_TLAST = 8, PP_KEPT = 16 };

// group / target boundaries in sorted order; tie[i] = position i continues a run of equal target_start inside its group
__global__ void k_pp_flags(uint64_t n, const uint64_t *__restrict__ ord, const uint64_t *__restrict__ ft, const uint64_t *__restrict__ fp,
                           const uint64_t *__restrict__ t_start, uint32_t *__restrict__ flags, uint32_t *__restrict__ ghead,
                           unsigned long long *n_ties) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { ghead[i] = 0; return; }
    const uint64_t r = ord[i];
    const bool gh = i == 0 || fp[ord[i - 1]] != fp[r], th = i == 0 || ft[ord[i - 1]] != ft[r];
    const bool gl = i + 1 == n || fp[ord[i + 1]] != fp[r], tl = i + 1 == n || ft[ord[i + 1]] != ft[r];
    flags[i] = (gh ? PP_GHEAD : 0) | (th ? PP_THEAD : 0) | (gl ? PP_GLAST : 0) | (tl ? PP_TLAST : 0);
    ghead[i] = gh ? 1u : 0u;
    if (!gh && t_start[ord[i - 1]] == t_start[r]) atomicAdd(n_ties, 1ull);
}

// slice::binary_search_by (core 1.52 .. 1.81) over [L smaller | E equal | G greater]: the index it returns
__device__ __forceinline__ uint64_t std_binary_search_pos(uint64_t L, uint64_t E, uint64_t G) {
    uint64_t size = L + E + G, left = 0, right = size;
    while (left < right) {
        const uint64_t mid = left + size / 2;
        if (mid < L) left = mid + 1;
        else if (mid >= L + E) right = mid;
        else return mid;
        size = right - left;
    }
    return left;
}
// pseudomaf.rs:88-95: the records of a (target, query) group are inserted one by one, in file order, at the index the
// binary search over target_start returns.  Records with distinct target_start end up sorted whatever the search does;
// a class of equal starts (file order in `ord`, the sorts are stable) is replayed here: the k-th record of the class
// sees L smaller and G greater records that came before it in the file, plus its k predecessors of the class, and is
// inserted in front of the class member the search hits.  One thread per tie class (classes are tiny; the counting
// loops run over the group).  ord_out = ord with every tie class permuted.
__global__ void k_pp_ties(uint64_t n, const uint64_t *__restrict__ ord, const uint64_t *__restrict__ fp, const uint64_t *__restrict__ t_start,
                          const uint32_t *__restrict__ gid, const uint64_t *__restrict__ gstart, uint64_t *__restrict__ ord_out) {
    uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const uint64_t r = ord[a], ts = t_start[r];
    const uint64_t g = gid[a + 1] - 1, ga = gstart[g], gb = gstart[g + 1];  // gid = exclusive scan of the head flags
    const bool cont = a > ga && t_start[ord[a - 1]] == ts;
    const bool more = a + 1 < gb && t_start[ord[a + 1]] == ts;
    if (!cont && !more) { ord_out[a] = r; return; }
    if (cont) return;  // the class head does the work
    uint64_t b = a + 1;
    while (b < gb && t_start[ord[b]] == ts) ++b;
    ord_out[a] = r;
    for (uint64_t k = 1; k < b - a; ++k) {
        const uint64_t f = ord[a + k];
        uint64_t L = 0, G = 0;
        for (uint64_t i = ga; i < a; ++i) L += ord[i] < f;
        for (uint64_t i = b; i < gb; ++i) G += ord[i] < f;
        const uint64_t p = std_binary_search_pos(L, k, G) - L;  // k >= 1 equal elements: the search always hits one
        for (uint64_t j = k; j > p; --j) ord_out[a + j] = ord_out[a + j - 1];
        ord_out[a + p] = f;
    }
}
__global__ void k_pp_group_starts(uint64_t n, const uint32_t *__restrict__ ghead, const uint32_t *__restrict__ gid, uint64_t *__restrict__ gstart,
                                  uint64_t n_groups) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { gstart[n_groups] = n; return; }
    if (ghead[i]) gstart[gid[i]] = i;
}

// pseudomaf.rs:108-205, one thread per (target, query) group: the only sequential state is last_target_end
__global__ void k_pp_groups(uint64_t n_groups, const uint64_t *__restrict__ gstart, const uint64_t *__restrict__ ord, PafRecs R, P2mRecs X,
                            PpRecs S, unsigned long long *err_key) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    uint64_t lte = 0, tsize = 0;
    const uint64_t i0 = gstart[g], i1 = gstart[g + 1];
    for (uint64_t i = i0; i < i1; ++i) {
        const uint64_t r = ord[i];
        const uint64_t ts = R.t_start[r], te = X.t_end[r];
        tsize = R.t_size[r];
        uint64_t gap = 0, ov = 0;
        bool keep = true;
        if (ts > lte) gap = ts - lte;
        else if (lte > te) keep = false;
        else ov = lte - ts;
        if (keep) lte = te;
        S.gap[i] = gap;
        S.overlap[i] = ov;
        if (keep) S.flags[i] |= PP_KEPT;
        S.tail[i] = 0;
    }
    if (i1 > i0) {
        if (lte > tsize) paf_report(err_key, i1 - 1, 0xFFFFFE, WGA_E_REF_PANIC);  // target_size - last_target_end underflows
        else S.tail[i1 - 1] = tsize - lte;
    }
}

// query FASTA spans of the kept records, [q_start, q_end - 1] (pseudomaf.rs:219-225); the whole target for target heads
__global__ void k_pp_spans(uint64_t n, const uint8_t *__restrict__ text, const uint64_t *__restrict__ ord, PafRecs R, P2mRecs X, PpRecs S,
                           FaView F, int base, uint64_t *__restrict__ t_src, uint64_t *__restrict__ t_cnt, unsigned long long *err_key) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t r = ord[i];
    const uint32_t fl = S.flags[i];
    S.q_src[i] = 0; S.LQ[i] = 0;
    t_src[i] = 0; t_cnt[i] = 0;
    if (fl & PP_THEAD) {  // pseudomaf.rs:121-132: the first processed record of the file fetches the target
        const uint64_t size = R.t_size[r];
        if (!base) t_cnt[i] = size;
        else if (size == 0) paf_report(err_key, i, 0, WGA_E_REF_PANIC);
        else {
            const uint64_t ti = fa_find(F, text + R.t_name_off[r], R.t_name_len[r]);
            if (ti == ~0ull) paf_report(err_key, i, 0, WGA_E_NOM);
            else {
                uint64_t first, count;
                fa_clip(F.meta[ti * kFaMeta + 4], 0, size - 1, &first, &count);
                t_src[i] = F.meta[ti * kFaMeta + 3] + first;
                t_cnt[i] = count;
            }
        }
    }
    if (!(fl & PP_KEPT) || !base) return;
    if (X.q_end[r] == 0) { paf_report(err_key, i, 1, WGA_E_REF_PANIC); return; }  // `q_end as usize - 1`
    const uint64_t qi = fa_find(F, text + R.q_name_off[r], R.q_name_len[r]);
    if (qi == ~0ull) { paf_report(err_key, i, 1, WGA_E_NOM); return; }
    uint64_t first, count;
    fa_clip(F.meta[qi * kFaMeta + 4], R.q_start[r], X.q_end[r] - 1, &first, &count);
    S.q_src[i] = F.meta[qi * kFaMeta + 3] + first;
    S.LQ[i] = count;
}

// reverse_complement validity of the kept '-' records (utils.rs:83-101; the reference walks from the END)
__global__ void k_pp_revcomp_check(uint64_t n, const uint64_t *__restrict__ ord, PafRecs R, PpRecs S, const uint8_t *__restrict__ seq,
                                   unsigned long long *err_key) {
    const uint64_t i = blockIdx.x;
    if (i >= n || !(S.flags[i] & PP_KEPT) || !(R.neg[ord[i]] & 1)) return;
    const uint8_t *p = seq + S.q_src[i];
    const uint64_t len = S.LQ[i];
    for (uint64_t k = threadIdx.x; k < len; k += blockDim.x) {
        const uint8_t c = p[len - 1 - k];
        const uint8_t u = c & 0xDFu;
        const bool ok = (u == 'A' || u == 'C' || u == 'G' || u == 'T' || u == 'N') && c >= 'A';
        if (!ok) { paf_report(err_key, i, 2 + (k > 0x7FFFF0ull ? 0x7FFFF0ull : k), WGA_E_VCF_BASE); return; }
    }
}

enum : uint8_t { PO_M = 0, PO_I = 1, PO_D = 2, PO_X = 3 };

// CIGAR of the kept records -> op count (PASS 0) / op table (PASS 1).  Tokens must be well formed (cst2cu), ops outside
// M = X I S D are ignored (cigar.rs:797).
template <int PASS>
__global__ void k_pp_ops(uint64_t n, const uint8_t *__restrict__ text, const uint64_t *__restrict__ ord, PafRecs R, PpRecs S,
                         uint64_t *__restrict__ n_ops, const uint64_t *__restrict__ op_base, uint64_t *__restrict__ op_len,
                         uint8_t *__restrict__ op_type, uint32_t *__restrict__ op_rec, uint32_t *__restrict__ op_pos,
                         unsigned long long *err_key) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { if (PASS == 0) n_ops[i] = 0; return; }
    if (!(S.flags[i] & PP_KEPT)) { if (PASS == 0) n_ops[i] = 0; return; }
    const uint64_t r = ord[i];
    const uint8_t negf = R.neg[r];
    const uint64_t len = R.cig_len[r];
    const uint8_t *p = text + R.cig_off[r];
    if ((negf & 2) || len == 0) {
        if (PASS == 0) {
            n_ops[i] = 0;
            // CigarTagNotFound / a panicking cs:Z: conversion were recorded per FILE-order record by k_paf_parse; re-report in sorted order
            if (negf & 2) paf_report(err_key, i, 0x800000, (negf & 4) ? WGA_E_REF_PANIC : WGA_E_CIGAR_TAG_NOT_FOUND);  // bit 2 left set: the cs:Z: conversion panics
            else paf_report(err_key, i, 0x800001, WGA_E_REF_PANIC);
        }
        return;
    }
    uint64_t o = PASS == 1 ? op_base[i] : 0, cnt = 0, val = 0;
    uint32_t nd = 0, op_run = 0;
    bool ovf = false, in_op = false, failed = false;
    uint8_t op = 0;
    auto close_token = [&](uint64_t pos) {
        if (failed) return;
        int code = 0;
        if (op_run != 1) code = WGA_E_CIGAR_OP_INVALID;
        else if (nd == 0 || ovf) code = WGA_E_PARSE_INT;
        if (code) {
            failed = true;
            if (PASS == 0) paf_report(err_key, i, 0x800002 + (pos > 0x7FFFF0ull ? 0x7FFFF0ull : pos), code);
            return;
        }
        uint8_t t;
        if (op == 'M' || op == '=') t = PO_M;
        else if (op == 'X') t = PO_X;
        else if (op == 'I' || op == 'S') t = PO_I;
        else if (op == 'D') t = PO_D;
        else return;
        if (PASS == 1) {
            op_len[o] = val; op_type[o] = t; op_rec[o] = (uint32_t)i;
            op_pos[o] = (uint32_t)(pos > 0x7FFFF0ull ? 0x7FFFF0ull : pos);
            ++o;
        }
        ++cnt;
    };
    for (uint64_t k = 0; k < len && !failed; ++k) {
        const uint8_t ch = p[k];
        if (ch >= '0' && ch <= '9') {
            if (in_op) { close_token(k); val = 0; nd = 0; ovf = false; op_run = 0; in_op = false; }
            const uint32_t d = ch - '0';
            if (val > (0xFFFFFFFFFFFFFFFFull - d) / 10) ovf = true;
            val = val * 10 + d;
            ++nd;
        } else { in_op = true; ++op_run; op = ch; }
    }
    if (!failed) close_token(len);
    if (PASS == 0) n_ops[i] = failed ? 0 : cnt;
}

// per-op contributions: c_out = projected chars produced (M,=,X,D), c_src = query bases consumed (M,=,X,I,S)
__global__ void k_pp_contrib(uint64_t n_ops, const uint64_t *__restrict__ op_len, const uint8_t *__restrict__ op_type,
                             uint64_t *__restrict__ c_out, uint64_t *__restrict__ c_src) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n_ops) return;
    if (k == n_ops) { c_out[k] = c_src[k] = 0; return; }
    const uint64_t l = op_len[k];
    const uint8_t t = op_type[k];
    c_out[k] = t == PO_I ? 0 : l;
    c_src[k] = t == PO_D ? 0 : l;
}

struct PpView {
    uint64_t n, n_ops;
    const uint8_t *text;
    const uint64_t *ord;
    PafRecs R;
    P2mRecs X;
    PpRecs S;
    const uint64_t *op_base;  // [n+1]
    const uint64_t *op_len;
    const uint8_t *op_type;
    const uint32_t *op_rec, *op_pos;
    const uint64_t *S_out, *S_src;  // exclusive scans over all ops [n_ops+1]
    const uint8_t *seq;             // FASTA bases (base mode)
    const uint64_t *t_src, *t_cnt;  // per sorted record (target heads only)
    int base;
    char *out;
};

// String::drain / insert_str past the end are the reference's panics (cigar.rs:775,783)
__global__ void k_pp_check(PpView V, unsigned long long *err_key) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= V.n_ops || !V.base) return;
    const uint8_t t = V.op_type[k];
    if (t != PO_I && t != PO_D) return;
    const uint32_t i = V.op_rec[k];
    const uint64_t src = V.S_src[k] - V.S_src[V.op_base[i]], LQ = V.S.LQ[i];
    const bool bad = t == PO_I ? (V.op_len[k] > LQ || src > LQ - V.op_len[k]) : src > LQ;
    if (bad) paf_report(err_key, i, 0x800002 + V.op_pos[k], WGA_E_REF_PANIC);
}

__device__ __forceinline__ uint32_t pp_hdr_len(uint32_t name_len, uint64_t size) {  // "s\t{name}\t0\t{size}\t+\t{size}\t"
    return 2 + name_len + 3 + ndigits_u64(size) + 3 + ndigits_u64(size) + 1;
}

// projected length, head-trim check, bytes per sorted record
__global__ void k_pp_layout(PpView V, unsigned long long *err_key) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > V.n) return;
    if (i == V.n) { V.S.seg_len[i] = 0; return; }
    const uint64_t r = V.ord[i];
    const uint32_t fl = V.S.flags[i];
    uint64_t bytes = 0, P = 0;
    if (fl & PP_KEPT) {
        const uint64_t b0 = V.op_base[i], b1 = V.op_base[i + 1];
        const uint64_t outc = V.S_out[b1] - V.S_out[b0], srcc = V.S_src[b1] - V.S_src[b0];
        // with bases: LQ - deleted + inserted = LQ - (src - kept) + (out - kept) = LQ + out - src
        P = V.base ? V.S.LQ[i] + outc - srcc : outc;
        if (V.S.overlap[i] > P) { paf_report(err_key, i, 0xFFFFFD, WGA_E_REF_PANIC); P = V.S.overlap[i]; }  // q_seq.drain(0..overlap)
        bytes += V.S.gap[i] + P - V.S.overlap[i];
    }
    V.S.P[i] = P;
    if (fl & PP_THEAD) bytes += 10 + pp_hdr_len(V.R.t_name_len[r], V.R.t_size[r]) + V.t_cnt[i] + 1;
    if (fl & PP_GHEAD) bytes += pp_hdr_len(V.R.q_name_len[r], V.R.q_size[r]);
    if (fl & PP_GLAST) bytes += V.S.tail[i] + 1;
    if (fl & PP_TLAST) bytes += 1;
    V.S.seg_len[i] = bytes;
}

__device__ __forceinline__ char *pp_put_hdr(char *o, const uint8_t *name, uint32_t nl, uint64_t size) {
    *o++ = 's'; *o++ = '\t';
    for (uint32_t k = 0; k < nl; ++k) *o++ = (char)name[k];
    *o++ = '\t'; *o++ = '0'; *o++ = '\t';
    o = put_u64(o, size);
    *o++ = '\t'; *o++ = '+'; *o++ = '\t';
    o = put_u64(o, size);
    *o++ = '\t';
    return o;
}
// offset of the record's piece (gap + projected chars) inside its segment
__device__ __forceinline__ uint64_t pp_piece_off(const PpView &V, uint64_t i) {
    const uint64_t r = V.ord[i];
    const uint32_t fl = V.S.flags[i];
    uint64_t o = V.S.seg_off[i];
    if (fl & PP_THEAD) o += 10 + pp_hdr_len(V.R.t_name_len[r], V.R.t_size[r]) + V.t_cnt[i] + 1;
    if (fl & PP_GHEAD) o += pp_hdr_len(V.R.q_name_len[r], V.R.q_size[r]);
    return o;
}

// fixed text: file head, s-line prefixes, line ends (the output was filled with '-' before)
__global__ void k_pp_fixed(PpView V) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V.n) return;
    const uint64_t r = V.ord[i];
    const uint32_t fl = V.S.flags[i];
    char *o = V.out + V.S.seg_off[i];
    if (fl & PP_THEAD) {
        const char a[10] = {'a', ' ', 's', 'c', 'o', 'r', 'e', '=', '0', '\n'};
        for (int k = 0; k < 10; ++k) *o++ = a[k];
        o = pp_put_hdr(o, V.text + V.R.t_name_off[r], V.R.t_name_len[r], V.R.t_size[r]);
        o += V.t_cnt[i];
        *o++ = '\n';
    }
    if (fl & PP_GHEAD) o = pp_put_hdr(o, V.text + V.R.q_name_off[r], V.R.q_name_len[r], V.R.q_size[r]);
    if (fl & PP_KEPT) o += V.S.gap[i] + V.S.P[i] - V.S.overlap[i];
    if (fl & PP_GLAST) { o += V.S.tail[i]; *o++ = '\n'; }
    if (fl & PP_TLAST) *o++ = '\n';
}

// whole target sequence (or 'N' x size) of every file: blockIdx.y walks the target heads
__global__ void k_pp_target_rows(PpView V, const uint64_t *__restrict__ thead_pos, uint64_t n_targets) {
    const uint64_t t = blockIdx.y;
    if (t >= n_targets) return;
    const uint64_t i = thead_pos[t], r = V.ord[i];
    char *dst = V.out + V.S.seg_off[i] + 10 + pp_hdr_len(V.R.t_name_len[r], V.R.t_size[r]);
    const uint64_t cnt = V.t_cnt[i];
    const uint8_t *src = V.base ? V.seq + V.t_src[i] : nullptr;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < cnt; b += (uint64_t)gridDim.x * blockDim.x)
        dst[b] = src ? (char)src[b] : 'N';
}

// one warp per CIGAR op (+ one virtual op per sorted record: the query bases left over after the last op); with a FASTA
// the warp of the first op of every gap-free stretch copies the whole stretch (16-byte stores)
__global__ void k_pp_rows(PpView V) {
    const uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (wid >= V.n_ops + V.n) return;
    uint64_t i, o, src, len;
    uint8_t type;
    if (wid < V.n_ops) {
        i = V.op_rec[wid];
        const uint64_t b0 = V.op_base[i];
        o = V.S_out[wid] - V.S_out[b0];
        src = V.S_src[wid] - V.S_src[b0];
        len = V.op_len[wid];
        type = V.op_type[wid];
        if (type == PO_I || type == PO_D) return;  // nothing to write: deleted bases / '-' is the fill value
        if (V.base) {
            // with a FASTA the M / = / X ops between two gap ops copy ONE contiguous stretch of the query to ONE contiguous
            // stretch of the row: the first op of the stretch copies all of it, the others have nothing left to do
            if (wid > b0) { const uint8_t pt = V.op_type[wid - 1]; if (pt != PO_I && pt != PO_D) return; }
            const uint64_t b1 = V.op_base[i + 1];
            uint64_t g = wid + 1;
            for (;; g += 32) {  // first gap op (or the record's end) after this op
                const uint64_t k = g + lane;
                const bool stop = k >= b1 || V.op_type[k] == PO_I || V.op_type[k] == PO_D;
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, stop);
                if (m) { g += (uint32_t)__ffs(m) - 1u; break; }
            }
            len = V.S_out[g] - V.S_out[wid];
        }
    } else {
        i = wid - V.n_ops;
        if (!V.base || !(V.S.flags[i] & PP_KEPT)) return;
        const uint64_t b0 = V.op_base[i], b1 = V.op_base[i + 1];
        o = V.S_out[b1] - V.S_out[b0];
        src = V.S_src[b1] - V.S_src[b0];
        len = ~0ull;
        type = PO_M;
    }
    const uint64_t ov = V.S.overlap[i];
    char *piece = V.out + pp_piece_off(V, i) + V.S.gap[i];  // projected char p (>= ov) lives at piece[p - ov]
    uint64_t n = len;
    if (V.base) {
        const uint64_t LQ = V.S.LQ[i];
        if (src >= LQ) return;
        if (LQ - src < n) n = LQ - src;
    }
    uint64_t j0 = o < ov ? ov - o : 0;  // first char of this op that survives the head trim
    if (!V.base) {
        const char c = type == PO_X ? '0' : '1';
        for (uint64_t j = j0 + lane; j < n; j += 32) piece[o + j - ov] = c;
        return;
    }
    const uint8_t *Q = V.seq + V.S.q_src[i];
    const uint64_t LQ = V.S.LQ[i];
    if (n <= j0) return;
    if (!(V.R.neg[V.ord[i]] & 1)) warp_copy16(piece + o + j0 - ov, Q + src + j0, n - j0, lane);
    else warp_revcomp16(piece + o + j0 - ov, Q + LQ - (src + j0), n - j0, lane);  // bases validated by k_pp_revcomp_check
}

// per target file: name and byte range, for the host-side index
__global__ void k_pp_files(PpView V, const uint64_t *__restrict__ thead_pos, uint64_t n_targets, uint64_t total,
                           uint64_t *__restrict__ info /* n_targets x 4: name_off, name_len, out_off, out_len */) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_targets) return;
    const uint64_t i = thead_pos[t], r = V.ord[i];
    const uint64_t start = V.S.seg_off[i], end = t + 1 < n_targets ? V.S.seg_off[thead_pos[t + 1]] : total;
    info[4 * t] = V.R.t_name_off[r];
    info[4 * t + 1] = V.R.t_name_len[r];
    info[4 * t + 2] = start;
    info[4 * t + 3] = end - start;
}
__global__ void k_pp_thead_pos(uint64_t n, const uint32_t *__restrict__ flags, const uint32_t *__restrict__ tid, uint64_t *__restrict__ thead_pos) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (flags[i] & PP_THEAD)) thead_pos[tid[i]] = i;
}
__global__ void k_pp_thead_flag(uint64_t n, const uint32_t *__restrict__ flags, uint32_t *__restrict__ f) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    f[i] = i < n && (flags[i] & PP_THEAD) ? 1u : 0u;
}

}  // namespace wga
