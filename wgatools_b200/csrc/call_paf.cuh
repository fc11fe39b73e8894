// call_paf.cuh -- `call -f paf`: PAF + 2 FASTA -> VCF rows.
//
// Replaces src/tools/caller.rs:268-302 call_var_paf and :610-822 call_within_var_paf (rows formatted as
// get_variant_rec + the noodles writer do, caller.rs:359-386).  Differences from the MAF caller that parity must keep
// (SURVEY Q12, Q13, Q15): sequences are fetched with an INCLUSIVE end ([start, end], one base more than the
// alignment), the query is NOT reverse-complemented for '-', the <INV> row is not gated by --inv, there is no
// chunking, and the Result of the CIGAR fold closure is thrown away: the first failing token or row of a record
// silently ends that record (rows before it are kept), while slices past the fetched strings are panics.
//
// The CIGAR ops are the runs: per-op prefix sums give t_pos / q_pos, `after_m` is "the previous op is M/=/X", every
// row is sized and written independently.  The sequential "stop at the first failure" rule is resolved with two
// per-record minima over row indices (first row with a bad base, first row that would panic).
#pragma once
#include "call.cuh"
#include "paf2maf.cuh"

namespace wga {

enum : uint8_t { CP_M = 0, CP_I = 1, CP_D = 2, CP_X = 3, CP_INV = 4 };

// CIGAR -> op count (PASS 0) / op table (PASS 1).  One leading virtual op per '-' record (its <INV> row).
// A token that cst2cu rejects, or an op outside M=XID, ends the record's op list without an error (caller.rs:673-815).
template <int PASS>
__global__ void k_cp_ops(uint64_t n_rec, const uint8_t *__restrict__ text, PafRecs R, uint64_t *__restrict__ n_ops,
                         const uint64_t *__restrict__ op_base, uint64_t *__restrict__ op_len, uint8_t *__restrict__ op_type,
                         uint32_t *__restrict__ op_rec, unsigned long long *err_key) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rec) return;
    if (r == n_rec) { if (PASS == 0) n_ops[r] = 0; return; }
    const uint64_t len = R.cig_len[r];
    const uint8_t *p = text + R.cig_off[r];
    const uint8_t negf = R.neg[r];
    uint64_t o = PASS == 1 ? op_base[r] : 0, cnt = 0;
    if (negf & 1) {
        if (PASS == 1) { op_len[o] = 1; op_type[o] = CP_INV; op_rec[o] = (uint32_t)r; ++o; }
        ++cnt;
    }
    if ((negf & 2) || len == 0) {
        if (PASS == 0) {
            n_ops[r] = cnt;
            if (!(negf & 2) && len == 0) paf_report(err_key, r, 0x800001, WGA_E_REF_PANIC);  // fold_many1 on "" (errors.rs:88-97)
        }
        return;
    }
    uint64_t val = 0;
    uint32_t nd = 0, op_run = 0;
    bool ovf = false, in_op = false, stopped = false;
    uint8_t op = 0;
    auto close_token = [&]() {
        if (stopped) return;
        uint8_t t;
        if (op_run != 1 || nd == 0 || ovf) { stopped = true; return; }
        if (op == 'M' || op == '=') t = CP_M;
        else if (op == 'X') t = CP_X;
        else if (op == 'I') t = CP_I;
        else if (op == 'D') t = CP_D;
        else { stopped = true; return; }
        if (PASS == 1) { op_len[o] = val; op_type[o] = t; op_rec[o] = (uint32_t)r; ++o; }
        ++cnt;
    };
    for (uint64_t i = 0; i < len && !stopped; ++i) {
        const uint8_t ch = p[i];
        if (ch >= '0' && ch <= '9') {
            if (in_op) { close_token(); val = 0; nd = 0; ovf = false; op_run = 0; in_op = false; }
            const uint32_t d = ch - '0';
            if (val > (0xFFFFFFFFFFFFFFFFull - d) / 10) ovf = true;
            val = val * 10 + d;
            ++nd;
        } else { in_op = true; ++op_run; op = ch; }
    }
    if (!stopped) close_token();
    if (PASS == 0) n_ops[r] = cnt;
}

// the <INV> row is built before the CIGAR is even looked up (caller.rs:641-660): its failures are real errors and precede
// CigarTagNotFound / the tokenizer's
__global__ void k_cp_inv_check(uint64_t n_rec, PafRecs R, P2mRecs X, const uint8_t *__restrict__ tseq, unsigned long long *err_key) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec || !(R.neg[r] & 1)) return;
    char c;
    if (X.LT[r] < 1) paf_report(err_key, r, 2, WGA_E_REF_PANIC);  // &t_seq_string[0..1] on ""
    else if (!vcf_base(tseq[X.t_src[r]], &c)) paf_report(err_key, r, 2, WGA_E_VCF_BASE);
}

struct CpView {
    uint64_t n_rec, n_ops;
    const uint8_t *text;
    PafRecs R;
    P2mRecs X;
    const uint64_t *op_base;   // [n_rec+1]
    const uint64_t *op_len;
    const uint8_t *op_type;
    const uint32_t *op_rec;
    const uint64_t *St, *Sq;   // exclusive scans of target / query bases consumed, over all ops [n_ops+1]
    const uint64_t *rowbase;   // exclusive scan of rows per op [n_ops+1]
    const uint8_t *tseq, *qseq;
    uint64_t svlen;
    int snp;
};

// bases consumed per op (for the prefix sums) and rows emitted per op
__global__ void k_cp_contrib(uint64_t n_ops, const uint64_t *__restrict__ op_len, const uint8_t *__restrict__ op_type,
                             const uint32_t *__restrict__ op_rec, const uint64_t *__restrict__ op_base, uint64_t svlen, int snp,
                             uint64_t *__restrict__ tcon, uint64_t *__restrict__ qcon, uint64_t *__restrict__ rows) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n_ops) return;
    if (k == n_ops) { tcon[k] = qcon[k] = rows[k] = 0; return; }
    const uint64_t l = op_len[k];
    const uint8_t t = op_type[k];
    tcon[k] = (t == CP_M || t == CP_X || t == CP_D) ? l : 0;
    qcon[k] = (t == CP_M || t == CP_X || t == CP_I) ? l : 0;
    uint64_t nrow = 0;
    if (t == CP_INV) nrow = 1;
    else if (t == CP_X) nrow = snp ? l : 0;
    else if (t == CP_I || t == CP_D) {
        // after_m: the previous op of the same record is M/=/X (caller.rs:688,715,721-729,765,772-789)
        const bool after_m = k > op_base[op_rec[k]] && (op_type[k - 1] == CP_M || op_type[k - 1] == CP_X);
        nrow = (l > svlen && after_m) ? 1 : 0;
    }
    rows[k] = nrow;
}

constexpr uint32_t kCpNoRow = 0xFFFFFFFFu;

// one VCF row per thread.  PASS 0: row_len + per-record minima (first bad-base row / first panicking row);
// PASS 1: write the rows that survive.
template <int PASS>
__global__ void k_cp_rows(CpView V, uint64_t n_rows, uint64_t *__restrict__ row_len, const uint64_t *__restrict__ row_off,
                          uint32_t *__restrict__ rec_bad, uint32_t *__restrict__ rec_panic, char *__restrict__ out,
                          unsigned long long *err_key) {
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (PASS == 0) {
        if (row > n_rows) return;
        if (row == n_rows) { row_len[row] = 0; return; }
    } else {
        if (row >= n_rows) return;
        if (row_off[row + 1] == row_off[row]) return;  // dropped
    }
    // op owning this row: last k with rowbase[k] <= row
    uint64_t lo = 0, hi = V.n_ops;
    while (hi - lo > 1) { uint64_t m = (lo + hi) >> 1; if (V.rowbase[m] <= row) lo = m; else hi = m; }
    const uint64_t k = lo, i = row - V.rowbase[k];
    const uint32_t r = V.op_rec[k];
    const uint64_t b0 = V.op_base[r];
    const uint64_t row_in_rec = row - V.rowbase[b0];
    const uint32_t rir = row_in_rec > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)row_in_rec;
    const uint8_t type = V.op_type[k];
    const uint64_t len = V.op_len[k];
    const bool neg = V.R.neg[r] & 1;
    const uint64_t t_st = V.R.t_start[r], q_st = V.R.q_start[r];
    const uint64_t LT = V.X.LT[r], LQ = V.X.LQ[r];
    const uint8_t *T = V.tseq + V.X.t_src[r], *Q = V.qseq + V.X.q_src[r];
    const uint64_t a = V.St[k] - V.St[b0] + (type == CP_X ? i : 0);  // t_pos - t_start
    const uint64_t b = V.Sq[k] - V.Sq[b0] + (type == CP_X ? i : 0);
    const uint8_t *qn = V.text + V.R.q_name_off[r];
    const uint32_t qnl = V.R.q_name_len[r];
    const char sfx = neg ? 'N' : 'P';
    RowOut O;
    O.o = PASS == 1 ? out + row_off[row] : nullptr;
    O.n = 0;
    O.bad_base = false;
    bool panic = false;
    O.bytes(V.text + V.R.t_name_off[r], V.R.t_name_len[r]);
    O.ch('\t');
    if (type == CP_INV) {  // caller.rs:641-658
        if (LT < 1) panic = true;
        else {
            O.u64(t_st + 1);
            O.lit("\t.\t");
            O.bases(T, 1);
            O.lit("\t<INV>\t.\t.\tSVTYPE=INV;END=");
            O.u64(V.X.t_end[r]);
            O.lit("\tGT:QI\t1|1:");
            O.bytes(qn, qnl);
            O.ch('@'); O.u64(q_st); O.ch('@'); O.u64(V.X.q_end[r]);
            O.lit("@N\n");
        }
    } else if (type == CP_X) {  // caller.rs:690-716
        if (a + 1 > LT || b + 1 > LQ) panic = true;
        else {
            O.u64(t_st + a + 1);
            O.lit("\t.\t");
            O.bases(T + a, 1);
            O.ch('\t');
            O.bases(Q + b, 1);
            O.lit("\t.\t.\t.\tGT:QI\t1|1:");
            O.bytes(qn, qnl);
            O.ch('@'); O.u64(q_st + b); O.ch('@'); O.ch(sfx); O.ch('\n');
        }
    } else {  // I / D with an anchor base (caller.rs:718-808)
        const bool ins = type == CP_I;
        if (a == 0 || b == 0) panic = true;  // `pos - start - 1` underflows
        else {
            const uint64_t ta = a - 1, qa = b - 1;
            const uint64_t tn = ins ? 1 : len + 1, qn_ = ins ? len + 1 : 1;
            if (len > (ins ? LQ : LT) || ta + tn > LT || qa + qn_ > LQ) panic = true;
            else {
                O.u64(t_st + a);
                O.lit("\t.\t");
                O.bases(T + ta, tn);
                O.ch('\t');
                O.bases(Q + qa, qn_);
                O.lit("\t.\t.\t");
                if (neg) O.lit("INV_NEST=TRUE;");
                if (ins) O.lit("SVTYPE=INS;SVLEN="); else O.lit("SVTYPE=DEL;SVLEN=");
                O.u64(len);
                O.lit(";END=");
                O.u64(ins ? t_st + a : t_st + a + len);
                O.lit("\tGT:QI\t1|1:");
                O.bytes(qn, qnl);
                O.ch('@'); O.u64(q_st + b); O.ch('@'); O.u64(ins ? q_st + b + len : q_st + b);
                O.ch('@'); O.ch(sfx); O.ch('\n');
            }
        }
    }
    if (PASS == 0) {
        row_len[row] = panic ? 0 : O.n;
        if (type == CP_INV) {  // outside the fold closure: both are real errors, before any CIGAR-level one
            if (panic) paf_report(err_key, r, 2, WGA_E_REF_PANIC);
            else if (O.bad_base) paf_report(err_key, r, 2, WGA_E_VCF_BASE);
        } else {
            if (panic) atomicMin(rec_panic + r, rir);
            else if (O.bad_base) atomicMin(rec_bad + r, rir);
        }
    }
}

// caller.rs:673-815: the first failing row ends the record silently; a panic before it aborts the command
__global__ void k_cp_resolve(CpView V, uint64_t n_rows, uint64_t *__restrict__ row_len, const uint32_t *__restrict__ rec_bad,
                             const uint32_t *__restrict__ rec_panic, unsigned long long *err_key) {
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    uint64_t lo = 0, hi = V.n_ops;
    while (hi - lo > 1) { uint64_t m = (lo + hi) >> 1; if (V.rowbase[m] <= row) lo = m; else hi = m; }
    const uint32_t r = V.op_rec[lo];
    const uint64_t row_in_rec = row - V.rowbase[V.op_base[r]];
    const uint32_t bad = rec_bad[r], pan = rec_panic[r];
    if (row_in_rec >= bad) row_len[row] = 0;
    if (pan < bad && row_in_rec == pan) paf_report(err_key, r, 0x800002ull + (pan > 0x7FFFF0u ? 0x7FFFF0u : pan), WGA_E_REF_PANIC);
}

}  // namespace wga
