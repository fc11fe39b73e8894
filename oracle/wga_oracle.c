/*
 * wga_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See wga_oracle.h.
 *
 * Literal sequential restatement of wgatools v1.1.0 (/root/reference, commit de13560d):
 *   src/parser/cigar.rs, src/parser/common.rs, src/parser/maf.rs, src/parser/paf.rs,
 *   src/converter.rs:29-54, src/tools/stat.rs, src/tools/caller.rs:42-265,304-608.
 * Written for obviousness, not speed: it mirrors the reference's control flow
 * (line reader -> s-line parser -> per-column group_by -> string building).
 * Integer arithmetic wraps like a Rust release build (no overflow-checks profile
 * in Cargo.toml).  Input is treated as ASCII (the reference iterates chars()).
 */
#include "wga_oracle.h"
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ buffers */
static void buf_reserve(orc_buf *b, size_t extra) {
    if (b->n + extra <= b->cap) return;
    size_t cap = b->cap ? b->cap : 256;
    while (cap < b->n + extra) cap *= 2;
    b->p = (char *)realloc(b->p, cap);
    b->cap = cap;
}
static void buf_put(orc_buf *b, const char *s, size_t n) {
    buf_reserve(b, n);
    memcpy(b->p + b->n, s, n);
    b->n += n;
}
static void buf_putc(orc_buf *b, char c) { buf_put(b, &c, 1); }
static void buf_puts(orc_buf *b, const char *s) { buf_put(b, s, strlen(s)); }
static void buf_putu(orc_buf *b, uint64_t v) { /* itoa */
    char tmp[24];
    int n = snprintf(tmp, sizeof tmp, "%llu", (unsigned long long)v);
    buf_put(b, tmp, (size_t)n);
}
void orc_buf_free(orc_buf *b) {
    free(b->p);
    b->p = NULL;
    b->n = b->cap = 0;
}

static void set_err(char *err, size_t cap, const char *fmt, ...) {
    if (!err || !cap) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, cap, fmt, ap);
    va_end(ap);
}

typedef struct { const char *p; size_t n; } sv;
static int sv_eq(sv a, sv b) { return a.n == b.n && memcmp(a.p, b.p, a.n) == 0; }
static int sv_eq_c(sv a, const char *c) { size_t n = strlen(c); return a.n == n && memcmp(a.p, c, n) == 0; }

/* ---------------------------------------------------------- number parsing */
/* utils.rs:69-74 parse_str2u64 == str::parse::<u64>: optional '+', >=1 digit, no overflow */
static int parse_u64(sv s, uint64_t *out) {
    size_t i = 0;
    if (s.n == 0) return 0;
    if (s.p[0] == '+') i = 1;
    if (i == s.n) return 0;
    uint64_t v = 0;
    for (; i < s.n; i++) {
        unsigned d = (unsigned char)s.p[i] - '0';
        if (d > 9) return 0;
        if (v > (UINT64_MAX - d) / 10) return 0;
        v = v * 10 + d;
    }
    *out = v;
    return 1;
}

/* ------------------------------------------------------------- line reader */
/* std::io::BufRead::lines(): split at '\n', drop it and a directly preceding '\r'. */
typedef struct { const char *t; size_t len, pos; } lreader;
static int next_line(lreader *r, sv *line) {
    if (r->pos >= r->len) return 0;
    const char *s = r->t + r->pos;
    const char *nl = (const char *)memchr(s, '\n', r->len - r->pos);
    size_t n;
    if (nl) {
        n = (size_t)(nl - s);
        r->pos += n + 1;
        if (n > 0 && s[n - 1] == '\r') n--;
    } else {
        n = r->len - r->pos;
        r->pos = r->len;
    }
    line->p = s;
    line->n = n;
    return 1;
}

/* char::is_whitespace restricted to ASCII: U+0009..U+000D, U+0020 */
static int is_ws(char c) { return c == ' ' || (c >= 9 && c <= 13); }

/* ------------------------------------------------------------------ s-line */
typedef struct {
    char mode;
    sv name;
    uint64_t start, align_size, size;
    int neg;
    sv seq;
} sline;

/* maf.rs:138-211 parse_sline: split_whitespace, exactly 7 fields */
static int parse_sline(sv line, sline *out, char *err, size_t errcap) {
    static const char *fname[7] = {"mode", "name", "start", "align_size", "strand", "size", "seq"};
    sv f[8];
    int nf = 0;
    size_t i = 0;
    while (i < line.n && nf < 8) {
        while (i < line.n && is_ws(line.p[i])) i++;
        if (i >= line.n) break;
        size_t b = i;
        while (i < line.n && !is_ws(line.p[i])) i++;
        f[nf].p = line.p + b;
        f[nf].n = i - b;
        nf++;
    }
    /* fields are consumed strictly in order, so the first problem in field order wins */
    for (int k = 0; k < 7; k++) {
        if (k >= nf) {
            set_err(err, errcap, "Parse MAF error by: S-line Filed `%s` Missing", fname[k]);
            return ORC_E_MAF_FIELD_MISSING;
        }
        switch (k) {
        case 0: out->mode = f[0].p[0]; break;
        case 1: out->name = f[1]; break;
        case 2:
            if (!parse_u64(f[2], &out->start)) goto bad_int;
            break;
        case 3:
            if (!parse_u64(f[3], &out->align_size)) goto bad_int;
            break;
        case 4:
            if (sv_eq_c(f[4], "+")) out->neg = 0;
            else if (sv_eq_c(f[4], "-")) out->neg = 1;
            else {
                set_err(err, errcap, "Parse Strand `%.*s` Error", (int)f[4].n, f[4].p);
                return ORC_E_PARSE_STRAND;
            }
            break;
        case 5:
            if (!parse_u64(f[5], &out->size)) goto bad_int;
            break;
        case 6: out->seq = f[6]; break;
        }
        continue;
    bad_int:
        set_err(err, errcap, "Parse `%.*s` Into Integer Error", (int)f[k].n, f[k].p);
        return ORC_E_PARSE_INT;
    }
    if (nf > 7) {
        set_err(err, errcap, "Parse MAF error by: Surplus Filed > 7");
        return ORC_E_MAF_SURPLUS_FIELD;
    }
    return ORC_OK;
}

/* -------------------------------------------------------------- MAF record */
typedef struct {
    sline *s;
    size_t n, cap;
    size_t query_idx;
} mafrec;

static void mafrec_push(mafrec *r, const sline *s) {
    if (r->n == r->cap) {
        r->cap = r->cap ? r->cap * 2 : 4;
        r->s = (sline *)realloc(r->s, r->cap * sizeof(sline));
    }
    r->s[r->n++] = *s;
}

/* maf.rs:25-36 MAFReader::new: the first line is always consumed as the header */
static void maf_reader_init(lreader *r, const char *text, size_t len) {
    r->t = text;
    r->len = len;
    r->pos = 0;
    sv hdr;
    next_line(r, &hdr);
}

/* maf.rs:371-421 MAFRecords::next.  Returns 1 record, 0 end, <0 error (negated code). */
static int maf_next(lreader *r, mafrec *rec, char *err, size_t errcap) {
    sv line;
    for (;;) {
        if (!next_line(r, &line)) return 0;
        if (line.n > 0 && line.p[0] == 's') break; /* maf.rs:378-379 skip non-'s' lines */
    }
    rec->n = 0;
    rec->query_idx = 1;
    sline s;
    int rc = parse_sline(line, &s, err, errcap);
    if (rc) return -rc;
    mafrec_push(rec, &s);
    while (next_line(r, &line)) { /* maf.rs:395-413 */
        if (line.n > 0 && line.p[0] == 's') {
            rc = parse_sline(line, &s, err, errcap);
            if (rc) return -rc;
            mafrec_push(rec, &s);
        } else {
            break; /* the terminating line is consumed */
        }
    }
    return 1;
}

/* maf.rs:277-285 set_query_idx_byname */
static int maf_set_query_byname(mafrec *rec, const char *qname, char *err, size_t errcap) {
    for (size_t i = 0; i < rec->n; i++)
        if (sv_eq_c(rec->s[i].name, qname)) {
            rec->query_idx = i;
            return ORC_OK;
        }
    set_err(err, errcap, "Query name:%s not found in MAF", qname);
    return ORC_E_QUERY_NAME_NOT_FOUND;
}

/* maf.rs:433-450: forward-strand query coordinates (wrapping u64 like release Rust) */
static uint64_t maf_query_start(const mafrec *r) {
    const sline *q = &r->s[r->query_idx];
    return q->neg ? q->size - q->start - q->align_size : q->start;
}
static uint64_t maf_query_end(const mafrec *r) {
    const sline *q = &r->s[r->query_idx];
    return q->neg ? q->size - q->start : q->start + q->align_size;
}

/* ----------------------------------------------------- column classifiers */
/* cigar.rs:298-308 */
static char cigar_cat_ext(char c1, char c2) {
    if (c1 == c2) return '=';
    if (c1 == '-') return 'I';
    if (c2 == '-') return 'D';
    return 'X';
}
/* cigar.rs:314-328 */
static char cigar_cat_ext_caller(char c1, char c2) {
    if (c1 == '-') return c2 == '-' ? 'W' : 'I';
    if (c2 == '-') return 'D';
    if (c1 == c2) return '=';
    return 'X';
}

/* cigar.rs:344-432 */
void orc_parse_maf_seq_to_cigar(const char *t, const char *q, size_t n, int neg, orc_counts *out,
                                orc_buf *cigar) {
    uint64_t *c = out->c;
    memset(c, 0, sizeof out->c);
    if (neg) c[10] = 1; /* inv_event, cigar.rs:370-376 */
    size_t i = 0;
    while (i < n) { /* group_by(cigar_cat_ext) */
        char k = cigar_cat_ext(t[i], q[i]);
        size_t j = i + 1;
        while (j < n && cigar_cat_ext(t[j], q[j]) == k) j++;
        uint64_t len = j - i;
        switch (k) {
        case '=': c[0] += len; break;
        case 'X': c[1] += len; break;
        case 'I':
            if (neg) { c[6] += 1; c[7] += len; } else { c[2] += 1; c[3] += len; }
            break;
        case 'D':
            if (neg) { c[8] += 1; c[9] += len; } else { c[4] += 1; c[5] += len; }
            break;
        }
        if (cigar) { /* cigar.rs:409-410 */
            buf_putu(cigar, len);
            buf_putc(cigar, k);
        }
        i = j;
    }
}

/* cigar.rs:43-75 tokenizer + cigar.rs:629-707 fold */
int orc_parse_paf_cigar(const char *cg, size_t n, int neg, orc_counts *out, char *err,
                        size_t errcap) {
    uint64_t *c = out->c;
    memset(c, 0, sizeof out->c);
    if (neg) c[10] = 1;
    size_t i = 0;
    if (n == 0) { /* fold_many1 needs one success; Eof on empty input -> nom error -> e.input[..10] panics */
        set_err(err, errcap, "panic: byte index 10 is out of bounds of ``");
        return ORC_E_PANIC;
    }
    while (i < n) {
        size_t d0 = i;
        while (i < n && cg[i] >= '0' && cg[i] <= '9') i++; /* take_while(is_dec_digit) */
        size_t o0 = i;
        while (i < n && !(cg[i] >= '0' && cg[i] <= '9')) i++; /* take_till(is_dec_digit) */
        sv digits = {cg + d0, o0 - d0}, op = {cg + o0, i - o0};
        /* cst2cu: op must be exactly one char, then the length must parse */
        if (op.n != 1) {
            set_err(err, errcap, "CIGAR OP `%.*s` invalid", (int)op.n, op.p);
            return ORC_E_CIGAR_OP_INVALID;
        }
        uint64_t len;
        if (!parse_u64(digits, &len)) {
            set_err(err, errcap, "Parse `%.*s` Into Integer Error", (int)digits.n, digits.p);
            return ORC_E_PARSE_INT;
        }
        switch (op.p[0]) {
        case 'M': case '=': c[0] += len; break;
        case 'X': c[1] += len; break;
        case 'I':
            if (neg) { c[6] += 1; c[7] += len; } else { c[2] += 1; c[3] += len; }
            break;
        case 'D':
            if (neg) { c[8] += 1; c[9] += len; } else { c[4] += 1; c[5] += len; }
            break;
        default:
            set_err(err, errcap, "CIGAR OP `%c` invalid", op.p[0]);
            return ORC_E_CIGAR_OP_INVALID;
        }
    }
    return ORC_OK;
}

/* ------------------------------------------------------------- csv writer */
/* csv 1.2.2 Writer, QuoteStyle::Necessary: quote a field iff it contains the
 * delimiter, a quote, CR or LF; quotes are doubled.  (parity unpinned) */
static void csv_field(orc_buf *b, sv f, int first) {
    if (!first) buf_putc(b, '\t');
    int need = 0;
    for (size_t i = 0; i < f.n; i++)
        if (f.p[i] == '\t' || f.p[i] == '"' || f.p[i] == '\r' || f.p[i] == '\n') need = 1;
    if (!need) {
        buf_put(b, f.p, f.n);
        return;
    }
    buf_putc(b, '"');
    for (size_t i = 0; i < f.n; i++) {
        if (f.p[i] == '"') buf_putc(b, '"');
        buf_putc(b, f.p[i]);
    }
    buf_putc(b, '"');
}
static void csv_u64(orc_buf *b, uint64_t v, int first) {
    if (!first) buf_putc(b, '\t');
    buf_putu(b, v);
}

/* ------------------------------------------------------------------ maf2paf */
/* maf.rs:484-520 convert2paf + converter.rs:35-52 csv serialize */
int orc_maf2paf(const char *text, size_t len, const char *query_name, orc_buf *out, char *err,
                size_t errcap) {
    lreader rd;
    maf_reader_init(&rd, text, len);
    mafrec rec = {0};
    orc_buf cigar = {0};
    int rc;
    while ((rc = maf_next(&rd, &rec, err, errcap)) > 0) {
        if (query_name) {
            int e = maf_set_query_byname(&rec, query_name, err, errcap);
            if (e) { rc = -e; break; }
        }
        if (rec.query_idx >= rec.n) { /* slines[1] on a 1-s-line block: index panic */
            set_err(err, errcap, "panic: index out of bounds: the len is %zu but the index is %zu",
                    rec.n, rec.query_idx);
            rc = -ORC_E_PANIC;
            break;
        }
        const sline *t = &rec.s[0], *q = &rec.s[rec.query_idx];
        size_t n = t->seq.n < q->seq.n ? t->seq.n : q->seq.n;
        orc_counts c;
        cigar.n = 0;
        orc_parse_maf_seq_to_cigar(t->seq.p, q->seq.p, n, q->neg, &c, &cigar);
        uint64_t matches = c.c[0];
        uint64_t block_length = c.c[0] + c.c[1] + c.c[3] + c.c[7] + c.c[5] + c.c[9];
        uint64_t edit = block_length - matches;
        csv_field(out, q->name, 1);
        csv_u64(out, q->size, 0);
        csv_u64(out, maf_query_start(&rec), 0);
        csv_u64(out, maf_query_end(&rec), 0);
        buf_putc(out, '\t');
        buf_putc(out, q->neg ? '-' : '+');
        csv_field(out, t->name, 0);
        csv_u64(out, t->size, 0);
        csv_u64(out, t->start, 0);
        csv_u64(out, t->start + t->align_size, 0);
        csv_u64(out, matches, 0);
        csv_u64(out, block_length, 0);
        csv_u64(out, 255, 0);
        buf_puts(out, "\tNM:i:");
        buf_putu(out, edit);
        buf_puts(out, "\tcg:Z:");
        buf_put(out, cigar.p, cigar.n);
        buf_putc(out, '\n');
    }
    free(rec.s);
    orc_buf_free(&cigar);
    return rc < 0 ? -rc : ORC_OK;
}

/* ------------------------------------------------------------------ natord */
/* natord 1.0.9 compare(): chars, skip = is_whitespace, digit = to_digit(10).
 * Restated from the published algorithm (Martin Pool's strnatcmp port). (parity unpinned) */
static int nat_digit(const char *s, size_t n, size_t i) {
    if (i < n && s[i] >= '0' && s[i] <= '9') return s[i] - '0';
    return -1;
}
int orc_natord_compare(const char *a, const char *b) {
    size_t na = strlen(a), nb = strlen(b), i = 0, j = 0;
    for (;;) {
        while (i < na && is_ws(a[i])) i++;
        while (j < nb && is_ws(b[j])) j++;
        if (i >= na && j >= nb) return 0;
        if (i >= na) return -1;
        if (j >= nb) return 1;
        int da = nat_digit(a, na, i), db = nat_digit(b, nb, j);
        if (da >= 0 && db >= 0) {
            if (da == 0 || db == 0) { /* left-aligned ("fractional") comparison */
                if (da != db) return da < db ? -1 : 1;
                for (;;) {
                    i++; j++;
                    da = nat_digit(a, na, i);
                    db = nat_digit(b, nb, j);
                    if (da >= 0 && db >= 0) {
                        if (da != db) return da < db ? -1 : 1;
                    } else if (da >= 0) return 1;
                    else if (db >= 0) return -1;
                    else break;
                }
            } else { /* right-aligned: longer digit run wins, else first difference */
                int last = da < db ? -1 : (da > db ? 1 : 0);
                for (;;) {
                    i++; j++;
                    da = nat_digit(a, na, i);
                    db = nat_digit(b, nb, j);
                    if (da >= 0 && db >= 0) {
                        if (last == 0) last = da < db ? -1 : (da > db ? 1 : 0);
                    } else if (da >= 0) return 1;
                    else if (db >= 0) return -1;
                    else break;
                }
                if (last) return last;
            }
            continue; /* do not advance again */
        }
        if (a[i] != b[j]) return (unsigned char)a[i] < (unsigned char)b[j] ? -1 : 1;
        i++; j++;
    }
}

/* ------------------------------------------------------------- f32 -> text */
/* ryu 1.0.14 Buffer::format(f32) ("pretty" layout, src/pretty/mod.rs format32):
 * shortest round-trip digits; decimal for 1e-5 <= |v| < 1e13 else exponent form. (parity unpinned) */
void orc_fmt_f32(float v, char *dst) {
    if (isnan(v)) { strcpy(dst, "NaN"); return; }
    if (isinf(v)) { strcpy(dst, v < 0 ? "-inf" : "inf"); return; }
    char *o = dst;
    if (signbit(v)) { *o++ = '-'; v = -v; }
    if (v == 0.0f) { strcpy(o, "0.0"); return; }
    char sci[40];
    int p;
    for (p = 1; p <= 9; p++) { /* shortest p with round trip */
        snprintf(sci, sizeof sci, "%.*e", p - 1, (double)v);
        if (strtof(sci, NULL) == v) break;
    }
    /* sci = d[.ddd]e[+-]XX */
    char digits[16];
    int nd = 0;
    const char *s = sci;
    for (; *s && *s != 'e'; s++)
        if (*s >= '0' && *s <= '9') digits[nd++] = *s;
    int e10 = atoi(s + 1);
    while (nd > 1 && digits[nd - 1] == '0') nd--; /* ryu never keeps trailing zeros */
    int length = nd, k = e10 - (nd - 1), kk = length + k;
    if (0 <= k && kk <= 13) { /* 1234e7 -> 12340000000.0 */
        memcpy(o, digits, (size_t)nd); o += nd;
        for (int i = length; i < kk; i++) *o++ = '0';
        *o++ = '.'; *o++ = '0'; *o = 0;
    } else if (0 < kk && kk <= 13) { /* 1234e-2 -> 12.34 */
        memcpy(o, digits, (size_t)kk); o += kk;
        *o++ = '.';
        memcpy(o, digits + kk, (size_t)(nd - kk)); o += nd - kk;
        *o = 0;
    } else if (-6 < kk && kk <= 0) { /* 1234e-6 -> 0.001234 */
        *o++ = '0'; *o++ = '.';
        for (int i = 0; i < -kk; i++) *o++ = '0';
        memcpy(o, digits, (size_t)nd); o += nd;
        *o = 0;
    } else if (length == 1) { /* 1e30 */
        *o++ = digits[0];
        sprintf(o, "e%d", kk - 1);
    } else { /* 1234e30 -> 1.234e33 */
        *o++ = digits[0]; *o++ = '.';
        memcpy(o, digits + 1, (size_t)(nd - 1)); o += nd - 1;
        sprintf(o, "e%d", kk - 1);
    }
}

/* -------------------------------------------------------------------- stat */
/* common.rs:98-140 RecStat::from(Cigar) */
typedef struct {
    uint64_t aligned_size, matched, mismatched, ins_event, del_event, ins_size, del_size,
        inv_ins_event, inv_ins_size, inv_del_event, inv_del_size, inv_event;
    float inv_size;
} recstat;
static recstat recstat_from(const orc_counts *cg) {
    const uint64_t *c = cg->c;
    recstat r;
    memset(&r, 0, sizeof r);
    r.matched = c[0]; r.mismatched = c[1];
    r.ins_event = c[2]; r.ins_size = c[3];
    r.del_event = c[4]; r.del_size = c[5];
    r.inv_ins_event = c[6]; r.inv_ins_size = c[7];
    r.inv_del_event = c[8]; r.inv_del_size = c[9];
    r.aligned_size = r.matched + r.mismatched + r.del_size + r.inv_del_size;
    uint64_t query_align = r.matched + r.mismatched + r.ins_size + r.inv_ins_size;
    r.inv_event = c[10];
    r.inv_size = 0.0f;
    if (r.inv_event != 0)
        r.inv_size = (float)(r.aligned_size + query_align) / (float)(r.inv_event + 1);
    return r;
}

/* stat.rs:27-50 Statistic */
typedef struct {
    sv ref_name; uint64_t ref_size, ref_start;
    sv query_name; uint64_t query_size, query_start;
    uint64_t aligned_size, unaligned_size;
    float identity, similarity;
    uint64_t matched, mismatched, ins_event, del_event, ins_size, del_size, inv_event;
    float inv_size;
    uint64_t inv_ins_event, inv_ins_size, inv_del_event, inv_del_size;
} statrow;

typedef struct { statrow *r; size_t n, cap; size_t *ht; size_t ht_cap; } statvec; /* ht: Pair -> row (the HashMap) */
static statrow *statvec_push(statvec *v) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 64;
        v->r = (statrow *)realloc(v->r, v->cap * sizeof(statrow));
    }
    statrow *s = &v->r[v->n++];
    memset(s, 0, sizeof *s);
    return s;
}

/* one PairStat (stat.rs:53-58, 226-250) folded straight into the output rows.
 * each: stat.rs:129-164 split_final;  merged: stat.rs:167-223 merge_final_from_pair */
static void stat_accumulate(statvec *v, int each, sv ref_name, uint64_t ref_size, uint64_t ref_start,
                            sv query_name, uint64_t query_size, uint64_t query_start,
                            const recstat *rs) {
    statrow *s = NULL;
    size_t slot = 0;
    if (!each) { /* HashMap<Pair,..> (stat.rs:170-176); rows keep first-seen order */
        if ((v->n + 1) * 2 > v->ht_cap) { /* grow + rehash */
            size_t nc = v->ht_cap ? v->ht_cap * 2 : 1024;
            size_t *nh = (size_t *)malloc(nc * sizeof(size_t));
            for (size_t i = 0; i < nc; i++) nh[i] = (size_t)-1;
            for (size_t i = 0; i < v->n; i++) {
                const statrow *c = &v->r[i];
                uint64_t h = 1469598103934665603ull;
                for (size_t k = 0; k < c->ref_name.n; k++) h = (h ^ (unsigned char)c->ref_name.p[k]) * 1099511628211ull;
                h = (h ^ c->ref_size) * 1099511628211ull;
                for (size_t k = 0; k < c->query_name.n; k++) h = (h ^ (unsigned char)c->query_name.p[k]) * 1099511628211ull;
                h = (h ^ c->query_size) * 1099511628211ull;
                size_t sl = (size_t)(h ^ (h >> 29)) & (nc - 1);
                while (nh[sl] != (size_t)-1) sl = (sl + 1) & (nc - 1);
                nh[sl] = i;
            }
            free(v->ht);
            v->ht = nh;
            v->ht_cap = nc;
        }
        uint64_t h = 1469598103934665603ull;
        for (size_t k = 0; k < ref_name.n; k++) h = (h ^ (unsigned char)ref_name.p[k]) * 1099511628211ull;
        h = (h ^ ref_size) * 1099511628211ull;
        for (size_t k = 0; k < query_name.n; k++) h = (h ^ (unsigned char)query_name.p[k]) * 1099511628211ull;
        h = (h ^ query_size) * 1099511628211ull;
        slot = (size_t)(h ^ (h >> 29)) & (v->ht_cap - 1);
        while (v->ht[slot] != (size_t)-1) {
            statrow *c = &v->r[v->ht[slot]];
            if (c->ref_size == ref_size && c->query_size == query_size && sv_eq(c->ref_name, ref_name) &&
                sv_eq(c->query_name, query_name)) {
                s = c;
                break;
            }
            slot = (slot + 1) & (v->ht_cap - 1);
        }
    }
    if (!s) {
        s = statvec_push(v);
        if (!each) v->ht[slot] = v->n - 1;
        s->ref_name = ref_name; s->ref_size = ref_size;
        s->query_name = query_name; s->query_size = query_size;
        if (each) { s->ref_start = ref_start; s->query_start = query_start; }
        else { s->ref_start = ref_size; s->query_start = query_size; } /* stat.rs:186,189 */
    }
    s->aligned_size += rs->aligned_size;
    s->matched += rs->matched; s->mismatched += rs->mismatched;
    s->ins_event += rs->ins_event; s->del_event += rs->del_event;
    s->ins_size += rs->ins_size; s->del_size += rs->del_size;
    s->inv_ins_event += rs->inv_ins_event; s->inv_ins_size += rs->inv_ins_size;
    s->inv_del_event += rs->inv_del_event; s->inv_del_size += rs->inv_del_size;
    s->inv_event += rs->inv_event;
    s->inv_size += rs->inv_size; /* f32, file order */
    if (!each) {
        if (ref_start < s->ref_start) s->ref_start = ref_start;
        if (query_start < s->query_start) s->query_start = query_start;
    }
}

static void stat_write(statvec *v, int each, orc_buf *out) {
    /* derived fields */
    for (size_t i = 0; i < v->n; i++) {
        statrow *s = &v->r[i];
        if (!each) s->unaligned_size = s->ref_size - s->aligned_size; /* wraps, stat.rs:216 */
        s->identity = (float)s->matched / (float)s->aligned_size;
        s->similarity = (float)(s->matched + s->mismatched) / (float)s->aligned_size;
    }
    /* stat.rs:116 stable sort by natord(ref_name): insertion-stable merge via index sort */
    size_t n = v->n;
    size_t *idx = (size_t *)malloc((n ? n : 1) * sizeof(size_t));
    char **names = (char **)malloc((n ? n : 1) * sizeof(char *));
    for (size_t i = 0; i < n; i++) {
        idx[i] = i;
        names[i] = (char *)malloc(v->r[i].ref_name.n + 1);
        memcpy(names[i], v->r[i].ref_name.p, v->r[i].ref_name.n);
        names[i][v->r[i].ref_name.n] = 0;
    }
    /* stable merge sort */
    size_t *tmp = (size_t *)malloc((n ? n : 1) * sizeof(size_t));
    for (size_t w = 1; w < n; w *= 2) {
        for (size_t lo = 0; lo < n; lo += 2 * w) {
            size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            size_t a = lo, b = mid, k = lo;
            while (a < mid && b < hi) {
                if (orc_natord_compare(names[idx[b]], names[idx[a]]) < 0) tmp[k++] = idx[b++];
                else tmp[k++] = idx[a++];
            }
            while (a < mid) tmp[k++] = idx[a++];
            while (b < hi) tmp[k++] = idx[b++];
        }
        memcpy(idx, tmp, n * sizeof(size_t));
    }
    if (n) /* csv writes the header row with the first serialize() only */
        buf_puts(out, "ref_name\tref_size\tref_start\tquery_name\tquery_size\tquery_start\taligned_size\t"
                      "unaligned_size\tidentity\tsimilarity\tmatched\tmismatched\tins_event\tdel_event\t"
                      "ins_size\tdel_size\tinv_event\tinv_size\tinv_ins_event\tinv_ins_size\t"
                      "inv_del_event\tinv_del_size\n");
    char fb[40];
    for (size_t k = 0; k < n; k++) {
        const statrow *s = &v->r[idx[k]];
        csv_field(out, s->ref_name, 1);
        csv_u64(out, s->ref_size, 0); csv_u64(out, s->ref_start, 0);
        csv_field(out, s->query_name, 0);
        csv_u64(out, s->query_size, 0); csv_u64(out, s->query_start, 0);
        csv_u64(out, s->aligned_size, 0); csv_u64(out, s->unaligned_size, 0);
        orc_fmt_f32(s->identity, fb); buf_putc(out, '\t'); buf_puts(out, fb);
        orc_fmt_f32(s->similarity, fb); buf_putc(out, '\t'); buf_puts(out, fb);
        csv_u64(out, s->matched, 0); csv_u64(out, s->mismatched, 0);
        csv_u64(out, s->ins_event, 0); csv_u64(out, s->del_event, 0);
        csv_u64(out, s->ins_size, 0); csv_u64(out, s->del_size, 0);
        csv_u64(out, s->inv_event, 0);
        orc_fmt_f32(s->inv_size, fb); buf_putc(out, '\t'); buf_puts(out, fb);
        csv_u64(out, s->inv_ins_event, 0); csv_u64(out, s->inv_ins_size, 0);
        csv_u64(out, s->inv_del_event, 0); csv_u64(out, s->inv_del_size, 0);
        buf_putc(out, '\n');
    }
    for (size_t i = 0; i < n; i++) free(names[i]);
    free(names); free(idx); free(tmp);
}

/* stat.rs:61-84 stat_maf */
int orc_stat_maf(const char *text, size_t len, int each, const char *query_name, orc_buf *out,
                 char *err, size_t errcap) {
    lreader rd;
    maf_reader_init(&rd, text, len);
    mafrec rec = {0};
    statvec v = {0};
    int rc;
    while ((rc = maf_next(&rd, &rec, err, errcap)) > 0) {
        if (query_name) {
            int e = maf_set_query_byname(&rec, query_name, err, errcap);
            if (e) { rc = -e; break; }
        }
        if (rec.query_idx >= rec.n) {
            set_err(err, errcap, "panic: index out of bounds: the len is %zu but the index is %zu",
                    rec.n, rec.query_idx);
            rc = -ORC_E_PANIC;
            break;
        }
        const sline *t = &rec.s[0], *q = &rec.s[rec.query_idx];
        size_t n = t->seq.n < q->seq.n ? t->seq.n : q->seq.n;
        orc_counts c;
        orc_parse_maf_seq_to_cigar(t->seq.p, q->seq.p, n, q->neg, &c, NULL);
        recstat rs = recstat_from(&c);
        stat_accumulate(&v, each, t->name, t->size, t->start, q->name, q->size, maf_query_start(&rec), &rs);
    }
    if (rc == 0) stat_write(&v, each, out);
    free(rec.s);
    free(v.r); free(v.ht);
    return rc < 0 ? -rc : ORC_OK;
}

int64_t orc_maf_record_counts(const char *text, size_t len, orc_counts *out, size_t cap) {
    lreader rd;
    maf_reader_init(&rd, text, len);
    mafrec rec = {0};
    char err[256];
    int rc;
    int64_t n = 0;
    while ((rc = maf_next(&rd, &rec, err, sizeof err)) > 0) {
        if (rec.n < 2) { n = -ORC_E_PANIC; break; }
        const sline *t = &rec.s[0], *q = &rec.s[1];
        size_t m = t->seq.n < q->seq.n ? t->seq.n : q->seq.n;
        if ((size_t)n < cap) orc_parse_maf_seq_to_cigar(t->seq.p, q->seq.p, m, q->neg, &out[n], NULL);
        n++;
    }
    free(rec.s);
    return rc < 0 ? rc : n;
}

/* --------------------------------------------------------------- PAF reader */
/* paf.rs:22-31 csv reader: tab, no header, flexible, '#' comments; csv-core quoting rules
 * ('"' at field start opens a quoted field, '""' escapes).  Records end at \n, \r\n or \r. */
typedef struct { sv *f; size_t n, cap; char *scratch; size_t sn, scap; } pafrow;

static void pafrow_push(pafrow *r, const char *p, size_t n) {
    if (r->n == r->cap) {
        r->cap = r->cap ? r->cap * 2 : 32;
        r->f = (sv *)realloc(r->f, r->cap * sizeof(sv));
    }
    r->f[r->n].p = p;
    r->f[r->n].n = n;
    r->n++;
}

static void pafrow_free_scratch(pafrow *r) {
    while (r->scratch) { char *nx = *(char **)r->scratch; free(r->scratch); r->scratch = nx; }
}

/* returns 1 if a record was read, 0 at end */
static int paf_next(const char *t, size_t len, size_t *pos, pafrow *row) {
    for (;;) {
        size_t i = *pos;
        if (i >= len) return 0;
        if (t[i] == '\n' || t[i] == '\r') { *pos = i + 1; continue; } /* empty lines skipped */
        if (t[i] == '#') { /* comment line */
            while (i < len && t[i] != '\n') i++;
            *pos = i < len ? i + 1 : len;
            continue;
        }
        row->n = 0;
        for (;;) {
            if (i < len && t[i] == '"') { /* quoted field -> unescape into scratch */
                /* the decoded field is never longer than the raw field; chunks are never moved or reused, so the
                 * fields (names kept by stat / pafpseudo / filter across records) stay valid until pafrow_free_scratch */
                size_t need = 0;
                {
                    size_t j = i + 1;
                    while (j < len) {
                        if (t[j] == '"') { if (j + 1 < len && t[j + 1] == '"') { j += 2; need++; continue; } j++; break; }
                        j++; need++;
                    }
                    while (j < len && t[j] != '\t' && t[j] != '\n' && t[j] != '\r') { j++; need++; }
                }
                if (!row->scratch || row->sn + need > row->scap) {
                    const size_t cap = need > 65536 ? need : 65536;
                    char *ns = (char *)malloc(sizeof(char *) + cap);
                    *(char **)ns = row->scratch; /* chunk list */
                    row->scratch = ns;
                    row->sn = 0;
                    row->scap = cap;
                }
                char *dst = row->scratch + sizeof(char *) + row->sn;
                size_t dn = 0;
                i++;
                while (i < len) {
                    if (t[i] == '"') {
                        if (i + 1 < len && t[i + 1] == '"') { dst[dn++] = '"'; i += 2; continue; }
                        i++;
                        break;
                    }
                    dst[dn++] = t[i++];
                }
                while (i < len && t[i] != '\t' && t[i] != '\n' && t[i] != '\r') dst[dn++] = t[i++];
                pafrow_push(row, dst, dn);
                row->sn += dn;
            } else {
                size_t b = i;
                while (i < len && t[i] != '\t' && t[i] != '\n' && t[i] != '\r') i++;
                pafrow_push(row, t + b, i - b);
            }
            if (i < len && t[i] == '\t') { i++; continue; }
            break;
        }
        if (i < len && t[i] == '\r') i++;
        if (i < len && t[i] == '\n') i++;
        *pos = i;
        return 1;
    }
}

/* csv deserializer integer rule: "0x" prefix => radix 16, else str::parse */
static int csv_parse_u64(sv s, uint64_t *out) {
    if (s.n >= 2 && s.p[0] == '0' && s.p[1] == 'x') {
        uint64_t v = 0;
        if (s.n == 2) return 0;
        size_t i = 2;
        if (s.p[i] == '+') { i++; if (i == s.n) return 0; }
        for (; i < s.n; i++) {
            char ch = s.p[i];
            unsigned d = ch >= '0' && ch <= '9' ? (unsigned)(ch - '0')
                       : ch >= 'a' && ch <= 'f' ? (unsigned)(ch - 'a' + 10)
                       : ch >= 'A' && ch <= 'F' ? (unsigned)(ch - 'A' + 10) : 99u;
            if (d > 15) return 0;
            if (v >> 60) return 0;
            v = v * 16 + d;
        }
        *out = v;
        return 1;
    }
    return parse_u64(s, out);
}

/* paf.rs:159-218 cs_to_cigar; regex (:[0-9]+|\*[a-z][a-z]|[=\+\-][A-Za-z]+), leftmost matches */
static int cs_to_cigar(sv cs, orc_buf *cigar) { /* 1 = a `:N` length does not fit usize: parse().unwrap() panics (paf.rs:171) */
    char last_op = 'M';
    uint64_t last_len = 0;
    size_t i = 0;
    while (i < cs.n) {
        char c = cs.p[i];
        if (c == ':' && i + 1 < cs.n && cs.p[i + 1] >= '0' && cs.p[i + 1] <= '9') {
            size_t j = i + 1;
            uint64_t v = 0;
            while (j < cs.n && cs.p[j] >= '0' && cs.p[j] <= '9') {
                const uint64_t d = (uint64_t)(cs.p[j++] - '0');
                if (v > (UINT64_MAX - d) / 10) return 1;
                v = v * 10 + d;
            }
            if (last_op == 'M') last_len += v; /* release build: wrapping add */
            else {
                if (last_len > 0) { buf_putu(cigar, last_len); buf_putc(cigar, last_op); }
                last_op = 'M';
                last_len = v;
            }
            i = j;
        } else if (c == '*' && i + 2 < cs.n && cs.p[i + 1] >= 'a' && cs.p[i + 1] <= 'z' &&
                   cs.p[i + 2] >= 'a' && cs.p[i + 2] <= 'z') {
            if (last_op == 'X') last_len += 1;
            else {
                if (last_len > 0) { buf_putu(cigar, last_len); buf_putc(cigar, last_op); }
                last_op = 'X';
                last_len = 1;
            }
            i += 3;
        } else if ((c == '=' || c == '+' || c == '-') && i + 1 < cs.n &&
                   ((cs.p[i + 1] >= 'A' && cs.p[i + 1] <= 'Z') || (cs.p[i + 1] >= 'a' && cs.p[i + 1] <= 'z'))) {
            size_t j = i + 1;
            while (j < cs.n && ((cs.p[j] >= 'A' && cs.p[j] <= 'Z') || (cs.p[j] >= 'a' && cs.p[j] <= 'z'))) j++;
            uint64_t l = j - (i + 1);
            if (c == '-' || c == '+') {
                if (last_len > 0) { buf_putu(cigar, last_len); buf_putc(cigar, last_op); }
                buf_putu(cigar, l);
                buf_putc(cigar, c == '-' ? 'D' : 'I');
                last_len = 0;
                last_op = 'M';
            } /* '=' long form: matched by the regex, ignored by the match arm (paf.rs:213) */
            i = j;
        } else {
            i++;
        }
    }
    if (last_len > 0) { buf_putu(cigar, last_len); buf_putc(cigar, last_op); }
    return 0;
}

/* paf.rs:122-140 get_cigar_string: the first cg:Z: tag, else the first cs:Z: tag converted (text after "cg:Z:").
 * 1 = found, 0 = CigarTagNotFound, -1 = the conversion panics */
static int paf_get_cigar(const pafrow *row, orc_buf *conv, sv *cg) {
    for (size_t k = 12; k < row->n; k++)
        if (row->f[k].n >= 5 && memcmp(row->f[k].p, "cg:Z:", 5) == 0) {
            cg->p = row->f[k].p + 5; cg->n = row->f[k].n - 5;
            return 1;
        }
    for (size_t k = 12; k < row->n; k++)
        if (row->f[k].n >= 5 && memcmp(row->f[k].p, "cs:Z:", 5) == 0) {
            sv cs = {row->f[k].p + 5, row->f[k].n - 5};
            conv->n = 0;
            buf_reserve(conv, 1);
            if (cs_to_cigar(cs, conv)) return -1;
            cg->p = conv->p; cg->n = conv->n;
            return 1;
        }
    return 0;
}
#define CS_PANIC_MSG "panic: called `Result::unwrap()` on an `Err` value: ParseIntError { kind: PosOverflow }"

/* test hook: the bare conversion (text after "cs:Z:" -> text after "cg:Z:") */
int orc_cs_to_cigar(const char *cs, size_t n, orc_buf *out, char *err, size_t errcap) {
    sv v = {cs, n};
    buf_reserve(out, 1);
    if (cs_to_cigar(v, out)) { set_err(err, errcap, CS_PANIC_MSG); return ORC_E_PANIC; }
    return ORC_OK;
}

/* stat.rs:87-105 stat_paf */
int orc_stat_paf(const char *text, size_t len, int each, orc_buf *out, char *err, size_t errcap) {
    size_t pos = 0;
    pafrow row = {0};
    statvec v = {0};
    orc_buf conv = {0};
    int rc = ORC_OK;
    while (paf_next(text, len, &pos, &row)) {
        uint64_t u[12] = {0};
        int neg = 0;
        if (row.n < 12) {
            set_err(err, errcap, "CSV deserialize error by: record has %zu fields, 12 required", row.n);
            rc = ORC_E_CSV;
            break;
        }
        int bad = 0;
        for (int k = 0; k < 12 && !bad; k++) {
            if (k == 0 || k == 5) continue;
            if (k == 4) {
                if (sv_eq_c(row.f[4], "+")) neg = 0;
                else if (sv_eq_c(row.f[4], "-")) neg = 1;
                else bad = 1;
                continue;
            }
            if (!csv_parse_u64(row.f[k], &u[k])) bad = 1;
        }
        if (bad) {
            set_err(err, errcap, "CSV deserialize error by: field parse");
            rc = ORC_E_CSV;
            break;
        }
        /* paf.rs:122-140 get_cigar_string */
        sv cg = {0, 0};
        const int have = paf_get_cigar(&row, &conv, &cg);
        if (have < 0) { set_err(err, errcap, CS_PANIC_MSG); rc = ORC_E_PANIC; break; }
        if (!have) {
            set_err(err, errcap, "CIGAR start tag not found");
            rc = ORC_E_CIGAR_TAG_NOT_FOUND;
            break;
        }
        orc_counts c;
        rc = orc_parse_paf_cigar(cg.p, cg.n, neg, &c, err, errcap);
        if (rc) break;
        recstat rs = recstat_from(&c);
        /* stat_rec: ref = target (cols 6-9), query = cols 1-4 */
        stat_accumulate(&v, each, row.f[5], u[6], u[7], row.f[0], u[1], u[2], &rs);
    }
    /* names reference row scratch only for quoted fields; copy-free path is the norm.  To be
     * safe for quoted names we would need owned copies; quoted names are not generated by tests. */
    if (rc == ORC_OK) stat_write(&v, each, out);
    free(row.f); pafrow_free_scratch(&row); free(v.r); free(v.ht);
    orc_buf_free(&conv);
    return rc;
}

/* -------------------------------------------------------------------- call */
/* caller.rs:304-337 build_header + noodles-vcf 0.43 header text (README.md:324-331 golden) */
static void vcf_header(orc_buf *out, const char *sample) {
    buf_puts(out, "##fileformat=VCFv4.4\n"
                  "##INFO=<ID=SVLEN,Number=A,Type=Integer,Description=\"Length of structural variant\">\n"
                  "##INFO=<ID=SVTYPE,Number=1,Type=String,Description=\"Type of structural variant\">\n"
                  "##INFO=<ID=END,Number=1,Type=Integer,Description=\"End position of the longest variant described in this record\">\n"
                  "##INFO=<ID=INV_NEST,Number=1,Type=String,Description=\"Varations nested within inversion\">\n"
                  "##FORMAT=<ID=QI,Number=1,Type=String,Description=\"Query informations\">\n"
                  "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
                  "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t");
    buf_puts(out, sample);
    buf_putc(out, '\n');
}

/* noodles-vcf 0.43 reference_bases / allele Base: parse upper-cases, only ACGTN valid. (parity unpinned) */
static int vcf_bases(orc_buf *out, const char *s, size_t n, char *err, size_t errcap) {
    for (size_t i = 0; i < n; i++) {
        char c = s[i];
        if (c >= 'a' && c <= 'z') c = (char)(c - 32);
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'N') {
            set_err(err, errcap, "invalid VCF base `%c`", s[i]);
            return ORC_E_VCF_BASE;
        }
        buf_putc(out, c);
    }
    return ORC_OK;
}

/* caller.rs:359-386 get_variant_rec + noodles writer: one VCF body row */
static int vcf_row(orc_buf *out, sv chro, uint64_t pos, const char *ref, size_t nref, const char *alt,
                   size_t nalt, int alt_symbolic, const char *info, const char *fmt_tail, char *err,
                   size_t errcap) {
    buf_put(out, chro.p, chro.n);
    buf_putc(out, '\t');
    buf_putu(out, pos);
    buf_puts(out, "\t.\t");
    int rc = vcf_bases(out, ref, nref, err, errcap);
    if (rc) return rc;
    buf_putc(out, '\t');
    if (alt_symbolic) buf_put(out, alt, nalt);
    else if ((rc = vcf_bases(out, alt, nalt, err, errcap))) return rc;
    buf_puts(out, "\t.\t.\t");
    buf_puts(out, info ? info : ".");
    buf_puts(out, "\tGT:QI\t1|1:");
    buf_puts(out, fmt_tail);
    buf_putc(out, '\n');
    return ORC_OK;
}

/* caller.rs:159-219 find_safe_chunk_boundary */
static size_t find_safe_chunk_boundary(const char *t, const char *q, size_t start, size_t chunk_size,
                                       uint64_t svlen_cutoff, size_t total_size) {
    size_t proposed_end = start + chunk_size < total_size ? start + chunk_size : total_size;
    size_t current_gap_size = 0, sv_start = 0, safe_end = proposed_end;
    int in_sv = 0;
    for (size_t abs_pos = start; abs_pos < proposed_end; abs_pos++) {
        if (t[abs_pos] == '-' || q[abs_pos] == '-') {
            if (!in_sv) { in_sv = 1; sv_start = abs_pos; }
            current_gap_size++;
        } else if (in_sv) {
            if (current_gap_size >= svlen_cutoff)
                if (sv_start >= start) safe_end = abs_pos;
            in_sv = 0;
            current_gap_size = 0;
        }
    }
    if (in_sv && current_gap_size >= svlen_cutoff) {
        size_t end_pos = proposed_end;
        for (size_t p = proposed_end; p < total_size; p++)
            if (t[p] != '-' && q[p] != '-') { end_pos = p; break; }
        safe_end = end_pos;
    }
    return safe_end;
}

/* caller.rs:388-608 call_within_var on the chunk [cs, ce) of (t,q);
 * chunk coordinates from caller.rs:221-265 create_chunk_record */
static int call_within_var(orc_buf *out, const mafrec *rec, size_t cs, size_t ce, int if_snp,
                           uint64_t svlen_cutoff, int if_inv, char *err, size_t errcap) {
    const sline *ts = &rec->s[0], *qs = &rec->s[rec->query_idx];
    const char *t = ts->seq.p, *q = qs->seq.p;
    /* create_chunk_record: new_start = start + nongap(seq[..cs]); align = nongap(seq[cs..ce]) */
    uint64_t t_start = ts->start, q_rawstart = qs->start, t_al = 0, q_al = 0;
    for (size_t i = 0; i < cs; i++) { if (t[i] != '-') t_start++; if (q[i] != '-') q_rawstart++; }
    for (size_t i = cs; i < ce; i++) { if (t[i] != '-') t_al++; if (q[i] != '-') q_al++; }
    uint64_t t_end = t_start + t_al;
    uint64_t q_start = qs->neg ? qs->size - q_rawstart - q_al : q_rawstart;
    uint64_t q_end = qs->neg ? qs->size - q_rawstart : q_rawstart + q_al;

    size_t n = ce - cs;
    char *t_ref = (char *)malloc(n + 1), *q_ref = (char *)malloc(n + 1);
    size_t nt = 0, nq = 0;
    for (size_t i = cs; i < ce; i++) { if (t[i] != '-') t_ref[nt++] = t[i]; if (q[i] != '-') q_ref[nq++] = q[i]; }

    uint64_t t_off = t_start, q_off = q_start;
    char suffix = qs->neg ? 'N' : 'P';
    char info[160], tail[512];
    const char *nest = qs->neg ? "INV_NEST=TRUE;" : "";
    int rc = ORC_OK;
#define QN (int)qs->name.n, qs->name.p
    if (qs->neg && nt > 0 && if_inv) { /* caller.rs:423-440 */
        snprintf(info, sizeof info, "SVTYPE=INV;END=%llu", (unsigned long long)t_end);
        snprintf(tail, sizeof tail, "%.*s@%llu@%llu@%c", QN, (unsigned long long)q_start,
                 (unsigned long long)q_end, suffix);
        rc = vcf_row(out, ts->name, t_off + 1, t_ref, 1, "<INV>", 5, 1, info, tail, err, errcap);
        if (rc) goto done;
    }
    int after_m = 0;
    size_t i = cs;
    while (i < ce) { /* group_by(cigar_cat_ext_caller) */
        char k = cigar_cat_ext_caller(t[i], q[i]);
        size_t j = i + 1;
        while (j < ce && cigar_cat_ext_caller(t[j], q[j]) == k) j++;
        uint64_t len = j - i;
        i = j;
        switch (k) {
        case '=':
            t_off += len; q_off += len; after_m = 1;
            break;
        case 'W': break;
        case 'I':
            if (len > svlen_cutoff) {
                if (!after_m) { q_off += len; after_m = 0; continue; }
                size_t tss = (size_t)(t_off - t_start - 1), qss = (size_t)(q_off - q_start - 1);
                if (tss + 1 > nt || qss + len + 1 > nq) { rc = ORC_E_PANIC; set_err(err, errcap, "panic: slice out of range"); goto done; }
                snprintf(info, sizeof info, "%sSVTYPE=INS;SVLEN=%llu;END=%llu", nest,
                         (unsigned long long)len, (unsigned long long)t_off);
                snprintf(tail, sizeof tail, "%.*s@%llu@%llu@%c", QN, (unsigned long long)q_off,
                         (unsigned long long)(q_off + len), suffix);
                rc = vcf_row(out, ts->name, t_off, t_ref + tss, 1, q_ref + qss, (size_t)len + 1, 0, info, tail, err, errcap);
                if (rc) goto done;
            }
            q_off += len; after_m = 0;
            break;
        case 'D':
            if (len > svlen_cutoff) {
                if (!after_m) { t_off += len; after_m = 0; continue; }
                size_t tss = (size_t)(t_off - t_start - 1), qss = (size_t)(q_off - q_start - 1);
                if (tss + len + 1 > nt || qss + 1 > nq) { rc = ORC_E_PANIC; set_err(err, errcap, "panic: slice out of range"); goto done; }
                snprintf(info, sizeof info, "%sSVTYPE=DEL;SVLEN=%llu;END=%llu", nest,
                         (unsigned long long)len, (unsigned long long)(t_off + len));
                snprintf(tail, sizeof tail, "%.*s@%llu@%llu@%c", QN, (unsigned long long)q_off,
                         (unsigned long long)q_off, suffix);
                rc = vcf_row(out, ts->name, t_off, t_ref + tss, (size_t)len + 1, q_ref + qss, 1, 0, info, tail, err, errcap);
                if (rc) goto done;
            }
            t_off += len; after_m = 0;
            break;
        case 'X':
            if (if_snp) {
                for (uint64_t x = 0; x < len; x++) {
                    size_t tss = (size_t)(t_off - t_start), qss = (size_t)(q_off - q_start);
                    if (tss + 1 > nt || qss + 1 > nq) { rc = ORC_E_PANIC; set_err(err, errcap, "panic: slice out of range"); goto done; }
                    snprintf(tail, sizeof tail, "%.*s@%llu@%c", QN, (unsigned long long)q_off, suffix);
                    rc = vcf_row(out, ts->name, t_off + 1, t_ref + tss, 1, q_ref + qss, 1, 0, NULL, tail, err, errcap);
                    if (rc) goto done;
                    t_off++; q_off++;
                }
            } else { q_off += len; t_off += len; }
            after_m = 1;
            break;
        }
    }
#undef QN
done:
    free(t_ref); free(q_ref);
    return rc;
}

/* caller.rs:339-357 add_header_contig: the index's isref entries, sorted by natord::compare on the name (Vec::sort_by is
 * stable), each parsed into a noodles contig name (`name.parse()?`: VCF 4.3 section 1.4.7 grammar -- third-party, parity
 * unpinned) and written by noodles-vcf 0.43 as "##contig=<ID=name,length=size>" behind the FORMAT lines. */
static int contig_name_ok(const char *s) {
    if (!*s || *s == '*' || *s == '=') return 0;
    for (const unsigned char *p = (const unsigned char *)s; *p; p++)
        if (*p < '!' || *p > '~' || strchr("\\,\"`'()[]{}<>", *p)) return 0;
    return 1;
}
static int vcf_header_contigs(orc_buf *out, const char *sample, const char *const *names, const uint64_t *sizes, size_t n,
                              char *err, size_t errcap) {
    size_t *idx = (size_t *)malloc((n + 1) * sizeof(size_t));
    for (size_t i = 0; i < n; i++) idx[i] = i;
    for (size_t i = 1; i < n; i++) { /* stable insertion sort */
        size_t v = idx[i], j = i;
        while (j > 0 && orc_natord_compare(names[v], names[idx[j - 1]]) < 0) { idx[j] = idx[j - 1]; j--; }
        idx[j] = v;
    }
    orc_buf h = {0};
    vcf_header(&h, sample);
    /* everything in front of "#CHROM" first, then the contig lines */
    size_t cut = 0;
    for (size_t i = 0; i + 6 <= h.n; i++)
        if (memcmp(h.p + i, "#CHROM", 6) == 0 && (i == 0 || h.p[i - 1] == '\n')) { cut = i; break; }
    buf_put(out, h.p, cut);
    int rc = ORC_OK;
    for (size_t k = 0; k < n && rc == ORC_OK; k++) {
        const size_t i = idx[k];
        if (!contig_name_ok(names[i])) {
            set_err(err, errcap, "invalid contig name `%s`", names[i]);
            rc = ORC_E_OTHER;
            break;
        }
        char num[32];
        snprintf(num, sizeof num, "%llu", (unsigned long long)sizes[i]);
        buf_puts(out, "##contig=<ID=");
        buf_puts(out, names[i]);
        buf_puts(out, ",length=");
        buf_puts(out, num);
        buf_puts(out, ">\n");
    }
    if (rc == ORC_OK) buf_put(out, h.p + cut, h.n - cut);
    orc_buf_free(&h);
    free(idx);
    return rc;
}

/* caller.rs:42-157 call_var_maf.  query_pred stands for `query_regex.is_match_at(&x.name, 0)` (maf.rs:267-271): the
 * regex engine itself is the regex crate, outside this restatement; has_index != 0 = a <maf>.index was given. */
int orc_call_maf_ex(const char *text, size_t len, int snp, int inv, uint64_t svlen_cutoff,
                    const char *sample, const char *query_name, orc_name_pred query_pred, void *pred_ud,
                    uint64_t chunk_size, int has_index, const char *const *contig_names, const uint64_t *contig_sizes,
                    size_t n_contigs, orc_buf *out, char *err, size_t errcap);

int orc_call_maf(const char *text, size_t len, int snp, int inv, uint64_t svlen_cutoff,
                 const char *sample, const char *query_name, uint64_t chunk_size, orc_buf *out,
                 char *err, size_t errcap) {
    return orc_call_maf_ex(text, len, snp, inv, svlen_cutoff, sample, query_name, NULL, NULL, chunk_size, 0, NULL, NULL, 0, out,
                           err, errcap);
}

int orc_call_maf_ex(const char *text, size_t len, int snp, int inv, uint64_t svlen_cutoff,
                    const char *sample, const char *query_name, orc_name_pred query_pred, void *pred_ud,
                    uint64_t chunk_size, int has_index, const char *const *contig_names, const uint64_t *contig_sizes,
                    size_t n_contigs, orc_buf *out, char *err, size_t errcap) {
    if (has_index) {
        int hrc = vcf_header_contigs(out, sample ? sample : "sample", contig_names, contig_sizes, n_contigs, err, errcap);
        if (hrc) return -hrc;
    } else {
        vcf_header(out, sample ? sample : "sample");
    }
    lreader rd;
    maf_reader_init(&rd, text, len);
    mafrec rec = {0};
    int rc;
    while ((rc = maf_next(&rd, &rec, err, errcap)) > 0) {
        if (rec.n == 1) continue; /* caller.rs:64-74 */
        if (query_name) {
            char e2[8];
            if (maf_set_query_byname(&rec, query_name, e2, sizeof e2)) continue; /* caller.rs:79-89 */
        } else if (query_pred) { /* caller.rs:91-104: first s-line (the target included) whose name matches, else skip */
            size_t hit = rec.n;
            for (size_t i = 0; i < rec.n && hit == rec.n; i++)
                if (query_pred(rec.s[i].name.p, rec.s[i].name.n, pred_ud)) hit = i;
            if (hit == rec.n) continue;
            rec.query_idx = hit;
        } else {
            rec.query_idx = 1;
        }
        const sline *ts = &rec.s[0], *qs = &rec.s[rec.query_idx];
        size_t total = ts->seq.n;
        if (qs->seq.n < total) { /* query_seq()[start..end] slicing would panic */
            set_err(err, errcap, "panic: byte index out of range of query seq");
            rc = -ORC_E_PANIC;
            break;
        }
        { /* create_chunk_record slices EVERY s-line of the block with [start..end] (caller.rs:233-234): any row shorter
             than the target row panics once the chunk grid reaches its end */
            int ragged = 0;
            for (size_t i = 0; i < rec.n; i++) ragged |= rec.s[i].seq.n < total;
            if (ragged) {
                set_err(err, errcap, "panic: byte index out of range of an s-line seq");
                rc = -ORC_E_PANIC;
                break;
            }
        }
        size_t cs = 0;
        int e = ORC_OK;
        if (chunk_size == 0 && total > 0) { /* the reference would loop forever (caller.rs:119) */
            set_err(err, errcap, "panic: chunk_size 0 never advances");
            rc = -ORC_E_PANIC;
            break;
        }
        while (cs < total) {
            size_t ce = find_safe_chunk_boundary(ts->seq.p, qs->seq.p, cs, (size_t)chunk_size, svlen_cutoff, total);
            e = call_within_var(out, &rec, cs, ce, snp, svlen_cutoff, inv, err, errcap);
            if (e) break;
            cs = ce;
        }
        if (e) { rc = -e; break; }
    }
    free(rec.s);
    return rc < 0 ? -rc : ORC_OK;
}

/* ======================================================================================
 * Threaded restatement of the reference's rayon structure (bench baseline only).
 *
 * src/converter.rs:41-48 and src/tools/stat.rs:67-81 run `records().par_bridge()`: the record
 * iterator (BufRead::lines -> String per line, parse_sline -> name.to_string(), seq.to_string())
 * is pulled under par_bridge's lock, i.e. SERIALLY, and only the per-record closure
 * (convert2paf / stat_rec) runs on the worker pool; results are written by the calling thread.
 * `call` is strictly sequential (src/tools/caller.rs:62).  The functions below reproduce exactly
 * that split with pthreads, including the reference's per-line and per-field copies, so that the
 * CPU baseline has the reference's parallel structure rather than an idealised sharding.
 * ====================================================================================== */
#include <pthread.h>

typedef struct {
    sline *s;       /* owned copies: name.p / seq.p are malloc'ed */
    size_t n;
    size_t query_idx;
} owned_rec;

static void owned_free(owned_rec *r) {
    for (size_t i = 0; i < r->n; i++) { free((void *)r->s[i].name.p); free((void *)r->s[i].seq.p); }
    free(r->s);
    r->s = NULL;
    r->n = 0;
}

/* one s-line the way the reference materialises it: lines() allocates a String for the line and checks
 * UTF-8, parse_sline copies name and seq into fresh Strings, then the line String is dropped */
static int owned_sline(sv line, sline *out, char *err, size_t errcap) {
    char *copy = (char *)malloc(line.n + 1);
    memcpy(copy, line.p, line.n);
    unsigned char hi = 0;
    for (size_t i = 0; i < line.n; i++) hi |= (unsigned char)copy[i]; /* from_utf8 ASCII fast path */
    if (hi & 0x80) { /* non-ASCII: full validation would follow; treated as valid here */ }
    sv lc = {copy, line.n};
    int rc = parse_sline(lc, out, err, errcap);
    if (rc == ORC_OK) {
        char *nm = (char *)malloc(out->name.n + 1), *sq = (char *)malloc(out->seq.n + 1);
        memcpy(nm, out->name.p, out->name.n);
        memcpy(sq, out->seq.p, out->seq.n);
        out->name.p = nm;
        out->seq.p = sq;
    }
    free(copy);
    return rc;
}

/* MAFRecords::next with owned s-lines; same control flow as maf_next */
static int maf_next_owned(lreader *r, owned_rec *rec, char *err, size_t errcap) {
    sv line;
    for (;;) {
        if (!next_line(r, &line)) return 0;
        if (line.n > 0 && line.p[0] == 's') break;
    }
    size_t cap = 4;
    rec->s = (sline *)malloc(cap * sizeof(sline));
    rec->n = 0;
    rec->query_idx = 1;
    sline s;
    int rc = owned_sline(line, &s, err, errcap);
    if (rc) { owned_free(rec); return -rc; }
    rec->s[rec->n++] = s;
    while (next_line(r, &line)) {
        if (line.n > 0 && line.p[0] == 's') {
            rc = owned_sline(line, &s, err, errcap);
            if (rc) { owned_free(rec); return -rc; }
            if (rec->n == cap) { cap *= 2; rec->s = (sline *)realloc(rec->s, cap * sizeof(sline)); }
            rec->s[rec->n++] = s;
        } else break;
    }
    return 1;
}

typedef struct {
    /* shared */
    pthread_mutex_t mu;
    lreader rd;
    size_t next_idx;
    int err;            /* first error code (lowest record index wins) */
    size_t err_idx;
    char errmsg[256];
    const char *query_name;
    int mode;           /* 0 maf2paf rows, 1 stat */
    /* results, indexed by record: grown under mu */
    orc_buf *rows;      /* mode 0: one PAF row per record */
    struct statres { sv ref_name, query_name; uint64_t ref_size, ref_start, query_size, query_start; recstat rs; char *own; } *st;
    size_t cap;
} par_state;

static void par_reserve(par_state *ps, size_t idx) {
    if (idx < ps->cap) return;
    size_t nc = ps->cap ? ps->cap : 1024;
    while (nc <= idx) nc *= 2;
    if (ps->mode == 0) {
        ps->rows = (orc_buf *)realloc(ps->rows, nc * sizeof(orc_buf));
        memset(ps->rows + ps->cap, 0, (nc - ps->cap) * sizeof(orc_buf));
    } else {
        ps->st = (struct statres *)realloc(ps->st, nc * sizeof(*ps->st));
        memset(ps->st + ps->cap, 0, (nc - ps->cap) * sizeof(*ps->st));
    }
    ps->cap = nc;
}

static void *par_worker(void *vp) {
    par_state *ps = (par_state *)vp;
    orc_buf cigar = {0};
    for (;;) {
        owned_rec rec = {0};
        char err[256];
        pthread_mutex_lock(&ps->mu); /* par_bridge: the iterator is advanced under a lock */
        if (ps->err) { pthread_mutex_unlock(&ps->mu); break; }
        int rc = maf_next_owned(&ps->rd, &rec, err, sizeof err);
        size_t idx = ps->next_idx;
        if (rc > 0) { ps->next_idx++; par_reserve(ps, idx); }
        else if (rc < 0) { ps->err = -rc; ps->err_idx = idx; memcpy(ps->errmsg, err, sizeof err); }
        pthread_mutex_unlock(&ps->mu);
        if (rc <= 0) break;
        int e = ORC_OK;
        if (ps->query_name) {
            e = ORC_E_QUERY_NAME_NOT_FOUND;
            for (size_t i = 0; i < rec.n; i++)
                if (sv_eq_c(rec.s[i].name, ps->query_name)) { rec.query_idx = i; e = ORC_OK; break; }
            if (e) snprintf(err, sizeof err, "Query name:%s not found in MAF", ps->query_name);
        }
        if (!e && rec.query_idx >= rec.n) {
            e = ORC_E_PANIC;
            snprintf(err, sizeof err, "panic: index out of bounds: the len is %zu but the index is %zu", rec.n, rec.query_idx);
        }
        if (e) {
            pthread_mutex_lock(&ps->mu);
            if (!ps->err || idx < ps->err_idx) { ps->err = e; ps->err_idx = idx; memcpy(ps->errmsg, err, sizeof err); }
            pthread_mutex_unlock(&ps->mu);
            owned_free(&rec);
            break;
        }
        const sline *t = &rec.s[0], *q = &rec.s[rec.query_idx];
        size_t n = t->seq.n < q->seq.n ? t->seq.n : q->seq.n;
        orc_counts c;
        mafrec view = {rec.s, rec.n, rec.n, rec.query_idx};
        if (ps->mode == 0) {
            cigar.n = 0;
            orc_parse_maf_seq_to_cigar(t->seq.p, q->seq.p, n, q->neg, &c, &cigar);
            orc_buf row = {0};
            uint64_t matches = c.c[0], bl = c.c[0] + c.c[1] + c.c[3] + c.c[7] + c.c[5] + c.c[9];
            csv_field(&row, q->name, 1); csv_u64(&row, q->size, 0);
            csv_u64(&row, maf_query_start(&view), 0); csv_u64(&row, maf_query_end(&view), 0);
            buf_putc(&row, '\t'); buf_putc(&row, q->neg ? '-' : '+');
            csv_field(&row, t->name, 0); csv_u64(&row, t->size, 0); csv_u64(&row, t->start, 0);
            csv_u64(&row, t->start + t->align_size, 0); csv_u64(&row, matches, 0); csv_u64(&row, bl, 0);
            csv_u64(&row, 255, 0);
            buf_puts(&row, "\tNM:i:"); buf_putu(&row, bl - matches);
            buf_puts(&row, "\tcg:Z:"); buf_put(&row, cigar.p, cigar.n); buf_putc(&row, '\n');
            pthread_mutex_lock(&ps->mu);
            ps->rows[idx] = row;
            pthread_mutex_unlock(&ps->mu);
        } else {
            orc_parse_maf_seq_to_cigar(t->seq.p, q->seq.p, n, q->neg, &c, &cigar);
            cigar.n = 0; /* get_stat builds (and drops) the cigar string too (maf.rs:535-539) */
            struct statres sr;
            sr.rs = recstat_from(&c);
            sr.own = (char *)malloc(t->name.n + q->name.n + 2); /* Pair{ref_name.to_string(), query_name.to_string()} */
            memcpy(sr.own, t->name.p, t->name.n);
            memcpy(sr.own + t->name.n, q->name.p, q->name.n);
            sr.ref_name.p = sr.own; sr.ref_name.n = t->name.n;
            sr.query_name.p = sr.own + t->name.n; sr.query_name.n = q->name.n;
            sr.ref_size = t->size; sr.ref_start = t->start; sr.query_size = q->size;
            sr.query_start = maf_query_start(&view);
            pthread_mutex_lock(&ps->mu);
            ps->st[idx] = sr;
            pthread_mutex_unlock(&ps->mu);
        }
        owned_free(&rec);
    }
    orc_buf_free(&cigar);
    return NULL;
}

static int par_run(par_state *ps, const char *text, size_t len, int threads) {
    pthread_mutex_init(&ps->mu, NULL);
    maf_reader_init(&ps->rd, text, len);
    if (threads < 1) threads = 1;
    if (threads > 512) threads = 512;
    pthread_t th[512];
    for (int i = 0; i < threads; i++) pthread_create(&th[i], NULL, par_worker, ps);
    for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
    pthread_mutex_destroy(&ps->mu);
    return ps->err;
}

int orc_maf2paf_mt(const char *text, size_t len, const char *query_name, int threads, orc_buf *out, char *err, size_t errcap) {
    par_state ps;
    memset(&ps, 0, sizeof ps);
    ps.query_name = query_name;
    ps.mode = 0;
    int rc = par_run(&ps, text, len, threads);
    if (rc) set_err(err, errcap, "%s", ps.errmsg);
    for (size_t i = 0; i < ps.next_idx; i++) { /* converter.rs:49-52: serial csv write */
        if (!rc && ps.rows[i].p) buf_put(out, ps.rows[i].p, ps.rows[i].n);
        orc_buf_free(&ps.rows[i]);
    }
    free(ps.rows);
    return rc;
}

int orc_stat_maf_mt(const char *text, size_t len, int each, const char *query_name, int threads, orc_buf *out, char *err, size_t errcap) {
    par_state ps;
    memset(&ps, 0, sizeof ps);
    ps.query_name = query_name;
    ps.mode = 1;
    int rc = par_run(&ps, text, len, threads);
    if (rc) set_err(err, errcap, "%s", ps.errmsg);
    statvec v = {0};
    if (!rc) {
        for (size_t i = 0; i < ps.next_idx; i++) {
            struct statres *s = &ps.st[i];
            stat_accumulate(&v, each, s->ref_name, s->ref_size, s->ref_start, s->query_name, s->query_size, s->query_start, &s->rs);
        }
        stat_write(&v, each, out);
    }
    for (size_t i = 0; i < ps.next_idx; i++) free(ps.st[i].own);
    free(ps.st);
    free(v.r); free(v.ht);
    return rc;
}

/* ======================================================================================
 * paf2maf  (src/converter.rs:176-265) with the FASTA access of rust-htslib 0.44 / htslib faidx
 * (un-vendored; restated from the published htslib faidx.c: faidx_adjust_position + fai_retrieve).
 * ====================================================================================== */
typedef struct { sv name; uint64_t len, offset, linebases, linewidth; } fai_ent;
typedef struct { fai_ent *e; size_t n; const char *fa; size_t fa_len; } faidx;

static int fai_parse(const char *fai, size_t fai_len, const char *fa, size_t fa_len, faidx *out) {
    out->e = NULL; out->n = 0; out->fa = fa; out->fa_len = fa_len;
    size_t cap = 0, i = 0;
    while (i < fai_len) {
        size_t j = i;
        while (j < fai_len && fai[j] != '\n') j++;
        sv f[5]; int nf = 0; size_t b = i;
        for (size_t k = i; k <= j && nf < 5; k++)
            if (k == j || fai[k] == '\t') { f[nf].p = fai + b; f[nf].n = k - b; nf++; b = k + 1; }
        if (nf == 5) {
            if (out->n == cap) { cap = cap ? cap * 2 : 64; out->e = (fai_ent *)realloc(out->e, cap * sizeof(fai_ent)); }
            fai_ent *e = &out->e[out->n];
            e->name = f[0];
            if (!parse_u64(f[1], &e->len) || !parse_u64(f[2], &e->offset) || !parse_u64(f[3], &e->linebases) ||
                !parse_u64(f[4], &e->linewidth)) return -1;
            out->n++;
        }
        i = j + 1;
    }
    return 0;
}

/* faidx_fetch_seq64(name, beg, end) with end INCLUSIVE; returns 0 and appends the bases, -1 if name is unknown */
static int fai_fetch(const faidx *fx, sv name, uint64_t beg_u, uint64_t end_u, orc_buf *out) {
    const fai_ent *e = NULL;
    for (size_t i = 0; i < fx->n; i++)
        if (sv_eq(fx->e[i].name, name)) { e = &fx->e[i]; break; }
    if (!e) return -1;
    int64_t beg = (int64_t)beg_u, end = (int64_t)end_u, len = (int64_t)e->len;
    /* faidx_adjust_position(end_adjust = 1) */
    if (end < beg) beg = end;
    if (beg < 0) beg = 0; else if (len <= beg) beg = len;
    if (end < 0) end = 0; else if (len <= end) end = len - 1;
    /* fai_retrieve(offset, beg, end + 1): seek, then take isgraph() bytes until the count is reached */
    int64_t want = end + 1 - beg;
    if (want <= 0 || e->linebases == 0) return 0;
    uint64_t pos = e->offset + (uint64_t)beg / e->linebases * e->linewidth + (uint64_t)beg % e->linebases;
    int64_t l = 0;
    while (l < want && pos < fx->fa_len) {
        unsigned char c = (unsigned char)fx->fa[pos++];
        if (c > 32 && c < 127) { buf_putc(out, (char)c); l++; }
    }
    return 0;
}

/* utils.rs:83-101 reverse_complement; returns the offending char or 0 */
static char revcomp_inplace(char *s, size_t n) {
    for (size_t i = 0; i < n; i++) { /* validate in the reference's iteration order: from the END of the input */
        char c = s[n - 1 - i];
        if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N' || c == 'a' || c == 'c' || c == 'g' || c == 't' || c == 'n'))
            return c;
    }
    for (size_t i = 0; i < n / 2; i++) { char t = s[i]; s[i] = s[n - 1 - i]; s[n - 1 - i] = t; }
    for (size_t i = 0; i < n; i++) {
        switch (s[i]) {
        case 'A': s[i] = 'T'; break; case 'C': s[i] = 'G'; break; case 'G': s[i] = 'C'; break; case 'T': s[i] = 'A'; break;
        case 'a': s[i] = 't'; break; case 'c': s[i] = 'g'; break; case 'g': s[i] = 'c'; break; case 't': s[i] = 'a'; break;
        default: break;
        }
    }
    return 0;
}

/* String::insert_str(idx, "-" * n) on an orc_buf */
static int buf_insert_gaps(orc_buf *b, uint64_t idx, uint64_t n) {
    if (idx > b->n) return -1; /* insert_str panics when idx is past the end */
    buf_reserve(b, n);
    memmove(b->p + idx + n, b->p + idx, b->n - idx);
    memset(b->p + idx, '-', n);
    b->n += n;
    return 0;
}

int orc_paf2maf(const char *paf, size_t paf_len, const char *t_fa, size_t t_fa_len, const char *t_fai, size_t t_fai_len,
                const char *q_fa, size_t q_fa_len, const char *q_fai, size_t q_fai_len, const char *t_path,
                const char *q_path, orc_buf *out, char *err, size_t errcap) {
    faidx tf, qf;
    if (fai_parse(t_fai, t_fai_len, t_fa, t_fa_len, &tf) || fai_parse(q_fai, q_fai_len, q_fa, q_fa_len, &qf)) {
        set_err(err, errcap, "HTS library error by bad .fai");
        return ORC_E_CSV;
    }
    /* converter.rs:190-195 header */
    buf_puts(out, "#maf version=1.6 convert_from=paf t_seq_path=");
    buf_puts(out, t_path);
    buf_puts(out, " q_seq_path=");
    buf_puts(out, q_path);
    buf_putc(out, '\n');
    size_t pos = 0;
    pafrow row = {0};
    orc_buf ts = {0}, qs = {0}, conv = {0};
    int rc = ORC_OK;
    while (rc == ORC_OK && paf_next(paf, paf_len, &pos, &row)) {
        uint64_t u[12] = {0};
        int neg = 0, bad = row.n < 12;
        for (int k = 0; k < 12 && !bad; k++) {
            if (k == 0 || k == 5) continue;
            if (k == 4) { if (sv_eq_c(row.f[4], "+")) neg = 0; else if (sv_eq_c(row.f[4], "-")) neg = 1; else bad = 1; continue; }
            if (!csv_parse_u64(row.f[k], &u[k])) bad = 1;
        }
        if (bad) { set_err(err, errcap, "CSV deserialize error by: field parse"); rc = ORC_E_CSV; break; }
        sv qname = row.f[0], tname = row.f[5];
        uint64_t q_size = u[1], q_st = u[2], q_en = u[3], t_size = u[6], t_st = u[7], t_en = u[8], mapq = u[11];
        uint64_t q_start_maf = neg ? q_size - q_en : q_st; /* converter.rs:213-216 */
        ts.n = 0; qs.n = 0;
        if (fai_fetch(&tf, tname, t_st, t_en - 1, &ts) || fai_fetch(&qf, qname, q_st, q_en - 1, &qs)) {
            set_err(err, errcap, "HTS library error by sequence not found");
            rc = ORC_E_NOM; break;
        }
        if (neg) {
            char c = revcomp_inplace(qs.p, qs.n);
            if (c) { set_err(err, errcap, "Invalid Base: `%c`", c); rc = ORC_E_VCF_BASE; break; }
        }
        /* cigar.rs:492-551 parse_cigar_to_insert */
        sv cg = {0, 0};
        const int have = paf_get_cigar(&row, &conv, &cg);
        if (have < 0) { set_err(err, errcap, CS_PANIC_MSG); rc = ORC_E_PANIC; break; }
        if (!have) { set_err(err, errcap, "CIGAR start tag not found"); rc = ORC_E_CIGAR_TAG_NOT_FOUND; break; }
        if (cg.n == 0) { set_err(err, errcap, "panic: byte index 10 is out of bounds of ``"); rc = ORC_E_PANIC; break; }
        uint64_t off = 0;
        size_t i = 0;
        while (i < cg.n && rc == ORC_OK) {
            size_t d0 = i;
            while (i < cg.n && cg.p[i] >= '0' && cg.p[i] <= '9') i++;
            size_t o0 = i;
            while (i < cg.n && !(cg.p[i] >= '0' && cg.p[i] <= '9')) i++;
            sv digits = {cg.p + d0, o0 - d0}, op = {cg.p + o0, i - o0};
            uint64_t cnt;
            if (op.n != 1) { set_err(err, errcap, "CIGAR OP `%.*s` invalid", (int)op.n, op.p); rc = ORC_E_CIGAR_OP_INVALID; break; }
            if (!parse_u64(digits, &cnt)) { set_err(err, errcap, "Parse `%.*s` Into Integer Error", (int)digits.n, digits.p); rc = ORC_E_PARSE_INT; break; }
            switch (op.p[0]) {
            case 'M': case '=': case 'X': off += cnt; break;
            case 'I': if (buf_insert_gaps(&ts, off, cnt)) { set_err(err, errcap, "panic: insert_str index out of range"); rc = ORC_E_PANIC; } off += cnt; break;
            case 'D': if (buf_insert_gaps(&qs, off, cnt)) { set_err(err, errcap, "panic: insert_str index out of range"); rc = ORC_E_PANIC; } off += cnt; break;
            default: set_err(err, errcap, "CIGAR OP `%c` invalid", op.p[0]); rc = ORC_E_CIGAR_OP_INVALID; break;
            }
        }
        if (rc) break;
        /* maf.rs:566-581 MAFWriter::write_record */
        buf_puts(out, "a score="); buf_putu(out, mapq); buf_putc(out, '\n');
        buf_puts(out, "s\t"); buf_put(out, tname.p, tname.n); buf_putc(out, '\t'); buf_putu(out, t_st); buf_putc(out, '\t');
        buf_putu(out, t_en - t_st); buf_puts(out, "\t+\t"); buf_putu(out, t_size); buf_putc(out, '\t');
        buf_put(out, ts.p, ts.n); buf_putc(out, '\n');
        buf_puts(out, "s\t"); buf_put(out, qname.p, qname.n); buf_putc(out, '\t'); buf_putu(out, q_start_maf); buf_putc(out, '\t');
        buf_putu(out, q_en - q_st); buf_putc(out, '\t'); buf_putc(out, neg ? '-' : '+'); buf_putc(out, '\t'); buf_putu(out, q_size);
        buf_putc(out, '\t'); buf_put(out, qs.p, qs.n); buf_putc(out, '\n');
        buf_putc(out, '\n');
    }
    free(row.f); pafrow_free_scratch(&row); free(tf.e); free(qf.e);
    orc_buf_free(&ts); orc_buf_free(&qs); orc_buf_free(&conv);
    return rc;
}

/* ------------------------------------------------------------- call -f paf */
/* src/tools/caller.rs:268-302 call_var_paf + :610-822 call_within_var_paf.
 * Sequences come from faidx with an INCLUSIVE end (paf.rs:221-237: [start, end], one base more than the
 * alignment -- Q13); the query is NOT reverse-complemented for '-' (Q12); the <INV> row is not gated (Q12).
 * The fold closure's Result is discarded (caller.rs:673-815: `let (_, _) = fold_many1(..)(cigar)?`), so the
 * first failing token / row of a record silently ENDS that record (rows before it are kept) -- Q15.
 * Slices past the fetched strings and `pos - start - 1` underflows are panics.  All rows are buffered and
 * the header is written after the loop (caller.rs:283-299), so any error leaves the output empty. */
int orc_call_paf(const char *paf, size_t paf_len, const char *t_fa, size_t t_fa_len, const char *t_fai, size_t t_fai_len,
                 const char *q_fa, size_t q_fa_len, const char *q_fai, size_t q_fai_len, int snp, uint64_t svlen_cutoff,
                 const char *sample, orc_buf *out, char *err, size_t errcap) {
    faidx tf, qf;
    if (fai_parse(t_fai, t_fai_len, t_fa, t_fa_len, &tf) || fai_parse(q_fai, q_fai_len, q_fa, q_fa_len, &qf)) {
        set_err(err, errcap, "HTS library error by bad .fai");
        return ORC_E_CSV;
    }
    size_t pos = 0;
    pafrow row = {0};
    orc_buf ts = {0}, qs = {0}, body = {0}, conv = {0}, rowb = {0};
    int rc = ORC_OK;
    char info[256], tail[1200], dis[64];
    while (rc == ORC_OK && paf_next(paf, paf_len, &pos, &row)) {
        uint64_t u[12] = {0};
        int neg = 0, bad = row.n < 12;
        for (int k = 0; k < 12 && !bad; k++) {
            if (k == 0 || k == 5) continue;
            if (k == 4) { if (sv_eq_c(row.f[4], "+")) neg = 0; else if (sv_eq_c(row.f[4], "-")) neg = 1; else bad = 1; continue; }
            if (!csv_parse_u64(row.f[k], &u[k])) bad = 1;
        }
        if (bad) { set_err(err, errcap, "CSV deserialize error by: field parse"); rc = ORC_E_CSV; break; }
        sv qname = row.f[0], tname = row.f[5];
        const uint64_t q_st = u[2], q_en = u[3], t_st = u[7], t_en = u[8];
        ts.n = 0; qs.n = 0;
        if (fai_fetch(&tf, tname, t_st, t_en, &ts) || fai_fetch(&qf, qname, q_st, q_en, &qs)) {
            set_err(err, errcap, "HTS library error by sequence not found");
            rc = ORC_E_NOM; break;
        }
        const char sfx = neg ? 'N' : 'P';
        if (qname.n > 1000) { set_err(err, errcap, "oracle: query name too long"); rc = ORC_E_PANIC; break; }
        if (neg) { /* caller.rs:641-658 */
            if (ts.n < 1) { set_err(err, errcap, "panic: byte index 1 is out of bounds of ``"); rc = ORC_E_PANIC; break; }
            snprintf(info, sizeof info, "SVTYPE=INV;END=%llu", (unsigned long long)t_en);
            snprintf(tail, sizeof tail, "%.*s@%llu@%llu@%c", (int)qname.n, qname.p, (unsigned long long)q_st,
                     (unsigned long long)q_en, sfx);
            rc = vcf_row(&body, tname, t_st + 1, ts.p, 1, "<INV>", 5, 1, info, tail, err, errcap);
            if (rc) break;
        }
        /* paf.rs:122-140 get_cigar_string */
        sv cg = {0, 0};
        const int have = paf_get_cigar(&row, &conv, &cg);
        if (have < 0) { set_err(err, errcap, CS_PANIC_MSG); rc = ORC_E_PANIC; break; }
        if (!have) { set_err(err, errcap, "CIGAR start tag not found"); rc = ORC_E_CIGAR_TAG_NOT_FOUND; break; }
        if (cg.n == 0) { set_err(err, errcap, "panic: byte index 10 is out of bounds of ``"); rc = ORC_E_PANIC; break; }
        uint64_t t_pos = t_st, q_pos = q_st;
        int after_m = 0, stopped = 0;
        size_t i = 0;
        while (i < cg.n && !stopped && rc == ORC_OK) {
            size_t d0 = i;
            while (i < cg.n && cg.p[i] >= '0' && cg.p[i] <= '9') i++;
            size_t o0 = i;
            while (i < cg.n && !(cg.p[i] >= '0' && cg.p[i] <= '9')) i++;
            sv digits = {cg.p + d0, o0 - d0}, op = {cg.p + o0, i - o0};
            uint64_t len;
            if (op.n != 1 || !parse_u64(digits, &len)) { stopped = 1; break; } /* cst2cu error, discarded */
            switch (op.p[0]) {
            case 'M': case '=': t_pos += len; q_pos += len; after_m = 1; break;
            case 'X':
                if (snp) {
                    for (uint64_t k = 0; k < len; k++) {
                        const uint64_t a = t_pos - t_st, b = q_pos - q_st;
                        if (a + 1 > ts.n || b + 1 > qs.n) { set_err(err, errcap, "panic: byte index out of range"); rc = ORC_E_PANIC; break; }
                        snprintf(tail, sizeof tail, "%.*s@%llu@%c", (int)qname.n, qname.p, (unsigned long long)q_pos, sfx);
                        rowb.n = 0;
                        if (vcf_row(&rowb, tname, t_pos + 1, ts.p + a, 1, qs.p + b, 1, 0, NULL, tail, dis, sizeof dis)) { stopped = 1; break; }
                        buf_put(&body, rowb.p, rowb.n);
                        t_pos++; q_pos++;
                    }
                } else { t_pos += len; q_pos += len; }
                after_m = 1;
                break;
            case 'I':
                if (len > svlen_cutoff && after_m) {
                    if (t_pos == t_st || q_pos == q_st) { set_err(err, errcap, "panic: attempt to subtract with overflow"); rc = ORC_E_PANIC; break; }
                    const uint64_t a = t_pos - t_st - 1, b = q_pos - q_st - 1;
                    if (a + 1 > ts.n || len > qs.n || b + len + 1 > qs.n) { set_err(err, errcap, "panic: byte index out of range"); rc = ORC_E_PANIC; break; }
                    snprintf(info, sizeof info, "%sSVTYPE=INS;SVLEN=%llu;END=%llu", neg ? "INV_NEST=TRUE;" : "",
                             (unsigned long long)len, (unsigned long long)t_pos);
                    snprintf(tail, sizeof tail, "%.*s@%llu@%llu@%c", (int)qname.n, qname.p, (unsigned long long)q_pos,
                             (unsigned long long)(q_pos + len), sfx);
                    rowb.n = 0;
                    if (vcf_row(&rowb, tname, t_pos, ts.p + a, 1, qs.p + b, len + 1, 0, info, tail, dis, sizeof dis)) { stopped = 1; break; }
                    buf_put(&body, rowb.p, rowb.n);
                }
                q_pos += len; after_m = 0;
                break;
            case 'D':
                if (len > svlen_cutoff && after_m) {
                    if (t_pos == t_st || q_pos == q_st) { set_err(err, errcap, "panic: attempt to subtract with overflow"); rc = ORC_E_PANIC; break; }
                    const uint64_t a = t_pos - t_st - 1, b = q_pos - q_st - 1;
                    if (len > ts.n || a + len + 1 > ts.n || b + 1 > qs.n) { set_err(err, errcap, "panic: byte index out of range"); rc = ORC_E_PANIC; break; }
                    snprintf(info, sizeof info, "%sSVTYPE=DEL;SVLEN=%llu;END=%llu", neg ? "INV_NEST=TRUE;" : "",
                             (unsigned long long)len, (unsigned long long)(t_pos + len));
                    snprintf(tail, sizeof tail, "%.*s@%llu@%llu@%c", (int)qname.n, qname.p, (unsigned long long)q_pos,
                             (unsigned long long)q_pos, sfx);
                    rowb.n = 0;
                    if (vcf_row(&rowb, tname, t_pos, ts.p + a, len + 1, qs.p + b, 1, 0, info, tail, dis, sizeof dis)) { stopped = 1; break; }
                    buf_put(&body, rowb.p, rowb.n);
                }
                t_pos += len; after_m = 0;
                break;
            default: stopped = 1; break; /* CigarOpInvalid, discarded */
            }
        }
    }
    if (rc == ORC_OK) { /* caller.rs:294-299 */
        vcf_header(out, sample ? sample : "sample");
        buf_put(out, body.p, body.n);
    }
    free(row.f); pafrow_free_scratch(&row); free(tf.e); free(qf.e);
    orc_buf_free(&ts); orc_buf_free(&qs); orc_buf_free(&body); orc_buf_free(&conv); orc_buf_free(&rowb);
    return rc;
}

/* ---------------------------------------------------------------- pafpseudo */
/* src/tools/pseudomaf.rs:18-237 generate_pesudo_maf / write_pmaf / get_sline_seq and
 * src/parser/cigar.rs:744-804 gen_pesudo_maf_by_cigar.
 * Iteration order of the two HashMaps (targets; queries inside a target) is nondeterministic in the reference; here
 * it is FIRST-OCCURRENCE order (the same convention as the stat rows).  Records with equal target_start
 * inside one (target, query) group are placed by slice::binary_search_by's tie behaviour (pseudomaf.rs:91-94); this
 * restates the three-way loop of core::slice 1.52-1.81; the reference's flake.lock pins nixpkgs-unstable of 2024-01-22,
 * i.e. rustc 1.75 -- parity unpinned beyond that (no golden vector).
 * out = the files' bytes concatenated in target order; index = one line "{target}\t{bytes}\n" per file
 * ({out_dir}/{target}.maf).  On error nothing is returned (the reference may leave other targets' files behind). */
typedef struct {
    sv qname, tname, cg, cs;
    uint64_t q_size, q_st, q_en, t_size, t_st, t_en;
    int neg, have_cg, have_cs;
} pprec;

static int pp_project(const pprec *r, orc_buf *q, int base, char *err, size_t errcap) {
    /* cigar.rs:744-804 */
    static __thread orc_buf conv; /* the converted cs:Z: of this record (paf.rs:129-135), reused between calls */
    sv cg = r->cg;
    if (!r->have_cg) {
        if (!r->have_cs) { set_err(err, errcap, "CIGAR start tag not found"); return ORC_E_CIGAR_TAG_NOT_FOUND; }
        conv.n = 0;
        buf_reserve(&conv, 1);
        if (cs_to_cigar(r->cs, &conv)) { set_err(err, errcap, CS_PANIC_MSG); return ORC_E_PANIC; }
        cg.p = conv.p; cg.n = conv.n;
    }
    if (cg.n == 0) { set_err(err, errcap, "panic: byte index 10 is out of bounds of ``"); return ORC_E_PANIC; }
    uint64_t off = 0;
    size_t i = 0;
    while (i < cg.n) {
        size_t d0 = i;
        while (i < cg.n && cg.p[i] >= '0' && cg.p[i] <= '9') i++;
        size_t o0 = i;
        while (i < cg.n && !(cg.p[i] >= '0' && cg.p[i] <= '9')) i++;
        sv digits = {cg.p + d0, o0 - d0}, op = {cg.p + o0, i - o0};
        uint64_t len;
        if (op.n != 1) { set_err(err, errcap, "CIGAR OP `%.*s` invalid", (int)op.n, op.p); return ORC_E_CIGAR_OP_INVALID; }
        if (!parse_u64(digits, &len)) { set_err(err, errcap, "Parse `%.*s` Into Integer Error", (int)digits.n, digits.p); return ORC_E_PARSE_INT; }
        switch (op.p[0]) {
        case 'M': case '=':
            if (base) off += len; else { buf_reserve(q, len); memset(q->p + q->n, '1', len); q->n += len; }
            break;
        case 'X':
            if (base) off += len; else { buf_reserve(q, len); memset(q->p + q->n, '0', len); q->n += len; }
            break;
        case 'I': case 'S':
            if (base) { /* String::drain(off..off+len) */
                if (len > q->n || off > q->n - len) { set_err(err, errcap, "panic: drain range out of bounds"); return ORC_E_PANIC; }
                memmove(q->p + off, q->p + off + len, q->n - off - len);
                q->n -= len;
            }
            break;
        case 'D':
            if (base) {
                if (buf_insert_gaps(q, off, len)) { set_err(err, errcap, "panic: insert_str index out of range"); return ORC_E_PANIC; }
                off += len;
            } else { buf_reserve(q, len); memset(q->p + q->n, '-', len); q->n += len; }
            break;
        default: break; /* other ops are ignored (cigar.rs:797) */
        }
    }
    return ORC_OK;
}

int orc_paf_pseudo(const char *paf, size_t paf_len, const char *fa, size_t fa_len, const char *fai, size_t fai_len,
                   const char *target, orc_buf *out, orc_buf *index, char *err, size_t errcap) {
    const int base = fa != NULL;
    faidx fx = {0};
    if (base && fai_parse(fai, fai_len, fa, fa_len, &fx)) { set_err(err, errcap, "HTS library error by bad .fai"); return ORC_E_CSV; }
    size_t pos = 0;
    pafrow row = {0};
    pprec *recs = NULL;
    size_t n = 0, cap = 0;
    int rc = ORC_OK;
    /* pseudomaf.rs:24-43: read everything (csv errors first) */
    while (paf_next(paf, paf_len, &pos, &row)) {
        uint64_t u[12] = {0};
        int neg = 0, bad = row.n < 12;
        for (int k = 0; k < 12 && !bad; k++) {
            if (k == 0 || k == 5) continue;
            if (k == 4) { if (sv_eq_c(row.f[4], "+")) neg = 0; else if (sv_eq_c(row.f[4], "-")) neg = 1; else bad = 1; continue; }
            if (!csv_parse_u64(row.f[k], &u[k])) bad = 1;
        }
        if (bad) { set_err(err, errcap, "CSV deserialize error by: field parse"); rc = ORC_E_CSV; break; }
        if (target && !sv_eq_c(row.f[5], target)) continue;
        if (n == cap) { cap = cap ? cap * 2 : 256; recs = (pprec *)realloc(recs, cap * sizeof(pprec)); }
        pprec *r = &recs[n++];
        r->qname = row.f[0]; r->tname = row.f[5];
        r->q_size = u[1]; r->q_st = u[2]; r->q_en = u[3]; r->t_size = u[6]; r->t_st = u[7]; r->t_en = u[8];
        r->neg = neg; r->have_cg = 0; r->cg.p = NULL; r->cg.n = 0; r->have_cs = 0; r->cs.p = NULL; r->cs.n = 0;
        for (size_t k = 12; k < row.n && !r->have_cg; k++)
            if (row.f[k].n >= 5 && memcmp(row.f[k].p, "cg:Z:", 5) == 0) { r->cg.p = row.f[k].p + 5; r->cg.n = row.f[k].n - 5; r->have_cg = 1; }
        for (size_t k = 12; k < row.n && !r->have_cg && !r->have_cs; k++)
            if (row.f[k].n >= 5 && memcmp(row.f[k].p, "cs:Z:", 5) == 0) { r->cs.p = row.f[k].p + 5; r->cs.n = row.f[k].n - 5; r->have_cs = 1; }
    }
    uint8_t *done = (uint8_t *)calloc(n ? n : 1, 1);
    size_t *grp = (size_t *)malloc((n ? n : 1) * sizeof(size_t));
    orc_buf q = {0};
    for (size_t a = 0; a < n && rc == ORC_OK; a++) { /* targets in first-occurrence order */
        if (done[a] & 1) continue;
        const sv tname = recs[a].tname;
        const size_t file_start = out->n;
        buf_puts(out, "a score=0\n");
        int first_flag = 1;
        uint64_t target_size = 0;
        for (size_t b = a; b < n && rc == ORC_OK; b++) { /* queries of this target in first-occurrence order */
            if ((done[b] & 1) || !sv_eq(recs[b].tname, tname)) continue;
            const sv qname = recs[b].qname;
            size_t m = 0;
            for (size_t c = b; c < n; c++) {
                if ((done[c] & 1) || !sv_eq(recs[c].tname, tname) || !sv_eq(recs[c].qname, qname)) continue;
                done[c] |= 1;
                /* pseudomaf.rs:91-94 binary_search_by(target_start) + insert */
                size_t size = m, left = 0, right = m, idx;
                int found = 0;
                while (left < right) {
                    size_t mid = left + size / 2;
                    uint64_t pv = recs[grp[mid]].t_st, kv = recs[c].t_st;
                    if (pv < kv) left = mid + 1; else if (pv > kv) right = mid; else { idx = mid; found = 1; break; }
                    size = right - left;
                }
                if (!found) idx = left;
                memmove(grp + idx + 1, grp + idx, (m - idx) * sizeof(size_t));
                grp[idx] = c;
                m++;
            }
            int first_query_flag = 1;
            uint64_t last_target_end = 0;
            for (size_t k = 0; k < m && rc == ORC_OK; k++) {
                const pprec *r = &recs[grp[k]];
                target_size = r->t_size;
                if (first_flag) { /* pseudomaf.rs:121-132 */
                    buf_puts(out, "s\t"); buf_put(out, tname.p, tname.n); buf_puts(out, "\t0\t"); buf_putu(out, target_size);
                    buf_puts(out, "\t+\t"); buf_putu(out, target_size); buf_putc(out, '\t');
                    if (base) {
                        if (target_size == 0) { set_err(err, errcap, "panic: attempt to subtract with overflow"); rc = ORC_E_PANIC; break; }
                        if (fai_fetch(&fx, tname, 0, target_size - 1, out)) { set_err(err, errcap, "HTS library error by sequence not found"); rc = ORC_E_NOM; break; }
                    } else { buf_reserve(out, target_size); memset(out->p + out->n, 'N', target_size); out->n += target_size; }
                    buf_putc(out, '\n');
                    first_flag = 0;
                }
                if (first_query_flag) {
                    buf_puts(out, "s\t"); buf_put(out, r->qname.p, r->qname.n); buf_puts(out, "\t0\t"); buf_putu(out, r->q_size);
                    buf_puts(out, "\t+\t"); buf_putu(out, r->q_size); buf_putc(out, '\t');
                }
                uint64_t overlap = 0;
                if (r->t_st > last_target_end) {
                    uint64_t gap = r->t_st - last_target_end;
                    buf_reserve(out, gap); memset(out->p + out->n, '-', gap); out->n += gap;
                } else {
                    if (last_target_end > r->t_en) continue; /* contained: skipped (first_query_flag stays) */
                    overlap = last_target_end - r->t_st;
                }
                last_target_end = r->t_en;
                q.n = 0;
                if (base) { /* get_sline_seq: [q_start, q_end - 1] */
                    if (r->q_en == 0) { set_err(err, errcap, "panic: attempt to subtract with overflow"); rc = ORC_E_PANIC; break; }
                    if (fai_fetch(&fx, qname, r->q_st, r->q_en - 1, &q)) { set_err(err, errcap, "HTS library error by sequence not found"); rc = ORC_E_NOM; break; }
                }
                if (r->neg) {
                    char c = revcomp_inplace(q.p, q.n);
                    if (c) { set_err(err, errcap, "Invalid Base: `%c`", c); rc = ORC_E_VCF_BASE; break; }
                }
                rc = pp_project(r, &q, base, err, errcap);
                if (rc) break;
                if (overlap > 0) {
                    if (overlap > q.n) { set_err(err, errcap, "panic: drain range out of bounds"); rc = ORC_E_PANIC; break; }
                    buf_put(out, q.p + overlap, q.n - overlap);
                } else buf_put(out, q.p, q.n);
                first_query_flag = 0;
            }
            if (rc) break;
            if (last_target_end > target_size) { set_err(err, errcap, "panic: attempt to subtract with overflow"); rc = ORC_E_PANIC; break; }
            uint64_t tail = target_size - last_target_end;
            buf_reserve(out, tail); memset(out->p + out->n, '-', tail); out->n += tail;
            buf_putc(out, '\n');
        }
        if (rc) break;
        buf_putc(out, '\n');
        buf_put(index, tname.p, tname.n); buf_putc(index, '\t'); buf_putu(index, out->n - file_start); buf_putc(index, '\n');
    }
    if (rc) { out->n = 0; index->n = 0; }
    free(done); free(grp); free(recs); free(row.f); pafrow_free_scratch(&row); free(fx.e);
    orc_buf_free(&q);
    return rc;
}

/* ---------------------------------------------------------------- validate */
/* src/tools/validate.rs:44-147 parallel_validatepaf: expected ends from the CIGAR counters.  The reduce order of the
 * rayon fold is unspecified; here (and in the CUDA path) the lists and the fixed records come in FILE order.
 * `rec.get_stat().unwrap()` (validate.rs:79) turns every CIGAR / tag error into a panic.  fix != 0: all records,
 * with corrected ends, re-serialised by csv (12 typed columns in decimal, then the tags). */
int orc_paf_validate(const char *text, size_t len, int fix, orc_buf *report, orc_buf *fixed, char *err, size_t errcap) {
    size_t pos = 0;
    pafrow row = {0};
    orc_buf ql = {0}, tl = {0}, conv = {0};
    uint64_t total = 0, qbad = 0, tbad = 0;
    int rc = ORC_OK;
    while (paf_next(text, len, &pos, &row)) {
        uint64_t u[12] = {0};
        int neg = 0, bad = row.n < 12;
        for (int k = 0; k < 12 && !bad; k++) {
            if (k == 0 || k == 5) continue;
            if (k == 4) { if (sv_eq_c(row.f[4], "+")) neg = 0; else if (sv_eq_c(row.f[4], "-")) neg = 1; else bad = 1; continue; }
            if (!csv_parse_u64(row.f[k], &u[k])) bad = 1;
        }
        if (bad) { set_err(err, errcap, "CSV deserialize error by: field parse"); rc = ORC_E_CSV; break; }
        total++;
        sv cg = {0, 0};
        const int have = paf_get_cigar(&row, &conv, &cg);
        if (have < 0) { set_err(err, errcap, CS_PANIC_MSG); rc = ORC_E_PANIC; break; }
        orc_counts c;
        char dis[256];
        if (!have || orc_parse_paf_cigar(cg.p, cg.n, neg, &c, dis, sizeof dis)) {
            set_err(err, errcap, "panic: called `Result::unwrap()` on an `Err` value");
            rc = ORC_E_PANIC;
            break;
        }
        const uint64_t exp_q = u[2] + c.c[0] + c.c[1] + c.c[3] + c.c[7], exp_t = u[7] + c.c[0] + c.c[1] + c.c[5] + c.c[9];
        uint64_t q_end = u[3], t_end = u[8];
        if (exp_q != q_end) {
            qbad++;
            buf_put(&ql, row.f[0].p, row.f[0].n); buf_putc(&ql, ':'); buf_putu(&ql, u[2]); buf_putc(&ql, '-'); buf_putu(&ql, q_end);
            buf_putc(&ql, '\n');
            q_end = exp_q;
        }
        if (exp_t != t_end) {
            tbad++;
            buf_put(&tl, row.f[5].p, row.f[5].n); buf_putc(&tl, ':'); buf_putu(&tl, u[7]); buf_putc(&tl, '-'); buf_putu(&tl, t_end);
            buf_putc(&tl, '\n');
            t_end = exp_t;
        }
        if (fix) { /* validate.rs:136-144 csv serialize of PafRecord */
            csv_field(fixed, row.f[0], 1); csv_u64(fixed, u[1], 0); csv_u64(fixed, u[2], 0); csv_u64(fixed, q_end, 0);
            buf_putc(fixed, '\t'); buf_putc(fixed, neg ? '-' : '+');
            csv_field(fixed, row.f[5], 0); csv_u64(fixed, u[6], 0); csv_u64(fixed, u[7], 0); csv_u64(fixed, t_end, 0);
            csv_u64(fixed, u[9], 0); csv_u64(fixed, u[10], 0); csv_u64(fixed, u[11], 0);
            for (size_t k = 12; k < row.n; k++) csv_field(fixed, row.f[k], 0);
            buf_putc(fixed, '\n');
        }
    }
    if (rc == ORC_OK) { /* impl Display for Validations + writeln!(writer, "{}", ..) */
        buf_puts(report, "Total records: "); buf_putu(report, total);
        buf_puts(report, "\nQuery invalid records: "); buf_putu(report, qbad);
        buf_puts(report, "\nTarget invalid records: "); buf_putu(report, tbad);
        buf_puts(report, "\nQuery invalid list:\n"); buf_put(report, ql.p, ql.n);
        buf_puts(report, "Target invalid list:\n"); buf_put(report, tl.p, tl.n);
        buf_putc(report, '\n');
    } else fixed->n = 0;
    free(row.f); pafrow_free_scratch(&row);
    orc_buf_free(&ql); orc_buf_free(&tl); orc_buf_free(&conv);
    return rc;
}


/* -------------------------------------------------------------------- filter -f paf */
/* src/tools/filter.rs:41-63 filter_paf + :92-106 filter_alignrec; :108-160 filter_paf_align_pair (have_align != 0: the
 * per-(query_name, target_name) sum of target_end - target_start must reach min_align; -b / -q are ignored then,
 * src/utils.rs:559-563).  Records come out in file order (the reference's par_bridge order is unspecified); the kept
 * records go through the csv writer.  u64 arithmetic wraps like release Rust.  filter_paf streams (rows before a bad
 * record are already written); nothing is returned on error here. */
int orc_paf_filter(const char *text, size_t len, uint64_t min_block, uint64_t min_query, int have_align, uint64_t min_align,
                   orc_buf *out, char *err, size_t errcap) {
    typedef struct { sv q, t; uint64_t sum; } fpair;
    fpair *pairs = NULL;
    size_t np = 0, capp = 0;
    int rc = ORC_OK;
    for (int pass = 0; pass < 2 && rc == ORC_OK; pass++) {
        size_t pos = 0;
        pafrow row = {0};
        while (paf_next(text, len, &pos, &row)) {
            uint64_t u[12] = {0};
            int neg = 0, bad = row.n < 12;
            for (int k = 0; k < 12 && !bad; k++) {
                if (k == 0 || k == 5) continue;
                if (k == 4) { if (sv_eq_c(row.f[4], "+")) neg = 0; else if (sv_eq_c(row.f[4], "-")) neg = 1; else bad = 1; continue; }
                if (!csv_parse_u64(row.f[k], &u[k])) bad = 1;
            }
            if (bad) { set_err(err, errcap, "CSV deserialize error by: field parse"); rc = ORC_E_CSV; break; }
            const uint64_t block = u[8] - u[7];
            int keep;
            if (have_align) {
                size_t i = 0;
                while (i < np && !(sv_eq(pairs[i].q, row.f[0]) && sv_eq(pairs[i].t, row.f[5]))) i++;
                if (pass == 0) {
                    if (i == np) {
                        if (np == capp) { capp = capp ? capp * 2 : 64; pairs = (fpair *)realloc(pairs, capp * sizeof(fpair)); }
                        char *cq = (char *)malloc(row.f[0].n + 1), *ct = (char *)malloc(row.f[5].n + 1); /* quoted names live in row scratch */
                        memcpy(cq, row.f[0].p, row.f[0].n); memcpy(ct, row.f[5].p, row.f[5].n);
                        pairs[np].q.p = cq; pairs[np].q.n = row.f[0].n; pairs[np].t.p = ct; pairs[np].t.n = row.f[5].n;
                        pairs[np].sum = 0; np++;
                    }
                    pairs[i].sum += block;
                    continue;
                }
                keep = pairs[i].sum >= min_align;
            } else {
                if (pass == 0) break; /* one pass is enough */
                keep = !((block < min_block) | (u[1] < min_query));
            }
            if (keep) {
                csv_field(out, row.f[0], 1); csv_u64(out, u[1], 0); csv_u64(out, u[2], 0); csv_u64(out, u[3], 0);
                buf_putc(out, '\t'); buf_putc(out, neg ? '-' : '+');
                csv_field(out, row.f[5], 0); csv_u64(out, u[6], 0); csv_u64(out, u[7], 0); csv_u64(out, u[8], 0);
                csv_u64(out, u[9], 0); csv_u64(out, u[10], 0); csv_u64(out, u[11], 0);
                for (size_t k = 12; k < row.n; k++) csv_field(out, row.f[k], 0);
                buf_putc(out, '\n');
            }
        }
        free(row.f); pafrow_free_scratch(&row);
    }
    if (rc != ORC_OK) out->n = 0;
    for (size_t i = 0; i < np; i++) { free((void *)pairs[i].q.p); free((void *)pairs[i].t.p); }
    free(pairs);
    return rc;
}


/* -------------------------------------------------------------------- dotplot -m base-level: the plot data */
/* src/parser/cigar.rs:807-914 reserve_query_start_end + emit_baseplotdatas, :917-952 parse_cigar_to_base_plotdata,
 * :955-985 parse_maf_to_base_plotdata; rendered as dotplot.rs:283-289 does for --out-format csv (csv::Writer with the
 * struct's header row, written when there is at least one segment; struct BasePlotdata dotplot.rs:181-190).  Records in
 * file order (the reference's par_bridge order is unspecified).  Pinned on test/test.html (the same data as JSON). */
typedef struct { uint64_t rs, re, qs, qe; char c; } plotseg;
typedef struct { plotseg *v; size_t n, cap; uint64_t ref, qry; int last_m; } plotstate;
static void plot_push(plotstate *st, uint64_t rs, uint64_t re, uint64_t qs, uint64_t qe, char c, int neg) {
    if (st->n == st->cap) { st->cap = st->cap ? st->cap * 2 : 16; st->v = (plotseg *)realloc(st->v, st->cap * sizeof(plotseg)); }
    plotseg g = {rs, re, qs, qe, c};
    if (neg) { g.qs = qe; g.qe = qs; } /* reserve_query_start_end */
    st->v[st->n++] = g;
}
static void plot_emit(plotstate *st, char op, uint64_t length, uint64_t cutoff, int neg) { /* cigar.rs:814-914 */
    if (op == 'M' || op == '=' || op == 'X') {
        const uint64_t re = st->ref + length, qe = st->qry + length;
        if (!st->last_m) plot_push(st, st->ref, re, st->qry, qe, 'M', neg);
        else {
            plotseg *m = &st->v[st->n - 1];
            m->re = re;
            if (neg) m->qs = qe; else m->qe = qe;
        }
        st->ref += length; st->qry += length;
        st->last_m = 1;
    } else if (op == 'I') {
        const uint64_t qe = st->qry + length;
        if (length > cutoff) { plot_push(st, st->ref, st->ref, st->qry, qe, 'I', neg); st->last_m = 0; }
        else if (st->last_m) { plotseg *m = &st->v[st->n - 1]; if (neg) m->qs = qe; else m->qe = qe; }
        st->qry += length;
    } else if (op == 'D') {
        const uint64_t re = st->ref + length;
        if (length > cutoff) { plot_push(st, st->ref, re, st->qry, st->qry, 'D', neg); st->last_m = 0; }
        else if (st->last_m) st->v[st->n - 1].re = re;
        st->ref += length;
    }
}
static void csvc_field(orc_buf *out, sv f) { /* csv::Writer default: ',' delimiter, QuoteStyle::Necessary */
    int need = 0;
    for (size_t i = 0; i < f.n; i++) if (f.p[i] == '"' || f.p[i] == ',' || f.p[i] == '\r' || f.p[i] == '\n') need = 1;
    if (!need) { buf_put(out, f.p, f.n); return; }
    buf_putc(out, '"');
    for (size_t i = 0; i < f.n; i++) { if (f.p[i] == '"') buf_putc(out, '"'); buf_putc(out, f.p[i]); }
    buf_putc(out, '"');
}
static void plot_write(orc_buf *out, const plotstate *st, sv ref_chro, sv query_chro) {
    for (size_t i = 0; i < st->n; i++) {
        const plotseg *g = &st->v[i];
        if (out->n == 0) buf_puts(out, "ref_start,ref_end,query_start,query_end,cigar,ref_chro,query_chro\n");
        buf_putu(out, g->rs); buf_putc(out, ','); buf_putu(out, g->re); buf_putc(out, ','); buf_putu(out, g->qs); buf_putc(out, ',');
        buf_putu(out, g->qe); buf_putc(out, ','); buf_putc(out, g->c); buf_putc(out, ',');
        csvc_field(out, ref_chro); buf_putc(out, ','); csvc_field(out, query_chro); buf_putc(out, '\n');
    }
}

int orc_maf_plotdata(const char *text, size_t len, const char *query_name, uint64_t cutoff, orc_buf *out, char *err, size_t errcap) {
    lreader rd;
    maf_reader_init(&rd, text, len);
    mafrec rec = {0};
    plotstate st = {0};
    int rc;
    while ((rc = maf_next(&rd, &rec, err, errcap)) > 0) {
        if (query_name) {
            int e = maf_set_query_byname(&rec, query_name, err, errcap);
            if (e) { rc = -e; break; }
        }
        if (rec.query_idx >= rec.n) {
            set_err(err, errcap, "panic: index out of bounds: the len is %zu but the index is %zu", rec.n, rec.query_idx);
            rc = -ORC_E_PANIC;
            break;
        }
        const sline *t = &rec.s[0], *q = &rec.s[rec.query_idx];
        const size_t n = t->seq.n < q->seq.n ? t->seq.n : q->seq.n;
        st.n = 0; st.ref = t->start; st.qry = maf_query_start(&rec); st.last_m = 0;
        for (size_t i = 0; i < n;) {
            char k = cigar_cat_ext(t->seq.p[i], q->seq.p[i]);
            size_t j = i;
            while (j < n && cigar_cat_ext(t->seq.p[j], q->seq.p[j]) == k) j++;
            plot_emit(&st, k, j - i, cutoff, q->neg);
            i = j;
        }
        plot_write(out, &st, t->name, q->name);
    }
    free(rec.s); free(st.v);
    if (rc < 0) out->n = 0;
    return rc < 0 ? -rc : ORC_OK;
}

int orc_paf_plotdata(const char *text, size_t len, uint64_t cutoff, orc_buf *out, char *err, size_t errcap) {
    size_t pos = 0;
    pafrow row = {0};
    orc_buf conv = {0};
    plotstate st = {0};
    int rc = ORC_OK;
    while (paf_next(text, len, &pos, &row)) {
        uint64_t u[12] = {0};
        int neg = 0, bad = row.n < 12;
        for (int k = 0; k < 12 && !bad; k++) {
            if (k == 0 || k == 5) continue;
            if (k == 4) { if (sv_eq_c(row.f[4], "+")) neg = 0; else if (sv_eq_c(row.f[4], "-")) neg = 1; else bad = 1; continue; }
            if (!csv_parse_u64(row.f[k], &u[k])) bad = 1;
        }
        if (bad) { set_err(err, errcap, "CSV deserialize error by: field parse"); rc = ORC_E_CSV; break; }
        sv cg = {0, 0};
        const int have = paf_get_cigar(&row, &conv, &cg);
        if (have < 0) { set_err(err, errcap, CS_PANIC_MSG); rc = ORC_E_PANIC; break; }
        if (!have) { set_err(err, errcap, "CIGAR start tag not found"); rc = ORC_E_CIGAR_TAG_NOT_FOUND; break; }
        if (cg.n == 0) { set_err(err, errcap, "panic: byte index 10 is out of bounds of ``"); rc = ORC_E_PANIC; break; }
        st.n = 0; st.ref = u[7]; st.qry = u[2]; st.last_m = 0;
        for (size_t i = 0; i < cg.n && rc == ORC_OK;) { /* cigar.rs:43-75 tokens: digits* then non-digits* */
            size_t d0 = i;
            while (i < cg.n && cg.p[i] >= '0' && cg.p[i] <= '9') i++;
            size_t o0 = i;
            while (i < cg.n && !(cg.p[i] >= '0' && cg.p[i] <= '9')) i++;
            sv digits = {cg.p + d0, o0 - d0}, op = {cg.p + o0, i - o0};
            uint64_t l;
            if (op.n != 1) { set_err(err, errcap, "CIGAR OP `%.*s` invalid", (int)op.n, op.p); rc = ORC_E_CIGAR_OP_INVALID; break; }
            if (!parse_u64(digits, &l)) { set_err(err, errcap, "Parse `%.*s` Into Integer Error", (int)digits.n, digits.p); rc = ORC_E_PARSE_INT; break; }
            plot_emit(&st, op.p[0], l, cutoff, neg); /* ops other than M = X I D: the `_ => {}` arm */
        }
        if (rc != ORC_OK) break;
        plot_write(out, &st, row.f[5], row.f[0]);
    }
    if (rc != ORC_OK) out->n = 0;
    free(row.f); pafrow_free_scratch(&row); free(st.v);
    orc_buf_free(&conv);
    return rc;
}

/* -------------------------------------------------------------------- chain */
/* src/converter.rs:57-92 maf2chain / :148-173 paf2chain; src/parser/chain.rs:92-204 (ChainDataLine / ChainHeader
 * Display, TryFrom<&MAFRecord> / TryFrom<&PafRecord> with the negative-strand quirk: query.end is computed from the
 * ALREADY UPDATED query.start, chain.rs:133-138,174-179); src/parser/cigar.rs:155-295 (head/tail indel trim,
 * parse_cigar_to_chain) and :435-490 (parse_maf_seq_to_chain, cigar_unit_chain).  u64 arithmetic wraps like release
 * Rust.  The reference streams: on an error the records before it are already written; here the buffer keeps them too. */
typedef struct { uint64_t size, qdiff, tdiff; } chainline;
static void chain_unit(orc_buf *out, chainline *d, char op, uint64_t count) { /* cigar.rs:458-490 */
    if (op == 'M') {
        if (d->size != 0 && d->tdiff + d->qdiff != 0) {
            buf_putc(out, '\n'); buf_putu(out, d->size); buf_putc(out, '\t'); buf_putu(out, d->qdiff); buf_putc(out, '\t');
            buf_putu(out, d->tdiff);
            d->size = 0;
        }
        d->size += count;
        d->tdiff = 0; d->qdiff = 0;
    } else if (op == 'I') d->tdiff += count;
    else d->qdiff += count;
}
static void chain_header(orc_buf *out, sv tname, uint64_t tsize, uint64_t tstart, uint64_t tend, sv qname, uint64_t qsize,
                         int neg, uint64_t qstart, uint64_t qend, uint64_t hi, uint64_t hd, uint64_t ti, uint64_t td, uint64_t id) {
    tstart += hd; tend -= td;
    if (!neg) { qstart += hi; qend -= ti; }
    else { qstart = qsize - (qend - hi); qend = qsize - (qstart + ti); } /* chain.rs:133-138: uses the NEW start */
    buf_puts(out, "chain\t255\t"); buf_put(out, tname.p, tname.n); buf_putc(out, '\t'); buf_putu(out, tsize); buf_puts(out, "\t+\t");
    buf_putu(out, tstart); buf_putc(out, '\t'); buf_putu(out, tend); buf_putc(out, '\t');
    buf_put(out, qname.p, qname.n); buf_putc(out, '\t'); buf_putu(out, qsize); buf_putc(out, '\t'); buf_putc(out, neg ? '-' : '+');
    buf_putc(out, '\t'); buf_putu(out, qstart); buf_putc(out, '\t'); buf_putu(out, qend); buf_putc(out, '\t'); buf_putu(out, id);
}

int orc_maf2chain(const char *text, size_t len, const char *query_name, orc_buf *out, char *err, size_t errcap) {
    lreader rd;
    maf_reader_init(&rd, text, len);
    mafrec rec = {0};
    int rc;
    uint64_t id = 0;
    while ((rc = maf_next(&rd, &rec, err, errcap)) > 0) {
        if (query_name) {
            int e = maf_set_query_byname(&rec, query_name, err, errcap);
            if (e) { rc = -e; break; }
        }
        if (rec.query_idx >= rec.n) {
            set_err(err, errcap, "panic: index out of bounds: the len is %zu but the index is %zu", rec.n, rec.query_idx);
            rc = -ORC_E_PANIC;
            break;
        }
        const sline *t = &rec.s[0], *q = &rec.s[rec.query_idx];
        const size_t n = t->seq.n < q->seq.n ? t->seq.n : q->seq.n;
        /* cigar.rs:155-199 parse_maf_seq_to_trim over cigar_cat_ext groups */
        uint64_t hi = 0, hd = 0, ti = 0, td = 0;
        int head = 1;
        for (size_t i = 0; i < n;) {
            char k = cigar_cat_ext(t->seq.p[i], q->seq.p[i]);
            size_t j = i;
            while (j < n && cigar_cat_ext(t->seq.p[j], q->seq.p[j]) == k) j++;
            uint64_t cnt = j - i;
            if (k == 'I') { if (head) hi += cnt; ti = cnt; }
            else if (k == 'D') { if (head) hd += cnt; td = cnt; }
            else { ti = 0; td = 0; head = 0; }
            i = j;
        }
        chain_header(out, t->name, t->size, t->start, t->start + t->align_size, q->name, q->size, q->neg, maf_query_start(&rec),
                     maf_query_end(&rec), hi, hd, ti, td, id);
        /* cigar.rs:435-456 parse_maf_seq_to_chain over cigar_cat groups (=/X/both-gap -> M) */
        chainline d = {0, 0, 0};
        for (size_t i = 0; i < n;) {
            char a = t->seq.p[i], b = q->seq.p[i];
            char k = a == b ? 'M' : a == '-' ? 'I' : b == '-' ? 'D' : 'M';
            size_t j = i;
            for (; j < n; j++) {
                char a2 = t->seq.p[j], b2 = q->seq.p[j];
                char k2 = a2 == b2 ? 'M' : a2 == '-' ? 'I' : b2 == '-' ? 'D' : 'M';
                if (k2 != k) break;
            }
            chain_unit(out, &d, k, j - i);
            i = j;
        }
        buf_putc(out, '\n'); buf_putu(out, d.size);
        buf_puts(out, "\n\n");
        id++;
    }
    free(rec.s);
    return rc < 0 ? -rc : ORC_OK;
}

int orc_paf2chain(const char *text, size_t len, orc_buf *out, char *err, size_t errcap) {
    size_t pos = 0;
    pafrow row = {0};
    orc_buf conv = {0};
    int rc = ORC_OK;
    uint64_t id = 0;
    while (rc == ORC_OK && paf_next(text, len, &pos, &row)) {
        uint64_t u[12] = {0};
        int neg = 0, bad = row.n < 12;
        for (int k = 0; k < 12 && !bad; k++) {
            if (k == 0 || k == 5) continue;
            if (k == 4) { if (sv_eq_c(row.f[4], "+")) neg = 0; else if (sv_eq_c(row.f[4], "-")) neg = 1; else bad = 1; continue; }
            if (!csv_parse_u64(row.f[k], &u[k])) bad = 1;
        }
        if (bad) { set_err(err, errcap, "CSV deserialize error by: field parse"); rc = ORC_E_CSV; break; }
        sv cg = {0, 0};
        const int have = paf_get_cigar(&row, &conv, &cg);
        if (have < 0) { set_err(err, errcap, CS_PANIC_MSG); rc = ORC_E_PANIC; break; }
        if (!have) { set_err(err, errcap, "CIGAR start tag not found"); rc = ORC_E_CIGAR_TAG_NOT_FOUND; break; }
        if (cg.n == 0) { set_err(err, errcap, "panic: byte index 10 is out of bounds of ``"); rc = ORC_E_PANIC; break; }
        /* cigar.rs:202-245 parse_cigar_to_trim: the closure's Result is discarded (a bad token just ends the trim) */
        uint64_t hi = 0, hd = 0, ti = 0, td = 0;
        int head = 1;
        for (size_t i = 0; i < cg.n;) {
            size_t d0 = i;
            while (i < cg.n && cg.p[i] >= '0' && cg.p[i] <= '9') i++;
            size_t o0 = i;
            while (i < cg.n && !(cg.p[i] >= '0' && cg.p[i] <= '9')) i++;
            sv digits = {cg.p + d0, o0 - d0};
            uint64_t cnt;
            if (i - o0 != 1 || !parse_u64(digits, &cnt)) break;
            char op = cg.p[o0];
            if (op == 'M' || op == '=' || op == 'X') { ti = 0; td = 0; head = 0; }
            else if (op == 'I') { if (head) hi += cnt; ti = cnt; }
            else if (op == 'D') { if (head) hd += cnt; td = cnt; }
            else break;
        }
        const size_t rec_start = out->n;
        chain_header(out, row.f[5], u[6], u[7], u[8], row.f[0], u[1], neg, u[2], u[3], hi, hd, ti, td, id);
        /* cigar.rs:252-295 parse_cigar_to_chain: errors are returned */
        chainline d = {0, 0, 0};
        for (size_t i = 0; i < cg.n && rc == ORC_OK;) {
            size_t d0 = i;
            while (i < cg.n && cg.p[i] >= '0' && cg.p[i] <= '9') i++;
            size_t o0 = i;
            while (i < cg.n && !(cg.p[i] >= '0' && cg.p[i] <= '9')) i++;
            sv digits = {cg.p + d0, o0 - d0}, op = {cg.p + o0, i - o0};
            uint64_t cnt;
            if (op.n != 1) { set_err(err, errcap, "CIGAR OP `%.*s` invalid", (int)op.n, op.p); rc = ORC_E_CIGAR_OP_INVALID; break; }
            if (!parse_u64(digits, &cnt)) { set_err(err, errcap, "Parse `%.*s` Into Integer Error", (int)digits.n, digits.p); rc = ORC_E_PARSE_INT; break; }
            char c = op.p[0];
            if (c == 'M' || c == '=' || c == 'X') chain_unit(out, &d, 'M', cnt);
            else if (c == 'I' || c == 'D') chain_unit(out, &d, c, cnt);
            else { set_err(err, errcap, "CIGAR OP `%c` invalid", c); rc = ORC_E_CIGAR_OP_INVALID; }
        }
        if (rc) { (void)rec_start; break; }
        buf_putc(out, '\n'); buf_putu(out, d.size);
        buf_puts(out, "\n\n");
        id++;
    }
    free(row.f); pafrow_free_scratch(&row);
    orc_buf_free(&conv);
    return rc;
}

/* ------------------------------------------------------------------ pafcov */
/* src/tools/pafcov.rs:13-60 + src/parser/cigar.rs:710-741 update_cov_vec.  Per target a coverage vector of the FIRST
 * record's target_length; M and = add 1 over [pos, pos+len) clipped to the vector, I and S do not move, every other op
 * (X, D, N, ...) advances pos (Q4).  Output: one BED line per base.  The reference folds per rayon worker and iterates
 * a HashMap; here targets come in first-occurrence order. */
int orc_pafcov(const char *text, size_t len, orc_buf *out, char *err, size_t errcap) {
    size_t pos = 0;
    pafrow row = {0};
    orc_buf conv = {0};
    struct tgt { sv name; uint64_t L; uint32_t *cov; } *T = NULL;
    size_t nt = 0, capt = 0;
    int rc = ORC_OK;
    while (rc == ORC_OK && paf_next(text, len, &pos, &row)) {
        uint64_t u[12] = {0};
        int bad = row.n < 12;
        for (int k = 0; k < 12 && !bad; k++) {
            if (k == 0 || k == 5) continue;
            if (k == 4) { if (!sv_eq_c(row.f[4], "+") && !sv_eq_c(row.f[4], "-")) bad = 1; continue; }
            if (!csv_parse_u64(row.f[k], &u[k])) bad = 1;
        }
        if (bad) { set_err(err, errcap, "CSV deserialize error by: field parse"); rc = ORC_E_CSV; break; }
        size_t ti = nt;
        for (size_t k = 0; k < nt; k++) if (sv_eq(T[k].name, row.f[5])) { ti = k; break; }
        if (ti == nt) {
            if (nt == capt) { capt = capt ? capt * 2 : 16; T = (struct tgt *)realloc(T, capt * sizeof *T); }
            T[nt].name = row.f[5]; T[nt].L = u[6]; T[nt].cov = (uint32_t *)calloc(u[6] ? u[6] : 1, sizeof(uint32_t));
            nt++;
        }
        sv cg = {0, 0};
        const int have = paf_get_cigar(&row, &conv, &cg);
        if (have < 0) { set_err(err, errcap, CS_PANIC_MSG); rc = ORC_E_PANIC; break; }
        if (!have) { set_err(err, errcap, "CIGAR start tag not found"); rc = ORC_E_CIGAR_TAG_NOT_FOUND; break; }
        if (cg.n == 0) { set_err(err, errcap, "panic: byte index 10 is out of bounds of ``"); rc = ORC_E_PANIC; break; }
        uint64_t p = u[7];
        for (size_t i = 0; i < cg.n;) {
            size_t d0 = i;
            while (i < cg.n && cg.p[i] >= '0' && cg.p[i] <= '9') i++;
            size_t o0 = i;
            while (i < cg.n && !(cg.p[i] >= '0' && cg.p[i] <= '9')) i++;
            sv digits = {cg.p + d0, o0 - d0}, op = {cg.p + o0, i - o0};
            uint64_t cnt;
            if (op.n != 1) { set_err(err, errcap, "CIGAR OP `%.*s` invalid", (int)op.n, op.p); rc = ORC_E_CIGAR_OP_INVALID; break; }
            if (!parse_u64(digits, &cnt)) { set_err(err, errcap, "Parse `%.*s` Into Integer Error", (int)digits.n, digits.p); rc = ORC_E_PARSE_INT; break; }
            char c = op.p[0];
            if (c == 'M' || c == '=') {
                uint64_t lo = p < T[ti].L ? p : T[ti].L, hi = (cnt > T[ti].L || p + cnt > T[ti].L) ? T[ti].L : p + cnt;
                for (uint64_t x = lo; x < hi; x++) T[ti].cov[x]++;
                p += cnt;
            } else if (c == 'I' || c == 'S') {
            } else p += cnt;
        }
    }
    if (rc == ORC_OK)
        for (size_t k = 0; k < nt; k++)
            for (uint64_t x = 0; x < T[k].L; x++) {
                buf_put(out, T[k].name.p, T[k].name.n); buf_putc(out, '\t'); buf_putu(out, x); buf_putc(out, '\t'); buf_putu(out, x + 1);
                buf_putc(out, '\t'); buf_putu(out, T[k].cov[x]); buf_putc(out, '\n');
            }
    for (size_t k = 0; k < nt; k++) free(T[k].cov);
    free(T); free(row.f); pafrow_free_scratch(&row);
    orc_buf_free(&conv);
    return rc;
}

/* ======================================================================================
 * chain reader, chain2paf, chain2maf
 *   src/parser/chain.rs:27-37,60-75 (ChainReader::records: the whole file is read into one String and
 *   chain_parser is applied to what is left of it), :206-383 (parse_header, line_not_chain,
 *   parse_line_to_cdl, parse_chain_data_line, chain_parser), :430-452 (ChainRecord::convert2paf),
 *   src/parser/cigar.rs:554-626 (parse_chain_to_cigar), src/converter.rs:268-357 (chain2maf),
 *   :360-388 (parse_chain_to_insert), :391-416 (chain2paf).
 * nom 7.1.3 combinators restated from their documented behaviour (parity unpinned beyond the round trips):
 *   tag("chain")            the rest must START with "chain" (no skipping before the first record)
 *   not_line_ending         header text up to "\n" / "\r\n" (a lone '\r' is an error)
 *   line_ending             "\n" or "\r\n" must follow the header (a header that ends the file is an error)
 *   fold_many1(terminated(is_not("chain\n"), line_ending))
 *                           a data line = >= 1 characters none of which is one of c,h,a,i,n,'\n', followed by a line
 *                           ending; the fold stops at the first line that is not one (a blank line, a line that
 *                           holds one of those letters, a last line WITHOUT its newline -- which is therefore
 *                           silently dropped) and needs at least one
 *   take_while(|x| x != 'c') everything up to the next 'c' is skipped (blank lines, comments without a 'c')
 * A data line is split_whitespace(): size, then query_diff (default 0), then target_diff (default 0) -- in that
 * order (chain.rs:333-350; the UCSC format has dt before dq, the reference reads AND writes them as it does).
 * ====================================================================================== */
typedef struct {
    sv tname, qname;
    uint64_t tsize, tstart, tend, qsize, qstart, qend, id;
    int tneg, qneg;
    chainline *lines;
    size_t n, cap;
} chainrec;

/* Rust's str::parse::<f64>() accepts: [+-] (digits [. digits?] | . digits) [(e|E) [+-] digits] | [+-] inf | infinity | nan
 * (case-insensitive words); only validity matters here (the score is never printed) */
static int rust_f64_ok(sv s) {
    size_t i = 0;
    if (s.n == 0) return 0;
    if (s.p[0] == '+' || s.p[0] == '-') i = 1;
    if (i == s.n) return 0;
    {
        char w[9]; size_t m = s.n - i;
        if (m <= 8) {
            for (size_t k = 0; k < m; k++) { char c = s.p[i + k]; w[k] = (c >= 'A' && c <= 'Z') ? (char)(c + 32) : c; }
            w[m] = 0;
            if (!strcmp(w, "inf") || !strcmp(w, "infinity") || !strcmp(w, "nan")) return 1;
        }
    }
    size_t nd = 0;
    while (i < s.n && s.p[i] >= '0' && s.p[i] <= '9') { i++; nd++; }
    if (i < s.n && s.p[i] == '.') { i++; while (i < s.n && s.p[i] >= '0' && s.p[i] <= '9') { i++; nd++; } }
    if (nd == 0) return 0;
    if (i < s.n && (s.p[i] == 'e' || s.p[i] == 'E')) {
        i++;
        if (i < s.n && (s.p[i] == '+' || s.p[i] == '-')) i++;
        size_t ne = 0;
        while (i < s.n && s.p[i] >= '0' && s.p[i] <= '9') { i++; ne++; }
        if (ne == 0) return 0;
    }
    return i == s.n;
}

static int ws_next(sv line, size_t *pos, sv *tok) { /* str::split_whitespace */
    size_t i = *pos;
    while (i < line.n && is_ws(line.p[i])) i++;
    if (i >= line.n) { *pos = i; return 0; }
    size_t b = i;
    while (i < line.n && !is_ws(line.p[i])) i++;
    tok->p = line.p + b; tok->n = i - b;
    *pos = i;
    return 1;
}

static int nom_err(char *err, size_t errcap, const char *code, const char *at, size_t left) {
    /* errors.rs:88-97: e.input[..10] -- itself a panic when fewer than 10 bytes are left */
    if (left < 10) { set_err(err, errcap, "panic: byte index 10 is out of bounds (nom error %s on a short rest)", code); return ORC_E_PANIC; }
    set_err(err, errcap, "Format error %s at: %.10s Parse Error by rust::nom, please check", code, at);
    return ORC_E_NOM;
}

/* one chain_parser() call on text[*pos..]; 1 = record, 0 = input exhausted, < 0 = -error */
static int chain_next(const char *t, size_t len, size_t *pos, chainrec *rec, char *err, size_t errcap) {
    size_t i = *pos;
    if (i >= len) return 0;
    if (len - i < 5 || memcmp(t + i, "chain", 5) != 0) return -nom_err(err, errcap, "Tag", t + i, len - i);
    i += 5;
    size_t h0 = i;
    while (i < len && t[i] != '\n' && t[i] != '\r') i++;
    if (i < len && t[i] == '\r' && !(i + 1 < len && t[i + 1] == '\n')) return -nom_err(err, errcap, "Tag", t + i, len - i);
    sv hl = {t + h0, i - h0};
    /* parse_header (chain.rs:206-315) */
    static const char *fname[12] = {"score", "target_name", "target_size", "target_strand", "target_start", "target_end",
                                    "query_name", "query_size", "query_strand", "query_start", "query_end", "chain_id"};
    sv f[12];
    size_t hp = 0;
    uint64_t u[12] = {0};
    for (int k = 0; k < 12; k++) {
        if (!ws_next(hl, &hp, &f[k])) {
            set_err(err, errcap, "Parse Chain Error By: Chain Line Field `%s` Missing", fname[k]);
            return -ORC_E_CHAIN_FIELD_MISSING;
        }
        if (k == 0) {
            if (!rust_f64_ok(f[k])) { set_err(err, errcap, "Parse `%.*s` Into Float Error", (int)f[k].n, f[k].p); return -ORC_E_PARSE_FLOAT; }
        } else if (k == 3 || k == 8) {
            if (!(sv_eq_c(f[k], "+") || sv_eq_c(f[k], "-"))) { set_err(err, errcap, "Parse Strand `%.*s` Error", (int)f[k].n, f[k].p); return -ORC_E_PARSE_STRAND; }
        } else if (k != 1 && k != 6) {
            if (!parse_u64(f[k], &u[k])) { set_err(err, errcap, "Parse `%.*s` Into Integer Error", (int)f[k].n, f[k].p); return -ORC_E_PARSE_INT; }
        }
    }
    rec->tname = f[1]; rec->tsize = u[2]; rec->tneg = sv_eq_c(f[3], "-"); rec->tstart = u[4]; rec->tend = u[5];
    rec->qname = f[6]; rec->qsize = u[7]; rec->qneg = sv_eq_c(f[8], "-"); rec->qstart = u[9]; rec->qend = u[10]; rec->id = u[11];
    /* line_ending */
    if (i >= len) return -nom_err(err, errcap, "CrLf", t + i, 0);
    i += t[i] == '\r' ? 2 : 1;
    /* fold_many1(line_not_chain): the closure's `?` parks the first parse error in the accumulator, later lines
     * are still consumed (chain.rs:352-369) */
    rec->n = 0;
    int line_rc = 0;
    size_t nl_lines = 0;
    for (;;) {
        size_t j = i;
        while (j < len && t[j] != 'c' && t[j] != 'h' && t[j] != 'a' && t[j] != 'i' && t[j] != 'n' && t[j] != '\n') j++;
        if (j == i) break;                       /* is_not needs one character */
        if (j >= len || t[j] != '\n') break;     /* terminated(.., line_ending): "\r\n" ends in '\n' too ('\r' is matched by is_not) */
        sv dl = {t + i, j - i};
        i = j + 1;
        nl_lines++;
        if (line_rc) continue;
        size_t dp = 0;
        sv tok;
        chainline c = {0, 0, 0};
        if (!ws_next(dl, &dp, &tok)) { set_err(err, errcap, "Parse Chain Error By: Chain Line Field `size` Missing"); line_rc = ORC_E_CHAIN_FIELD_MISSING; continue; }
        if (!parse_u64(tok, &c.size)) { set_err(err, errcap, "Parse `%.*s` Into Integer Error", (int)tok.n, tok.p); line_rc = ORC_E_PARSE_INT; continue; }
        if (ws_next(dl, &dp, &tok) && !parse_u64(tok, &c.qdiff)) { set_err(err, errcap, "Parse `%.*s` Into Integer Error", (int)tok.n, tok.p); line_rc = ORC_E_PARSE_INT; continue; }
        if (dp < dl.n && ws_next(dl, &dp, &tok) && !parse_u64(tok, &c.tdiff)) { set_err(err, errcap, "Parse `%.*s` Into Integer Error", (int)tok.n, tok.p); line_rc = ORC_E_PARSE_INT; continue; }
        if (rec->n == rec->cap) { rec->cap = rec->cap ? rec->cap * 2 : 64; rec->lines = (chainline *)realloc(rec->lines, rec->cap * sizeof(chainline)); }
        rec->lines[rec->n++] = c;
    }
    if (nl_lines == 0) return -nom_err(err, errcap, "Many1", t + i, len - i);
    if (line_rc) return -line_rc;
    while (i < len && t[i] != 'c') i++;  /* take_while(|x| x != 'c') */
    *pos = i;
    return 1;
}

/* chain.rs:430-451 convert2paf + cigar.rs:554-626 parse_chain_to_cigar + the csv writer of converter.rs:396-414:
 * 12 columns + the one cg:Z: tag; block_length = match + mismatch + del + inv_del with del = query_diff */
int orc_chain2paf(const char *text, size_t len, orc_buf *out, char *err, size_t errcap) {
    size_t pos = 0;
    chainrec rec = {0};
    int rc;
    while ((rc = chain_next(text, len, &pos, &rec, err, errcap)) > 0) {
        uint64_t match = 0, del = 0;
        for (size_t k = 0; k < rec.n; k++) { match += rec.lines[k].size; del += rec.lines[k].qdiff; }
        csv_field(out, rec.qname, 1);
        csv_u64(out, rec.qsize, 0); csv_u64(out, rec.qstart, 0); csv_u64(out, rec.qend, 0);
        buf_putc(out, '\t'); buf_putc(out, rec.qneg ? '-' : '+');
        csv_field(out, rec.tname, 0);
        csv_u64(out, rec.tsize, 0); csv_u64(out, rec.tstart, 0); csv_u64(out, rec.tend, 0);
        csv_u64(out, match, 0); csv_u64(out, match + del, 0); csv_u64(out, 255, 0);
        buf_puts(out, "\tcg:Z:");
        for (size_t k = 0; k < rec.n; k++) {
            buf_putu(out, rec.lines[k].size); buf_putc(out, 'M');
            if (rec.lines[k].tdiff) { buf_putu(out, rec.lines[k].tdiff); buf_putc(out, 'I'); }
            if (rec.lines[k].qdiff) { buf_putu(out, rec.lines[k].qdiff); buf_putc(out, 'D'); }
        }
        buf_putc(out, '\n');
    }
    free(rec.lines);
    return rc < 0 ? -rc : ORC_OK;
}

/* converter.rs:268-388 chain2maf: score 255, target fetched [t_start, t_end-1], query [q_start, q_end-1] with the header's
 * own query coordinates, reverse-complemented for '-'; s-line start of a '-' query = q_size - q_end; the target strand is
 * printed as the header has it; parse_chain_to_insert walks one shared offset: += size, target_diff x '-' into the TARGET
 * row, query_diff x '-' into the QUERY row (insert_str panics past the end of the row) */
int orc_chain2maf(const char *text, size_t len, const char *t_fa, size_t t_fa_len, const char *t_fai, size_t t_fai_len,
                  const char *q_fa, size_t q_fa_len, const char *q_fai, size_t q_fai_len, const char *t_path,
                  const char *q_path, orc_buf *out, char *err, size_t errcap) {
    faidx tf, qf;
    if (fai_parse(t_fai, t_fai_len, t_fa, t_fa_len, &tf) || fai_parse(q_fai, q_fai_len, q_fa, q_fa_len, &qf)) {
        set_err(err, errcap, "HTS library error by bad .fai");
        return ORC_E_CSV;
    }
    buf_puts(out, "#maf version=1.6 convert_from=chain t_seq_path=");
    buf_puts(out, t_path);
    buf_puts(out, " q_seq_path=");
    buf_puts(out, q_path);
    buf_putc(out, '\n');
    size_t pos = 0;
    chainrec rec = {0};
    orc_buf ts = {0}, qs = {0};
    int rc, e = ORC_OK;
    while ((rc = chain_next(text, len, &pos, &rec, err, errcap)) > 0) {
        ts.n = 0; qs.n = 0;
        /* `target_end() - 1` / `query_end() - 1` are u64 subtractions: 0 - 1 panics in debug, wraps in release; the wrapped
         * value then clamps to the sequence end in faidx -- the release build's behaviour is kept */
        if (fai_fetch(&tf, rec.tname, rec.tstart, rec.tend - 1, &ts) || fai_fetch(&qf, rec.qname, rec.qstart, rec.qend - 1, &qs)) {
            set_err(err, errcap, "HTS library error by sequence not found");
            e = ORC_E_NOM; break;
        }
        if (rec.qneg) {
            char c = revcomp_inplace(qs.p, qs.n);
            if (c) { set_err(err, errcap, "Invalid Base: `%c`", c); e = ORC_E_VCF_BASE; break; }
        }
        uint64_t off = 0;
        for (size_t k = 0; k < rec.n && !e; k++) {
            off += rec.lines[k].size;
            if (rec.lines[k].tdiff) {
                if (buf_insert_gaps(&ts, off, rec.lines[k].tdiff)) { set_err(err, errcap, "panic: insert_str index out of range"); e = ORC_E_PANIC; break; }
                off += rec.lines[k].tdiff;
            }
            if (rec.lines[k].qdiff) {
                if (buf_insert_gaps(&qs, off, rec.lines[k].qdiff)) { set_err(err, errcap, "panic: insert_str index out of range"); e = ORC_E_PANIC; break; }
                off += rec.lines[k].qdiff;
            }
        }
        if (e) break;
        uint64_t q_start_maf = rec.qneg ? rec.qsize - rec.qend : rec.qstart;
        buf_puts(out, "a score=255\n");
        buf_puts(out, "s\t"); buf_put(out, rec.tname.p, rec.tname.n); buf_putc(out, '\t'); buf_putu(out, rec.tstart); buf_putc(out, '\t');
        buf_putu(out, rec.tend - rec.tstart); buf_putc(out, '\t'); buf_putc(out, rec.tneg ? '-' : '+'); buf_putc(out, '\t');
        buf_putu(out, rec.tsize); buf_putc(out, '\t'); buf_put(out, ts.p, ts.n); buf_putc(out, '\n');
        buf_puts(out, "s\t"); buf_put(out, rec.qname.p, rec.qname.n); buf_putc(out, '\t'); buf_putu(out, q_start_maf); buf_putc(out, '\t');
        buf_putu(out, rec.qend - rec.qstart); buf_putc(out, '\t'); buf_putc(out, rec.qneg ? '-' : '+'); buf_putc(out, '\t');
        buf_putu(out, rec.qsize); buf_putc(out, '\t'); buf_put(out, qs.p, qs.n); buf_putc(out, '\n');
        buf_putc(out, '\n');
    }
    free(rec.lines); free(tf.e); free(qf.e);
    orc_buf_free(&ts); orc_buf_free(&qs);
    if (e) return e;
    return rc < 0 ? -rc : ORC_OK;
}
