/*
 * wga_synth.c -- seeded synthetic MAF / PAF generators for bench.py and the parity tests
 * (bench/test infrastructure; the shapes follow SURVEY.md section 8(d), configs 2-4).
 *
 * Two passes so that threads can write disjoint byte ranges of one (pinned) buffer:
 *   plan : per block, replay the event stream (cheap: geometric skipping, no bases) to get the
 *          non-gap counts -> coordinates -> exact byte size of every block -> offsets;
 *   fill : replay the same per-block RNG stream and write the text.
 * Every block owns an independent splitmix64/xoshiro stream derived from (seed, block index), so the
 * output is identical for any thread count.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t seed;
    uint64_t n_blocks;
    double len_median, len_sigma; /* lognormal block columns */
    uint64_t len_min, len_max;
    double p_mis;                 /* per-column mismatch probability */
    double p_gap;                 /* per-column small-indel open probability (50/50 I/D) */
    double gap_mean;              /* geometric mean length of small indels */
    double p_sv;                  /* per-column SV open probability (length log-uniform) */
    uint64_t sv_min, sv_max;
    double p_neg;                 /* fraction of '-' strand blocks */
    double p_bothgap;             /* per-column probability of a both-gap ('W') column */
    double p_lower;               /* per-event probability of lower-casing the query base */
    uint32_t n_chr;
    uint64_t chr_size;
    int32_t threads;
} synth_maf_params;

typedef struct {
    uint64_t ncols, t_al, q_al, t_start, q_start, off, size;
    uint32_t chr;
    uint8_t neg;
} blk;

typedef struct {
    synth_maf_params p;
    blk *b;
    uint64_t total;
    uint64_t total_cols;
} plan_t;

/* ---- rng */
typedef struct { uint64_t s[4]; } rng_t;
static inline uint64_t splitmix(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline void rng_seed(rng_t *r, uint64_t seed, uint64_t stream) {
    uint64_t x = seed ^ (stream * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull);
    for (int i = 0; i < 4; i++) r->s[i] = splitmix(&x);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(rng_t *r) { /* xoshiro256** */
    uint64_t *s = r->s, res = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return res;
}
static inline double rng_u01(rng_t *r) { return ((rng_next(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static inline uint64_t rng_geom(rng_t *r, double p) { /* failures before first success, >= 0 */
    if (p <= 0) return UINT64_MAX / 4;
    if (p >= 1) return 0;
    double g = floor(log(rng_u01(r)) / log1p(-p));
    return g > 1e18 ? (uint64_t)1e18 : (uint64_t)g;
}
static inline double rng_normal(rng_t *r) {
    double u1 = rng_u01(r), u2 = rng_u01(r);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

/* ---- event stream of one block: calls ev(kind, len) in column order; kinds:
 * 0 match run, 1 mismatch column, 2 ins (target gaps), 3 del (query gaps), 4 both-gap column */
typedef void (*ev_fn)(void *ctx, int kind, uint64_t len, rng_t *base_rng, int lower);

static void block_events(const synth_maf_params *p, uint64_t bi, uint64_t ncols, ev_fn ev, void *ctx) {
    rng_t er, br;
    rng_seed(&er, p->seed, bi * 2 + 1);
    rng_seed(&br, p->seed ^ 0xA5A5A5A5ull, bi * 2 + 2);
    double p_ev = p->p_mis + p->p_gap + p->p_sv + p->p_bothgap;
    uint64_t c = 0;
    while (c < ncols) {
        uint64_t skip = rng_geom(&er, p_ev);
        if (skip > ncols - c) skip = ncols - c;
        if (skip) ev(ctx, 0, skip, &br, 0);
        c += skip;
        if (c >= ncols) break;
        double u = rng_u01(&er) * p_ev;
        int lower = p->p_lower > 0 && rng_u01(&er) < p->p_lower;
        if (u < p->p_mis) {
            ev(ctx, 1, 1, &br, lower);
            c += 1;
        } else if (u < p->p_mis + p->p_bothgap) {
            ev(ctx, 4, 1, &br, 0);
            c += 1;
        } else {
            uint64_t len;
            if (u < p->p_mis + p->p_bothgap + p->p_gap) {
                len = 1 + rng_geom(&er, 1.0 / (p->gap_mean > 1 ? p->gap_mean : 1.0));
            } else { /* log-uniform SV length */
                double lo = log((double)p->sv_min), hi = log((double)p->sv_max + 1.0);
                len = (uint64_t)exp(lo + (hi - lo) * rng_u01(&er));
                if (len < p->sv_min) len = p->sv_min;
                if (len > p->sv_max) len = p->sv_max;
            }
            if (len > ncols - c) len = ncols - c;
            int ins = rng_next(&er) & 1;
            ev(ctx, ins ? 2 : 3, len, &br, lower);
            c += len;
        }
    }
}

typedef struct { uint64_t t_al, q_al; } cnt_ctx;
static void ev_count(void *vc, int kind, uint64_t len, rng_t *br, int lower) {
    (void)br; (void)lower;
    cnt_ctx *c = (cnt_ctx *)vc;
    if (kind == 0 || kind == 1) { c->t_al += len; c->q_al += len; }
    else if (kind == 2) c->q_al += len;
    else if (kind == 3) c->t_al += len;
}

typedef struct { char *t, *q; uint64_t bits; int nbits; } fill_ctx;
static const char BASES[4] = {'A', 'C', 'G', 'T'};
static inline unsigned base2(fill_ctx *f, rng_t *br) {
    if (f->nbits == 0) { f->bits = rng_next(br); f->nbits = 32; }
    unsigned v = (unsigned)(f->bits & 3);
    f->bits >>= 2;
    f->nbits--;
    return v;
}
static void ev_fill(void *vc, int kind, uint64_t len, rng_t *br, int lower) {
    fill_ctx *f = (fill_ctx *)vc;
    char *t = f->t, *q = f->q;
    switch (kind) {
    case 0: {
        uint64_t i = 0;
        /* 32 bases per 64-bit draw */
        while (i + 32 <= len) {
            uint64_t w = rng_next(br);
            for (int k = 0; k < 32; k++) { char b = BASES[(w >> (2 * k)) & 3]; t[i + k] = b; q[i + k] = b; }
            i += 32;
        }
        for (; i < len; i++) { char b = BASES[base2(f, br)]; t[i] = b; q[i] = b; }
        break;
    }
    case 1: {
        unsigned a = base2(f, br), d = 1 + (unsigned)(rng_next(br) % 3);
        t[0] = BASES[a];
        q[0] = BASES[(a + d) & 3];
        if (lower) q[0] = (char)(q[0] + 32);
        break;
    }
    case 2:
        for (uint64_t i = 0; i < len; i++) { t[i] = '-'; q[i] = BASES[base2(f, br)]; }
        break;
    case 3:
        for (uint64_t i = 0; i < len; i++) { q[i] = '-'; t[i] = BASES[base2(f, br)]; }
        break;
    case 4: t[0] = '-'; q[0] = '-'; break;
    }
    f->t += len;
    f->q += len;
}

static int ndig(uint64_t v) { int n = 1; while (v >= 10) { v /= 10; n++; } return n; }
static char *put_u64(char *o, uint64_t v) {
    char tmp[24]; int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *o++ = tmp[--n];
    return o;
}

typedef struct { plan_t *pl; uint64_t lo, hi; char *dst; int pass; } job_t;

static void *worker(void *vj) {
    job_t *j = (job_t *)vj;
    plan_t *pl = j->pl;
    const synth_maf_params *p = &pl->p;
    for (uint64_t bi = j->lo; bi < j->hi; bi++) {
        blk *b = &pl->b[bi];
        if (j->pass == 1) {
            cnt_ctx c = {0, 0};
            block_events(p, bi, b->ncols, ev_count, &c);
            b->t_al = c.t_al;
            b->q_al = c.q_al;
        } else {
            char *o = j->dst + b->off;
            memcpy(o, "a score=0\n", 10); o += 10;
            o += sprintf(o, "s\tref.chr%u\t", b->chr + 1);
            o = put_u64(o, b->t_start); *o++ = '\t';
            o = put_u64(o, b->t_al); *o++ = '\t'; *o++ = '+'; *o++ = '\t';
            o = put_u64(o, p->chr_size); *o++ = '\t';
            char *tseq = o;
            o += b->ncols; *o++ = '\n';
            o += sprintf(o, "s\tqry.chr%u\t", b->chr + 1);
            o = put_u64(o, b->q_start); *o++ = '\t';
            o = put_u64(o, b->q_al); *o++ = '\t'; *o++ = b->neg ? '-' : '+'; *o++ = '\t';
            o = put_u64(o, p->chr_size); *o++ = '\t';
            char *qseq = o;
            o += b->ncols; *o++ = '\n'; *o++ = '\n';
            fill_ctx f = {tseq, qseq, 0, 0};
            block_events(p, bi, b->ncols, ev_fill, &f);
        }
    }
    return NULL;
}

static void run_jobs(plan_t *pl, char *dst, int pass) {
    int nt = pl->p.threads > 0 ? pl->p.threads : 1;
    if ((uint64_t)nt > pl->p.n_blocks) nt = (int)(pl->p.n_blocks ? pl->p.n_blocks : 1);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nt);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)nt);
    /* balance by columns */
    uint64_t per = pl->total_cols / (uint64_t)nt + 1, acc = 0, lo = 0;
    int k = 0;
    for (uint64_t bi = 0; bi < pl->p.n_blocks && k < nt; bi++) {
        acc += pl->b[bi].ncols;
        if (acc >= per * (uint64_t)(k + 1) || bi + 1 == pl->p.n_blocks) {
            jobs[k] = (job_t){pl, lo, bi + 1, dst, pass};
            lo = bi + 1;
            k++;
        }
    }
    if (k > 0) jobs[k - 1].hi = pl->p.n_blocks;
    for (int i = 0; i < k; i++) pthread_create(&th[i], NULL, worker, &jobs[i]);
    for (int i = 0; i < k; i++) pthread_join(th[i], NULL);
    free(th);
    free(jobs);
}

#define HEADER "##maf version=1 scoring=synthetic\n"

void *wga_synth_maf_plan(const synth_maf_params *p, uint64_t *total_bytes, uint64_t *total_cols) {
    plan_t *pl = (plan_t *)calloc(1, sizeof(plan_t));
    pl->p = *p;
    pl->b = (blk *)calloc(p->n_blocks ? p->n_blocks : 1, sizeof(blk));
    rng_t r;
    rng_seed(&r, p->seed, 0);
    double mu = log(p->len_median);
    for (uint64_t i = 0; i < p->n_blocks; i++) {
        double l = exp(mu + p->len_sigma * rng_normal(&r));
        uint64_t n = l < 1 ? 1 : (l > 1e15 ? (uint64_t)1e15 : (uint64_t)l);
        if (n < p->len_min) n = p->len_min;
        if (n > p->len_max) n = p->len_max;
        pl->b[i].ncols = n;
        pl->b[i].neg = rng_u01(&r) < p->p_neg;
        pl->b[i].chr = (uint32_t)((i * (uint64_t)p->n_chr) / (p->n_blocks ? p->n_blocks : 1));
        pl->total_cols += n;
    }
    run_jobs(pl, NULL, 1);
    /* coordinates: increasing within a chromosome, small random spacer between blocks */
    uint64_t off = strlen(HEADER), tpos = 0, qpos = 0;
    uint32_t cur = UINT32_MAX;
    for (uint64_t i = 0; i < p->n_blocks; i++) {
        blk *b = &pl->b[i];
        if (b->chr != cur) { cur = b->chr; tpos = rng_next(&r) % 10000; qpos = rng_next(&r) % 10000; }
        b->t_start = tpos;
        uint64_t qfwd = qpos;
        b->q_start = b->neg ? p->chr_size - qfwd - b->q_al : qfwd;
        tpos += b->t_al + rng_next(&r) % 500;
        qpos += b->q_al + rng_next(&r) % 500;
        char nm[32];
        int nl = snprintf(nm, sizeof nm, "%u", b->chr + 1);
        b->size = 10 /* a-line */
                  + (uint64_t)(10 + nl + ndig(b->t_start) + 1 + ndig(b->t_al) + 3 + ndig(p->chr_size) + 1) + b->ncols + 1
                  + (uint64_t)(10 + nl + ndig(b->q_start) + 1 + ndig(b->q_al) + 3 + ndig(p->chr_size) + 1) + b->ncols + 1
                  + 1;
        b->off = off;
        off += b->size;
    }
    pl->total = off;
    if (total_bytes) *total_bytes = off;
    if (total_cols) *total_cols = pl->total_cols;
    return pl;
}

void wga_synth_maf_fill(void *vpl, char *dst) {
    plan_t *pl = (plan_t *)vpl;
    memcpy(dst, HEADER, strlen(HEADER));
    run_jobs(pl, dst, 2);
}

void wga_synth_plan_free(void *vpl) {
    plan_t *pl = (plan_t *)vpl;
    if (!pl) return;
    free(pl->b);
    free(pl);
}

/* ------------------------------------------------------------------ PAF (config 4 shape) */
typedef struct {
    uint64_t seed;
    uint64_t n_records;
    double ops_median, ops_sigma; /* lognormal ops per record */
    uint64_t ops_min, ops_max;
    double eq_mean;               /* geometric mean '=' run length */
    double p_neg;
    uint32_t n_genomes, n_chr;
    uint64_t chr_size;
    int32_t m_only;               /* 1: emit M-only CIGARs (=/X merged) */
    int32_t threads;
} synth_paf_params;

typedef struct { synth_paf_params p; uint64_t *off; uint64_t total; } pplan_t;

/* one record -> text at o (or only its size when o == NULL); returns bytes */
static uint64_t paf_record(const synth_paf_params *p, uint64_t ri, char *o) {
    rng_t r;
    rng_seed(&r, p->seed, ri + 17);
    double l = exp(log(p->ops_median) + p->ops_sigma * rng_normal(&r));
    uint64_t nops = l < 1 ? 1 : (uint64_t)l;
    if (nops < p->ops_min) nops = p->ops_min;
    if (nops > p->ops_max) nops = p->ops_max;
    uint32_t gq = (uint32_t)(rng_next(&r) % p->n_genomes), gt = (uint32_t)(rng_next(&r) % p->n_genomes);
    uint32_t cq = (uint32_t)(rng_next(&r) % p->n_chr), ct = (uint32_t)(rng_next(&r) % p->n_chr);
    int neg = rng_u01(&r) < p->p_neg;
    /* first pass over ops to get spans; ops regenerated from a forked stream */
    rng_t ro = r, ro2 = r;
    uint64_t tspan = 0, qspan = 0, nmatch = 0, blen = 0;
    for (int pass = 0; pass < 2; pass++) {
        rng_t *rr = pass ? &ro2 : &ro;
        char *c = NULL;
        if (pass) {
            if (!o) break;
            uint64_t qs = rng_next(&r) % (p->chr_size > qspan ? p->chr_size - qspan : 1), ts = rng_next(&r) % (p->chr_size > tspan ? p->chr_size - tspan : 1);
            c = o;
            c += sprintf(c, "g%u#chr%u\t", gq, cq + 1); c = put_u64(c, p->chr_size); *c++ = '\t';
            c = put_u64(c, qs); *c++ = '\t'; c = put_u64(c, qs + qspan); *c++ = '\t';
            *c++ = neg ? '-' : '+'; *c++ = '\t';
            c += sprintf(c, "g%u#chr%u\t", gt, ct + 1); c = put_u64(c, p->chr_size); *c++ = '\t';
            c = put_u64(c, ts); *c++ = '\t'; c = put_u64(c, ts + tspan); *c++ = '\t';
            c = put_u64(c, nmatch); *c++ = '\t'; c = put_u64(c, blen); *c++ = '\t';
            memcpy(c, "60\tNM:i:", 8); c += 8; c = put_u64(c, blen - nmatch);
            memcpy(c, "\tcg:Z:", 6); c += 6;
        }
        uint64_t bytes = 0, pend_m = 0;
        for (uint64_t k = 0; k < nops; k++) {
            char op; uint64_t len;
            if ((k & 1) == 0) { op = '='; len = 1 + rng_geom(rr, 1.0 / p->eq_mean); }
            else {
                uint64_t u = rng_next(rr) % 10;
                if (u < 6) { op = 'X'; len = 1 + rng_next(rr) % 3; }
                else { op = (u & 1) ? 'I' : 'D'; len = 1 + rng_geom(rr, 0.25); }
            }
            if (!pass) {
                if (op == '=' ) { tspan += len; qspan += len; nmatch += len; }
                else if (op == 'X') { tspan += len; qspan += len; }
                else if (op == 'I') qspan += len;
                else tspan += len;
                blen += len;
            }
            if (p->m_only && (op == '=' || op == 'X')) { pend_m += len; if (k + 1 < nops) continue; }
            if (p->m_only && pend_m) {
                bytes += (uint64_t)ndig(pend_m) + 1;
                if (c) { c = put_u64(c, pend_m); *c++ = 'M'; }
                pend_m = 0;
                if (op == '=' || op == 'X') continue;
            }
            bytes += (uint64_t)ndig(len) + 1;
            if (c) { c = put_u64(c, len); *c++ = op; }
        }
        if (pass) { *c++ = '\n'; return (uint64_t)(c - o); }
        if (!o) {
            /* size only: replicate the header digits */
            uint64_t qs = rng_next(&r) % (p->chr_size > qspan ? p->chr_size - qspan : 1), ts = rng_next(&r) % (p->chr_size > tspan ? p->chr_size - tspan : 1);
            char tmp[64];
            uint64_t h = (uint64_t)snprintf(tmp, sizeof tmp, "g%u#chr%u\t", gq, cq + 1) + (uint64_t)ndig(p->chr_size) + 1 +
                         (uint64_t)ndig(qs) + 1 + (uint64_t)ndig(qs + qspan) + 1 + 2 +
                         (uint64_t)snprintf(tmp, sizeof tmp, "g%u#chr%u\t", gt, ct + 1) + (uint64_t)ndig(p->chr_size) + 1 +
                         (uint64_t)ndig(ts) + 1 + (uint64_t)ndig(ts + tspan) + 1 + (uint64_t)ndig(nmatch) + 1 +
                         (uint64_t)ndig(blen) + 1 + 8 + (uint64_t)ndig(blen - nmatch) + 6;
            return h + bytes + 1;
        }
    }
    return 0;
}

typedef struct { pplan_t *pl; uint64_t lo, hi; char *dst; } pjob_t;
static void *pworker(void *vj) {
    pjob_t *j = (pjob_t *)vj;
    for (uint64_t i = j->lo; i < j->hi; i++) {
        if (j->dst) paf_record(&j->pl->p, i, j->dst + j->pl->off[i]);
        else j->pl->off[i] = paf_record(&j->pl->p, i, NULL);
    }
    return NULL;
}
static void prun(pplan_t *pl, char *dst) {
    int nt = pl->p.threads > 0 ? pl->p.threads : 1;
    pthread_t th[256];
    pjob_t jobs[256];
    if (nt > 256) nt = 256;
    uint64_t n = pl->p.n_records, per = (n + (uint64_t)nt - 1) / (uint64_t)nt;
    int k = 0;
    for (uint64_t lo = 0; lo < n; lo += per, k++) {
        jobs[k] = (pjob_t){pl, lo, lo + per < n ? lo + per : n, dst};
        pthread_create(&th[k], NULL, pworker, &jobs[k]);
    }
    for (int i = 0; i < k; i++) pthread_join(th[i], NULL);
}

void *wga_synth_paf_plan(const synth_paf_params *p, uint64_t *total_bytes) {
    pplan_t *pl = (pplan_t *)calloc(1, sizeof(pplan_t));
    pl->p = *p;
    pl->off = (uint64_t *)calloc(p->n_records + 1, sizeof(uint64_t));
    prun(pl, NULL);
    uint64_t off = 0;
    for (uint64_t i = 0; i < p->n_records; i++) { uint64_t s = pl->off[i]; pl->off[i] = off; off += s; }
    pl->off[p->n_records] = off;
    pl->total = off;
    if (total_bytes) *total_bytes = off;
    return pl;
}
void wga_synth_paf_fill(void *vpl, char *dst) { prun((pplan_t *)vpl, dst); }
void wga_synth_paf_plan_free(void *vpl) {
    pplan_t *pl = (pplan_t *)vpl;
    if (!pl) return;
    free(pl->off);
    free(pl);
}
