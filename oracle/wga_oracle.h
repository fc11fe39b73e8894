/*
 * wga_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A literal, sequential C restatement of the reference's (wjwei-handsome/wgatools
 * v1.1.0, commit de13560d) hot path: the per-record CIGAR / gapped-sequence walk
 * behind `maf2paf`, `stat` and `call`.  Every function cites the reference
 * file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (wgatools_b200/) never links, imports or executes anything under oracle/.
 *
 * Pinning status: the reference cannot be compiled here (no cargo/rustc, crates
 * not vendored), so oracle/_ref does not exist.  The oracle is pinned against
 * the golden vectors the reference ships (test/test.maf, test/testdotplot.paf,
 * README.md:322-343 VCF excerpt) -- see tests/test_oracle_golden.py -- and by
 * code review.  Behaviour that lives in un-vendored crates (noodles-vcf 0.43
 * record text, csv 1.2.2 quoting, ryu 1.0.14 f32 text, natord 1.0.9 ordering)
 * is restated from their published algorithms and is flagged "parity unpinned"
 * wherever no golden vector covers it.
 */
#ifndef WGA_ORACLE_H
#define WGA_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes; messages reproduce `impl Display for WGAError` (src/errors.rs:8-86) */
enum {
    ORC_OK = 0,
    ORC_E_PARSE_INT = 1,          /* "Parse `{0}` Into Integer Error"    errors.rs:50-52 */
    ORC_E_PARSE_STRAND = 2,       /* "Parse `{0}` Error"... see errors.rs:48-49 */
    ORC_E_MAF_FIELD_MISSING = 3,  /* "Parse MAF error by: S-line Filed `{0}` Missing" */
    ORC_E_MAF_SURPLUS_FIELD = 4,  /* "Parse MAF error by: Surplus Filed > 7" */
    ORC_E_CIGAR_OP_INVALID = 5,   /* "CIGAR OP `{0}` invalid"            errors.rs:58-59 */
    ORC_E_CIGAR_TAG_NOT_FOUND = 6,/* "CIGAR start tag not found"         errors.rs:56-57 */
    ORC_E_QUERY_NAME_NOT_FOUND = 7,/* "Query name:{0} not found in MAF"  errors.rs:80-81 */
    ORC_E_PANIC = 8,              /* the reference would panic (index out of bounds, slice) */
    ORC_E_CSV = 9,                /* "CSV deserialize error by: ..."     errors.rs:20-21 */
    ORC_E_VCF_BASE = 10,          /* noodles REF/ALT base parse failure (parity unpinned) */
    ORC_E_NOM = 11,               /* nom tag("cg:Z:") / tag("chain") / line_ending failure */
    ORC_E_CHAIN_FIELD_MISSING = 12, /* "Parse Chain Error By: Chain Line Field `{0}` Missing"  errors.rs:47-48,109-112 */
    ORC_E_PARSE_FLOAT = 13,       /* "Parse `{0}` Into Float Error"      errors.rs:54-56 */
    ORC_E_OTHER = 14              /* anyhow::Error passed through (e.g. noodles rejecting a contig name) */
};

typedef struct {
    char  *p;
    size_t n, cap;
} orc_buf;

void orc_buf_free(orc_buf *b);

/* the 11 counters of `struct Cigar` (src/parser/cigar.rs:16-29) in declaration order:
 * match, mismatch, ins_event, ins_count, del_event, del_count,
 * inv_ins_event, inv_ins_count, inv_del_event, inv_del_count, inv_event */
typedef struct { uint64_t c[11]; } orc_counts;

/* src/parser/cigar.rs:344-432 parse_maf_seq_to_cigar(rec, with_h=false).
 * n = min(len_t, len_q) (zip truncation, cigar.rs:359-361).  cigar may be NULL. */
void orc_parse_maf_seq_to_cigar(const char *t, const char *q, size_t n, int neg_strand,
                                orc_counts *out, orc_buf *cigar);

/* src/parser/cigar.rs:629-707 parse_paf_to_cigar on the text AFTER "cg:Z:".
 * Returns ORC_OK or an error code; err (cap >= 256) receives the Display text. */
int orc_parse_paf_cigar(const char *cg, size_t n, int neg_strand, orc_counts *out,
                        char *err, size_t errcap);

/* src/converter.rs:29-54 maf2paf, -t 1 order (file order). */
int orc_maf2paf(const char *text, size_t len, const char *query_name, orc_buf *out,
                char *err, size_t errcap);

/* src/tools/stat.rs:61-126 stat_maf / stat_paf.  Row order: stable natord sort on
 * ref_name over first-occurrence order (the reference's HashMap order is
 * nondeterministic within equal ref_name -- SURVEY Q8). */
int orc_stat_maf(const char *text, size_t len, int each, const char *query_name,
                 orc_buf *out, char *err, size_t errcap);
/* filter.rs:41-63,92-160: filter -f paf [-b -q | -a] */
int orc_paf_filter(const char *text, size_t len, uint64_t min_block, uint64_t min_query, int have_align, uint64_t min_align,
                   orc_buf *out, char *err, size_t errcap);

/* cigar.rs:807-985 + dotplot.rs:283-289: the base-level plot data as csv (header row when there is at least one segment) */
int orc_maf_plotdata(const char *text, size_t len, const char *query_name, uint64_t cutoff, orc_buf *out, char *err, size_t errcap);
int orc_paf_plotdata(const char *text, size_t len, uint64_t cutoff, orc_buf *out, char *err, size_t errcap);

/* paf.rs:159-218 cs_to_cigar alone (test hook) */
int orc_cs_to_cigar(const char *cs, size_t n, orc_buf *out, char *err, size_t errcap);

int orc_stat_paf(const char *text, size_t len, int each, orc_buf *out,
                 char *err, size_t errcap);

/* src/tools/caller.rs:42-157 call_var_maf (no <maf>.index => no ##contig lines). */
int orc_call_maf(const char *text, size_t len, int snp, int inv, uint64_t svlen_cutoff,
                 const char *sample, const char *query_name, uint64_t chunk_size,
                 orc_buf *out, char *err, size_t errcap);

/* the same with --query-regex and a <maf>.index: query_pred(name, n, ud) != 0 stands for
 * `query_regex.is_match_at(&x.name, 0)` (maf.rs:267-271; the regex crate is outside this restatement), has_index != 0
 * adds the ##contig lines of caller.rs:339-357 for the given isref entries. */
typedef int (*orc_name_pred)(const char *name, size_t n, void *ud);
int orc_call_maf_ex(const char *text, size_t len, int snp, int inv, uint64_t svlen_cutoff,
                    const char *sample, const char *query_name, orc_name_pred query_pred, void *pred_ud,
                    uint64_t chunk_size, int has_index, const char *const *contig_names, const uint64_t *contig_sizes,
                    size_t n_contigs, orc_buf *out, char *err, size_t errcap);

/* the reference's rayon structure with `-t threads` (serial reader under par_bridge's lock + worker pool);
 * output identical to the sequential functions (file order).  Used as the CPU baseline of bench.py. */
int orc_maf2paf_mt(const char *text, size_t len, const char *query_name, int threads, orc_buf *out,
                   char *err, size_t errcap);
int orc_stat_maf_mt(const char *text, size_t len, int each, const char *query_name, int threads, orc_buf *out,
                    char *err, size_t errcap);

/* src/converter.rs:176-265 paf2maf; FASTA access = htslib faidx_fetch_seq64 (inclusive end, clipped to the sequence).
 * The header (and every block before a failing record) would already have been written by the reference; on error the
 * buffer holds that partial output. */
int orc_paf2maf(const char *paf, size_t paf_len, const char *t_fa, size_t t_fa_len, const char *t_fai, size_t t_fai_len,
                const char *q_fa, size_t q_fa_len, const char *q_fai, size_t q_fai_len, const char *t_path,
                const char *q_path, orc_buf *out, char *err, size_t errcap);

/* src/tools/caller.rs:268-302,610-822 call_var_paf (`call -f paf`): PAF + 2 FASTA -> VCF.  Any error leaves out empty
 * (the reference buffers every row and writes the header last). */
int orc_call_paf(const char *paf, size_t paf_len, const char *t_fa, size_t t_fa_len, const char *t_fai, size_t t_fai_len,
                 const char *q_fa, size_t q_fa_len, const char *q_fai, size_t q_fai_len, int snp, uint64_t svlen_cutoff,
                 const char *sample, orc_buf *out, char *err, size_t errcap);

/* src/tools/pseudomaf.rs:18-237 + src/parser/cigar.rs:744-804 (`pafpseudo`).  fa/fai NULL = no FASTA ('N' target row,
 * 1/0/- query rows); target NULL = all targets.  out = files concatenated in first-occurrence target order,
 * index = "{target}\t{bytes}\n" per file. */
int orc_paf_pseudo(const char *paf, size_t paf_len, const char *fa, size_t fa_len, const char *fai, size_t fai_len,
                   const char *target, orc_buf *out, orc_buf *index, char *err, size_t errcap);

/* src/tools/validate.rs:44-147 (`validate [-f fix.paf]`): report text + (fix != 0) the re-serialised records */
int orc_paf_validate(const char *text, size_t len, int fix, orc_buf *report, orc_buf *fixed, char *err, size_t errcap);

/* src/converter.rs:57-92 maf2chain, :148-173 paf2chain */
int orc_maf2chain(const char *text, size_t len, const char *query_name, orc_buf *out, char *err, size_t errcap);
int orc_paf2chain(const char *text, size_t len, orc_buf *out, char *err, size_t errcap);

/* src/tools/pafcov.rs:13-60 + cigar.rs:710-741 (`pafcov`): per-base BED coverage of every target */
int orc_pafcov(const char *text, size_t len, orc_buf *out, char *err, size_t errcap);
/* chain.rs:206-383 reader + converter.rs:391-416 chain2paf / :268-388 chain2maf */
int orc_chain2paf(const char *text, size_t len, orc_buf *out, char *err, size_t errcap);
int orc_chain2maf(const char *text, size_t len, const char *t_fa, size_t t_fa_len, const char *t_fai, size_t t_fai_len,
                  const char *q_fa, size_t q_fa_len, const char *q_fai, size_t q_fai_len, const char *t_path,
                  const char *q_path, orc_buf *out, char *err, size_t errcap);

/* helpers exposed for unit tests */
int  orc_natord_compare(const char *a, const char *b);       /* natord 1.0.9 compare */
void orc_fmt_f32(float v, char *dst /* >= 32 bytes */);      /* ryu 1.0.14 Buffer::format(f32) */
/* per-record counters of a MAF/PAF text, file order; returns number of records or <0 */
int64_t orc_maf_record_counts(const char *text, size_t len, orc_counts *out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
