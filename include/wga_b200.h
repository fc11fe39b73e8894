/*
 * wga_b200.h -- C ABI of the B200-native hot path of wgatools (maf2paf / stat / call).
 *
 * The reference (wjwei-handsome/wgatools v1.1.0) has no FFI of its own; the path sits behind two
 * Rust surfaces:
 *   - trait AlignRecord            src/parser/common.rs:142-176 (get_stat, convert2paf, get_cigar_string)
 *   - the L3 tool/converter fns    src/converter.rs:29-33 (maf2paf), src/tools/stat.rs:61-66,87-91
 *                                  (stat_maf, stat_paf), src/tools/caller.rs:42-54 (call_var_maf)
 * A per-record FFI call would serialise the GPU, so the boundary is batched at the level the L3
 * functions work at: a slab of MAF/PAF text in, the command's output bytes (or stat partials) out.
 * INTEGRATION.md shows the `extern "C"` block a maintainer adds to the Rust crate.
 *
 * Conventions
 *   - every entry returns int: 0 = ok, otherwise a WGA_E_* code; never aborts or throws.
 *     Codes 1..11 mirror WGAError variants (src/errors.rs:8-86); wga_last_error() returns the
 *     reference's Display text for them.  First-error-wins with the LOWEST file position, i.e.
 *     what the reference's sequential `-t 1` run reports.
 *   - inputs are caller-owned and read-only for the duration of the call;
 *     outputs are library-owned and valid until the next call on the same ctx.
 *   - one ctx per (host thread, GPU); a ctx is not thread-safe; different ctxs may run concurrently.
 *   - there is NO CPU fallback: without a CUDA device wga_ctx_create fails with WGA_E_CUDA.
 *   - device-resident inputs (mem == WGA_MEM_DEVICE) must be readable up to the next multiple of 16
 *     past `len` (true for any cudaMalloc'ed / torch buffer); a pointer that is not 16-byte aligned
 *     costs one device-to-device copy of the slab.
 */
#ifndef WGA_B200_H
#define WGA_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum wga_status {
    WGA_OK = 0,
    /* mirrors of WGAError (src/errors.rs) */
    WGA_E_PARSE_INT = 1,            /* ParseIntError        errors.rs:50-52 */
    WGA_E_PARSE_STRAND = 2,         /* ParseStrand          errors.rs:48-49 */
    WGA_E_MAF_FIELD_MISSING = 3,    /* ParseMaf(FiledMissing) errors.rs:99-105 */
    WGA_E_MAF_SURPLUS_FIELD = 4,    /* ParseMaf(SurplusField) */
    WGA_E_CIGAR_OP_INVALID = 5,     /* CigarOpInvalid       errors.rs:58-59 */
    WGA_E_CIGAR_TAG_NOT_FOUND = 6,  /* CigarTagNotFound     errors.rs:56-57 */
    WGA_E_QUERY_NAME_NOT_FOUND = 7, /* QueryNameNotFound    errors.rs:80-81 */
    WGA_E_REF_PANIC = 8,            /* input on which the reference panics (e.g. 1-s-line block in
                                       stat/maf2paf: slines[1], maf.rs:426-453) */
    WGA_E_CSV = 9,                  /* CsvDeserialize       errors.rs:20-21 */
    WGA_E_VCF_BASE = 10,            /* noodles REF/ALT base outside ACGTNacgtn (caller.rs:379-382) */
    WGA_E_NOM = 11,                 /* NomErr               errors.rs:45-46,88-97 */
    WGA_E_CHAIN_FIELD_MISSING = 12, /* ParseChain(FiledMissing) errors.rs:47-48,109-112 */
    WGA_E_PARSE_FLOAT = 13,         /* ParseFloatError      errors.rs:54-56 */
    /* library errors */
    WGA_E_CUDA = 100,
    WGA_E_INVALID_ARG = 101,
    WGA_E_UNSUPPORTED = 102,        /* input the reference accepts and this library reports instead of parsing: a csv-quoted PAF
                                     * field whose content holds a quote, a tab or a line break (plain quoted fields are fine);
                                     * `chain` in the middle of a line */
    WGA_E_NOMEM = 103,
    WGA_E_INTERNAL = 104,
    WGA_E_NCCL = 105
};

enum { WGA_MEM_HOST = 0, WGA_MEM_DEVICE = 1 };
/* the slab starts at byte 0 of the file: its first line is the header MAFReader::new swallows
 * unconditionally (src/parser/maf.rs:25-36).  Slabs of one file must be cut before a line that
 * does not start with 's' (wga_maf_safe_cut) so that s-line groups are never split. */
enum {
    WGA_F_FIRST_SLAB = 1u,
    /* with mem == WGA_MEM_DEVICE: still hand the output back in (pinned) host memory */
    WGA_F_OUT_HOST = 2u,
    /* with mem == WGA_MEM_DEVICE, on wga_maf2paf / wga_maf_stat / wga_maf_call: the caller asserts that (text, len)
     * is the slab the previous MAF command on this ctx was given and that its bytes have not changed; the on-device
     * tokenisation of that call (delimiter index, line table, s-line fields) is then reused -- the reference parses
     * the file once per command -- and only the command-specific record building and walk are run (not even those
     * when the previous command left them behind, see WGA_F_SHARE_WALK).  Ignored when no matching index is held, and
     * by the PAF entry points. */
    WGA_F_REUSE_INDEX = 4u,
    /* on wga_maf2paf, when further MAF commands on the SAME slab follow (with WGA_F_REUSE_INDEX): run the dense column
     * walk once in the finer classification `call` needs (both-gap columns kept apart, cigar.rs:314-328) and keep what
     * it produced -- the 11 counters of every record and the run list.  A following wga_maf_stat then needs no walk
     * at all (RecStat::from(Cigar), common.rs:116-140, is a function of those counters) and a following wga_maf_call
     * starts at its chunk grid.  The reference computes the same walk once per command. */
    WGA_F_SHARE_WALK = 8u
};

typedef struct wga_ctx wga_ctx;

typedef struct {
    const void *data; /* same memory kind as the input text (pinned host or device); host with WGA_F_OUT_HOST */
    uint64_t len;
} wga_span;

/* ---- context ------------------------------------------------------------------------------- */
/* stream: a cudaStream_t to enqueue all work on (e.g. torch's current stream), or NULL for a
 * private stream.  Every entry point is synchronous at return. */
int wga_ctx_create(int device, void *stream, wga_ctx **out);
void wga_ctx_destroy(wga_ctx *ctx);
/* Display text of the last error (WGAError wording for codes 1..11); *where = line index (MAF) or
 * record index (PAF) it was raised at. */
const char *wga_last_error(const wga_ctx *ctx, uint64_t *where);
/* number of kernels this ctx launched since creation (hand-written kernels / CUB utility scans) */
void wga_ctx_launch_counts(const wga_ctx *ctx, uint64_t *own_kernels, uint64_t *cub_calls);
/* NCCL collectives / point-to-point groups this ctx enqueued since creation */
uint64_t wga_ctx_nccl_calls(const wga_ctx *ctx);
/* elapsed device milliseconds of the dense column-walk kernel(s) of the last command and the
 * algorithmic bytes they covered (sum over launches) -- measured with CUDA events on the ctx stream */
void wga_ctx_last_walk(const wga_ctx *ctx, float *ms, uint64_t *alg_bytes, uint32_t *launches);

/* pinned host memory for callers that want full PCIe rate on the host-buffer entry points */
void *wga_host_alloc(uint64_t bytes);
void wga_host_free(void *p);
/* offset of the last line in text[0..len) (other than the first) that does not start with 's':
 * cutting a file there never splits an s-line group (maf.rs:395-413).  0 if there is none. */
uint64_t wga_maf_safe_cut(const char *text, uint64_t len);

/* ---- resident text: one upload, several commands ------------------------------------------------
 * The reference runs maf2paf / stat / call as separate passes over the same file (each re-reads it,
 * src/parser/maf.rs:15-44).  Here the file is uploaded ONCE, asynchronously and slab by slab on a
 * private copy stream; wga_text_slab() makes the ctx stream wait for slab k only, so the commands
 * run on slab k while slab k+1 is still crossing PCIe.  Slabs are cut with wga_maf_safe_cut and lie
 * contiguously in HBM, so after the last slab the whole text can also be used in one call.
 * host_text must stay valid until wga_text_close (pinned memory gives full PCIe rate).  One resident
 * text per ctx at a time; slab_bytes == 0 picks 256 MiB. */
typedef struct wga_text wga_text;
int wga_text_open(wga_ctx *ctx, const void *host_text, uint64_t len, uint64_t slab_bytes, wga_text **out);
uint32_t wga_text_slabs(const wga_text *t);
/* device pointer / length / WGA_F_FIRST_SLAB flag of slab k, to be passed with WGA_MEM_DEVICE */
int wga_text_slab(wga_text *t, uint32_t k, const void **dev, uint64_t *len, uint32_t *flags);
int wga_text_whole(wga_text *t, const void **dev, uint64_t *len);
void wga_text_close(wga_text *t);

/* ---- maf2paf: src/converter.rs:29-54 + MAFRecord::convert2paf src/parser/maf.rs:484-520 ------ */
/* out = PAF rows (14 tab-separated fields, '\n'-terminated), one per MAF record, file order.
 * query_name = Some(name) selects the query s-line by name (maf.rs:277-285), NULL = s-line 1. */
int wga_maf2paf(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags,
                const char *query_name, wga_span *out);

/* ---- stat: src/tools/stat.rs:61-126 ------------------------------------------------------------ */
/* Host-side accumulator of PairStat partials; slabs (and ranks) are added in file order. */
typedef struct wga_stat_acc wga_stat_acc;
wga_stat_acc *wga_stat_acc_create(int each);
void wga_stat_acc_destroy(wga_stat_acc *acc);
/* stat_maf (stat.rs:61-84) over one slab: per-record counters on the GPU, pair reduction on the GPU,
 * partial table merged into acc. */
int wga_maf_stat(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags,
                 const char *query_name, wga_stat_acc *acc);
/* stat_paf (stat.rs:87-105) over one slab of PAF text (whole lines). */
int wga_paf_stat(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags,
                 wga_stat_acc *acc);
/* write_style_result (stat.rs:107-126): TSV with header; text valid until the next write/destroy */
int wga_stat_acc_write(wga_stat_acc *acc, wga_span *out);
/* flat export/import for the cross-rank reduction (8(e)): keys = n_pairs records of
 * "ref_name\0query_name\0" + {ref_size, query_size}; vals = n_pairs x 13 u64
 * (11 counters of struct Cigar order + ref_start_min + query_start_min);
 * inv = per '-'-strand-record (pair index, (aligned+query_aligned) numerator) in file order. */
int wga_stat_acc_export(wga_stat_acc *acc, wga_span *blob);
int wga_stat_acc_import(wga_stat_acc *acc, const void *blob, uint64_t len);
uint64_t wga_stat_acc_npairs(const wga_stat_acc *acc);

/* ---- cross-rank reduction of stat partials: SURVEY.md 8(e), src/tools/stat.rs:167-223 --------------------------------
 * One rank per GPU; ranks hold contiguous byte ranges of the input in rank order and feed their own accumulator with
 * wga_maf_stat / wga_paf_stat.  wga_stat_allreduce then turns every rank's (merged-mode) accumulator into the
 * accumulator of the whole file: the ranks agree on the pair list (first-seen order over ranks 0..N-1 = file order),
 * ONE all-reduce(sum, u64) of the [n_pairs x 12] table (11 counters + the integer sum of the '-' records' inv_size
 * numerators), one all-reduce(min, u64) of the [n_pairs x 2] starts.  inv_size is an f32 sum in file order; for a pair
 * whose numerators total less than 2^24 every partial sum is exact, the order cannot matter and the integer column
 * settles it; only pairs beyond that bound have their f32 partial carried rank to rank (4 B per such pair and hop;
 * no hop at all when there is none), so that the result is bit-identical to a single process reading the whole file.
 * wga_stat_acc_write on any rank prints the global TSV. */
typedef struct wga_comm wga_comm;
/* 128 bytes (ncclUniqueId): produced on one rank, distributed by the host's own means (MPI, torch.distributed, ...) */
int wga_nccl_unique_id(void *id128);
/* ncclCommInitRank on ctx's device; collectives run on the ctx stream */
int wga_comm_init_nccl(wga_ctx *ctx, const void *id128, int rank, int world, wga_comm **out);
/* adopt a communicator the host already owns (an ncclComm_t; not destroyed by wga_comm_destroy) */
int wga_comm_from_nccl(wga_ctx *ctx, void *nccl_comm, wga_comm **out);
/* hosts with another kind of communicator (and the CPU tests, over gloo) provide the three primitives on HOST buffers;
 * each returns 0 on success.  op: 0 = sum, 1 = min.  sendrecv is one ring hop: send to dst and receive from src, either
 * of which may be -1 (none). */
typedef struct {
    void *user;
    int rank, world;
    int (*allgather)(void *user, const void *send, void *recv, uint64_t bytes_per_rank);
    int (*allreduce_u64)(void *user, uint64_t *buf, uint64_t n, int op);
    int (*sendrecv)(void *user, const void *send, uint64_t send_bytes, int dst, void *recv, uint64_t recv_bytes, int src);
} wga_transport;
int wga_comm_from_transport(const wga_transport *t, wga_comm **out);
void wga_comm_destroy(wga_comm *c);
int wga_comm_rank(const wga_comm *c);
int wga_comm_world(const wga_comm *c);
const char *wga_comm_last_error(const wga_comm *c);
/* collective: every rank calls it with its own accumulator (merged mode); on return all accumulators are identical */
int wga_stat_allreduce(wga_comm *c, wga_stat_acc *acc);

/* ---- call: src/tools/caller.rs:42-157, 159-265, 388-608 ------------------------------------- */
typedef struct {
    int32_t snp;           /* -s/--snp   (cli.rs:146-147) */
    int32_t inv;           /* -i/--inv   (cli.rs:149-150) */
    uint64_t svlen_cutoff; /* -l/--svlen, default 50 (cli.rs:152-153) */
    uint64_t chunk_size;   /* -c/--chunk-size, default 1000000 (cli.rs:171-173); 0 is rejected */
    const char *query_name; /* --query-name or NULL */
    /* --query-regex (cli.rs:166-170; caller.rs:91-104, maf.rs:267-271 `position(|x| regex.is_match_at(&x.name, 0))`): the
     * regex engine stays with the host.  The host asks for the slab's distinct s-line names (wga_maf_sline_names), keeps
     * those its Regex matches and passes them here, '\n'-separated (NULL = no regex; "" = nothing matched, every block is
     * skipped).  Per block the query is then the first s-line whose name is in the list.  Ignored when query_name is set
     * ("Use regex only if query_name not provided", caller.rs:76). */
    const char *query_match;
} wga_call_opts;
/* VCF header text for `sample` (caller.rs:304-337) without ##contig lines (no <maf>.index) */
int wga_vcf_header(wga_ctx *ctx, const char *sample, wga_span *out);
/* the same with the ##contig lines of a <maf>.index (caller.rs:339-357 add_header_contig): `names` / `sizes` = the
 * index's entries with isref == true, in any order; they are natord-sorted here and written as
 * "##contig=<ID=name,length=size>" behind the FORMAT lines (noodles-vcf 0.43 header order).  A name that is not a valid
 * VCF contig ID fails like the reference's `name.parse()?` (WGA_E_INVALID_ARG).  Host-side text: ctx may be NULL. */
int wga_vcf_header_contigs(wga_ctx *ctx, const char *sample, const char *const *names, const uint64_t *sizes, uint64_t n,
                           wga_span *out);
/* distinct s-line names of one slab, first-occurrence order, each followed by '\n' (host memory, valid until the next
 * call on this thread): what the host evaluates --query-regex on.  Blocks are the ones `call` builds (maf.rs:371-421). */
int wga_maf_sline_names(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, wga_span *out);
/* VCF body rows of one slab, file order */
int wga_maf_call(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags,
                 const wga_call_opts *opts, wga_span *out);

/* ---- paf2maf: src/converter.rs:176-265 (second wave, config 5) ------------------------------------- */
/* A FASTA resident in HBM (rust-htslib faidx::Reader stand-in): `fa` = file bytes, `fai` = its samtools .fai text
 * (the reference refuses to run without the .fai, utils.rs:480-492).  Host buffers.  Well-formed FASTA assumed
 * (uniform line width per record, which is what a .fai describes). */
typedef struct wga_fasta wga_fasta;
int wga_fasta_load(wga_ctx *ctx, const void *fa, uint64_t fa_len, const void *fai, uint64_t fai_len, wga_fasta **out);
void wga_fasta_free(wga_fasta *fa);
/* "#maf version=1.6 convert_from=paf t_seq_path=.. q_seq_path=..\n" (converter.rs:190-195) */
int wga_paf2maf_header(wga_ctx *ctx, const char *t_fa_path, const char *q_fa_path, wga_span *out);
/* MAF blocks of one slab of PAF lines, file order: fetch [t_start,t_end-1] / [q_start,q_end-1], reverse-complement
 * '-' queries (utils.rs:83-101), expand the CIGAR into gapped rows (cigar.rs:492-551), maf.rs:566-581 text. */
int wga_paf2maf(wga_ctx *ctx, const void *paf, uint64_t len, int mem, uint32_t flags, const wga_fasta *target,
                const wga_fasta *query, wga_span *out);

/* ---- call -f paf: src/tools/caller.rs:268-302 call_var_paf + :610-822 call_within_var_paf -------------------- */
/* VCF body rows of one slab of PAF lines, file order (the header comes from wga_vcf_header).  Sequences are
 * fetched as [start, end] INCLUSIVE (paf.rs:221-237), '-' queries are not reverse-complemented, every '-' record
 * gets an <INV> row, and a failing CIGAR token / VCF row silently ends its record (the reference discards the
 * fold closure's Result, caller.rs:673-815).  The reference buffers all rows and writes the header last, so on
 * any error there is no output at all. */
int wga_paf_call(wga_ctx *ctx, const void *paf, uint64_t len, int mem, uint32_t flags, const wga_fasta *target,
                 const wga_fasta *query, int snp, uint64_t svlen_cutoff, wga_span *out);

/* ---- pafpseudo: src/tools/pseudomaf.rs:18-237 + src/parser/cigar.rs:744-804 gen_pesudo_maf_by_cigar ------------ */
/* One pseudo-MAF per target: `out` = the files' bytes concatenated, `index` (always host memory) = one line
 * "{target}\t{bytes}\n" per file, i.e. what the reference writes to {out_dir}/{target}.maf.  fa = the FASTA holding
 * targets and queries, or NULL ('N' target row, 1/0/- query rows); target = --target filter or NULL.
 * Targets, and queries inside a target, come in first-occurrence order (the reference iterates HashMaps).  Records of
 * one (target, query) group with equal target_start are placed as the reference's sorted insert places them
 * (pseudomaf.rs:91-94: slice::binary_search_by + Vec::insert, core::slice of rustc 1.52-1.81).  On error nothing is
 * returned. */
int wga_paf_pseudo(wga_ctx *ctx, const void *paf, uint64_t len, int mem, uint32_t flags, const wga_fasta *fa,
                   const char *target, wga_span *out, wga_span *index);

/* ---- chain: src/converter.rs:57-92 maf2chain, :148-173 paf2chain --------------------------------- */
/* One chain per record, file order: "chain 255 tName tSize + tStart tEnd qName qSize qStrand qStart qEnd id" with the
 * head / tail indels trimmed off the coordinates (cigar.rs:155-245; the '-' strand keeps the reference's arithmetic,
 * chain.rs:133-138), "size qdiff tdiff" lines, the last block's size and a blank line.  id = id_base + index of the
 * record in this slab (the reference enumerates the records of the file); *n_records (may be NULL) = records in the
 * slab, i.e. what to add to id_base for the next one. */
int wga_maf2chain(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, const char *query_name,
                  uint64_t id_base, wga_span *out, uint64_t *n_records);
int wga_paf2chain(wga_ctx *ctx, const void *paf, uint64_t len, int mem, uint32_t flags, uint64_t id_base, wga_span *out,
                  uint64_t *n_records);

/* ---- chain2paf / chain2maf: src/converter.rs:391-416, :268-388 with the reader of src/parser/chain.rs:27-75,206-383 ----
 * `text` = whole chain records (it must start with a `chain` header line).  The reader's nom grammar is reproduced
 * line-locally on the device (wgatools_b200/csrc/chain_in.cuh): a data line is a line without any of the letters
 * c,h,a,i,n that ends in a newline; lines between a record's data lines and the next header are skipped; a last data
 * line without its newline is dropped, as the reference drops it.
 * chain2paf: one PAF row per chain, 12 columns + "cg:Z:<size>M[<target_diff>I][<query_diff>D]..." (chain.rs:430-451,
 * cigar.rs:554-626; data lines are read as "size query_diff target_diff", chain.rs:333-350).
 * chain2maf: "a score=255" blocks; target [t_start, t_end-1] and query [q_start, q_end-1] from the FASTAs, '-' queries
 * reverse-complemented, gaps inserted by the data lines (converter.rs:360-388); header text from wga_chain2maf_header. */
int wga_chain2paf(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, wga_span *out);
int wga_chain2maf_header(wga_ctx *ctx, const char *t_fa_path, const char *q_fa_path, wga_span *out);
int wga_chain2maf(wga_ctx *ctx, const void *text, uint64_t len, int mem, uint32_t flags, const wga_fasta *target,
                  const wga_fasta *query, wga_span *out);

/* ---- pafcov: src/tools/pafcov.rs:13-60 + src/parser/cigar.rs:710-741 ------------------------------- */
/* BED lines "{target}\t{pos}\t{pos+1}\t{count}\n" for every base of every target (vector length = target_length of
 * the target's first record; M and = cover, I and S do not move, every other op advances).  Targets come in
 * first-occurrence order.  One slab = the whole PAF (coverage is a reduction over all records). */
int wga_paf_cov(wga_ctx *ctx, const void *paf, uint64_t len, int mem, uint32_t flags, wga_span *out);

/* ---- filter -f paf: src/tools/filter.rs:41-63 filter_paf, :92-106 filter_alignrec, :108-160 filter_paf_align_pair
 *  ---------------------------------------------------------------------------------------------------- */
/* min_align_size == NULL: a record is dropped when target_end - target_start < min_block_size or query_length <
 * min_query_size.  min_align_size != NULL (`-a`, all-to-all PAF): records are grouped by (query_name, target_name), the
 * group's summed target_end - target_start must reach *min_align_size, and the two other cut-offs are ignored
 * (src/utils.rs:559-563).  out = the kept records re-serialised by the csv writer, in file order (the reference's
 * par_bridge order is unspecified).  One slab = the whole PAF (the pair sums are a reduction over all records). */
int wga_paf_filter(wga_ctx *ctx, const void *paf, uint64_t len, int mem, uint32_t flags, uint64_t min_block_size,
                   uint64_t min_query_size, const uint64_t *min_align_size, wga_span *out);

/* ---- dotplot -m base-level, the plot data: src/parser/cigar.rs:807-985 (emit_baseplotdatas, parse_cigar_to_base_plotdata,
 * parse_maf_to_base_plotdata) + the csv rendering of src/tools/dotplot.rs:283-289 -------------------- */
/* out = one csv row "ref_start,ref_end,query_start,query_end,cigar,ref_chro,query_chro\n" per plot segment (struct
 * BasePlotdata, dotplot.rs:181-190), records in file order (the reference's par_bridge order is unspecified).  The csv
 * writer's header row (WGA_PLOTDATA_CSV_HEADER) precedes the first row of a non-empty result and is left to the caller,
 * who may be concatenating slabs.  skip_cutoff = `-l` (an I / D longer than it becomes a segment of its own).  PAF: ops
 * other than M = X I D are ignored, as the reference's match arm does. */
#define WGA_PLOTDATA_CSV_HEADER "ref_start,ref_end,query_start,query_end,cigar,ref_chro,query_chro\n"
int wga_maf_plotdata(wga_ctx *ctx, const void *maf, uint64_t len, int mem, uint32_t flags, const char *query_name,
                     uint64_t skip_cutoff, wga_span *out);
int wga_paf_plotdata(wga_ctx *ctx, const void *paf, uint64_t len, int mem, uint32_t flags, uint64_t skip_cutoff, wga_span *out);

/* ---- validate: src/tools/validate.rs:44-147 (a free rider on the PAF CIGAR counters) ------------ */
/* report (always host memory) = the reference's report text: counts, then "{name}:{start}-{end}" of every record whose
 * query / target end differs from start + the CIGAR's consumed bases; fix != 0: `fixed` = every record re-serialised
 * with the expected ends (the reference's -f file).  Lists and records come in file order (the reference's rayon fold
 * has no defined order).  A CIGAR error is the reference's `get_stat().unwrap()` panic -> WGA_E_REF_PANIC. */
int wga_paf_validate(wga_ctx *ctx, const void *paf, uint64_t len, int mem, uint32_t flags, int fix, wga_span *report,
                     wga_span *fixed);

/* ---- AlignRecord-level batch access (common.rs:142-176 get_stat / get_cigar_string) --------- */
/* per-record counters of the last wga_maf2paf / wga_maf_stat / wga_paf_stat call, struct Cigar order
 * (cigar.rs:16-29): match, mismatch, ins_event, ins, del_event, del, inv_ins_event, inv_ins,
 * inv_del_event, inv_del, inv_event.  Copies n_records x 11 u64 into dst (host). */
int wga_last_record_counts(wga_ctx *ctx, uint64_t *dst, uint64_t cap_records, uint64_t *n_records);

#ifdef __cplusplus
}
#endif
#endif
