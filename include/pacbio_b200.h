/*
 * pacbio_b200.h -- C ABI of the B200-native read-to-reference hot path.
 *
 * Drop-in boundary for vmingchen/PacBioAssembly's spaced-seed lookup + banded-DP verify path.
 * The reference has no FFI layer; its boundary is the header-level class API.  Each entry point
 * below names the reference interface it replaces (file:line under the reference's src/).  The C++
 * classes in pacbioassembly_b200/host/ (dna_seq, seq_accessor, seq_aligner<>, hash_table view, the
 * locator driver) are thin wrappers over exactly these calls -- see INTEGRATION.md.
 *
 * Conventions
 *   - plain C, plain pointers and sizes; no CUDA or torch types appear in any signature.
 *   - every call returns PB_OK (0) or a negative pb_status; pb_last_error() gives the text.
 *   - all work runs on the GPU (sm_100a).  There is NO CPU fallback: without a device every compute
 *     call fails with PB_ERR_NO_DEVICE.
 *   - calls are synchronous on return; internally they are ordered on the context's CUDA stream.
 *   - one context per host thread per GPU (the reference is single-threaded with global state).
 *   - host buffers are owned by the caller; handles are owned by the library until *_free.
 */
#ifndef PACBIO_B200_H
#define PACBIO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_ABI_VERSION 1

#if defined(__GNUC__)
#define PB_API __attribute__((visibility("default")))
#else
#define PB_API
#endif

typedef enum {
    PB_OK = 0,
    PB_ERR_NO_DEVICE = -1,   /* no CUDA device / driver: the product path never falls back to the CPU */
    PB_ERR_CUDA = -2,        /* a CUDA call failed; see pb_last_error */
    PB_ERR_ARG = -3,         /* invalid argument */
    PB_ERR_NOMEM = -4,       /* device or host allocation failed */
    PB_ERR_DOMAIN = -5,      /* outside the supported domain (e.g. band half-width >= PB_MAX_BAND) */
    PB_ERR_ALPHABET = -6     /* a sequence holds bytes outside {A,C,G,T} and the requested path cannot take them */
} pb_status;

/* widest supported band half-width (max_dst); the reference's MAX_DIFF_LEN is 6000 (common.h:35) */
#define PB_MAX_BAND 8191

typedef struct pb_ctx pb_ctx;
typedef struct pb_seqset pb_seqset; /* device-resident sequences: 2-bit packed (dna_seq.h:113-127 layout) + bit planes */
typedef struct pb_index pb_index;   /* device-resident seed index: replaces hash_table (common.h:54) */

/* edit operations, seq_aligner.h:32-36 */
enum { PB_MATCH = 1, PB_INSERT = 2, PB_DELETE = 3 };

/* seed-index build policy */
enum {
    PB_POLICY_LOCATOR = 0, /* locator.cpp:62-66 : every position 0..len-1, tail reads past the end as code 3 */
    PB_POLICY_REFSEQ = 1   /* ref_seq.h:291-311 : head min(len-16,20000) ascending, then tail descending */
};

/* ---- context ------------------------------------------------------------------------------- */

PB_API int pb_abi_version(void);
PB_API int pb_device_count(void); /* 0 when no usable device; never an error */
PB_API int pb_ctx_create(int device, pb_ctx **out);
PB_API void pb_ctx_destroy(pb_ctx *ctx);
/* text of the last error raised through ctx (or through pb_ctx_create when ctx is NULL) */
PB_API const char *pb_last_error(const pb_ctx *ctx);
/* the context's cudaStream_t as an opaque pointer, so callers can bracket calls with their own events */
PB_API void *pb_ctx_stream(pb_ctx *ctx);
/* number of kernels launched by this context since creation (bench.py reports the per-step difference) */
PB_API int64_t pb_ctx_launch_count(const pb_ctx *ctx);
/* upper bound on device scratch (bytes) the aligner may claim for parent matrices; 0 = default (50% of free) */
PB_API int pb_ctx_set_scratch_limit(pb_ctx *ctx, size_t bytes);

/* named device timings (milliseconds, CUDA events on the context's streams) of the most recent call.  For a pipelined step
 * (pb_locate_submit*) the stages before the aligner run on a second stream under the aligner of the step before, so PB_T_TOTAL
 * -- first kernel of the step to the end of its aligner -- includes the wait behind that step. */
enum {
    PB_T_H2D = 0, PB_T_INGEST, PB_T_SEED, PB_T_INDEX, PB_T_PROBE, PB_T_PREFILTER, PB_T_ALIGN, PB_T_D2H, PB_T_TOTAL,
    PB_T_COUNT
};
PB_API int pb_ctx_timings(const pb_ctx *ctx, float *ms /* [PB_T_COUNT] */);
/* Measured peak of the integer ALU pipe (LOP3 / SHF / IADD3, what the banded aligner issues): a register-only kernel of
 * independent chains on every SM; *warp_instr_per_s = warp-level instructions retired per second.  The denominator of the
 * aligner's int-pipe fraction (BASELINE.md section 3 asks for a measured figure, not a datasheet one). */
PB_API int pb_int_pipe_peak(pb_ctx *ctx, double *warp_instr_per_s);
/* Measured peak of independent random 4-byte reads of a table of `table_bytes` (0: 64 MB, the bucket table of a weight-12
 * mask; rounded down to a power of two): no key stream, no result stream, eight reads in flight per thread.  The ceiling the
 * bulk seed probe (one random bucket read per query, locator.cpp:75-79) is measured against, next to the HBM peak. */
PB_API int pb_random_gather_peak(pb_ctx *ctx, size_t table_bytes, double *reads_per_s);

/* ---- L0: sequence representation (dna_seq.h) -------------------------------------------------- */

/* dna_seq::encode (dna_seq.h:86-96) for n windows: out[i] = seed word of text[off[i] .. off[i]+16), where bytes
 * at or beyond text_len read as NUL (-> code 3), which is what locator.cpp:63 sees at the contig tail. */
PB_API int pb_encode_batch(pb_ctx *ctx, const char *text, size_t text_len, const int64_t *off, int64_t n, uint32_t *out);
/* dna_seq::decode (dna_seq.h:101-107): out16[16*i .. 16*i+16) */
PB_API int pb_decode_batch(pb_ctx *ctx, const uint32_t *codes, int64_t n, char *out16);
/* dna_seq::text2bin (dna_seq.h:113-127): record = u32 length + ceil(tlen/4) bytes, 4 bases per byte, first base in
 * bits 7:6.  Returns the record length through *written; PB_ERR_ARG if cap is too small (the reference asserts). */
PB_API int pb_text2bin(pb_ctx *ctx, const char *text, size_t tlen, uint8_t *out, size_t cap, size_t *written);
/* dna_seq::bin2text (dna_seq.h:133-145); writes tlen chars + NUL; needs cap > tlen */
PB_API int pb_bin2text(pb_ctx *ctx, const uint8_t *rec, char *out, size_t cap, size_t *tlen);
/* dna_seq::seed_at (dna_seq.h:62-76) for n positions of one packed record.  Canonical value = encode(text+pos);
 * quirk != 0 reproduces the reference's pos%4==0 branch, which reads the word at byte offset pos (SURVEY Q-S1);
 * bytes beyond rec_bytes read as 0. */
PB_API int pb_seed_at_batch(pb_ctx *ctx, const uint8_t *rec, size_t rec_bytes, const int32_t *pos, int64_t n, int quirk,
                     uint32_t *out);
/* parse_pattern (spaced_seed.cpp:166-180; locator.cpp:51-54): '1' -> care.  Pure string -> mask, no device needed. */
PB_API uint32_t pb_parse_pattern(const char *pattern);

/* Upload n sequences.  Sequence i has len[i] elements; element k is text[off[i] + k*stride[i]] (stride +1 or -1:
 * seq_accessor forward / backward views, dna_seq.h:185-233).  stride == NULL means all +1. */
PB_API int pb_seqset_from_text(pb_ctx *ctx, const char *text, const int64_t *off, const int32_t *len, const int32_t *stride,
                        int64_t n, pb_seqset **out);
/* Same, but `d_text` is a DEVICE pointer to the text blob of text_bytes bytes (inputs already resident in HBM);
 * off/len/stride are host arrays. */
PB_API int pb_seqset_from_device_text(pb_ctx *ctx, const void *d_text, size_t text_bytes, const int64_t *off,
                               const int32_t *len, const int32_t *stride, int64_t n, pb_seqset **out);
/* Walk a .bin image (binary_test.cpp:55-63 writer, spaced_seed.cpp:330-342 reader): records back to back, keep those
 * with min_excl < length < max_excl (the reference keeps 500 < len < 20000).  ids = rank among kept. */
PB_API int pb_seqset_from_bin(pb_ctx *ctx, const uint8_t *bin, size_t nbytes, int min_excl, int max_excl, pb_seqset **out);
PB_API void pb_seqset_free(pb_seqset *s);
PB_API int64_t pb_seqset_count(const pb_seqset *s);
PB_API int32_t pb_seqset_length(const pb_seqset *s, int64_t i);
/* decode sequence i back to text (bin2text semantics: non-ACG bytes come back as 'T'); needs cap > length */
PB_API int pb_seqset_text(pb_ctx *ctx, const pb_seqset *s, int64_t i, char *out, size_t cap);
/* the packed body (4 bases/byte, MSB first) of sequence i, ceil(len/4) bytes */
PB_API int pb_seqset_packed(pb_ctx *ctx, const pb_seqset *s, int64_t i, uint8_t *out, size_t cap);

/* ---- L1: seeds and the seed index ------------------------------------------------------------ */

/* K1 bulk: keys[p] = encode(text+p) & mask for every position p in [0,len) of sequence i (positions within 15 of the
 * end see code 3 past the end, SURVEY Q-S3).  keys is a host buffer of len entries. */
PB_API int pb_seed_extract(pb_ctx *ctx, const pb_seqset *s, int64_t i, uint32_t mask, uint32_t *keys);
/* K1 bulk over the whole set, keys left on the device (bandwidth measurement).  Returns the number of keys produced
 * through *nkeys and the kernel time in ms through *kernel_ms (either may be NULL). */
PB_API int pb_seed_extract_all_device(pb_ctx *ctx, const pb_seqset *s, uint32_t mask, int64_t *nkeys, float *kernel_ms);

/* K1 + K2 bulk over every position of `s` as a query against `ix`, everything left on the device (bandwidth measurement of the
 * probe/gather kernels).  Algorithmic bytes: 4 B key + 8 B bucket header per query (count pass); key + header + 4 B per
 * position read + 8 B per candidate written (gather pass).  Any output pointer may be NULL. */
PB_API int pb_probe_bulk_device(pb_ctx *ctx, const pb_index *ix, const pb_seqset *s, int64_t *nqueries, int64_t *ncand,
                                float *ms_seed, float *ms_count, float *ms_gather);

/* key = encode(ref+i) & mask; if (key) map[key].push_back(i)   (locator.cpp:62-66 / ref_seq.h:291-311) */
PB_API int pb_index_build(pb_ctx *ctx, const pb_seqset *ref, int64_t seq, uint32_t mask, int policy, pb_index **out);
/* The seed map filled the way the reference's drivers fill it, one seedmap[key].push_back(pos) at a time (locator.cpp:65,
 * ref_seq.h:299,307; common.h:54): n (key, position) pairs in insertion order.  pb_index_find_batch returns each key's
 * positions in that order.  Zero keys are skipped (the drivers never insert one, locator.cpp:64).  Such an index answers
 * lookups; the batched locate / overlap pipelines want one built by pb_index_build over a resident sequence. */
PB_API int pb_index_build_pairs(pb_ctx *ctx, const uint32_t *keys, const int32_t *pos, int64_t n, pb_index **out);
PB_API void pb_index_free(pb_index *ix);
PB_API int64_t pb_index_nkeys(const pb_index *ix);    /* hash_table::size() */
PB_API int64_t pb_index_nentries(const pb_index *ix); /* positions stored; get_seedmap's return value is nhead+ntail, see pb_index_nscanned */
PB_API int64_t pb_index_nscanned(const pb_index *ix); /* positions visited (ref_seq.h:310 return value under PB_POLICY_REFSEQ) */
PB_API uint32_t pb_index_mask(const pb_index *ix);
/* hash_table::find (locator.cpp:76, spaced_seed.cpp:265) for n keys at once.  Positions of key i, in the
 * reference's list order, are written to pos[pos_off[i] .. pos_off[i] + min(count[i], cap_each)); count[i]
 * is the full list length (0 == end()).  pos/pos_off may be NULL to get counts only. */
PB_API int pb_index_find_batch(pb_ctx *ctx, const pb_index *ix, const uint32_t *keys, int64_t n, int64_t *count, int32_t *pos,
                        const int64_t *pos_off, int64_t cap_each);

/* ---- L2: banded edit-distance aligner (seq_aligner.h) ---------------------------------------- */

/* Result of one seq_aligner<MAXN,MAXM>::align call (seq_aligner.h:73-81,92-125), fresh-state semantics
 * (cells the call never writes read as 0, SURVEY Q-D2). */
typedef struct {
    int32_t ret;               /* matlen_b, or -1 */
    int32_t len_a, len_b, max_dst;
    int32_t matlen_a, matlen_b;
    int32_t cost;              /* final_cost() */
    int32_t diag_cost;         /* get_cost(|a|,|a|) if that cell was computed, else 0 (locator.cpp:86) */
    int32_t nedit;
    int32_t fail_row;          /* row at which the early-failure test (seq_aligner.h:185) fired, else 0 */
    int64_t cells;             /* DP cells of the reference's recurrence covered by this call */
} pb_align_out;

/* n independent align(seg_a, seg_b) calls with ratio R and template limits (maxn, maxm).
 * Sequences are given as accessor views like pb_seqset_from_text.  ops (may be NULL) receives the forward-ordered
 * edit operations of pair i at ops[ops_off[i] ..), nedit[i] bytes (PB_MATCH/INSERT/DELETE); the caller sizes each
 * slot with at least a_len+b_len+1 bytes.  The transcripts come back in ONE copy: every byte of ops between the lowest
 * ops_off[i] and the end of the highest slot is overwritten -- nedit bytes of operations per successful pair, zeros
 * everywhere else (failed pairs, slack behind a transcript, gaps between slots); keep nothing of your own in that range.
 * The same holds for the ops arguments of pb_locate_batch / pb_locate_fetch / pb_overlap_batch / pb_overlap_subset.
 * edit.val (seq_aligner.h:42) is seg_b's element under each MATCH/INSERT and is filled in by the C++ wrapper from the
 * caller's own text. */
PB_API int pb_align_batch(pb_ctx *ctx, const char *a_text, const int64_t *a_off, const int32_t *a_len, const int32_t *a_stride,
                   const char *b_text, const int64_t *b_off, const int32_t *b_len, const int32_t *b_stride, int64_t n,
                   double R, int maxn, int maxm, pb_align_out *out, uint8_t *ops, const int64_t *ops_off);

/* EXTENSION -- quality-weighted scoring (BASELINE config 3).  The reference has no quality-aware DP: its scoring hooks are
 * hard-wired to unit costs (seq_aligner.h:136-137) and quality.cpp only prints the mean ASCII code of a line.  This entry
 * point is seq_aligner::align with those two hooks replaced by per-element table lookups and everything else kept (band,
 * tie-breaking, early failure, goal cell, coverage test, traceback):
 *     match(a_i, b_j) = (a_i != b_j) ? a_w[i] : 0      DELETE (a_i skipped) costs a_w[i]      INSERT (b_j skipped) costs b_w[j]
 * a_w / b_w: one weight (1..4) per byte of a_text / b_text, same offsets.  Forward views only.  The early-failure line is
 * cost(i,i) > i*R*fail_scale (1 = the reference's).  With all weights 1 and fail_scale 1 the results are pb_align_batch's. */
PB_API int pb_align_weighted_batch(pb_ctx *ctx, const char *a_text, const uint8_t *a_w, const int64_t *a_off, const int32_t *a_len,
                                   const char *b_text, const uint8_t *b_w, const int64_t *b_off, const int32_t *b_len, int64_t n,
                                   double R, double fail_scale, int maxn, int maxm, pb_align_out *out, uint8_t *ops,
                                   const int64_t *ops_off);

/* ---- the locate loop (locator.cpp:70-92) ----------------------------------------------------- */

typedef struct {
    int32_t nseq;      /* rank among kept reads (len >= minlen), column 1 */
    int32_t found;     /* 0/1 */
    int32_t j;         /* read offset of the winning seed */
    int32_t pos;       /* contig position, column 2 */
    int32_t cost;      /* final_cost(), column 3 */
    int32_t seg_len;   /* len - j, column 4 */
    int32_t diag_cost; /* get_cost(len-j,len-j), column 5 */
    int32_t matlen_a, matlen_b, nedit;
    int32_t ncand;     /* align() calls the reference would have made for this read */
    int32_t _pad;
    int64_t cells;     /* DP cells the reference would have evaluated for this read */
} pb_locate_rec;

typedef struct {
    double R;          /* 0.15 in locator.cpp:68 */
    int32_t ntrial;    /* 50, locator.cpp:74 */
    int32_t minlen;    /* 500, locator.cpp:72 */
    int32_t maxn;      /* 40000, locator.cpp:23 */
    int32_t maxm;      /* 6000, locator.cpp:24 */
    int32_t want_ops;  /* also produce the winning transcripts */
    int32_t reserved;
} pb_locate_params;

PB_API void pb_locate_default_params(pb_locate_params *p);

/* For every read with len >= minlen: j = 0..ntrial-1 until found; key = encode(read+j) & mask; candidates in
 * list order; first align(read[j:], ref[pos:]) > 0 wins.  recs has one entry per kept read, in input order;
 * *nkept receives their number.  ops/ops_off as in pb_align_batch, indexed by kept rank, slots of at least
 * 2*len + maxm + 16 bytes (ignored unless want_ops). */
PB_API int pb_locate_batch(pb_ctx *ctx, const pb_index *ix, const pb_seqset *ref, int64_t ref_seq, const char *reads,
                    const int64_t *off, const int32_t *len, int64_t nreads, const pb_locate_params *prm,
                    pb_locate_rec *recs, int64_t *nkept, uint8_t *ops, const int64_t *ops_off);

/* Device-resident variant used to time the path without host<->device copies: reads already uploaded as a seqset;
 * results stay on the device until pb_locate_fetch.  ops_off (host array indexed by kept rank, may be NULL) fixes the
 * transcript layout; with NULL the library packs slots of 2*len + maxm + 16 bytes (rounded up to 16) back to back and
 * pb_locate_job_ops_layout reports them.  pb_locate_fetch copies the records and, if ops != NULL, the byte range
 * [0, extent) of the transcript buffer straight into `ops` (one copy; bytes between slots are overwritten too). */
typedef struct pb_locate_job pb_locate_job;
PB_API int pb_locate_run(pb_ctx *ctx, const pb_index *ix, const pb_seqset *ref, int64_t ref_seq, const pb_seqset *reads,
                  const pb_locate_params *prm, const int64_t *ops_off, pb_locate_job **job);
PB_API int64_t pb_locate_job_nkept(const pb_locate_job *job);
PB_API int64_t pb_locate_job_ncand(const pb_locate_job *job); /* seed hits (candidates) gathered for the batch */
/* after pb_locate_fetch: out[0] = candidates gathered (K2), out[1] = alignments the banded aligner (K3) ran,
 * out[2] = reference cell count of those alignments (what seq_aligner.h:151-190 would have filled for them), out[3] = band
 * cells K3 actually computed (its first pass runs a certified strip of the band), out[4] = reads the strip
 * could not certify and the full-band pass ran again, out[5] = integer-ALU warp instructions of K3's row loops (rows x the
 * per-row count read off the SASS of each band class), out[6] = traceback rounds of the strip pass (32 parent
 * tiles recomputed from their checkpoints per round), out[7] = cold starts of its window ring */
PB_API int pb_locate_job_stats(const pb_locate_job *job, int64_t *out /* [8] */);
/* Diagonal-bin tally of each kept read's seed hits (K2): votes[k] = hits whose diagonal pos - j falls in the fullest
 * 256-base bin, best_diag[k] = that bin's first diagonal.  A diagnostic of how concentrated the hits are; candidates are
 * never reordered or pruned by it (the reference takes the first success in list order).  Either pointer may be NULL. */
PB_API int pb_locate_job_votes(pb_ctx *ctx, const pb_locate_job *job, int32_t *votes, int32_t *best_diag);
PB_API int pb_locate_job_ops_layout(const pb_locate_job *job, int64_t *ops_off /* [nkept] or NULL */, int64_t *extent);
PB_API int pb_locate_fetch(pb_ctx *ctx, const pb_locate_job *job, pb_locate_rec *recs, uint8_t *ops);

/* Pipelined locate: what a caller streaming batches through locator.cpp:70-92 uses so that the host->device copy of batch
 * k+1 runs under the alignment of batch k.  pb_locate_submit starts the copy of the batch's text (pinned host memory makes
 * it asynchronous) on the context's copy stream, queues the whole step behind the step submitted before it and returns
 * once its kernels are queued; pb_locate_collect waits for that step only, copies its records (and, with want_ops, the
 * byte range [0, extent) of its transcripts, layout as pb_locate_run with ops_off == NULL) to the host and frees the step.
 * Steps complete in submission order; submit k+1 before collecting k.  pb_locate_submit_bin takes the batch as a .bin image
 * (binary_test.cpp:55-63: u32 length + ceil(len/4) packed bytes per record, 4x fewer bytes to copy) and keeps records with
 * min_excl < len < max_excl like spaced_seed.cpp:336 -- pass prm->minlen - 1 and INT32_MAX for locator.cpp's rule. */
typedef struct pb_locate_step pb_locate_step;
PB_API int pb_locate_submit(pb_ctx *ctx, const pb_index *ix, const pb_seqset *ref, int64_t ref_seq, const char *reads,
                            const int64_t *off, const int32_t *len, int64_t nreads, const pb_locate_params *prm,
                            pb_locate_step **step);
PB_API int pb_locate_submit_bin(pb_ctx *ctx, const pb_index *ix, const pb_seqset *ref, int64_t ref_seq, const uint8_t *bin,
                                size_t nbytes, int min_excl, int max_excl, const pb_locate_params *prm, pb_locate_step **step);
/* the same for a batch whose text is already in device memory (as pb_seqset_from_device_text takes it): no staging copy; the
 * caller keeps d_text alive until the step has been collected */
PB_API int pb_locate_submit_device(pb_ctx *ctx, const pb_index *ix, const pb_seqset *ref, int64_t ref_seq, const void *d_text,
                                   size_t text_bytes, const int64_t *off, const int32_t *len, int64_t nreads,
                                   const pb_locate_params *prm, pb_locate_step **step);
PB_API int64_t pb_locate_step_nkept(const pb_locate_step *step);
PB_API int64_t pb_locate_step_ops_extent(const pb_locate_step *step);
/* out[8] as pb_locate_job_stats; timings[PB_T_COUNT] (may be NULL) = this step's stage times as pb_ctx_timings */
PB_API int pb_locate_collect(pb_ctx *ctx, pb_locate_step *step, pb_locate_rec *recs, int64_t *nkept, uint8_t *ops,
                             int64_t *stats /* [8] or NULL */);
PB_API void pb_locate_step_free(pb_locate_step *step); /* only for a step that will not be collected */
PB_API void pb_locate_job_free(pb_locate_job *job);

/* ---- assembler-side probe / verify (spaced_seed.cpp:261-299, 424-436; ref_seq.h:259-266) -------------------- */

typedef struct {
    int32_t id;        /* rank of the read in the set (seq_index::id) */
    int32_t found;     /* 0/1 */
    int32_t j;         /* trial number of the success */
    int32_t ref_pos;   /* seed-map position (*it) of the success */
    int32_t cost;      /* final_cost(), the "found <id> at cost c" line (spaced_seed.cpp:429-431) */
    int32_t read_pos;  /* the pos argument of try_align: j (forward) or len-j-16 (backward) */
    int32_t dir;       /* +1 head / forward, -1 tail / backward */
    int32_t matlen_a;  /* ref_ml */
    int32_t matlen_b;  /* seg_ml */
    int32_t nedit;
    int32_t ncand;     /* ref_seq::try_align calls the reference would have made for this read */
    int32_t _pad;
    int64_t cells;
} pb_overlap_rec;

typedef struct {
    double R;              /* 0.3 (MAXR, common.h:37; spaced_seed -r) */
    int32_t max_trial;     /* 32, spaced_seed.cpp:93 (-t) */
    int32_t min_overlap;   /* 64, OVERLAP_MIN (common.h:39): segment length and matlen_a gates */
    int32_t maxn, maxm;    /* t_aligner = seq_aligner<26000,6000>, seq_aligner.h:260 */
    int32_t seed_at_quirk; /* 1: dna_seq::seed_at's shipped pos%4==0 branch (byte offset pos, SURVEY Q-S1); needs a set made
                              by pb_seqset_from_bin (the raw image is what that branch reads); 0: encode(text+pos) */
    int32_t want_ops;
    int32_t ref_shift;     /* 0, or beg - pre of a grown reference (ref_seq.h:363-367): `ix` was built over the text [beg, end) of
                              the current iteration while `ref` holds the whole readable text [pre, post); seed-map positions are
                              then relative to beg and views are taken at position + ref_shift (ref_seq::get_accessor, :282-286) */
} pb_overlap_params;

PB_API void pb_overlap_default_params(pb_overlap_params *p);

/* For every sequence of `reads`: for j < max_trial: try_align(read, j, +1) || try_align(read, len-j-16, -1), i.e. probe
 * seed_at(read,pos) & mask in `ix` (built with PB_POLICY_REFSEQ over ref/ref_seq), and take the first list entry for which
 * align(ref view, read view) >= 0 and matlen_a >= min_overlap -- forward views, or backward views anchored at the seed's
 * last base (spaced_seed.cpp:274-285).  The reference is treated as locked (no voting / growth, ref_seq.h:266).
 * recs: one entry per sequence of `reads`.  ops/ops_off as in pb_align_batch, slots of 3*len + 2*maxm + 16 bytes. */
PB_API int pb_overlap_batch(pb_ctx *ctx, const pb_index *ix, const pb_seqset *ref, int64_t ref_seq, const pb_seqset *reads,
                            const pb_overlap_params *prm, pb_overlap_rec *recs, uint8_t *ops, const int64_t *ops_off);
/* The same for the nids sequences ids[0..nids) of `reads` only, in that order (the pool of reads still unmatched,
 * spaced_seed.cpp:419-437): recs[k] / ops_off[k] belong to read ids[k], recs[k].id = ids[k].  The set stays the whole .bin image, so
 * the shipped seed_at (which reads raw image bytes past the record at pos%4==0, SURVEY Q-S1) sees what the reference sees. */
PB_API int pb_overlap_subset(pb_ctx *ctx, const pb_index *ix, const pb_seqset *ref, int64_t ref_seq, const pb_seqset *reads,
                             const int32_t *ids, int64_t nids, const pb_overlap_params *prm, pb_overlap_rec *recs, uint8_t *ops,
                             const int64_t *ops_off);

/* ---- consensus voting (ref_seq.h:25-41 apply_edits, :47-183 base_vote / vote_box, :207-256 ctor / append / prepend,
 *      :317-362 evolve / elect) ---------------------------------------------------------------------------------- */

typedef struct pb_consensus pb_consensus; /* ref_seq's voting state on the device: the text [pre, post) and one vote box per base */

/* ref_seq(const char *ptxt, int len, bool l, int w): every base starts with `weight` votes and total = 1 (ref_seq.h:218-225) */
PB_API int pb_consensus_create(pb_ctx *ctx, const char *text, int64_t len, int weight, pb_consensus **out);
PB_API void pb_consensus_free(pb_consensus *c);
PB_API int64_t pb_consensus_length(const pb_consensus *c); /* ref_seq::length() = end - beg */
/* *before = beg - pre (text grown in front of the current iteration's origin), *total = post - pre */
PB_API int pb_consensus_extent(const pb_consensus *c, int64_t *before, int64_t *total);
/* the text [beg, end) (full == 0) or [pre, post) (full != 0), NUL-terminated; cap > its length */
PB_API int pb_consensus_text(pb_ctx *ctx, const pb_consensus *c, int full, char *out, size_t cap);
/* the same text as a one-sequence pb_seqset (device to device): what pb_index_build / pb_overlap_batch take as reference */
PB_API int pb_consensus_seqset(pb_ctx *ctx, const pb_consensus *c, int full, pb_seqset **out);
/* ref_seq::append / prepend (ref_seq.h:227-243): the text a read contributes beyond an end of the reference */
PB_API int pb_consensus_append(pb_ctx *ctx, pb_consensus *c, const char *seg, int32_t len);
PB_API int pb_consensus_prepend(pb_ctx *ctx, pb_consensus *c, const char *seg, int32_t len);
/* ref_seq::elect (ref_seq.h:351-361) for every found record of a pb_overlap_batch call made with want_ops: recs / ops / ops_off
 * as that call returned them, `reads` the set it ran on (edit.val = the read's base).  All matches must have been aligned against
 * the text as it is now; votes commute, so the batch is one launch. */
PB_API int pb_consensus_elect_batch(pb_ctx *ctx, pb_consensus *c, const pb_seqset *reads, const pb_overlap_rec *recs, int64_t n,
                                    const uint8_t *ops, const int64_t *ops_off);
/* ref_seq::evolve (ref_seq.h:317-348): suppliments above half the votes become bases, bases at or below half are dropped
 * (their votes move to the suppliment in front of them), the text is rewritten from the winners; pre = beg, post = end after */
PB_API int pb_consensus_evolve(pb_ctx *ctx, pb_consensus *c);
/* the vote boxes of [pre, post), nine ints each: selection A,C,G,T, suppliment A,C,G,T, total */
PB_API int pb_consensus_votes(pb_ctx *ctx, const pb_consensus *c, int32_t *out9, int64_t cap_boxes);

/* ---- all-vs-all overlap detection (BASELINE config 5) --------------------------------- */

/* The reference has no all-vs-all driver.  This is its trial loop (spaced_seed.cpp:424-436) with every sequence T of the set
 * taking the locked reference's place in turn: for each ordered pair (T, Q), T != Q, the result is what
 *     seedmap = T.get_seedmap(mask);  for j < max_trial: try_align(Q, j, +1) || try_align(Q, len-j-16, -1)
 * returns (try_align: spaced_seed.cpp:261-299, ref_seq::try_align: ref_seq.h:259-266, get_seedmap: ref_seq.h:291-311).  One
 * index over the whole set replaces the per-T seed maps. */

/* seed index of EVERY sequence of `set`, each indexed as ref_seq::get_seedmap indexes a reference of at most
 * MAX_READ_LEN+16 = 20016 bases (head pass only, positions 0..len-17 ascending; longer sequences: PB_ERR_DOMAIN).
 * pb_index_nscanned = sum of the per-sequence get_seedmap return values.  Usable with pb_overlap_all_run only. */
PB_API int pb_index_build_set(pb_ctx *ctx, const pb_seqset *set, uint32_t mask, pb_index **out);

typedef struct {
    int32_t read_id;   /* Q: the read whose head / tail seeds were probed (rank in the set) */
    int32_t found;     /* 0/1 */
    int32_t j;         /* trial number of the success */
    int32_t ref_pos;   /* seed-map position (*it) inside T */
    int32_t cost;      /* final_cost() */
    int32_t read_pos;  /* the pos argument of try_align: j (forward) or len-j-16 (backward) */
    int32_t dir;       /* +1 head / forward, -1 tail / backward */
    int32_t matlen_a;  /* ref_ml (elements of T consumed) */
    int32_t matlen_b;  /* seg_ml (elements of Q consumed) */
    int32_t nedit;
    int32_t ncand;     /* ref_seq::try_align calls the reference would have made for this (T, Q) pair */
    int32_t ref_id;    /* T: the sequence acting as the locked reference (rank in the set) */
    int64_t cells;     /* DP cells the reference would have evaluated for this pair */
} pb_pair_rec;

typedef struct pb_pairs_job pb_pairs_job;

/* Probe the head / tail trial seeds of reads [q_first, q_first+q_count) of `set` against `ix` (pb_index_build_set over the
 * same set) and verify every (T, Q) pair that shares a seed, first success in (trial, direction, list) order per pair.
 * A query sub-range is how callers bound memory per call and shard the work over GPUs (the index is replicated).
 * prm: as pb_overlap_batch (want_ops must be 0).  Synchronous; results stay on the device until pb_pairs_job_fetch. */
PB_API int pb_overlap_all_run(pb_ctx *ctx, const pb_index *ix, const pb_seqset *set, int64_t q_first, int64_t q_count,
                              const pb_overlap_params *prm, pb_pairs_job **job);
/* out[0] seed hits gathered (K2), [1] (T,Q) pairs with at least one hit, [2] pairs that reached the banded aligner (a
 * candidate survived the exact 32-row prefix filter), [3] pairs found, [4] try_align calls the reference would have made
 * over all pairs, [5] DP cells it would have evaluated, [6] alignments K3 ran, [7] DP cells K3 computed */
PB_API int pb_pairs_job_stats(const pb_pairs_job *job, int64_t *out /* [8] */);
/* found_only != 0: the successful pairs; 0: every pair that reached the aligner (found or not, with ncand / cells).
 * Records come in ascending (read_id, ref_id).  recs == NULL: only *n is set.  cap = room in recs (records). */
PB_API int pb_pairs_job_fetch(pb_ctx *ctx, const pb_pairs_job *job, int found_only, pb_pair_rec *recs, int64_t cap, int64_t *n);
PB_API void pb_pairs_job_free(pb_pairs_job *job);

#ifdef __cplusplus
}
#endif
#endif /* PACBIO_B200_H */
