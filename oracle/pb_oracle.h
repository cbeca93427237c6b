/*
 * pb_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A plain-C restatement of the read-to-reference hot path of
 * vmingchen/PacBioAssembly.  It exists only so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs can CHECK the CUDA path.
 * Nothing under pacbioassembly_b200/ may include, link or call it.
 *
 * Parity status: PINNED.  Every function here is differential-tested against the
 * unmodified reference compiled from /root/reference (oracle/_ref, see
 * oracle/Makefile + oracle/ref_shim.cpp) and against the golden vectors the
 * reference's own tests hold (test/dna_test.cpp, test/aligner_test.cpp,
 * test/ref_test.cpp:119-128); tests/test_oracle.py runs both checks.
 *
 * Each function cites the reference file:line it restates.
 */
#ifndef PB_ORACLE_H
#define PB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- L0: sequence representation (src/dna_seq.h) ------------------------- */

/* C2I, dna_seq.h:21 : A->0 C->1 G->2 anything else->3 */
int pbo_c2i(int ch);

/* dna_seq::encode, dna_seq.h:86-96,147-159 : 16 text bases -> seed word.
 * `avail` = number of readable chars at text; chars beyond behave as NUL
 * (-> code 3), which is what locator.cpp:62-66 sees at the contig tail (Q-S3). */
uint32_t pbo_encode(const char *text, size_t avail);

/* dna_seq::decode, dna_seq.h:101-107 */
void pbo_decode(uint32_t code, char *out16);

/* dna_seq::text2bin, dna_seq.h:113-127 : returns record length 4+ceil(tlen/4),
 * or 0 if cap is too small (the reference asserts). */
size_t pbo_text2bin(const char *text, size_t tlen, uint8_t *out, size_t cap);

/* dna_seq::bin2text, dna_seq.h:133-145 : returns tlen (0 if cap <= tlen). */
size_t pbo_bin2text(const uint8_t *rec, char *out, size_t cap);

/* dna_seq::seed_at, dna_seq.h:62-76.  Canonical = value of encode(text+pos)
 * (what the unaligned branch computes; reads past the record behave as zero
 * bytes -> 'A').  `quirk`!=0 reproduces Q-S1 (pos%4==0 reads the u32 at byte
 * offset pos instead of pos/4); rec_bytes bounds the quirk read (0 beyond). */
uint32_t pbo_seed_at(const uint8_t *rec, size_t rec_bytes, int pos, int quirk);

/* parse_pattern, spaced_seed.cpp:166-180 (same as locator.cpp:51-54 for
 * 16-char patterns): '1' -> care (bits 11), anything else -> 00, padded to 16. */
uint32_t pbo_parse_pattern(const char *pat);

/* ---- L1: seed index (locator.cpp:62-66, ref_seq.h:291-311) --------------- */

enum { PBO_POLICY_LOCATOR = 0, PBO_POLICY_REFSEQ = 1 };

typedef struct pbo_index pbo_index;

/* key = encode(ref+i) & mask ; if (key) map[key].push_back(i).
 * LOCATOR: i in [0,len) ascending, tail NUL-padded.
 * REFSEQ : head i in [0,min(len-16,20000)) ascending, then tail
 *          len-16-i for i in [0,min(len-20016,20000)) (descending positions). */
pbo_index *pbo_index_build(const char *ref, size_t len, uint32_t mask, int policy);
void pbo_index_free(pbo_index *ix);
size_t pbo_index_nkeys(const pbo_index *ix);    /* hash_table::size() */
size_t pbo_index_nentries(const pbo_index *ix); /* total positions stored */
/* hash_table::find : returns count (0 = end()), *pos -> list in insertion order */
size_t pbo_index_find(const pbo_index *ix, uint32_t key, const int32_t **pos);

/* ---- L2: banded edit-distance aligner (src/seq_aligner.h) ---------------- */

enum { PBO_MATCH = 1, PBO_INSERT = 2, PBO_DELETE = 3 }; /* seq_aligner.h:32-36 */

typedef struct {
    int32_t ret;       /* align() return: matlen_b or -1 */
    int32_t len_a, len_b, max_dst;         /* seq_aligner.h:94-102 */
    int32_t matlen_a, matlen_b;            /* goal_cell :191-213 */
    int32_t cost;                          /* final_cost() :130 */
    int32_t diag_cost;                     /* get_cost(|a|,|a|), 0 if unwritten (Q-L2) */
    int32_t nedit;                         /* find_path :214-233 */
    int32_t fail_row;  /* first row whose early-failure test fired, else 0 */
    int64_t cells;     /* DP cells evaluated (rows actually run)            */
} pbo_align_out;

/* seq_aligner<MAXN,MAXM>::align with FRESH-state semantics (Q-D2): cells the
 * call never writes read as 0.  a/b are seq_accessor views: element k is
 * a[k*a_stride] (stride +1 forward, -1 backward).  ops/vals (cap entries each,
 * may be NULL) receive the forward-ordered transcript; vals[k] is b.at(j-1) for
 * MATCH/INSERT and 0 for DELETE.  Domain (Q-D3): len_a < maxn, max_dst < maxm,
 * else -1. */
int pbo_align(const char *a, int a_len, int a_stride,
              const char *b, int b_len, int b_stride,
              double R, int maxn, int maxm,
              pbo_align_out *out, uint8_t *ops, char *vals, size_t cap);

/* EXTENSION (BASELINE config 3, "quality-weighted scoring"; no counterpart in the reference, whose scoring hooks
 * seq_aligner.h:136-137 are hard-wired to unit costs and whose quality.cpp is a stand-alone mean-of-ASCII tool): the same
 * banded recurrence, tie-breaking, goal cell and traceback with the two hooks replaced by per-element table lookups --
 * match(a_i, b_j) = (a_i != b_j) ? wa[i] : 0; indel = wa[i] when a_i is skipped (DELETE), wb[j] when b_j is skipped
 * (INSERT) -- weights in 1..4, forward views, and the early-failure line cost(i,i) > i*R*fail_scale.  With all weights 1
 * and fail_scale 1 it is pbo_align (tests/test_oracle.py checks that), which is what pins it. */
int pbo_align_weighted(const char *a, int a_len, const uint8_t *wa, const char *b, int b_len, const uint8_t *wb,
                       double R, double fail_scale, int maxn, int maxm, pbo_align_out *out, uint8_t *ops, size_t cap);

/* ---- locate loop (locator.cpp:70-92) -------------------------------------- */

typedef struct {
    int32_t nseq;      /* rank among kept reads (len >= minlen), Q-L1 */
    int32_t found;     /* 0/1 */
    int32_t j;         /* read offset of the winning seed */
    int32_t pos;       /* contig position (col 2) */
    int32_t cost;      /* final_cost (col 3) */
    int32_t seg_len;   /* len - j (col 4) */
    int32_t diag_cost; /* get_cost(len-j,len-j) (col 5) */
    int32_t matlen_a, matlen_b, nedit;
    int32_t ncand;     /* align() calls made for this read */
    int64_t cells;     /* DP cells evaluated for this read */
} pbo_locate_rec;

/* For each read r (text at reads+offs[r], length lens[r]) with len >= minlen:
 * j = 0..ntrial-1 until found; key = encode(read+j)&mask; candidates in list
 * order; align(read[j:], ref[pos:]) with ratio R; first ret>0 wins.
 * recs has one entry per KEPT read, in order; returns number of kept reads.
 * nthreads>1 shards reads over pthreads (each with its own DP buffers).
 * If ops_out != NULL, the winning transcript ops of kept read k are stored at
 * ops_out + ops_off[k] (caller sizes with lens[r]*2+maxm). */
int64_t pbo_locate(const pbo_index *ix, const char *ref, size_t ref_len,
                   const char *reads, const int64_t *offs, const int32_t *lens,
                   int64_t nreads, uint32_t mask, double R, int ntrial, int minlen,
                   int maxn, int maxm, int nthreads, pbo_locate_rec *recs,
                   uint8_t *ops_out, const int64_t *ops_off);

/* ---- assembler-side probe / verify: try_align + trial loop (spaced_seed.cpp:261-299, 424-436) ---------------- */

typedef struct {
    int32_t id;        /* rank among kept reads (seq_index::id) */
    int32_t found;
    int32_t j;         /* trial number of the success */
    int32_t ref_pos;   /* seed-map position *it of the success */
    int32_t cost;      /* final_cost() */
    int32_t read_pos;  /* the `pos` argument of try_align */
    int32_t dir;       /* +1 head / forward, -1 tail / backward */
    int32_t matlen_a;  /* ref_ml */
    int32_t matlen_b;  /* seg_ml */
    int32_t nedit;
    int32_t ncand;     /* ref_seq::try_align calls made for this read */
    int32_t _pad;
    int64_t cells;
} pbo_overlap_rec;

/* For every record of the .bin image `bin` with min_excl < len < max_excl (open_binary, spaced_seed.cpp:330-342):
 * for j < max_trial: try_align(read, j, +1) || try_align(read, slen-j-16, -1)   (spaced_seed.cpp:424-426), where
 * try_align probes seed_at(read,pos) & mask in the REFSEQ-policy index `ix` of `ref`, requires a segment of at least
 * min_overlap elements, and accepts the first list entry for which align(ref_view, read_view) >= 0 and matlen_a >=
 * min_overlap (ref_seq.h:264-265, locked reference: no voting).  quirk != 0 reproduces dna_seq::seed_at's pos%4==0
 * branch, reading the u32 at byte offset pos of the record body straight from the image (zero past its end). */
int64_t pbo_overlap(const pbo_index *ix, const char *ref, size_t ref_len, const uint8_t *bin, size_t bin_bytes,
                    int min_excl, int max_excl, uint32_t mask, double R, int max_trial, int min_overlap, int maxn,
                    int maxm, int quirk, int nthreads, pbo_overlap_rec *recs);

/* ---- consensus voting and the unlocked assembler rounds (ref_seq.h:25-41,47-183,207-276,317-362; spaced_seed.cpp:408-453) ---- */

typedef struct pbo_cons pbo_cons; /* ref_seq's voting state: text buffer with pre/beg/end/post and the vote_box list */

pbo_cons *pbo_cons_create(const char *text, size_t len, int weight, size_t cap); /* ref_seq(text, len, false, w) */
void pbo_cons_free(pbo_cons *c);
size_t pbo_cons_length(const pbo_cons *c);                 /* ref_seq::length() = end - beg */
size_t pbo_cons_extent(const pbo_cons *c, long *before);   /* post - pre; *before = beg - pre */
const char *pbo_cons_text(const pbo_cons *c);              /* txt_buf + beg */
void pbo_cons_append(pbo_cons *c, const char *seg, int len);
void pbo_cons_prepend(pbo_cons *c, const char *seg, int len);
/* elect(pos, edits, nedit, forward): ops/vals as produced by pbo_align (vals[k] = seg_b's element under MATCH / INSERT) */
void pbo_cons_elect(pbo_cons *c, int pos, const uint8_t *ops, const char *vals, int nedit, int forward);
void pbo_cons_evolve(pbo_cons *c);
int64_t pbo_cons_votes(const pbo_cons *c, int32_t *out9);  /* per box: sel[4], sup[4], total; returns the number of boxes */

int64_t pbo_assemble(const char *ref_text, size_t ref_len, int weight, const uint8_t *bin, size_t bin_bytes, int min_excl,
                     int max_excl, const uint32_t *round_masks, int nrounds, double R, int max_trial, int min_overlap,
                     int maxn, int maxm, int quirk, char *cons_out, size_t cons_stride, int32_t *cons_len,
                     int32_t *found_round, pbo_overlap_rec *recs);

#ifdef __cplusplus
}
#endif
#endif
