// pb_internal.cuh -- shared declarations of the B200 read-to-reference library (not installed).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/pacbio_b200.h"

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------

struct LocatePlan; // pb_align.cu: the host-side plan of a locate step (band class of every item)
void pb_locate_plan_free(LocatePlan *lp);

struct pb_ctx {
    int device = 0;
    LocatePlan *planned = nullptr; // made ahead of the step by the pipelined entry points (pb_align_locate_prepare)
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr; // host<->device copies of the pipelined entry points (pb_locate_submit / _collect)
    // Pipelined steps run everything BEFORE the aligner (ingest, seeds, probe, gather, prefilter and the host round trips between
    // them) on a second, high-priority stream, so that it overlaps the aligner kernels of the step before; `stream` points at it
    // while that part is being queued, main_pending holds the real stream until pb_join_main() switches back
    cudaStream_t prep_stream = nullptr;
    cudaEvent_t prep_event = nullptr;
    cudaStream_t main_pending = nullptr;
    cudaMemPool_t pool = nullptr;       // private pool behind DevBuf (bounded release threshold)
    // two grow-only device staging buffers for the batches of pb_locate_submit (the copy of batch k+1 lands in one while
    // batch k is still being ingested from the other); stage_ev[i] = the last ingest that read buffer i has finished
    void *stage[2] = {nullptr, nullptr};
    size_t stage_bytes[2] = {0, 0};
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    int stage_next = 0;
    cudaEvent_t ev[2 * PB_T_COUNT] = {};
    float times[PB_T_COUNT] = {};
    bool timed[PB_T_COUNT] = {};
    // pipelined steps (pb_locate_submit) keep their own stage events: while one is being queued the stage timers record here too
    cudaEvent_t *step_ev = nullptr;
    bool *step_timed = nullptr;
    int64_t launches = 0;
    size_t scratch_limit = 0;
    int sm_count = 148;
    std::string err;
    // pinned staging for small host<->device exchanges
    void *h_pin = nullptr;
    size_t h_pin_bytes = 0;
    // one extra stream per band class of the aligner (forked from / joined back into `stream`)
    std::vector<cudaStream_t> aux_streams;
    std::vector<cudaEvent_t> aux_events;
    cudaEvent_t fork_event = nullptr;
    // grow-only scratch for the aligner's parent planes (kept across calls: no per-step allocation)
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    size_t scratch_budget_cached = 0;
    cudaEvent_t wait_event = nullptr; // blocking-sync event behind pb_wait_stream
};

void pb_set_error(pb_ctx *ctx, const char *fmt, ...);
int pb_fail(pb_ctx *ctx, int code, const char *fmt, ...);

#define PB_CUDA(ctx, call)                                                                            \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return pb_fail((ctx), _e == cudaErrorMemoryAllocation ? PB_ERR_NOMEM : PB_ERR_CUDA,       \
                           "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
    } while (0)

#define PB_TRY(expr)                 \
    do {                             \
        int _r = (expr);             \
        if (_r != PB_OK) return _r;  \
    } while (0)

#define PB_LAUNCH_CHECK(ctx)                        \
    do {                                            \
        (ctx)->launches++;                          \
        PB_CUDA((ctx), cudaGetLastError());         \
    } while (0)

// stage timers: CUDA events on the context stream; times are read back by pb_timer_collect
void pb_timer_reset(pb_ctx *ctx);
void pb_timer_begin(pb_ctx *ctx, int which);
void pb_timer_end(pb_ctx *ctx, int which);
void pb_timer_collect(pb_ctx *ctx); // requires the stream to be idle

// stream-ordered device buffer (cudaMallocAsync pool: reuse across calls is cheap)
struct DevBuf {
    pb_ctx *ctx = nullptr;
    void *p = nullptr;
    size_t bytes = 0;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    int alloc(pb_ctx *c, size_t n);
    int alloc_zero(pb_ctx *c, size_t n);
    void release();
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

int pb_h2d(pb_ctx *ctx, void *dst, const void *src, size_t bytes);
int pb_d2h(pb_ctx *ctx, void *dst, const void *src, size_t bytes);
int pb_sync(pb_ctx *ctx);
// waits for `st`: polls for a short while (the common short waits cost no wake-up latency), then sleeps on a blocking-sync
// event instead of spinning -- a 45 ms aligner step no longer burns a host core per process (PB_SPIN_WAIT=1: plain spinning)
int pb_wait_stream(pb_ctx *ctx, cudaStream_t st);
int pb_join_main(pb_ctx *ctx); // back from the prep stream to the context's own: what follows waits for what was queued there

// ---------------------------------------------------------------------------------------------
// device-resident sequences
// ---------------------------------------------------------------------------------------------
//
// Padded layout: sequence i occupies base offsets [base[i], base[i]+len[i]) of one long coordinate line;
// base[i] is a multiple of 32 and at least 16 padding bases (code 3, what a NUL/T/N maps to under C2I,
// dna_seq.h:21) follow every sequence, so a 16-base seed window starting at any in-sequence position
// never sees a neighbour (SURVEY Q-S3).  Three views of the same line are kept in HBM:
//   packed : 4 bases per byte, first base in bits 7:6 (dna_seq.h:113-127)          -> seeds (K1)
//   hi, lo : bit planes, bit (g & 31) of word (g >> 5) = code bit of base g          -> banded DP (K3)
// flags[i] bit 0 = the text held a byte outside {A,C,G,T} (it maps to code 3 for seeding exactly like
// the reference, but DP compares raw bytes, so such sequences take the byte-exact DP path).

struct pb_seqset {
    pb_ctx *ctx = nullptr;
    int64_t n = 0;
    int64_t total = 0; // padded bases, multiple of 128
    std::vector<int64_t> base;
    std::vector<int32_t> len;
    std::vector<uint32_t> flags;
    DevBuf d_base, d_len, d_flags, d_hi, d_lo, d_packed;
    // Bytes outside {A,C,G,T}: they seed as code 3 exactly like the reference (C2I), but the DP compares raw bytes
    // (seq_aligner.h:136).  d_irr marks their positions on the line (bit plane); the bytes themselves are kept as a
    // sorted exception list (line position, value); tab[i] holds the (at most 4) distinct such values of sequence i.
    DevBuf d_irr, d_exc_pos, d_exc_val, d_tab;
    // sets made from a .bin image keep it (the reference's seed_at quirk reads raw image bytes, SURVEY Q-S1)
    DevBuf d_image, d_recoff;
    int64_t image_bytes = 0;
    int64_t nexc = 0;
    std::vector<uint32_t> tab;      // 4 byte values packed little-endian
    std::vector<uint8_t> tab_count; // 0..4, 255 = more than four distinct values
    int64_t nwords() const { return total / 32; }
};

#define PB_FLAG_IRREGULAR 1u

// ---------------------------------------------------------------------------------------------
// seed index
// ---------------------------------------------------------------------------------------------

#define PB_MAX_RUNS 16

struct BucketFn {          // key -> bucket id
    int exact;             // 1: bit-compress under the mask (injective); 0: multiplicative hash + key check
    int nruns;
    int bits;              // log2(number of buckets)
    uint32_t run_mask[PB_MAX_RUNS];
    uint8_t run_shift[PB_MAX_RUNS];
};

struct pb_index {
    pb_ctx *ctx = nullptr;
    uint32_t mask = 0;
    int policy = 0;
    int64_t nbuckets = 0;
    int64_t nentries = 0, nkeys = 0, nscanned = 0;
    int64_t ref_len = 0;
    bool whole_set = false; // built by pb_index_build_set: entries are positions on the set's padded line
    BucketFn fn;
    DevBuf d_start; // [nbuckets+1] u32
    DevBuf d_pos;   // [nentries] i32, per bucket in the reference's list order
    DevBuf d_key;   // [nentries] u32 (only when !fn.exact)
    // packed bucket headers (exact indexes): pk[b] = start[b] | min(count, pk_esc) << pk_shift -- one random 4-byte read per
    // probe instead of two; a count of pk_esc or more is re-read exactly from start[]
    DevBuf d_pk;
    int pk_shift = 0;
    uint32_t pk_esc = 0;
};

// ---------------------------------------------------------------------------------------------
// kernels / stages implemented across the .cu files
// ---------------------------------------------------------------------------------------------

// pb_seq.cu
int pb_seqset_build(pb_ctx *ctx, const void *d_text, const int64_t *h_toff, const int32_t *h_len,
                    const int32_t *h_stride, int64_t n, int src_mode, pb_seqset **out, int64_t text_bytes = -1);
#define PB_SRC_TEXT 0     // bytes of text, element k at toff + k*stride
#define PB_SRC_PACKED 1   // 4 bases per byte (dna_seq.h:113-127 body), first byte at toff
#define PB_SRC_REVLINE 2  // another set's packed line, sequence read backwards; toff = its base offset (in bases)
int pb_seqset_reversed(pb_ctx *ctx, const pb_seqset *s, pb_seqset **out);

// pb_seed.cu
int pb_seed_bulk_device(pb_ctx *ctx, const pb_seqset *s, int64_t first_base, int64_t count, uint32_t mask,
                        uint32_t *d_keys);
int pb_scan_u32(pb_ctx *ctx, const uint32_t *d_in, uint32_t *d_out, int64_t n, DevBuf &tmp);   // exclusive, out[n] = total
int pb_scan_i64(pb_ctx *ctx, const uint32_t *d_in, int64_t *d_out, int64_t n, DevBuf &tmp);    // exclusive, out[n] = total

struct ProbeOut {
    DevBuf d_qoff;     // [nq+1] i64 exclusive offsets into the candidate arrays
    DevBuf d_cand_pos; // [ncand] i32
    DevBuf d_cand_q;   // [ncand] i32 query id
    int64_t ncand = 0;
};
// keys of the first ntrial offsets of every kept read, then probe + gather
int pb_locate_seed_probe(pb_ctx *ctx, const pb_index *ix, const pb_seqset *reads, const int32_t *d_kept, int64_t nkept,
                         int ntrial, ProbeOut *po);
// diagonal-bin tally per read (diagnostic; never used to prune): votes in the fullest 256-base bin and that bin's diagonal
int pb_vote(pb_ctx *ctx, const ProbeOut *po, int64_t nkept, int ntrial, int32_t *d_votes, int32_t *d_best_diag);
// overlap mode: per read 2*max_trial queries (j forward at pos j, j backward at pos len-j-16), spaced_seed.cpp:424-426
int pb_overlap_seed_probe(pb_ctx *ctx, const pb_index *ix, const pb_seqset *reads, const int32_t *d_kept, int64_t nkept,
                          int max_trial, int min_overlap, int quirk, ProbeOut *po);

// pb_align.cu
struct LocateView { // everything the aligner needs to derive candidate (a,b) views in locate mode
    const int32_t *d_kept;
    const int64_t *d_qoff;
    const int32_t *d_cand_pos;
    const int32_t *d_cand_q;
    int ntrial;       // queries per read: ntrial (locator) or 2*max_trial (overlap: head/forward and tail/backward per j)
    int64_t ref_base; // base offset of the reference sequence inside its seqset
    int32_t ref_len;
    int mode;         // PB_MODE_LOCATE: a = read[j:], b = ref[pos:] (locator.cpp:78-82)
                      // PB_MODE_OVERLAP: a = ref view, b = read view, forward or backward (spaced_seed.cpp:274-285, ref_seq.h:264)
    int min_overlap;  // OVERLAP_MIN gate on matlen_a (ref_seq.h:265), overlap mode only
    int ref_shift = 0; // overlap mode: candidate positions are relative to ref position ref_shift (a reference grown in front, beg - pre)
    // all-vs-all (pb_overlap_all_run): work item k = one (reference sequence, read) pair.  d_kept[k] = the read, d_item_ref[k] =
    // the sequence acting as reference (in ss.ref), candidates [d_item_beg[k], d_item_end[k]) in (trial, list) order with
    // d_cand_q = trial number and d_cand_pos = position inside that reference; d_cand_item (prefilter only) = item of a
    // candidate slot, -1 for a dropped one.  All NULL in the single-reference modes.
    const int32_t *d_item_ref = nullptr;
    const int64_t *d_item_beg = nullptr, *d_item_end = nullptr;
    const int32_t *d_cand_item = nullptr;
};
#define PB_MODE_LOCATE 0
#define PB_MODE_OVERLAP 1
// reads / ref and (overlap mode only, else NULL) their reversed copies for the backward views
struct SeqSets { const pb_seqset *reads, *ref, *reads_rev, *ref_rev; };
int pb_prefilter(pb_ctx *ctx, const SeqSets &ss, const LocateView &lv, int64_t ncand, double R,
                 int maxn, int maxm, uint8_t *d_survive, int32_t *d_rej_cells);
int pb_align_locate(pb_ctx *ctx, const SeqSets &ss, const LocateView &lv, int64_t nkept,
                    const std::vector<int32_t> &kept_lens, const std::vector<uint8_t> &kept_irr, double R, int maxn, int maxm,
                    const uint8_t *d_survive,
                    const int32_t *d_rej_cells, pb_locate_rec *d_recs, uint8_t *d_ops, const int64_t *d_ops_off,
                    unsigned long long *d_stats);
int pb_align_locate_prepare(pb_ctx *ctx, const std::vector<int32_t> &kept_lens, double R, int maxn, int maxm, int mode);
int pb_align_pairs(pb_ctx *ctx, const pb_seqset *A, const pb_seqset *B, int64_t n, double R, int maxn, int maxm,
                   pb_align_out *d_out, uint8_t *d_ops, const int64_t *d_ops_off);

// host-side mirror of seq_aligner.h:94-102 (also used to size scratch)
static inline void pb_align_params(int a_len, int b_len, double R, int *len_a, int *len_b, int *max_dst)
{
    if (b_len >= a_len) {
        *len_a = a_len;
        *max_dst = 1 + (int)(a_len * R);
        *len_b = b_len < *len_a + *max_dst ? b_len : *len_a + *max_dst;
    } else {
        *len_b = b_len;
        *max_dst = 1 + (int)(b_len * R);
        *len_a = a_len < *len_b + *max_dst ? a_len : *len_b + *max_dst;
    }
}
