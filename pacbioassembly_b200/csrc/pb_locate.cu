// pb_locate.cu -- the locate loop (locator.cpp:70-92) and the batched align entry point, as pipelines over
// K1 (seeds) -> K2 (probe/gather) -> K3a (prefix filter) -> K3 (banded aligner, first success in list order).
#include <algorithm>
#include <chrono>

#include "pb_internal.cuh"

// PB_HOST_TRACE=1: host wall time of the phases of pb_locate_run on stderr (where the GPU waits for the host)
struct HostTrace {
    bool on = getenv("PB_HOST_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    std::string line;
    void mark(const char *what)
    {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        char buf[64];
        snprintf(buf, sizeof buf, " %s %.2f", what, std::chrono::duration<double, std::milli>(now - t).count());
        line += buf;
        t = now;
    }
    ~HostTrace() { if (on) fprintf(stderr, "[pb_host_trace] pb_locate_run (ms):%s\n", line.c_str()); }
};

extern "C" void pb_locate_default_params(pb_locate_params *p)
{
    if (!p) return;
    p->R = 0.15;     // locator.cpp:68
    p->ntrial = 50;  // locator.cpp:74
    p->minlen = 500; // locator.cpp:72
    p->maxn = 40000; // locator.cpp:23
    p->maxm = 6000;  // locator.cpp:24
    p->want_ops = 0;
    p->reserved = 0;
}

// The byte-exact aligner variant keeps one Eq plane per distinct non-ACGT byte value of seg_a's sequence, four at most.
static int check_tables(pb_ctx *ctx, const pb_seqset *s, const char *what)
{
    for (int64_t i = 0; i < s->n; ++i)
        if (s->tab_count[i] == 255)
            return pb_fail(ctx, PB_ERR_ALPHABET, "%s %lld holds more than four distinct byte values outside {A,C,G,T}", what, (long long)i);
    return PB_OK;
}

// ---------------------------------------------------------------------------------------------
// pb_align_batch
// ---------------------------------------------------------------------------------------------

extern "C" int pb_align_batch(pb_ctx *ctx, const char *a_text, const int64_t *a_off, const int32_t *a_len, const int32_t *a_stride,
                              const char *b_text, const int64_t *b_off, const int32_t *b_len, const int32_t *b_stride, int64_t n,
                              double R, int maxn, int maxm, pb_align_out *out, uint8_t *ops, const int64_t *ops_off)
{
    if (!ctx || n < 0 || (n && (!a_text || !a_off || !a_len || !b_text || !b_off || !b_len || !out)) || (ops && !ops_off))
        return pb_fail(ctx, PB_ERR_ARG, "pb_align_batch: bad argument");
    if (!n) return PB_OK;
    if (n > INT32_MAX) return pb_fail(ctx, PB_ERR_ARG, "pb_align_batch: too many pairs in one call");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    pb_timer_reset(ctx);
    pb_timer_begin(ctx, PB_T_TOTAL);
    pb_seqset *A = nullptr, *B = nullptr;
    int r = pb_seqset_from_text(ctx, a_text, a_off, a_len, a_stride, n, &A);
    if (r == PB_OK) r = pb_seqset_from_text(ctx, b_text, b_off, b_len, b_stride, n, &B);
    // seg_a sequences with more than four distinct bytes outside {A,C,G,T} do not fit the byte-exact bit-parallel variant (one
    // extra Eq plane per such value): those pairs are redone below by the wavefront aligner, which compares raw bytes
    std::vector<int64_t> wide;
    if (r == PB_OK)
        for (int64_t i = 0; i < n; ++i)
            if (A->tab_count[i] == 255) wide.push_back(i);
    DevBuf d_out, d_ops, d_ops_off;
    int64_t extent = 0;
    if (r == PB_OK) r = d_out.alloc_zero(ctx, (size_t)n * sizeof(pb_align_out));
    if (r == PB_OK && ops) {
        for (int64_t i = 0; i < n; ++i) {
            if (ops_off[i] < 0) { r = pb_fail(ctx, PB_ERR_ARG, "negative ops offset"); break; }
            extent = std::max<int64_t>(extent, ops_off[i] + a_len[i] + b_len[i] + 1);
        }
        if (r == PB_OK) r = d_ops.alloc_zero(ctx, (size_t)extent + 16);
        if (r == PB_OK) r = d_ops_off.alloc(ctx, (size_t)n * 8);
        if (r == PB_OK) r = pb_h2d(ctx, d_ops_off.p, ops_off, (size_t)n * 8);
    }
    if (r == PB_OK) {
        // PB_T_ALIGN starts inside the aligner, after its host-side planning: the stage is the kernels' time
        r = pb_align_pairs(ctx, A, B, n, R, maxn, maxm, d_out.as<pb_align_out>(), ops ? d_ops.as<uint8_t>() : nullptr,
                           ops ? d_ops_off.as<int64_t>() : nullptr);
        pb_timer_end(ctx, PB_T_ALIGN);
    }
    if (r == PB_OK) {
        pb_timer_begin(ctx, PB_T_D2H);
        r = pb_d2h(ctx, out, d_out.p, (size_t)n * sizeof(pb_align_out));
        // transcripts: copy each pair's slot (nedit is only known after the records arrive, so copy whole slots)
        if (r == PB_OK && ops) {
            int64_t lo = INT64_MAX;
            for (int64_t i = 0; i < n; ++i) lo = std::min(lo, ops_off[i]);
            r = pb_d2h(ctx, ops + lo, d_ops.as<uint8_t>() + lo, (size_t)(extent - lo));
        }
        pb_timer_end(ctx, PB_T_D2H);
    }
    pb_timer_end(ctx, PB_T_TOTAL);
    int rs = pb_sync(ctx);
    if (r == PB_OK) r = rs;
    if (r == PB_OK) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) r = pb_fail(ctx, PB_ERR_CUDA, "align kernels failed: %s", cudaGetErrorString(e));
    }
    pb_timer_collect(ctx);
    if (A) pb_seqset_free(A);
    if (B) pb_seqset_free(B);
    if (r == PB_OK && !wide.empty()) { // unit weights: seq_aligner::align itself, any alphabet (pb_alignw.cu); views made forward here
        std::vector<char> ta, tb;
        std::vector<int64_t> oa, ob, oo;
        std::vector<int32_t> la, lb;
        for (int64_t i : wide) {
            oa.push_back((int64_t)ta.size()); ob.push_back((int64_t)tb.size());
            la.push_back(a_len[i]); lb.push_back(b_len[i]);
            const int64_t sa = a_stride ? a_stride[i] : 1, sb = b_stride ? b_stride[i] : 1;
            for (int32_t q = 0; q < a_len[i]; ++q) ta.push_back(a_text[a_off[i] + q * sa]);
            for (int32_t q = 0; q < b_len[i]; ++q) tb.push_back(b_text[b_off[i] + q * sb]);
            oo.push_back(oo.empty() ? 0 : oo.back() + la[la.size() - 2] + lb[lb.size() - 2] + 1);
        }
        std::vector<uint8_t> wops(ops ? (size_t)(oo.back() + la.back() + lb.back() + 16) : 0);
        std::vector<uint8_t> wa(ta.size() + 1, 1), wb(tb.size() + 1, 1);
        std::vector<pb_align_out> wo(wide.size());
        float keep_times[PB_T_COUNT];
        memcpy(keep_times, ctx->times, sizeof keep_times);
        r = pb_align_weighted_batch(ctx, ta.data(), wa.data(), oa.data(), la.data(), tb.data(), wb.data(), ob.data(), lb.data(),
                                    (int64_t)wide.size(), R, 1.0, maxn, maxm, wo.data(), ops ? wops.data() : nullptr,
                                    ops ? oo.data() : nullptr);
        for (int q = 0; q < PB_T_COUNT; ++q) ctx->times[q] += keep_times[q];
        for (size_t q = 0; r == PB_OK && q < wide.size(); ++q) {
            out[wide[q]] = wo[q];
            if (ops) { // the pair's own slot only: the neighbours' transcripts stay
                memset(ops + ops_off[wide[q]], 0, (size_t)a_len[wide[q]] + b_len[wide[q]] + 1);
                if (wo[q].ret >= 0) memcpy(ops + ops_off[wide[q]], wops.data() + oo[q], (size_t)wo[q].nedit);
            }
        }
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// locate
// ---------------------------------------------------------------------------------------------

struct pb_locate_job {
    pb_ctx *ctx = nullptr;
    int64_t nkept = 0, ncand = 0;
    int want_ops = 0;
    std::vector<int64_t> ops_off;
    int64_t extent = 0;
    DevBuf d_recs, d_ops, d_stats, d_votes, d_best_diag;
    mutable unsigned long long stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

extern "C" int pb_locate_run(pb_ctx *ctx, const pb_index *ix, const pb_seqset *ref, int64_t ref_seq, const pb_seqset *reads,
                             const pb_locate_params *prm, const int64_t *ops_off, pb_locate_job **out)
{
    if (!ctx || !ix || !ref || !reads || !prm || !out || ref_seq < 0 || ref_seq >= ref->n)
        return pb_fail(ctx, PB_ERR_ARG, "pb_locate_run: bad argument");
    if (prm->ntrial < 1 || prm->ntrial > 4096) return pb_fail(ctx, PB_ERR_ARG, "ntrial %d out of range", prm->ntrial);
    if (prm->minlen < prm->ntrial + 15)
        return pb_fail(ctx, PB_ERR_ARG, "minlen %d < ntrial+15: encode(read+j) would read past the read (locator.cpp:75)", prm->minlen);
    if (reads->n * (int64_t)prm->ntrial > INT32_MAX) return pb_fail(ctx, PB_ERR_DOMAIN, "too many reads in one batch");
    if (ref->len[ref_seq] != ix->ref_len) return pb_fail(ctx, PB_ERR_ARG, "index was built over a different sequence");
    *out = nullptr;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    PB_TRY(check_tables(ctx, reads, "read"));

    HostTrace ht;
    pb_locate_job *job = new pb_locate_job();
    job->ctx = ctx;
    job->want_ops = prm->want_ops;
    int r = PB_OK;
#define TRYJ(x) do { r = (x); if (r != PB_OK) { delete job; return r; } } while (0)
    // kept reads: len >= minlen, nseq = rank among kept (locator.cpp:72, SURVEY Q-L1)
    std::vector<int32_t> kept, kept_lens;
    std::vector<uint8_t> kept_irr;
    kept.reserve((size_t)reads->n); kept_lens.reserve((size_t)reads->n); kept_irr.reserve((size_t)reads->n);
    const bool ref_irr = (ref->flags[ref_seq] & PB_FLAG_IRREGULAR) != 0;
    for (int64_t i = 0; i < reads->n; ++i)
        if (reads->len[i] >= prm->minlen) {
            kept.push_back((int32_t)i);
            kept_lens.push_back(reads->len[i]);
            kept_irr.push_back((ref_irr || (reads->flags[i] & PB_FLAG_IRREGULAR)) ? 1 : 0);
        }
    const int64_t nkept = (int64_t)kept.size();
    job->nkept = nkept;
    TRYJ(job->d_recs.alloc_zero(ctx, (size_t)std::max<int64_t>(nkept, 1) * sizeof(pb_locate_rec)));
    TRYJ(job->d_stats.alloc_zero(ctx, 64));
    if (nkept == 0) { *out = job; return PB_OK; }

    DevBuf d_kept, d_survive, d_rej, d_ops_off;
    TRYJ(d_kept.alloc(ctx, (size_t)nkept * 4));
    TRYJ(pb_h2d(ctx, d_kept.p, kept.data(), (size_t)nkept * 4));
    if (prm->want_ops) {
        job->ops_off.resize((size_t)nkept);
        int64_t ext = 0;
        for (int64_t k = 0; k < nkept; ++k) {
            const int64_t slot = ((int64_t)2 * kept_lens[k] + prm->maxm + 16 + 15) & ~(int64_t)15;
            if (ops_off) {
                if (ops_off[k] < 0) { delete job; return pb_fail(ctx, PB_ERR_ARG, "negative ops offset"); }
                job->ops_off[k] = ops_off[k];
                ext = std::max(ext, ops_off[k] + slot);
            } else {
                job->ops_off[k] = ext;
                ext += slot;
            }
        }
        job->extent = ext;
        TRYJ(job->d_ops.alloc_zero(ctx, (size_t)ext + 16));
        TRYJ(d_ops_off.alloc(ctx, (size_t)nkept * 8));
        TRYJ(pb_h2d(ctx, d_ops_off.p, job->ops_off.data(), (size_t)nkept * 8));
    }
    ht.mark("kept");
    ProbeOut po;
    TRYJ(pb_locate_seed_probe(ctx, ix, reads, d_kept.as<int32_t>(), nkept, prm->ntrial, &po));
    job->ncand = po.ncand;
    ht.mark("seed+probe(sync)");
    TRYJ(job->d_votes.alloc(ctx, (size_t)nkept * 4));
    TRYJ(job->d_best_diag.alloc(ctx, (size_t)nkept * 4));
    TRYJ(pb_vote(ctx, &po, nkept, prm->ntrial, job->d_votes.as<int32_t>(), job->d_best_diag.as<int32_t>()));
    LocateView lv;
    lv.d_kept = d_kept.as<int32_t>();
    lv.d_qoff = po.d_qoff.as<int64_t>();
    lv.d_cand_pos = po.d_cand_pos.as<int32_t>();
    lv.d_cand_q = po.d_cand_q.as<int32_t>();
    lv.ntrial = prm->ntrial;
    lv.ref_base = ref->base[ref_seq];
    lv.ref_len = ref->len[ref_seq];
    lv.mode = PB_MODE_LOCATE;
    lv.min_overlap = 0;
    SeqSets ss = {reads, ref, nullptr, nullptr};
    TRYJ(d_survive.alloc(ctx, (size_t)std::max<int64_t>(po.ncand, 1)));
    TRYJ(d_rej.alloc(ctx, (size_t)std::max<int64_t>(po.ncand, 1) * 4));
    pb_timer_begin(ctx, PB_T_PREFILTER);
    TRYJ(pb_prefilter(ctx, ss, lv, po.ncand, prm->R, prm->maxn, prm->maxm, d_survive.as<uint8_t>(), d_rej.as<int32_t>()));
    pb_timer_end(ctx, PB_T_PREFILTER);
    ht.mark("vote+prefilter");
    // a pipelined step has queued everything so far on the prep stream, under the aligner of the step before: the aligner
    // itself goes to the context's own stream, behind that one
    TRYJ(pb_join_main(ctx));
    // PB_T_ALIGN starts inside the aligner, after its host-side planning: the stage is the kernels' time
    TRYJ(pb_align_locate(ctx, ss, lv, nkept, kept_lens, kept_irr, prm->R, prm->maxn, prm->maxm, d_survive.as<uint8_t>(),
                         d_rej.as<int32_t>(), job->d_recs.as<pb_locate_rec>(), prm->want_ops ? job->d_ops.as<uint8_t>() : nullptr,
                         prm->want_ops ? d_ops_off.as<int64_t>() : nullptr, job->d_stats.as<unsigned long long>()));
    pb_timer_end(ctx, PB_T_ALIGN);
    ht.mark("align(plan+launch)");
#undef TRYJ
    *out = job;
    return PB_OK;
}

extern "C" int64_t pb_locate_job_nkept(const pb_locate_job *job) { return job ? job->nkept : 0; }
extern "C" int64_t pb_locate_job_ncand(const pb_locate_job *job) { return job ? job->ncand : 0; }

extern "C" int pb_locate_job_stats(const pb_locate_job *job, int64_t *out)
{
    if (!job || !out) return PB_ERR_ARG;
    out[0] = job->ncand;
    out[1] = (int64_t)job->stats[1];
    out[2] = (int64_t)job->stats[0];
    out[3] = (int64_t)job->stats[2];
    out[4] = (int64_t)job->stats[3];
    out[5] = (int64_t)job->stats[4];
    out[6] = (int64_t)job->stats[5];
    out[7] = (int64_t)job->stats[6];
    return PB_OK;
}

extern "C" int pb_locate_job_votes(pb_ctx *ctx, const pb_locate_job *job, int32_t *votes, int32_t *best_diag)
{
    if (!ctx || !job) return pb_fail(ctx, PB_ERR_ARG, "pb_locate_job_votes: bad argument");
    if (job->nkept == 0) return PB_OK;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (votes) PB_TRY(pb_d2h(ctx, votes, job->d_votes.p, (size_t)job->nkept * 4));
    if (best_diag) PB_TRY(pb_d2h(ctx, best_diag, job->d_best_diag.p, (size_t)job->nkept * 4));
    return pb_sync(ctx);
}

extern "C" int pb_locate_job_ops_layout(const pb_locate_job *job, int64_t *ops_off, int64_t *extent)
{
    if (!job) return PB_ERR_ARG;
    if (ops_off) for (size_t k = 0; k < job->ops_off.size(); ++k) ops_off[k] = job->ops_off[k];
    if (extent) *extent = job->extent;
    return PB_OK;
}

extern "C" int pb_locate_fetch(pb_ctx *ctx, const pb_locate_job *job, pb_locate_rec *recs, uint8_t *ops)
{
    if (!ctx || !job || (job->nkept && !recs)) return pb_fail(ctx, PB_ERR_ARG, "pb_locate_fetch: bad argument");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    pb_timer_begin(ctx, PB_T_D2H);
    PB_TRY(pb_d2h(ctx, recs, job->d_recs.p, (size_t)job->nkept * sizeof(pb_locate_rec)));
    PB_TRY(pb_d2h(ctx, job->stats, job->d_stats.p, 64));
    if (ops && job->want_ops && job->extent) PB_TRY(pb_d2h(ctx, ops, job->d_ops.p, (size_t)job->extent));
    pb_timer_end(ctx, PB_T_D2H);
    PB_TRY(pb_sync(ctx));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return pb_fail(ctx, PB_ERR_CUDA, "locate kernels failed: %s", cudaGetErrorString(e));
    pb_timer_collect(ctx);
    return PB_OK;
}

extern "C" void pb_locate_job_free(pb_locate_job *job)
{
    if (!job) return;
    cudaSetDevice(job->ctx->device);
    delete job;
}

extern "C" int pb_locate_batch(pb_ctx *ctx, const pb_index *ix, const pb_seqset *ref, int64_t ref_seq, const char *reads_text,
                               const int64_t *off, const int32_t *len, int64_t nreads, const pb_locate_params *prm,
                               pb_locate_rec *recs, int64_t *nkept, uint8_t *ops, const int64_t *ops_off)
{
    if (!ctx || !prm || nreads < 0 || (nreads && (!reads_text || !off || !len)))
        return pb_fail(ctx, PB_ERR_ARG, "pb_locate_batch: bad argument");
    if (prm->want_ops && (!ops || !ops_off)) return pb_fail(ctx, PB_ERR_ARG, "want_ops needs ops and ops_off");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    pb_timer_reset(ctx);
    pb_timer_begin(ctx, PB_T_TOTAL);
    pb_seqset *rs = nullptr;
    PB_TRY(pb_seqset_from_text(ctx, reads_text, off, len, nullptr, nreads, &rs));
    pb_locate_job *job = nullptr;
    int r = pb_locate_run(ctx, ix, ref, ref_seq, rs, prm, prm->want_ops ? ops_off : nullptr, &job);
    if (r == PB_OK) {
        if (nkept) *nkept = job->nkept;
        r = pb_locate_fetch(ctx, job, recs, prm->want_ops ? ops : nullptr);
    }
    pb_timer_end(ctx, PB_T_TOTAL);
    if (r == PB_OK) r = pb_sync(ctx);
    pb_timer_collect(ctx);
    pb_locate_job_free(job);
    pb_seqset_free(rs);
    return r;
}

// ---------------------------------------------------------------------------------------------
// pipelined locate: the copy of batch k+1 runs under the alignment of batch k
// ---------------------------------------------------------------------------------------------

struct pb_locate_step {
    pb_ctx *ctx = nullptr;
    pb_seqset *reads = nullptr;
    pb_locate_job *job = nullptr;
    cudaEvent_t done = nullptr;
    cudaEvent_t ev[2 * PB_T_COUNT] = {};
    bool timed[PB_T_COUNT] = {};
};

extern "C" void pb_locate_step_free(pb_locate_step *st)
{
    if (!st) return;
    cudaSetDevice(st->ctx->device);
    if (st->job) pb_locate_job_free(st->job);
    if (st->reads) pb_seqset_free(st->reads);
    if (st->done) cudaEventDestroy(st->done);
    for (auto e : st->ev)
        if (e) cudaEventDestroy(e);
    delete st;
}

// h_src: the batch as the host holds it (text blob or .bin image); off/len: where each sequence starts in it.
// on_device: h_src is device memory the caller keeps alive until the step is collected -- no staging buffer, no copy
static int submit_common(pb_ctx *ctx, const pb_index *ix, const pb_seqset *ref, int64_t ref_seq, const void *h_src, size_t nbytes,
                         const int64_t *off, const int32_t *len, int64_t n, int src_mode, const pb_locate_params *prm,
                         pb_locate_step **out, bool on_device = false)
{
    *out = nullptr;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    pb_locate_step *st = new pb_locate_step();
    st->ctx = ctx;
    int r = PB_OK;
    void *d_src = nullptr;
    auto cu = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && r == PB_OK)
            r = pb_fail(ctx, e == cudaErrorMemoryAllocation ? PB_ERR_NOMEM : PB_ERR_CUDA, "pb_locate_submit: %s: %s", what, cudaGetErrorString(e));
    };
    cu(cudaEventCreateWithFlags(&st->done, cudaEventDisableTiming), "event");
    for (int i = 0; i < 2 * PB_T_COUNT && r == PB_OK; ++i) cu(cudaEventCreate(&st->ev[i]), "event");
    if (r == PB_OK) { // plan the aligner's band classes now, while the previous step still runs (they depend on the lengths only)
        std::vector<int32_t> kept_lens;
        for (int64_t i = 0; i < n; ++i)
            if (len[i] >= prm->minlen) kept_lens.push_back(len[i]);
        if (!kept_lens.empty()) r = pb_align_locate_prepare(ctx, kept_lens, prm->R, prm->maxn, prm->maxm, PB_MODE_LOCATE);
    }
    if (r == PB_OK && !ctx->prep_stream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi); // hi = the greatest priority: the prep kernels take the first SM slots that
        cu(cudaStreamCreateWithPriority(&ctx->prep_stream, cudaStreamNonBlocking, hi), "prep stream"); // the aligner's CTAs give up
        cu(cudaEventCreateWithFlags(&ctx->prep_event, cudaEventDisableTiming), "event");
    }
    if (r == PB_OK) { // from here to the aligner everything is queued on the prep stream (pb_locate_run switches back)
        ctx->main_pending = ctx->stream;
        ctx->stream = ctx->prep_stream;
    }
    // the batch lands in one of the context's two staging buffers (no allocation on the pipelined path)
    const int slot = ctx->stage_next;
    if (!on_device) ctx->stage_next ^= 1;
    if (r == PB_OK && !ctx->stage_ev[slot]) cu(cudaEventCreateWithFlags(&ctx->stage_ev[slot], cudaEventDisableTiming), "event");
    if (on_device) {
        d_src = const_cast<void *>(h_src);
    } else if (r == PB_OK && ctx->stage_bytes[slot] < nbytes + 16) {
        if (ctx->stage[slot]) {
            cu(cudaEventSynchronize(ctx->stage_ev[slot]), "wait for the staging buffer");
            cudaFree(ctx->stage[slot]);
            ctx->stage[slot] = nullptr;
            ctx->stage_bytes[slot] = 0;
        }
        const size_t want = nbytes + nbytes / 16 + 4096;
        cu(cudaMalloc(&ctx->stage[slot], want), "device staging buffer for the batch");
        if (r == PB_OK) ctx->stage_bytes[slot] = want;
    } else if (r == PB_OK && ctx->stage[slot]) {
        cu(cudaStreamWaitEvent(ctx->copy_stream, ctx->stage_ev[slot], 0), "wait"); // its previous batch has been ingested
    }
    if (!on_device) d_src = ctx->stage[slot];
    if (r == PB_OK && !on_device) {
        cu(cudaEventRecord(st->ev[2 * PB_T_H2D], ctx->copy_stream), "record");
        cu(cudaMemcpyAsync(d_src, h_src, nbytes, cudaMemcpyHostToDevice, ctx->copy_stream), "host->device copy");
        cu(cudaEventRecord(st->ev[2 * PB_T_H2D + 1], ctx->copy_stream), "record");
        st->timed[PB_T_H2D] = true;
        cu(cudaStreamWaitEvent(ctx->stream, st->ev[2 * PB_T_H2D + 1], 0), "wait");
    }
    if (r == PB_OK) {
        ctx->step_ev = st->ev;
        ctx->step_timed = st->timed;
        pb_timer_begin(ctx, PB_T_TOTAL);
        r = pb_seqset_build(ctx, d_src, off, len, nullptr, n, src_mode, &st->reads, (int64_t)nbytes);
        if (!on_device) cu(cudaEventRecord(ctx->stage_ev[slot], ctx->stream), "record"); // ingest has read the staging buffer by then
        if (r == PB_OK) r = pb_locate_run(ctx, ix, ref, ref_seq, st->reads, prm, nullptr, &st->job);
        if (pb_join_main(ctx) != PB_OK && r == PB_OK) r = PB_ERR_CUDA; // a step that never reached the aligner (nothing kept, an error)
        pb_timer_end(ctx, PB_T_TOTAL);
        ctx->step_ev = nullptr;
        ctx->step_timed = nullptr;
        st->timed[PB_T_H2D] = !on_device; // the copy-stream pair recorded above (pb_seqset_build does not touch it)
    }
    if (ctx->main_pending) { ctx->stream = ctx->main_pending; ctx->main_pending = nullptr; } // an error before anything was queued
    if (r == PB_OK) cu(cudaEventRecord(st->done, ctx->stream), "record");
    if (r != PB_OK) { pb_locate_step_free(st); return r; }
    *out = st;
    return PB_OK;
}

extern "C" int pb_locate_submit(pb_ctx *ctx, const pb_index *ix, const pb_seqset *ref, int64_t ref_seq, const char *reads_text,
                                const int64_t *off, const int32_t *len, int64_t nreads, const pb_locate_params *prm,
                                pb_locate_step **step)
{
    if (!ctx || !ix || !ref || !prm || !step || nreads < 0 || (nreads && (!reads_text || !off || !len)))
        return pb_fail(ctx, PB_ERR_ARG, "pb_locate_submit: bad argument");
    int64_t lo = INT64_MAX, hi = 0;
    for (int64_t i = 0; i < nreads; ++i) {
        if (len[i] <= 0) continue;
        if (off[i] < 0) return pb_fail(ctx, PB_ERR_ARG, "negative text offset");
        lo = std::min(lo, off[i]);
        hi = std::max(hi, off[i] + len[i]);
    }
    if (lo > hi) lo = hi = 0;
    std::vector<int64_t> rel(off, off + nreads);
    for (auto &x : rel) x -= lo;
    return submit_common(ctx, ix, ref, ref_seq, reads_text + lo, (size_t)(hi - lo), rel.data(), len, nreads, PB_SRC_TEXT, prm, step);
}

extern "C" int pb_locate_submit_device(pb_ctx *ctx, const pb_index *ix, const pb_seqset *ref, int64_t ref_seq, const void *d_text,
                                       size_t text_bytes, const int64_t *off, const int32_t *len, int64_t nreads,
                                       const pb_locate_params *prm, pb_locate_step **step)
{
    if (!ctx || !ix || !ref || !prm || !step || nreads < 0 || (nreads && (!d_text || !off || !len)))
        return pb_fail(ctx, PB_ERR_ARG, "pb_locate_submit_device: bad argument");
    for (int64_t i = 0; i < nreads; ++i)
        if (len[i] > 0 && (off[i] < 0 || (size_t)(off[i] + len[i]) > text_bytes))
            return pb_fail(ctx, PB_ERR_ARG, "views reach outside the device text blob");
    return submit_common(ctx, ix, ref, ref_seq, d_text, text_bytes, off, len, nreads, PB_SRC_TEXT, prm, step, true);
}

extern "C" int pb_locate_submit_bin(pb_ctx *ctx, const pb_index *ix, const pb_seqset *ref, int64_t ref_seq, const uint8_t *bin,
                                    size_t nbytes, int min_excl, int max_excl, const pb_locate_params *prm, pb_locate_step **step)
{
    if (!ctx || !ix || !ref || !prm || !step || (!bin && nbytes)) return pb_fail(ctx, PB_ERR_ARG, "pb_locate_submit_bin: bad argument");
    // record walk, spaced_seed.cpp:330-342: u32 length, ceil(len/4) body bytes, records back to back
    std::vector<int64_t> off;
    std::vector<int32_t> len;
    size_t p = 0;
    while (p + 4 <= nbytes) {
        uint32_t l;
        memcpy(&l, bin + p, 4);
        const size_t body = ((size_t)l + 3) / 4;
        if (p + 4 + body > nbytes) return pb_fail(ctx, PB_ERR_ARG, "truncated .bin record at byte %zu", p);
        if ((int64_t)l > min_excl && (int64_t)l < max_excl) {
            off.push_back((int64_t)p + 4);
            len.push_back((int32_t)l);
        }
        p += 4 + body;
    }
    return submit_common(ctx, ix, ref, ref_seq, bin, nbytes, off.data(), len.data(), (int64_t)off.size(), PB_SRC_PACKED, prm, step);
}

extern "C" int64_t pb_locate_step_nkept(const pb_locate_step *st) { return st && st->job ? st->job->nkept : 0; }
extern "C" int64_t pb_locate_step_ops_extent(const pb_locate_step *st) { return st && st->job ? st->job->extent : 0; }

extern "C" int pb_locate_collect(pb_ctx *ctx, pb_locate_step *st, pb_locate_rec *recs, int64_t *nkept, uint8_t *ops, int64_t *stats)
{
    if (!ctx || !st || !st->job || (st->job->nkept && !recs)) return pb_fail(ctx, PB_ERR_ARG, "pb_locate_collect: bad argument");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    pb_locate_job *job = st->job;
    int r = PB_OK;
    auto cu = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && r == PB_OK) r = pb_fail(ctx, PB_ERR_CUDA, "pb_locate_collect: %s: %s", what, cudaGetErrorString(e));
    };
    cu(cudaStreamWaitEvent(ctx->copy_stream, st->done, 0), "wait");
    cu(cudaEventRecord(st->ev[2 * PB_T_D2H], ctx->copy_stream), "record");
    if (job->nkept) cu(cudaMemcpyAsync(recs, job->d_recs.p, (size_t)job->nkept * sizeof(pb_locate_rec), cudaMemcpyDeviceToHost, ctx->copy_stream), "records");
    cu(cudaMemcpyAsync(job->stats, job->d_stats.p, 64, cudaMemcpyDeviceToHost, ctx->copy_stream), "stats");
    if (ops && job->want_ops && job->extent) cu(cudaMemcpyAsync(ops, job->d_ops.p, (size_t)job->extent, cudaMemcpyDeviceToHost, ctx->copy_stream), "transcripts");
    cu(cudaEventRecord(st->ev[2 * PB_T_D2H + 1], ctx->copy_stream), "record");
    if (r == PB_OK) r = pb_wait_stream(ctx, ctx->copy_stream);
    if (r == PB_OK) {
        st->timed[PB_T_D2H] = true;
        if (nkept) *nkept = job->nkept;
        if (stats) pb_locate_job_stats(job, stats);
        for (int i = 0; i < PB_T_COUNT; ++i) { // this step's stage times become the context's current timings
            float ms = 0.f;
            ctx->times[i] = 0.f;
            if (st->timed[i] && cudaEventElapsedTime(&ms, st->ev[2 * i], st->ev[2 * i + 1]) == cudaSuccess) ctx->times[i] = ms;
            else cudaGetLastError();
        }
    }
    pb_locate_step_free(st);
    return r;
}

// ---------------------------------------------------------------------------------------------
// assembler-side probe / verify: the trial loop of spaced_seed.cpp:424-436 with try_align (:261-299)
// ---------------------------------------------------------------------------------------------

extern "C" void pb_overlap_default_params(pb_overlap_params *p)
{
    if (!p) return;
    p->R = 0.3;            // MAXR, common.h:37 (spaced_seed -r)
    p->max_trial = 32;     // spaced_seed.cpp:93
    p->min_overlap = 64;   // OVERLAP_MIN, common.h:39
    p->maxn = 26000;       // t_aligner = seq_aligner<MAX_READ_LEN+MAX_DIFF_LEN, MAX_DIFF_LEN>, seq_aligner.h:260
    p->maxm = 6000;
    p->seed_at_quirk = 0;
    p->want_ops = 0;
    p->ref_shift = 0;
}

extern "C" int pb_overlap_batch(pb_ctx *ctx, const pb_index *ix, const pb_seqset *ref, int64_t ref_seq, const pb_seqset *reads,
                                const pb_overlap_params *prm, pb_overlap_rec *recs, uint8_t *ops, const int64_t *ops_off)
{
    return pb_overlap_subset(ctx, ix, ref, ref_seq, reads, nullptr, reads ? reads->n : 0, prm, recs, ops, ops_off);
}

extern "C" int pb_overlap_subset(pb_ctx *ctx, const pb_index *ix, const pb_seqset *ref, int64_t ref_seq, const pb_seqset *reads,
                                 const int32_t *ids, int64_t nids, const pb_overlap_params *prm, pb_overlap_rec *recs, uint8_t *ops,
                                 const int64_t *ops_off)
{
    if (!ctx || !ix || !ref || !reads || !prm || ref_seq < 0 || ref_seq >= ref->n || nids < 0 || (nids && !recs) || (!ids && nids != reads->n))
        return pb_fail(ctx, PB_ERR_ARG, "pb_overlap_batch: bad argument");
    if (prm->max_trial < 1 || prm->max_trial > 2048) return pb_fail(ctx, PB_ERR_ARG, "max_trial %d out of range", prm->max_trial);
    if (prm->want_ops && (!ops || !ops_off)) return pb_fail(ctx, PB_ERR_ARG, "want_ops needs ops and ops_off");
    // the reference may be longer than the indexed text at both ends: a grown ref_seq is indexed over [beg, end) only (ref_seq.h:291-293)
    if (prm->ref_shift < 0 || ix->ref_len < 0 || ref->len[ref_seq] < ix->ref_len + prm->ref_shift)
        return pb_fail(ctx, PB_ERR_ARG, "index was built over a different sequence");
    if (nids * (int64_t)prm->max_trial * 2 > INT32_MAX) return pb_fail(ctx, PB_ERR_DOMAIN, "too many reads in one batch");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    pb_timer_reset(ctx);
    pb_timer_begin(ctx, PB_T_TOTAL);
    const int64_t n = nids;
    if (n == 0) return PB_OK;
    for (int64_t k = 0; k < n; ++k) {
        const int64_t i = ids ? ids[k] : k;
        if (i < 0 || i >= reads->n) return pb_fail(ctx, PB_ERR_ARG, "read id %lld outside the set of %lld", (long long)i, (long long)reads->n);
        if (reads->len[i] < prm->max_trial + 16)
            return pb_fail(ctx, PB_ERR_ARG, "read %lld is shorter than max_trial+16: seed_at(read, len-j-16) would start before the read "
                           "(the reference only keeps reads longer than 500, spaced_seed.cpp:336)", (long long)i);
    }
    // backward trials align reversed views: build the reversed copies once per call
    pb_seqset *reads_rev = nullptr, *ref_rev = nullptr;
    int r = pb_seqset_reversed(ctx, reads, &reads_rev);
    if (r == PB_OK) r = pb_seqset_reversed(ctx, ref, &ref_rev);
    DevBuf d_kept, d_survive, d_rej, d_recs, d_ops, d_ops_off, d_stats;
    ProbeOut po;
    std::vector<int32_t> kept((size_t)n), kept_lens((size_t)n);
    std::vector<uint8_t> kept_irr((size_t)n, 0);
    for (int64_t k = 0; k < n; ++k) {
        kept[k] = ids ? ids[k] : (int32_t)k;
        kept_lens[k] = reads->len[kept[k]];
    }
    int64_t extent = 0;
    if (r == PB_OK) r = d_kept.alloc(ctx, (size_t)n * 4);
    if (r == PB_OK) r = pb_h2d(ctx, d_kept.p, kept.data(), (size_t)n * 4);
    if (r == PB_OK) r = d_recs.alloc_zero(ctx, (size_t)n * sizeof(pb_overlap_rec));
    if (r == PB_OK) r = d_stats.alloc_zero(ctx, 64);
    if (r == PB_OK && prm->want_ops) {
        for (int64_t k = 0; k < n && r == PB_OK; ++k) {
            if (ops_off[k] < 0) r = pb_fail(ctx, PB_ERR_ARG, "negative ops offset");
            extent = std::max<int64_t>(extent, ops_off[k] + 3 * (int64_t)kept_lens[k] + 2 * prm->maxm + 16);
        }
        if (r == PB_OK) r = d_ops.alloc_zero(ctx, (size_t)extent + 16);
        if (r == PB_OK) r = d_ops_off.alloc(ctx, (size_t)n * 8);
        if (r == PB_OK) r = pb_h2d(ctx, d_ops_off.p, ops_off, (size_t)n * 8);
    }
    if (r == PB_OK) r = pb_overlap_seed_probe(ctx, ix, reads, d_kept.as<int32_t>(), n, prm->max_trial, prm->min_overlap, prm->seed_at_quirk, &po);
    if (r == PB_OK) {
        LocateView lv;
        lv.d_kept = d_kept.as<int32_t>();
        lv.d_qoff = po.d_qoff.as<int64_t>();
        lv.d_cand_pos = po.d_cand_pos.as<int32_t>();
        lv.d_cand_q = po.d_cand_q.as<int32_t>();
        lv.ntrial = 2 * prm->max_trial;
        lv.ref_base = ref->base[ref_seq];
        lv.ref_len = ref->len[ref_seq];
        lv.mode = PB_MODE_OVERLAP;
        lv.min_overlap = prm->min_overlap;
        lv.ref_shift = prm->ref_shift;
        SeqSets ss = {reads, ref, reads_rev, ref_rev};
        r = d_survive.alloc(ctx, (size_t)std::max<int64_t>(po.ncand, 1));
        if (r == PB_OK) r = d_rej.alloc(ctx, (size_t)std::max<int64_t>(po.ncand, 1) * 4);
        pb_timer_begin(ctx, PB_T_PREFILTER);
        if (r == PB_OK) r = pb_prefilter(ctx, ss, lv, po.ncand, prm->R, prm->maxn, prm->maxm, d_survive.as<uint8_t>(), d_rej.as<int32_t>());
        pb_timer_end(ctx, PB_T_PREFILTER);
        // PB_T_ALIGN starts inside the aligner, after its host-side planning: the stage is the kernels' time
        // the record layouts of the two modes coincide field by field (pb_locate_rec / pb_overlap_rec, both 56 bytes)
        if (r == PB_OK)
            r = pb_align_locate(ctx, ss, lv, n, kept_lens, kept_irr, prm->R, prm->maxn, prm->maxm, d_survive.as<uint8_t>(), d_rej.as<int32_t>(),
                                reinterpret_cast<pb_locate_rec *>(d_recs.p), prm->want_ops ? d_ops.as<uint8_t>() : nullptr,
                                prm->want_ops ? d_ops_off.as<int64_t>() : nullptr, d_stats.as<unsigned long long>());
        pb_timer_end(ctx, PB_T_ALIGN);
    }
    if (r == PB_OK) {
        pb_timer_begin(ctx, PB_T_D2H);
        r = pb_d2h(ctx, recs, d_recs.p, (size_t)n * sizeof(pb_overlap_rec));
        if (r == PB_OK && prm->want_ops && extent) r = pb_d2h(ctx, ops, d_ops.p, (size_t)extent);
        pb_timer_end(ctx, PB_T_D2H);
    }
    pb_timer_end(ctx, PB_T_TOTAL);
    int rs = pb_sync(ctx);
    if (r == PB_OK) r = rs;
    if (r == PB_OK) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) r = pb_fail(ctx, PB_ERR_CUDA, "overlap kernels failed: %s", cudaGetErrorString(e));
    }
    pb_timer_collect(ctx);
    if (r == PB_OK && ids)
        for (int64_t k = 0; k < n; ++k) recs[k].id = ids[k]; // the kernels number the records by rank in the batch
    if (reads_rev) pb_seqset_free(reads_rev);
    if (ref_rev) pb_seqset_free(ref_rev);
    return r;
}
