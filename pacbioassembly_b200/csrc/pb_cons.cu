// pb_cons.cu -- consensus voting: ref_seq's vote boxes as a GPU pile-up.
//
// The reference keeps a std::list<vote_box> beside its text (ref_seq.h:47-183, 362): every successful try_align votes its
// transcript into the boxes under the aligned region (elect -> apply_edits, ref_seq.h:25-41,351-361), a read that runs past an
// end grows text and list (append / prepend, :227-243), and evolve() (:317-348) rewrites the text from the votes at the end of
// a round.  Votes are counter increments, so all matches that were aligned against the same text commute: here they are
// applied in one launch (one thread walks one transcript, 16-bit counters packed two per word, atomic adds), and evolve is a
// flag / scan / scatter over the boxes.  Growth changes the text later alignments see, so the caller applies matches in
// batches that end at a growing read (host/src/spaced_seed.cpp does exactly that) -- that order dependence is the reference's.
//
// Layout: struct of arrays over box index (box 0 <-> text position `pre`), with MAX_SEQ_LEN slots of slack on both sides like
// the reference's txt_buf[3*MAX_SEQ_LEN]: sel01/sel23/sup01/sup23 (two unsigned shorts per u32: A|C<<16, G|T<<16), total (int),
// text (u8).  A counter would have to pass 65535 votes before the packing differs from the reference's unsigned short.
#include <algorithm>

#include "pb_internal.cuh"

#define CONS_SLACK 800000 // MAX_SEQ_LEN, common.h:31

struct pb_consensus {
    pb_ctx *ctx = nullptr;
    int64_t cap = 0;              // slots in every array
    int64_t beg = 0, end = 0, pre = 0, post = 0; // ref_seq.h:363-367
    DevBuf sel01, sel23, sup01, sup23, total, text;
    DevBuf n_sel01, n_sel23, n_sup01, n_sup23, n_total, n_text; // evolve's output side
};

__device__ __forceinline__ int c2i_dev(uint8_t c) { return c == 'A' ? 0 : (c == 'C' ? 1 : (c == 'G' ? 2 : 3)); } // dna_seq.h:21

// vote_box(c, w): selection[C2I(c)] = w, total = 1 (ref_seq.h:124); slots [first, first+n) from text already on the device
__global__ void cons_init_kernel(const uint8_t *__restrict__ text, int64_t first, int64_t n, uint32_t w, uint32_t *sel01, uint32_t *sel23,
                                 uint32_t *sup01, uint32_t *sup23, int32_t *total)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = c2i_dev(text[first + i]);
    const uint32_t v = (w & 0xFFFFu) << (16 * (c & 1));
    sel01[first + i] = c < 2 ? v : 0u;
    sel23[first + i] = c < 2 ? 0u : v;
    sup01[first + i] = 0u;
    sup23[first + i] = 0u;
    total[first + i] = 1;
}

struct ElectView {
    const pb_overlap_rec *recs;
    const uint8_t *ops;
    const int64_t *ops_off;
    const uint32_t *read_packed; // the reads' packed line (4 bases per byte, first base in bits 7:6)
    const int64_t *read_base;
    int64_t box0;                // box index of text position 0 (= beg - pre ... as an absolute slot: beg)
    int64_t lo, hi;              // valid slots [pre, post)
};

// 2-bit code of base g of the packed line
__device__ __forceinline__ int packed_code(const uint32_t *pw, int64_t g)
{
    const uint8_t byte = reinterpret_cast<const uint8_t *>(pw)[g >> 2];
    return (byte >> (6 - 2 * (int)(g & 3))) & 3;
}

// elect (ref_seq.h:351-361) + apply_edits (:25-41) for every found record: one thread per match
__global__ void __launch_bounds__(128)
cons_elect_kernel(ElectView v, int64_t n, uint32_t *sel01, uint32_t *sel23, uint32_t *sup01, uint32_t *sup23, int32_t *total)
{
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const pb_overlap_rec r = v.recs[m];
    if (!r.found) return;
    const bool forward = r.dir == 1;
    const int step = forward ? 1 : -1;
    int64_t idx = v.box0 + (forward ? r.ref_pos : r.ref_pos + 15);               // box of the view's first element
    int64_t g = v.read_base[r.id] + (forward ? r.read_pos : r.read_pos + 15);    // line position of seg_b's element 0
    const uint8_t *ops = v.ops + v.ops_off[m];
    for (int k = 0; k < r.nedit; ++k) {
        const int op = ops[k];
        if (op == PB_DELETE) {
            atomicAdd(&total[idx], 1);
            idx += step;
        } else {
            const int c = packed_code(v.read_packed, g); // edit.val = seg_b's element (seq_aligner.h:219,225)
            g += step;
            const uint32_t inc = 1u << (16 * (c & 1));
            if (op == PB_MATCH) {
                atomicAdd(c < 2 ? &sel01[idx] : &sel23[idx], inc);
                atomicAdd(&total[idx], 1);
                idx += step;
            } else { // INSERT: "--it; supply; ++it" forward, the box under a reverse_iterator backward
                const int64_t t = forward ? idx - 1 : idx;
                if (t >= v.lo && t < v.hi) atomicAdd(c < 2 ? &sup01[t] : &sup23[t], inc);
            }
        }
    }
}

__device__ __forceinline__ uint32_t max4(uint32_t p01, uint32_t p23)
{
    return max(max(p01 & 0xFFFFu, p01 >> 16), max(p23 & 0xFFFFu, p23 >> 16));
}
__device__ __forceinline__ char winner4(uint32_t p01, uint32_t p23) // base_vote::winner, ref_seq.h:94-98: first of A,C,G,T at the maximum
{
    const uint32_t m = max4(p01, p23);
    return m == (p01 & 0xFFFFu) ? 'A' : (m == (p01 >> 16) ? 'C' : (m == (p23 & 0xFFFFu) ? 'G' : 'T'));
}

// evolve, pass 1: boxes each input box leaves behind -- itself if is_valid(0.5), plus its suppliment as a box of its own if
// has_supply(0.5) (ref_seq.h:325-339)
__global__ void cons_evolve_count_kernel(const uint32_t *sel01, const uint32_t *sel23, const uint32_t *sup01, const uint32_t *sup23,
                                         const int32_t *total, int64_t first, int64_t n, uint32_t *cnt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t s = first + i;
    const int t = total[s];
    cnt[i] = (2 * (int)max4(sel01[s], sel23[s]) > t ? 1u : 0u) + (2 * (int)max4(sup01[s], sup23[s]) > t ? 1u : 0u);
}

// pass 2: write the surviving boxes and the new text at their scanned positions
__global__ void cons_evolve_write_kernel(const uint32_t *sel01, const uint32_t *sel23, const uint32_t *sup01, const uint32_t *sup23,
                                         const int32_t *total, int64_t first, int64_t n, const int64_t *off, int64_t obase,
                                         uint32_t *o_sel01, uint32_t *o_sel23, uint32_t *o_sup01, uint32_t *o_sup23, int32_t *o_total,
                                         uint8_t *o_text)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t s = first + i;
    const int t = total[s];
    const uint32_t a01 = sel01[s], a23 = sel23[s], b01 = sup01[s], b23 = sup23[s];
    const bool split = 2 * (int)max4(b01, b23) > t, valid = 2 * (int)max4(a01, a23) > t;
    int64_t o = obase + off[i];
    if (valid) {
        o_sel01[o] = a01; o_sel23[o] = a23;
        o_sup01[o] = split ? 0u : b01; o_sup23[o] = split ? 0u : b23; // split() resets the suppliment; otherwise it stays for the next round
        o_total[o] = t;
        o_text[o] = (uint8_t)winner4(a01, a23);
        ++o;
    }
    if (split) { // vote_box::split, ref_seq.h:157-161: selection = the suppliment, same total, empty suppliment
        o_sel01[o] = b01; o_sel23[o] = b23; o_sup01[o] = 0u; o_sup23[o] = 0u;
        o_total[o] = t;
        o_text[o] = (uint8_t)winner4(b01, b23);
    }
}

// pass 3: an erased box hands its selection to the suppliment of the box in front of it in the NEW list (ref_seq.h:341-345)
__global__ void cons_evolve_absorb_kernel(const uint32_t *sel01, const uint32_t *sel23, const uint32_t *sup01, const uint32_t *sup23,
                                          const int32_t *total, int64_t first, int64_t n, const int64_t *off, int64_t obase, uint32_t *o_sup01,
                                          uint32_t *o_sup23)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t s = first + i;
    const int t = total[s];
    const uint32_t a01 = sel01[s], a23 = sel23[s];
    if (2 * (int)max4(a01, a23) > t) return; // valid: kept
    if (off[i] == 0) return;                 // nothing in front of it: its votes are dropped (cur == begin)
    const int64_t prev = obase + off[i] - 1;
    if (a01) atomicAdd(&o_sup01[prev], a01);
    if (a23) atomicAdd(&o_sup23[prev], a23);
}

static inline unsigned grid_of(int64_t n, int b = 256) { return (unsigned)std::max<int64_t>(1, (n + b - 1) / b); }

static int cons_alloc(pb_ctx *ctx, pb_consensus *c)
{
    const size_t n = (size_t)c->cap;
    DevBuf *u32s[] = {&c->sel01, &c->sel23, &c->sup01, &c->sup23, &c->total, &c->n_sel01, &c->n_sel23, &c->n_sup01, &c->n_sup23, &c->n_total};
    for (DevBuf *b : u32s) PB_TRY(b->alloc_zero(ctx, n * 4));
    PB_TRY(c->text.alloc_zero(ctx, n));
    PB_TRY(c->n_text.alloc_zero(ctx, n));
    return PB_OK;
}

extern "C" int pb_consensus_create(pb_ctx *ctx, const char *text, int64_t len, int weight, pb_consensus **out)
{
    if (!ctx || !out || len < 0 || (len && !text)) return pb_fail(ctx, PB_ERR_ARG, "pb_consensus_create: bad argument");
    if (len >= CONS_SLACK) return pb_fail(ctx, PB_ERR_DOMAIN, "reference of %lld bases: the reference's buffers hold MAX_SEQ_LEN = %d (common.h:31)", (long long)len, CONS_SLACK);
    if (weight < 0 || weight > 65535) return pb_fail(ctx, PB_ERR_ARG, "weight %d does not fit an unsigned short vote counter", weight);
    *out = nullptr;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    pb_consensus *c = new pb_consensus();
    c->ctx = ctx;
    c->cap = 3 * (int64_t)CONS_SLACK;
    c->beg = c->pre = CONS_SLACK;
    c->end = c->post = c->beg + len;
    int r = cons_alloc(ctx, c);
    if (r == PB_OK && len) r = pb_h2d(ctx, c->text.as<uint8_t>() + c->beg, text, (size_t)len);
    if (r == PB_OK && len) {
        cons_init_kernel<<<grid_of(len), 256, 0, ctx->stream>>>(c->text.as<uint8_t>(), c->beg, len, (uint32_t)weight, c->sel01.as<uint32_t>(),
                                                               c->sel23.as<uint32_t>(), c->sup01.as<uint32_t>(), c->sup23.as<uint32_t>(), c->total.as<int32_t>());
        ctx->launches++;
    }
    if (r == PB_OK) r = pb_sync(ctx);
    if (r != PB_OK) { delete c; return r; }
    *out = c;
    return PB_OK;
}

extern "C" void pb_consensus_free(pb_consensus *c)
{
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    delete c;
}

extern "C" int64_t pb_consensus_length(const pb_consensus *c) { return c ? c->end - c->beg : 0; }

extern "C" int pb_consensus_extent(const pb_consensus *c, int64_t *before, int64_t *total)
{
    if (!c) return PB_ERR_ARG;
    if (before) *before = c->beg - c->pre;
    if (total) *total = c->post - c->pre;
    return PB_OK;
}

extern "C" int pb_consensus_text(pb_ctx *ctx, const pb_consensus *c, int full, char *out, size_t cap)
{
    if (!ctx || !c || !out) return pb_fail(ctx, PB_ERR_ARG, "pb_consensus_text: bad argument");
    const int64_t lo = full ? c->pre : c->beg, hi = full ? c->post : c->end;
    if ((size_t)(hi - lo) + 1 > cap) return pb_fail(ctx, PB_ERR_ARG, "pb_consensus_text: %lld characters + NUL do not fit %zu bytes", (long long)(hi - lo), cap);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (hi > lo) PB_TRY(pb_d2h(ctx, out, c->text.as<uint8_t>() + lo, (size_t)(hi - lo)));
    PB_TRY(pb_sync(ctx));
    out[hi - lo] = '\0';
    return PB_OK;
}

extern "C" int pb_consensus_seqset(pb_ctx *ctx, const pb_consensus *c, int full, pb_seqset **out)
{
    if (!ctx || !c || !out) return pb_fail(ctx, PB_ERR_ARG, "pb_consensus_seqset: bad argument");
    const int64_t lo = full ? c->pre : c->beg, hi = full ? c->post : c->end;
    const int64_t off = lo;
    const int32_t len = (int32_t)(hi - lo);
    return pb_seqset_from_device_text(ctx, c->text.p, (size_t)c->cap, &off, &len, nullptr, 1, out);
}

static int cons_grow(pb_ctx *ctx, pb_consensus *c, const char *seg, int32_t len, bool back)
{
    if (!ctx || !c || len < 0 || (len && !seg)) return pb_fail(ctx, PB_ERR_ARG, "pb_consensus_append/prepend: bad argument");
    if (len == 0) return PB_OK;
    if (back ? c->post + len > c->cap : c->pre - len < 0)
        return pb_fail(ctx, PB_ERR_DOMAIN, "the consensus outgrew the reference's txt_buf[3*MAX_SEQ_LEN] (ref_seq.h:369)");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t first = back ? c->post : c->pre - len;
    PB_TRY(pb_h2d(ctx, c->text.as<uint8_t>() + first, seg, (size_t)len));
    cons_init_kernel<<<grid_of(len), 256, 0, ctx->stream>>>(c->text.as<uint8_t>(), first, len, 1u, c->sel01.as<uint32_t>(), c->sel23.as<uint32_t>(),
                                                           c->sup01.as<uint32_t>(), c->sup23.as<uint32_t>(), c->total.as<int32_t>());
    PB_LAUNCH_CHECK(ctx);
    PB_TRY(pb_sync(ctx)); // `seg` is the caller's
    if (back) c->post += len; else c->pre -= len;
    return PB_OK;
}

extern "C" int pb_consensus_append(pb_ctx *ctx, pb_consensus *c, const char *seg, int32_t len) { return cons_grow(ctx, c, seg, len, true); }
extern "C" int pb_consensus_prepend(pb_ctx *ctx, pb_consensus *c, const char *seg, int32_t len) { return cons_grow(ctx, c, seg, len, false); }

extern "C" int pb_consensus_elect_batch(pb_ctx *ctx, pb_consensus *c, const pb_seqset *reads, const pb_overlap_rec *recs, int64_t n,
                                        const uint8_t *ops, const int64_t *ops_off)
{
    if (!ctx || !c || !reads || n < 0 || (n && (!recs || !ops || !ops_off))) return pb_fail(ctx, PB_ERR_ARG, "pb_consensus_elect_batch: bad argument");
    if (n == 0) return PB_OK;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    int64_t lo = INT64_MAX, hi = 0;
    for (int64_t m = 0; m < n; ++m) {
        if (!recs[m].found) continue;
        if (recs[m].id < 0 || recs[m].id >= reads->n) return pb_fail(ctx, PB_ERR_ARG, "record %lld names read %d, the set holds %lld", (long long)m, recs[m].id, (long long)reads->n);
        if (reads->flags[recs[m].id] & PB_FLAG_IRREGULAR) return pb_fail(ctx, PB_ERR_ALPHABET, "read %d holds bytes outside {A,C,G,T}", recs[m].id);
        const int64_t r_off = recs[m].dir == 1 ? recs[m].ref_pos : recs[m].ref_pos + 15;
        if (c->beg + r_off < c->pre || c->beg + r_off >= c->post) return pb_fail(ctx, PB_ERR_ARG, "record %lld starts outside the reference (ref_seq::contained)", (long long)m);
        lo = std::min(lo, ops_off[m]);
        hi = std::max(hi, ops_off[m] + recs[m].nedit);
    }
    if (hi <= lo) return PB_OK; // nothing found
    DevBuf d_recs, d_ops, d_off;
    PB_TRY(d_recs.alloc(ctx, (size_t)n * sizeof(pb_overlap_rec)));
    PB_TRY(d_ops.alloc(ctx, (size_t)(hi - lo) + 16));
    PB_TRY(d_off.alloc(ctx, (size_t)n * 8));
    std::vector<int64_t> rel((size_t)n);
    for (int64_t m = 0; m < n; ++m) rel[m] = ops_off[m] - lo;
    PB_TRY(pb_h2d(ctx, d_recs.p, recs, (size_t)n * sizeof(pb_overlap_rec)));
    PB_TRY(pb_h2d(ctx, d_ops.p, ops + lo, (size_t)(hi - lo)));
    PB_TRY(pb_h2d(ctx, d_off.p, rel.data(), (size_t)n * 8));
    ElectView v;
    v.recs = d_recs.as<pb_overlap_rec>();
    v.ops = d_ops.as<uint8_t>();
    v.ops_off = d_off.as<int64_t>();
    v.read_packed = reads->d_packed.as<uint32_t>();
    v.read_base = reads->d_base.as<int64_t>();
    v.box0 = c->beg;
    v.lo = c->pre;
    v.hi = c->post;
    cons_elect_kernel<<<grid_of(n, 128), 128, 0, ctx->stream>>>(v, n, c->sel01.as<uint32_t>(), c->sel23.as<uint32_t>(), c->sup01.as<uint32_t>(),
                                                               c->sup23.as<uint32_t>(), c->total.as<int32_t>());
    PB_LAUNCH_CHECK(ctx);
    return pb_sync(ctx); // host buffers (rel, the caller's) may go
}

extern "C" int pb_consensus_evolve(pb_ctx *ctx, pb_consensus *c)
{
    if (!ctx || !c) return pb_fail(ctx, PB_ERR_ARG, "pb_consensus_evolve: bad argument");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n = c->post - c->pre, first = c->pre;
    int64_t nout = 0;
    if (n > 0) {
        DevBuf d_cnt, d_off, tmp;
        PB_TRY(d_cnt.alloc(ctx, (size_t)n * 4));
        PB_TRY(d_off.alloc(ctx, (size_t)(n + 2) * 8));
        const uint32_t *s01 = c->sel01.as<uint32_t>(), *s23 = c->sel23.as<uint32_t>(), *p01 = c->sup01.as<uint32_t>(), *p23 = c->sup23.as<uint32_t>();
        const int32_t *tot = c->total.as<int32_t>();
        cons_evolve_count_kernel<<<grid_of(n), 256, 0, ctx->stream>>>(s01, s23, p01, p23, tot, first, n, d_cnt.as<uint32_t>());
        PB_LAUNCH_CHECK(ctx);
        PB_TRY(pb_scan_i64(ctx, d_cnt.as<uint32_t>(), d_off.as<int64_t>(), n, tmp));
        PB_TRY(pb_d2h(ctx, &nout, d_off.as<int64_t>() + n, 8));
        PB_TRY(pb_sync(ctx));
        if (CONS_SLACK + nout > c->cap) return pb_fail(ctx, PB_ERR_DOMAIN, "the consensus outgrew the reference's txt_buf[3*MAX_SEQ_LEN]");
        cons_evolve_write_kernel<<<grid_of(n), 256, 0, ctx->stream>>>(s01, s23, p01, p23, tot, first, n, d_off.as<int64_t>(), CONS_SLACK,
                                                                     c->n_sel01.as<uint32_t>(), c->n_sel23.as<uint32_t>(), c->n_sup01.as<uint32_t>(),
                                                                     c->n_sup23.as<uint32_t>(), c->n_total.as<int32_t>(), c->n_text.as<uint8_t>());
        PB_LAUNCH_CHECK(ctx);
        cons_evolve_absorb_kernel<<<grid_of(n), 256, 0, ctx->stream>>>(s01, s23, p01, p23, tot, first, n, d_off.as<int64_t>(), CONS_SLACK,
                                                                      c->n_sup01.as<uint32_t>(), c->n_sup23.as<uint32_t>());
        PB_LAUNCH_CHECK(ctx);
        PB_TRY(pb_sync(ctx));
    }
    std::swap(c->sel01.p, c->n_sel01.p); std::swap(c->sel23.p, c->n_sel23.p);
    std::swap(c->sup01.p, c->n_sup01.p); std::swap(c->sup23.p, c->n_sup23.p);
    std::swap(c->total.p, c->n_total.p); std::swap(c->text.p, c->n_text.p);
    c->beg = c->pre = CONS_SLACK; // ref_seq.h:319
    c->end = c->post = c->beg + nout;
    return PB_OK;
}

// the vote boxes of [pre, post), nine ints per box: selection A,C,G,T, suppliment A,C,G,T, total (tests / inspection)
extern "C" int pb_consensus_votes(pb_ctx *ctx, const pb_consensus *c, int32_t *out9, int64_t cap_boxes)
{
    if (!ctx || !c || !out9) return pb_fail(ctx, PB_ERR_ARG, "pb_consensus_votes: bad argument");
    const int64_t n = c->post - c->pre;
    if (cap_boxes < n) return pb_fail(ctx, PB_ERR_ARG, "pb_consensus_votes: room for %lld boxes, %lld to return", (long long)cap_boxes, (long long)n);
    if (n == 0) return PB_OK;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<uint32_t> a((size_t)n), b((size_t)n), s((size_t)n), t((size_t)n);
    std::vector<int32_t> tot((size_t)n);
    PB_TRY(pb_d2h(ctx, a.data(), c->sel01.as<uint32_t>() + c->pre, (size_t)n * 4));
    PB_TRY(pb_d2h(ctx, b.data(), c->sel23.as<uint32_t>() + c->pre, (size_t)n * 4));
    PB_TRY(pb_d2h(ctx, s.data(), c->sup01.as<uint32_t>() + c->pre, (size_t)n * 4));
    PB_TRY(pb_d2h(ctx, t.data(), c->sup23.as<uint32_t>() + c->pre, (size_t)n * 4));
    PB_TRY(pb_d2h(ctx, tot.data(), c->total.as<int32_t>() + c->pre, (size_t)n * 4));
    PB_TRY(pb_sync(ctx));
    for (int64_t i = 0; i < n; ++i) {
        int32_t *o = out9 + 9 * i;
        o[0] = a[i] & 0xFFFF; o[1] = a[i] >> 16; o[2] = b[i] & 0xFFFF; o[3] = b[i] >> 16;
        o[4] = s[i] & 0xFFFF; o[5] = s[i] >> 16; o[6] = t[i] & 0xFFFF; o[7] = t[i] >> 16;
        o[8] = tot[i];
    }
    return PB_OK;
}
