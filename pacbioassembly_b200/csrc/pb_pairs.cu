// pb_pairs.cu -- all-vs-all overlap detection (BASELINE config 5).
//
// The reference has no all-vs-all driver; what it has is the assembler's trial loop (spaced_seed.cpp:424-436) of every read
// against ONE locked reference (try_align :261-299, ref_seq::try_align ref_seq.h:259-266).  All-vs-all is that loop with every
// read T of the set taking the reference's place in turn: for each ordered pair (T, Q), T != Q,
//     for j < max_trial: try_align(Q, j, +1) || try_align(Q, len-j-16, -1)        against get_seedmap(T)
// and the first success is the pair's overlap.  Running it pair by pair would build one seed map per T; here ONE index over
// the whole set answers every Q's 2*max_trial probes at once, and the hits are regrouped per (T, Q):
//
//   K1' overlap_seed_kernel   keys of the head / tail trials of the query reads              (pb_seed.cu)
//   K2  probe + gather        candidates (line position, query) in (Q, trial, list) order    (pb_seed.cu)
//   P1  pairs_owner_kernel    line position -> owning sequence T; sort key (T, rank in Q's run); self hits dropped
//   P2  pairs_sort_kernel     one CTA per Q: bitonic sort of its run by (T, trial, list) -> runs of equal T = work items
//   P3  pairs_emit_kernel     item tables (Q, T, candidate range), candidate -> item
//   K3a prefilter_kernel      exact early-failure test on rows 1..32                           (pb_align.cu)
//   P4  pairs_live_kernel     items with a surviving candidate go on; the others only add to the try_align / cell totals
//   K3  align_locate_kernel<S,false,true>  first success per item in (trial, list) order      (pb_align.cu)
//   P5  compaction of the found records
#include <algorithm>

#include "pb_internal.cuh"

struct pb_pairs_job {
    pb_ctx *ctx = nullptr;
    int64_t ncand = 0, nitems = 0, nlive = 0, nfound = 0;
    int64_t tot_ncand = 0, tot_cells = 0, k3_aligns = 0, k3_cells = 0;
    DevBuf d_recs, d_found;
};

#define PAIRS_DROPPED 0xFFFFFFFFFFFFFFFFull
#define PAIRS_SORT_SMEM 8192 // keys sorted in shared memory (64 KB); longer runs sort in place in global memory

// P1: owner of each hit + sort key.  Candidate c belongs to query q = cand_q[c] = k*ntr + t of read Q = kept[k]; its run starts at
// qoff[k*ntr].  key = (T << 32) | (c - run start): ascending keys = ascending T, and inside one T the original (trial, list) order.
__global__ void __launch_bounds__(256)
pairs_owner_kernel(const int32_t *__restrict__ cand_pos, const int32_t *__restrict__ cand_q, int64_t ncand, const int64_t *__restrict__ qoff,
                   const int32_t *__restrict__ kept, int ntr, const int64_t *__restrict__ base, int64_t nseq,
                   unsigned long long *__restrict__ key)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand) return;
    const int64_t g = cand_pos[c];
    int64_t lo = 0, hi = nseq; // largest i with base[i] <= g
    while (hi - lo > 1) {
        const int64_t m = (lo + hi) >> 1;
        if (__ldg(base + m) <= g) lo = m; else hi = m;
    }
    const int k = cand_q[c] / ntr;
    const unsigned long long rank = (unsigned long long)(c - qoff[(int64_t)k * ntr]);
    key[c] = (int)lo == kept[k] ? PAIRS_DROPPED : (((unsigned long long)lo << 32) | rank);
}

__device__ __forceinline__ void cmpxchg(unsigned long long *t, uint32_t i, uint32_t l)
{
    const unsigned long long x = t[i], y = t[l];
    if (x > y) { t[i] = y; t[l] = x; }
}

// ascending bitonic network in its "flip" form: every comparator leaves the smaller key at the lower index, so keys at
// indices >= n (virtual +inf) never move and any n sorts in place
__device__ __forceinline__ void bitonic_any(unsigned long long *t, uint32_t n)
{
    uint32_t P = 1;
    while (P < n) P <<= 1;
    for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t l = i ^ (k - 1);
            if (l > i && l < n) cmpxchg(t, i, l);
        }
        __syncthreads();
        for (uint32_t j = k >> 2; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                const uint32_t l = i ^ j;
                if (l > i && l < n) cmpxchg(t, i, l);
            }
            __syncthreads();
        }
    }
}

// P2: one CTA per query read.  Sorts the read's run of keys, then rewrites the run in sorted order:
//   s_pos = position inside T, s_t = trial number, s_ref = T, s_rank = rank of the candidate's item among the read's items;
// nvalid[k] = candidates left after dropping self hits, nitems[k] = distinct T.
__global__ void __launch_bounds__(256)
pairs_sort_kernel(unsigned long long *__restrict__ key, const int32_t *__restrict__ cand_pos, const int32_t *__restrict__ cand_q,
                  const int64_t *__restrict__ qoff, int64_t nq_reads, int ntr, const int64_t *__restrict__ base,
                  int32_t *__restrict__ s_pos, int32_t *__restrict__ s_t, int32_t *__restrict__ s_ref, int32_t *__restrict__ s_rank,
                  uint32_t *__restrict__ nvalid, uint32_t *__restrict__ nitems)
{
    extern __shared__ unsigned long long sk[];
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t carry_s, valid_s;
    for (int64_t k = blockIdx.x; k < nq_reads; k += gridDim.x) {
        const int64_t c0 = qoff[k * ntr], c1 = qoff[(k + 1) * ntr];
        const uint32_t n = (uint32_t)(c1 - c0);
        if (threadIdx.x == 0) { carry_s = 0u; valid_s = 0u; }
        __syncthreads();
        if (n == 0) {
            if (threadIdx.x == 0) { nvalid[k] = 0u; nitems[k] = 0u; }
            continue;
        }
        unsigned long long *t = key + c0;
        const bool in_smem = n <= PAIRS_SORT_SMEM;
        if (in_smem) {
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) sk[i] = t[i];
            __syncthreads();
            t = sk;
        }
        bitonic_any(t, n);
        // sorted run -> candidate arrays; item heads ranked with a chunked block scan
        for (uint32_t cb = 0; cb < n; cb += blockDim.x) {
            const uint32_t i = cb + threadIdx.x;
            unsigned long long me = PAIRS_DROPPED;
            uint32_t head = 0u;
            if (i < n) {
                me = t[i];
                if (me != PAIRS_DROPPED) head = (i == 0 || (uint32_t)(t[i - 1] >> 32) != (uint32_t)(me >> 32)) ? 1u : 0u;
            }
            uint32_t x = head;
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += y;
            }
            if (lane == 31) wsum[wid] = x;
            __syncthreads();
            uint32_t off = carry_s, tot = 0u;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                if (w < wid) off += wsum[w];
                tot += wsum[w];
            }
            if (me != PAIRS_DROPPED) {
                const uint32_t T = (uint32_t)(me >> 32), rk = (uint32_t)me;
                const int64_t src = c0 + rk;
                s_pos[c0 + i] = (int32_t)((int64_t)cand_pos[src] - base[T]);
                s_t[c0 + i] = cand_q[src] - (int32_t)(k * ntr);
                s_ref[c0 + i] = (int32_t)T;
                s_rank[c0 + i] = (int32_t)(off + x - 1u); // inclusive count of heads up to here, minus one
                atomicAdd(&valid_s, 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) carry_s += tot;
            __syncthreads();
        }
        if (threadIdx.x == 0) { nvalid[k] = valid_s; nitems[k] = carry_s; }
        __syncthreads();
    }
}

// P3: item tables.  Items of query read k are numbered item_base[k] + rank, in ascending T.
__global__ void __launch_bounds__(256)
pairs_emit_kernel(const int64_t *__restrict__ qoff, int64_t nq_reads, int ntr, const int32_t *__restrict__ kept,
                  const uint32_t *__restrict__ nvalid, const int64_t *__restrict__ item_base, const int32_t *__restrict__ s_ref,
                  const int32_t *__restrict__ s_rank, int32_t *__restrict__ cand_item, int32_t *__restrict__ item_q,
                  int32_t *__restrict__ item_ref, int64_t *__restrict__ item_beg, int64_t *__restrict__ item_end)
{
    for (int64_t k = blockIdx.x; k < nq_reads; k += gridDim.x) {
        const int64_t c0 = qoff[k * ntr], c1 = qoff[(k + 1) * ntr];
        const uint32_t n = (uint32_t)(c1 - c0), nv = nvalid[k];
        const int64_t ib = item_base[k];
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const int64_t c = c0 + i;
            if (i >= nv) { cand_item[c] = -1; continue; } // slots freed by dropped self hits
            const int T = s_ref[c];
            const int64_t it = ib + s_rank[c];
            cand_item[c] = (int32_t)it;
            if (i == 0 || s_ref[c - 1] != T) { item_q[it] = kept[k]; item_ref[it] = T; item_beg[it] = c; }
            if (i + 1 == nv || s_ref[c + 1] != T) item_end[it] = c + 1;
        }
    }
}

// P4: an item whose candidates all fail the prefix filter is finished: it adds its try_align calls and their DP cells to the
// totals [0], [1]; the others are flagged for the aligner.
// For a live item, bound[it] = the longest min(|a|,|b|) over its surviving candidate views (spaced_seed.cpp:274-276, ref_seq.h:282-286):
// it fixes the widest band (max_dst = 1 + min*R, seq_aligner.h:94-102) the item can ask of the aligner, i.e. its band class.
__global__ void __launch_bounds__(256)
pairs_live_kernel(const int64_t *__restrict__ item_beg, const int64_t *__restrict__ item_end, int64_t nitems,
                  const int32_t *__restrict__ item_q, const int32_t *__restrict__ item_ref, const int32_t *__restrict__ len,
                  const int32_t *__restrict__ s_pos, const int32_t *__restrict__ s_t,
                  const uint8_t *__restrict__ survive, const int32_t *__restrict__ rej_cells, uint32_t *__restrict__ live,
                  int32_t *__restrict__ bound, unsigned long long *__restrict__ totals)
{
    const int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long nc = 0, cells = 0;
    if (it < nitems) {
        bool any = false;
        int bnd = 0;
        const int lenq = len[item_q[it]], lent = len[item_ref[it]];
        for (int64_t c = item_beg[it]; c < item_end[it]; ++c) {
            if (survive[c]) {
                any = true;
                const int t = s_t[c], j = t >> 1, pos = s_pos[c];
                const bool forward = (t & 1) == 0;
                const int a_len = forward ? lent - pos : pos + 16;
                const int b_len = lenq - j; // forward: len - pos with pos = j; backward: pos + 16 with pos = len-j-16
                bnd = max(bnd, min(a_len, b_len));
            }
            cells += (unsigned long long)rej_cells[c];
            ++nc;
        }
        live[it] = any ? 1u : 0u;
        bound[it] = bnd;
        if (any) { nc = 0; cells = 0; }
    }
    for (int d = 16; d; d >>= 1) {
        nc += __shfl_xor_sync(0xffffffffu, nc, d);
        cells += __shfl_xor_sync(0xffffffffu, cells, d);
    }
    if ((threadIdx.x & 31) == 0 && nc) { atomicAdd(totals, nc); atomicAdd(totals + 1, cells); }
}

__global__ void __launch_bounds__(256)
pairs_compact_kernel(const uint32_t *__restrict__ live, const int64_t *__restrict__ live_off, int64_t nitems, const int32_t *__restrict__ item_q,
                     const int32_t *__restrict__ item_ref, const int64_t *__restrict__ item_beg, const int64_t *__restrict__ item_end,
                     const int32_t *__restrict__ bound, int32_t *__restrict__ l_q, int32_t *__restrict__ l_ref, int64_t *__restrict__ l_beg,
                     int64_t *__restrict__ l_end, int32_t *__restrict__ l_bound)
{
    const int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= nitems || !live[it]) return;
    const int64_t o = live_off[it];
    l_q[o] = item_q[it]; l_ref[o] = item_ref[it]; l_beg[o] = item_beg[it]; l_end[o] = item_end[it]; l_bound[o] = bound[it];
}

// P5: totals over the aligned items ([0] try_align calls, [1] cells) and the found flags
__global__ void __launch_bounds__(256)
pairs_found_flag_kernel(const pb_pair_rec *__restrict__ recs, int64_t n, uint32_t *__restrict__ flag, unsigned long long *__restrict__ totals)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long nc = 0, cells = 0;
    if (i < n) {
        flag[i] = recs[i].found ? 1u : 0u;
        nc = (unsigned long long)recs[i].ncand;
        cells = (unsigned long long)recs[i].cells;
    }
    for (int d = 16; d; d >>= 1) {
        nc += __shfl_xor_sync(0xffffffffu, nc, d);
        cells += __shfl_xor_sync(0xffffffffu, cells, d);
    }
    if ((threadIdx.x & 31) == 0 && nc) { atomicAdd(totals, nc); atomicAdd(totals + 1, cells); }
}

__global__ void __launch_bounds__(256)
pairs_found_scatter_kernel(const pb_pair_rec *__restrict__ recs, int64_t n, const uint32_t *__restrict__ flag, const int64_t *__restrict__ off,
                           pb_pair_rec *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) out[off[i]] = recs[i];
}

static inline unsigned grid_for(int64_t n) { return (unsigned)std::max<int64_t>(1, (n + 255) / 256); }

extern "C" int pb_overlap_all_run(pb_ctx *ctx, const pb_index *ix, const pb_seqset *set, int64_t q_first, int64_t q_count,
                                  const pb_overlap_params *prm, pb_pairs_job **out)
{
    static_assert(sizeof(pb_pair_rec) == sizeof(pb_locate_rec) && sizeof(pb_pair_rec) == 56, "record layouts must coincide");
    if (!ctx || !ix || !set || !prm || !out || q_first < 0 || q_count < 0 || q_first + q_count > set->n)
        return pb_fail(ctx, PB_ERR_ARG, "pb_overlap_all_run: bad argument");
    if (!ix->whole_set || ix->nbuckets == 0) return pb_fail(ctx, PB_ERR_ARG, "pb_overlap_all_run needs an index made by pb_index_build_set");
    if (prm->max_trial < 1 || prm->max_trial > 2048) return pb_fail(ctx, PB_ERR_ARG, "max_trial %d out of range", prm->max_trial);
    if (prm->want_ops) return pb_fail(ctx, PB_ERR_ARG, "pb_overlap_all_run returns records only; re-align a pair with pb_align_batch for its transcript");
    if (q_count * (int64_t)prm->max_trial * 2 > INT32_MAX) return pb_fail(ctx, PB_ERR_DOMAIN, "too many query reads in one call");
    *out = nullptr;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int64_t i = 0; i < set->n; ++i)
        if (set->flags[i] & PB_FLAG_IRREGULAR)
            return pb_fail(ctx, PB_ERR_ALPHABET, "sequence %lld holds bytes outside {A,C,G,T}; all-vs-all runs on packed reads (.bin records)", (long long)i);
    for (int64_t i = q_first; i < q_first + q_count; ++i)
        if (set->len[i] < prm->max_trial + 16)
            return pb_fail(ctx, PB_ERR_ARG, "read %lld is shorter than max_trial+16 (the reference only keeps reads longer than 500, "
                           "spaced_seed.cpp:336)", (long long)i);
    pb_timer_reset(ctx);
    pb_timer_begin(ctx, PB_T_TOTAL);
    pb_pairs_job *job = new pb_pairs_job();
    job->ctx = ctx;
    pb_seqset *rev = nullptr;
    int r = PB_OK;
    const int ntr = 2 * prm->max_trial;
    const int64_t nq = q_count;
    DevBuf d_kept, d_key, d_spos, d_st, d_sref, d_srank, d_nvalid, d_nitems, d_ibase, tmp, d_cand_item, d_iq, d_iref, d_ibeg, d_iend;
    DevBuf d_survive, d_rej, d_live, d_loff, d_lq, d_lref, d_lbeg, d_lend, d_tot, d_stats, d_flag, d_foff, d_bound, d_lbound;
    ProbeOut po;
    unsigned long long tot[2] = {0, 0}, tot2[2] = {0, 0}, k3[2] = {0, 0};
#define STEP(x) do { if (r == PB_OK) r = (x); } while (0)
#define CHECK_LAUNCH() do { if (r == PB_OK) { ctx->launches++; cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) r = pb_fail(ctx, PB_ERR_CUDA, "launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); } } while (0)
    if (nq > 0) {
        std::vector<int32_t> kept((size_t)nq);
        for (int64_t k = 0; k < nq; ++k) kept[k] = (int32_t)(q_first + k);
        STEP(d_kept.alloc(ctx, (size_t)nq * 4));
        STEP(pb_h2d(ctx, d_kept.p, kept.data(), (size_t)nq * 4));
        STEP(pb_sync(ctx)); // `kept` leaves scope
        STEP(pb_overlap_seed_probe(ctx, ix, set, d_kept.as<int32_t>(), nq, prm->max_trial, prm->min_overlap, prm->seed_at_quirk, &po));
    }
    job->ncand = po.ncand;
    const int64_t nc = po.ncand;
    if (r == PB_OK && nc > 0) {
        // ---- regroup the hits per (reference, read) pair
        pb_timer_begin(ctx, PB_T_INDEX);
        STEP(d_key.alloc(ctx, (size_t)nc * 8));
        STEP(d_spos.alloc(ctx, (size_t)nc * 4));
        STEP(d_st.alloc(ctx, (size_t)nc * 4));
        STEP(d_sref.alloc(ctx, (size_t)nc * 4));
        STEP(d_srank.alloc(ctx, (size_t)nc * 4));
        STEP(d_cand_item.alloc(ctx, (size_t)nc * 4));
        STEP(d_nvalid.alloc(ctx, (size_t)nq * 4));
        STEP(d_nitems.alloc(ctx, (size_t)nq * 4));
        STEP(d_ibase.alloc(ctx, (size_t)(nq + 2) * 8));
        if (r == PB_OK) {
            pairs_owner_kernel<<<grid_for(nc), 256, 0, ctx->stream>>>(po.d_cand_pos.as<int32_t>(), po.d_cand_q.as<int32_t>(), nc, po.d_qoff.as<int64_t>(),
                                                                       d_kept.as<int32_t>(), ntr, set->d_base.as<int64_t>(), set->n,
                                                                       d_key.as<unsigned long long>());
            CHECK_LAUNCH();
        }
        if (r == PB_OK) {
            const size_t smem = (size_t)PAIRS_SORT_SMEM * 8;
            cudaError_t e = cudaFuncSetAttribute(pairs_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) r = pb_fail(ctx, PB_ERR_CUDA, "pairs_sort_kernel attribute: %s", cudaGetErrorString(e));
            const unsigned g = (unsigned)std::min<int64_t>(nq, (int64_t)ctx->sm_count * 3);
            if (r == PB_OK) {
                pairs_sort_kernel<<<g, 256, smem, ctx->stream>>>(d_key.as<unsigned long long>(), po.d_cand_pos.as<int32_t>(), po.d_cand_q.as<int32_t>(),
                                                                 po.d_qoff.as<int64_t>(), nq, ntr, set->d_base.as<int64_t>(), d_spos.as<int32_t>(),
                                                                 d_st.as<int32_t>(), d_sref.as<int32_t>(), d_srank.as<int32_t>(),
                                                                 d_nvalid.as<uint32_t>(), d_nitems.as<uint32_t>());
                CHECK_LAUNCH();
            }
        }
        STEP(pb_scan_i64(ctx, d_nitems.as<uint32_t>(), d_ibase.as<int64_t>(), nq, tmp));
        int64_t nitems = 0;
        STEP(pb_d2h(ctx, &nitems, d_ibase.as<int64_t>() + nq, 8));
        STEP(pb_sync(ctx));
        job->nitems = nitems;
        d_key.release();
        if (r == PB_OK && nitems > 0) {
            STEP(d_iq.alloc(ctx, (size_t)nitems * 4));
            STEP(d_iref.alloc(ctx, (size_t)nitems * 4));
            STEP(d_ibeg.alloc(ctx, (size_t)nitems * 8));
            STEP(d_iend.alloc(ctx, (size_t)nitems * 8));
            if (r == PB_OK) {
                const unsigned g = (unsigned)std::min<int64_t>(nq, (int64_t)ctx->sm_count * 8);
                pairs_emit_kernel<<<g, 256, 0, ctx->stream>>>(po.d_qoff.as<int64_t>(), nq, ntr, d_kept.as<int32_t>(), d_nvalid.as<uint32_t>(),
                                                              d_ibase.as<int64_t>(), d_sref.as<int32_t>(), d_srank.as<int32_t>(),
                                                              d_cand_item.as<int32_t>(), d_iq.as<int32_t>(), d_iref.as<int32_t>(),
                                                              d_ibeg.as<int64_t>(), d_iend.as<int64_t>());
                CHECK_LAUNCH();
            }
            pb_timer_end(ctx, PB_T_INDEX);
            // ---- verify: prefix filter over every candidate, then the banded aligner over the items that still have one
            STEP(pb_seqset_reversed(ctx, set, &rev));
            LocateView lv;
            lv.d_kept = d_iq.as<int32_t>();
            lv.d_qoff = nullptr;
            lv.d_cand_pos = d_spos.as<int32_t>();
            lv.d_cand_q = d_st.as<int32_t>();
            lv.ntrial = ntr;
            lv.ref_base = 0;
            lv.ref_len = 0;
            lv.mode = PB_MODE_OVERLAP;
            lv.min_overlap = prm->min_overlap;
            lv.d_item_ref = d_iref.as<int32_t>();
            lv.d_item_beg = d_ibeg.as<int64_t>();
            lv.d_item_end = d_iend.as<int64_t>();
            lv.d_cand_item = d_cand_item.as<int32_t>();
            SeqSets ss = {set, set, rev, rev};
            STEP(d_survive.alloc(ctx, (size_t)nc));
            STEP(d_rej.alloc(ctx, (size_t)nc * 4));
            pb_timer_begin(ctx, PB_T_PREFILTER);
            STEP(pb_prefilter(ctx, ss, lv, nc, prm->R, prm->maxn, prm->maxm, d_survive.as<uint8_t>(), d_rej.as<int32_t>()));
            STEP(d_live.alloc(ctx, (size_t)nitems * 4));
            STEP(d_bound.alloc(ctx, (size_t)nitems * 4));
            STEP(d_loff.alloc(ctx, (size_t)(nitems + 2) * 8));
            STEP(d_tot.alloc_zero(ctx, 32));
            if (r == PB_OK) {
                pairs_live_kernel<<<grid_for(nitems), 256, 0, ctx->stream>>>(d_ibeg.as<int64_t>(), d_iend.as<int64_t>(), nitems, d_iq.as<int32_t>(),
                                                                          d_iref.as<int32_t>(), set->d_len.as<int32_t>(), d_spos.as<int32_t>(),
                                                                          d_st.as<int32_t>(), d_survive.as<uint8_t>(), d_rej.as<int32_t>(),
                                                                          d_live.as<uint32_t>(), d_bound.as<int32_t>(), d_tot.as<unsigned long long>());
                CHECK_LAUNCH();
            }
            STEP(pb_scan_i64(ctx, d_live.as<uint32_t>(), d_loff.as<int64_t>(), nitems, tmp));
            int64_t nlive = 0;
            STEP(pb_d2h(ctx, &nlive, d_loff.as<int64_t>() + nitems, 8));
            STEP(pb_d2h(ctx, tot, d_tot.p, 16));
            STEP(pb_sync(ctx));
            pb_timer_end(ctx, PB_T_PREFILTER);
            job->nlive = nlive;
            if (r == PB_OK && nlive > INT32_MAX) r = pb_fail(ctx, PB_ERR_DOMAIN, "%lld pairs to align in one call; split the query range", (long long)nlive);
            if (r == PB_OK && nlive > 0) {
                STEP(d_lq.alloc(ctx, (size_t)nlive * 4));
                STEP(d_lref.alloc(ctx, (size_t)nlive * 4));
                STEP(d_lbeg.alloc(ctx, (size_t)nlive * 8));
                STEP(d_lend.alloc(ctx, (size_t)nlive * 8));
                STEP(d_lbound.alloc(ctx, (size_t)nlive * 4));
                if (r == PB_OK) {
                    pairs_compact_kernel<<<grid_for(nitems), 256, 0, ctx->stream>>>(d_live.as<uint32_t>(), d_loff.as<int64_t>(), nitems, d_iq.as<int32_t>(),
                                                                                 d_iref.as<int32_t>(), d_ibeg.as<int64_t>(), d_iend.as<int64_t>(),
                                                                                 d_bound.as<int32_t>(), d_lq.as<int32_t>(), d_lref.as<int32_t>(),
                                                                                 d_lbeg.as<int64_t>(), d_lend.as<int64_t>(), d_lbound.as<int32_t>());
                    CHECK_LAUNCH();
                }
                // the host plans the band classes from each pair's bound (the aligner takes min(|a|,|b|) as its `read length`)
                std::vector<int32_t> bound((size_t)nlive);
                STEP(pb_d2h(ctx, bound.data(), d_lbound.p, (size_t)nlive * 4));
                STEP(pb_sync(ctx));
                std::vector<uint8_t> irr((size_t)nlive, 0);
                STEP(job->d_recs.alloc_zero(ctx, (size_t)nlive * sizeof(pb_pair_rec)));
                STEP(d_stats.alloc_zero(ctx, 64));
                lv.d_kept = d_lq.as<int32_t>();
                lv.d_item_ref = d_lref.as<int32_t>();
                lv.d_item_beg = d_lbeg.as<int64_t>();
                lv.d_item_end = d_lend.as<int64_t>();
                lv.d_cand_item = nullptr;
                // PB_T_ALIGN starts inside the aligner, after its host-side planning: the stage is the kernels' time
                STEP(pb_align_locate(ctx, ss, lv, nlive, bound, irr, prm->R, prm->maxn, prm->maxm, d_survive.as<uint8_t>(), d_rej.as<int32_t>(),
                                     reinterpret_cast<pb_locate_rec *>(job->d_recs.p), nullptr, nullptr, d_stats.as<unsigned long long>()));
                pb_timer_end(ctx, PB_T_ALIGN);
                // ---- found records, compacted in (read, reference) order
                STEP(d_flag.alloc(ctx, (size_t)nlive * 4));
                STEP(d_foff.alloc(ctx, (size_t)(nlive + 2) * 8));
                STEP(d_tot.alloc_zero(ctx, 32));
                if (r == PB_OK) {
                    pairs_found_flag_kernel<<<grid_for(nlive), 256, 0, ctx->stream>>>(job->d_recs.as<pb_pair_rec>(), nlive, d_flag.as<uint32_t>(),
                                                                                   d_tot.as<unsigned long long>());
                    CHECK_LAUNCH();
                }
                STEP(pb_scan_i64(ctx, d_flag.as<uint32_t>(), d_foff.as<int64_t>(), nlive, tmp));
                int64_t nfound = 0;
                STEP(pb_d2h(ctx, &nfound, d_foff.as<int64_t>() + nlive, 8));
                STEP(pb_d2h(ctx, tot2, d_tot.p, 16));
                STEP(pb_d2h(ctx, k3, d_stats.p, 16));
                STEP(pb_sync(ctx));
                job->nfound = nfound;
                STEP(job->d_found.alloc(ctx, (size_t)std::max<int64_t>(nfound, 1) * sizeof(pb_pair_rec)));
                if (r == PB_OK && nfound > 0) {
                    pairs_found_scatter_kernel<<<grid_for(nlive), 256, 0, ctx->stream>>>(job->d_recs.as<pb_pair_rec>(), nlive, d_flag.as<uint32_t>(),
                                                                                      d_foff.as<int64_t>(), job->d_found.as<pb_pair_rec>());
                    CHECK_LAUNCH();
                }
            }
        } else {
            pb_timer_end(ctx, PB_T_INDEX);
        }
    }
#undef STEP
#undef CHECK_LAUNCH
    pb_timer_end(ctx, PB_T_TOTAL);
    int rs = pb_sync(ctx);
    if (r == PB_OK) r = rs;
    if (r == PB_OK) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) r = pb_fail(ctx, PB_ERR_CUDA, "all-vs-all kernels failed: %s", cudaGetErrorString(e));
    }
    pb_timer_collect(ctx);
    if (rev) pb_seqset_free(rev);
    job->tot_ncand = (int64_t)(tot[0] + tot2[0]);
    job->tot_cells = (int64_t)(tot[1] + tot2[1]);
    job->k3_cells = (int64_t)k3[0];
    job->k3_aligns = (int64_t)k3[1];
    if (r != PB_OK) { delete job; return r; }
    *out = job;
    return PB_OK;
}

extern "C" int pb_pairs_job_stats(const pb_pairs_job *job, int64_t *out)
{
    if (!job || !out) return PB_ERR_ARG;
    out[0] = job->ncand; out[1] = job->nitems; out[2] = job->nlive; out[3] = job->nfound;
    out[4] = job->tot_ncand; out[5] = job->tot_cells; out[6] = job->k3_aligns; out[7] = job->k3_cells;
    return PB_OK;
}

extern "C" int pb_pairs_job_fetch(pb_ctx *ctx, const pb_pairs_job *job, int found_only, pb_pair_rec *recs, int64_t cap, int64_t *n)
{
    if (!ctx || !job || cap < 0 || (cap && !recs)) return pb_fail(ctx, PB_ERR_ARG, "pb_pairs_job_fetch: bad argument");
    const int64_t have = found_only ? job->nfound : job->nlive;
    if (n) *n = have;
    if (!recs) return PB_OK;
    if (cap < have) return pb_fail(ctx, PB_ERR_ARG, "pb_pairs_job_fetch: room for %lld records, %lld to return", (long long)cap, (long long)have);
    if (have == 0) return PB_OK;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    pb_timer_begin(ctx, PB_T_D2H);
    PB_TRY(pb_d2h(ctx, recs, found_only ? job->d_found.p : job->d_recs.p, (size_t)have * sizeof(pb_pair_rec)));
    pb_timer_end(ctx, PB_T_D2H);
    PB_TRY(pb_sync(ctx));
    pb_timer_collect(ctx);
    return PB_OK;
}

extern "C" void pb_pairs_job_free(pb_pairs_job *job)
{
    if (!job) return;
    cudaSetDevice(job->ctx->device);
    delete job;
}
