// pb_seed.cu -- L1: spaced-seed extraction (K1), seed index build, probe/gather (K2).
//
// Replaces dna_seq::encode called in a loop (locator.cpp:62-66, ref_seq.h:291-311), the
// __gnu_cxx::hash_map<unsigned, std::list<int>> seed map (common.h:54) and hash_table::find
// (locator.cpp:76, spaced_seed.cpp:265).
//
// Index layout in HBM: direct-address CSR.  bucket(key) compresses the key's bits under the spaced
// mask (weight 11/12 seeds -> 2^22 / 2^24 buckets, injective, so no key compare on probe); masks with
// more than 24 care bits or more than PB_MAX_RUNS runs fall back to a multiplicative hash plus a stored
// key per entry.  Within a bucket, entries are in the reference's list order (insertion order), which
// is what makes "first successful candidate" (locator.cpp:79-88) reproducible.
#include <stdlib.h>

#include <algorithm>

#include "pb_internal.cuh"

// ---------------------------------------------------------------------------------------------
// K1: bulk seed extraction.  HBM-bound: 0.25 B read + 4 B written per position.
// A warp turns 32 coalesced packed words (512 bases) into 512 keys; every lane stores uint4
// (4 consecutive keys), so each store instruction writes one contiguous 512-byte run.
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// seed word at base offset o (0..15) of the big-endian-bit stream (h0:h1): equals dna_seq::encode(text+p)
__device__ __forceinline__ uint32_t seed_from_be(uint32_t h0, uint32_t h1, int o)
{
    return bswap32(__funnelshift_l(h1, h0, 2 * o));
}

__global__ void __launch_bounds__(256)
seed_bulk_kernel(const uint32_t *__restrict__ pw, int64_t first_word, int64_t last_word, int64_t count, uint32_t mask,
                 uint32_t *__restrict__ keys)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nchunks = (count + 511) >> 9;
    for (int64_t c = warp0; c < nchunks; c += nwarps) {
        const int64_t w0 = first_word + c * 32;
        const uint32_t hw = bswap32(__ldg(pw + min(w0 + lane, last_word)));
        const uint32_t hx = bswap32(__ldg(pw + min(w0 + 32, last_word)));
        const int o = 4 * (lane & 3);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int src = 8 * r + (lane >> 2);
            const uint32_t wa = __shfl_sync(0xffffffffu, hw, src);
            uint32_t wb = __shfl_sync(0xffffffffu, hw, (src + 1) & 31);
            if (src == 31) wb = hx;
            uint4 k;
            k.x = seed_from_be(wa, wb, o) & mask;
            k.y = seed_from_be(wa, wb, o + 1) & mask;
            k.z = seed_from_be(wa, wb, o + 2) & mask;
            k.w = seed_from_be(wa, wb, o + 3) & mask;
            const int64_t idx = (c << 9) + 128 * r + 4 * lane;
            if (idx + 3 < count) {
                *reinterpret_cast<uint4 *>(keys + idx) = k;
            } else {
                if (idx < count) keys[idx] = k.x;
                if (idx + 1 < count) keys[idx + 1] = k.y;
                if (idx + 2 < count) keys[idx + 2] = k.z;
            }
        }
    }
}

// keys for positions [first_base, first_base+count) of the padded line (first_base % 16 == 0)
int pb_seed_bulk_device(pb_ctx *ctx, const pb_seqset *s, int64_t first_base, int64_t count, uint32_t mask, uint32_t *d_keys)
{
    if (count <= 0) return PB_OK;
    if (first_base & 15) return pb_fail(ctx, PB_ERR_ARG, "seed extraction must start on a 16-base boundary");
    const int64_t nchunks = (count + 511) >> 9;
    int64_t blocks = std::min<int64_t>((nchunks + 7) / 8, (int64_t)ctx->sm_count * 8);
    seed_bulk_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(s->d_packed.as<uint32_t>(), first_base >> 4,
                                                                2 * s->nwords() + 7, count, mask, d_keys);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}

extern "C" int pb_seed_extract(pb_ctx *ctx, const pb_seqset *s, int64_t i, uint32_t mask, uint32_t *keys)
{
    if (!ctx || !s || i < 0 || i >= s->n || !keys) return pb_fail(ctx, PB_ERR_ARG, "pb_seed_extract: bad argument");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t L = s->len[i];
    if (!L) return PB_OK;
    DevBuf d;
    PB_TRY(d.alloc(ctx, (size_t)L * 4 + 64));
    pb_timer_begin(ctx, PB_T_SEED);
    PB_TRY(pb_seed_bulk_device(ctx, s, s->base[i], L, mask, d.as<uint32_t>()));
    pb_timer_end(ctx, PB_T_SEED);
    PB_TRY(pb_d2h(ctx, keys, d.p, (size_t)L * 4));
    PB_TRY(pb_sync(ctx));
    pb_timer_collect(ctx);
    return PB_OK;
}

extern "C" int pb_seed_extract_all_device(pb_ctx *ctx, const pb_seqset *s, uint32_t mask, int64_t *nkeys, float *kernel_ms)
{
    if (!ctx || !s) return pb_fail(ctx, PB_ERR_ARG, "pb_seed_extract_all_device: bad argument");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t count = s->base[s->n];
    DevBuf d;
    PB_TRY(d.alloc(ctx, (size_t)count * 4 + 64));
    pb_timer_reset(ctx);
    pb_timer_begin(ctx, PB_T_SEED);
    PB_TRY(pb_seed_bulk_device(ctx, s, 0, count, mask, d.as<uint32_t>()));
    pb_timer_end(ctx, PB_T_SEED);
    PB_TRY(pb_sync(ctx));
    pb_timer_collect(ctx);
    if (nkeys) *nkeys = count;
    if (kernel_ms) *kernel_ms = ctx->times[PB_T_SEED];
    return PB_OK;
}

// ---------------------------------------------------------------------------------------------
// exclusive scan (three phases; block = 256 threads x 16 items)
// ---------------------------------------------------------------------------------------------

#define SCAN_ITEMS 16
#define SCAN_BLOCK 256
#define SCAN_TILE (SCAN_ITEMS * SCAN_BLOCK)

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long *total)
{
    __shared__ unsigned long long wsum[SCAN_BLOCK / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    unsigned long long off = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < SCAN_BLOCK / 32; ++k) {
        unsigned long long t = wsum[k];
        if (k < wid) off += t;
        tot += t;
    }
    __syncthreads();
    *total = tot;
    return off + x - v;
}

__global__ void __launch_bounds__(SCAN_BLOCK) scan_sums_kernel(const uint32_t *__restrict__ in, int64_t n, unsigned long long *bsum)
{
    const int64_t t0 = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (t0 + k < n) s += in[t0 + k];
    unsigned long long tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void scan_bsum_kernel(unsigned long long *bsum, int64_t nb)
{ // one warp, sequential over chunks of 32
    const int lane = threadIdx.x;
    unsigned long long carry = 0;
    for (int64_t b0 = 0; b0 < nb; b0 += 32) {
        unsigned long long v = b0 + lane < nb ? bsum[b0 + lane] : 0, x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (b0 + lane < nb) bsum[b0 + lane] = carry + x - v;
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) bsum[nb] = carry;
}

template <class OutT>
__global__ void __launch_bounds__(SCAN_BLOCK)
scan_write_kernel(const uint32_t *__restrict__ in, int64_t n, const unsigned long long *__restrict__ bsum, int64_t nb,
                  OutT *__restrict__ out)
{
    const int64_t t0 = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = t0 + k < n ? in[t0 + k] : 0u;
        s += v[k];
    }
    unsigned long long tot;
    unsigned long long off = block_exclusive_scan(s, &tot) + bsum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (t0 + k < n) out[t0 + k] = (OutT)off;
        off += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = (OutT)bsum[nb];
}

template <class OutT> static int scan_impl(pb_ctx *ctx, const uint32_t *d_in, OutT *d_out, int64_t n, DevBuf &tmp)
{
    const int64_t nb = std::max<int64_t>(1, (n + SCAN_TILE - 1) / SCAN_TILE);
    PB_TRY(tmp.alloc(ctx, (size_t)(nb + 1) * sizeof(unsigned long long)));
    scan_sums_kernel<<<(unsigned)nb, SCAN_BLOCK, 0, ctx->stream>>>(d_in, n, tmp.as<unsigned long long>());
    PB_LAUNCH_CHECK(ctx);
    scan_bsum_kernel<<<1, 32, 0, ctx->stream>>>(tmp.as<unsigned long long>(), nb);
    PB_LAUNCH_CHECK(ctx);
    scan_write_kernel<OutT><<<(unsigned)nb, SCAN_BLOCK, 0, ctx->stream>>>(d_in, n, tmp.as<unsigned long long>(), nb, d_out);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}

int pb_scan_u32(pb_ctx *ctx, const uint32_t *d_in, uint32_t *d_out, int64_t n, DevBuf &tmp) { return scan_impl<uint32_t>(ctx, d_in, d_out, n, tmp); }
int pb_scan_i64(pb_ctx *ctx, const uint32_t *d_in, int64_t *d_out, int64_t n, DevBuf &tmp) { return scan_impl<int64_t>(ctx, d_in, d_out, n, tmp); }

// ---------------------------------------------------------------------------------------------
// index build
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t bucket_of(const BucketFn &fn, uint32_t key)
{
    if (fn.exact) {
        uint32_t b = 0;
#pragma unroll 4
        for (int r = 0; r < fn.nruns; ++r) b |= (key & fn.run_mask[r]) >> fn.run_shift[r];
        return b;
    }
    return (key * 0x9E3779B1u) >> (32 - fn.bits);
}

static BucketFn make_bucket_fn(uint32_t mask)
{
    BucketFn fn;
    memset(&fn, 0, sizeof fn);
    int pop = __builtin_popcount(mask), nruns = 0, outbit = 0;
    bool ok = pop <= 24;
    uint32_t m = mask;
    while (m && ok) {
        int lo = __builtin_ctz(m);
        uint32_t run = m >> lo;
        int rl = __builtin_ctz(~run);
        uint32_t rm = (rl >= 32 ? 0xFFFFFFFFu : ((1u << rl) - 1u)) << lo;
        if (nruns == PB_MAX_RUNS) { ok = false; break; }
        fn.run_mask[nruns] = rm;
        fn.run_shift[nruns] = (uint8_t)(lo - outbit);
        ++nruns;
        outbit += rl;
        m &= ~rm;
    }
    if (ok) {
        fn.exact = 1;
        fn.nruns = nruns;
        fn.bits = pop;
    } else {
        fn.exact = 0;
        fn.nruns = 0;
        fn.bits = 24;
    }
    return fn;
}

struct EntryMap { // scan order e -> reference position (ref_seq.h:291-311 for REFSEQ; identity for LOCATOR)
    int64_t nhead, len;
    // explicit (key, position) pairs in insertion order (pb_index_build_pairs: the map filled through hash_table::operator[])
    const uint32_t *xkeys = nullptr;
    const int32_t *xpos = nullptr;
    __host__ __device__ int64_t pos(int64_t e) const { return xpos ? (int64_t)xpos[e] : (e < nhead ? e : len - 16 - (e - nhead)); }
    __device__ uint32_t key(const uint32_t *__restrict__ keys_all, int64_t e) const { return xkeys ? xkeys[e] : keys_all[pos(e)]; }
};

__global__ void index_count_kernel(const uint32_t *__restrict__ keys_all, EntryMap em, int64_t nscan, BucketFn fn,
                                   uint32_t *__restrict__ count)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nscan) return;
    const uint32_t key = em.key(keys_all, e);
    if (key) atomicAdd(&count[bucket_of(fn, key)], 1u);
}

__global__ void index_scatter_kernel(const uint32_t *__restrict__ keys_all, EntryMap em, int64_t nscan, BucketFn fn,
                                     uint32_t *__restrict__ cursor, int32_t *__restrict__ ent)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nscan) return;
    const uint32_t key = em.key(keys_all, e);
    if (!key) return;
    const uint32_t slot = atomicAdd(&cursor[bucket_of(fn, key)], 1u);
    ent[slot] = (int32_t)e;
}

// Whole-set variants (all-vs-all, pb_index_build_set): every sequence is indexed the way ref_seq::get_seedmap indexes a
// reference no longer than MAX_READ_LEN+16 (ref_seq.h:291-300: head positions 0..len-17, no tail pass); an entry is the
// position on the set's padded line, so ascending entries = ascending (sequence, position) = each sequence's list order.
// One CTA walks one sequence at a time: no owner search, coalesced key loads.
__global__ void __launch_bounds__(256)
set_index_count_kernel(const uint32_t *__restrict__ keys_all, const int64_t *__restrict__ base, const int32_t *__restrict__ len,
                       int64_t nseq, BucketFn fn, uint32_t *__restrict__ count)
{
    for (int64_t i = blockIdx.x; i < nseq; i += gridDim.x) {
        const int64_t b = base[i];
        const int nhead = len[i] - 16;
        for (int o = threadIdx.x; o < nhead; o += blockDim.x) {
            const uint32_t key = keys_all[b + o];
            if (key) atomicAdd(&count[bucket_of(fn, key)], 1u);
        }
    }
}

__global__ void __launch_bounds__(256)
set_index_scatter_kernel(const uint32_t *__restrict__ keys_all, const int64_t *__restrict__ base, const int32_t *__restrict__ len,
                         int64_t nseq, BucketFn fn, uint32_t *__restrict__ cursor, int32_t *__restrict__ ent)
{
    for (int64_t i = blockIdx.x; i < nseq; i += gridDim.x) {
        const int64_t b = base[i];
        const int nhead = len[i] - 16;
        for (int o = threadIdx.x; o < nhead; o += blockDim.x) {
            const uint32_t key = keys_all[b + o];
            if (!key) continue;
            const uint32_t slot = atomicAdd(&cursor[bucket_of(fn, key)], 1u);
            ent[slot] = (int32_t)(b + o);
        }
    }
}

#define SMALL_BUCKET 48

// restore insertion order inside each bucket (the atomic scatter is unordered)
__global__ void index_sort_small_kernel(const uint32_t *__restrict__ start, int64_t nbuckets, int32_t *__restrict__ ent,
                                        uint32_t *__restrict__ big_list, uint32_t *__restrict__ nbig)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbuckets) return;
    const uint32_t s = start[b], c = start[b + 1] - s;
    if (c < 2) return;
    if (c > SMALL_BUCKET) {
        big_list[atomicAdd(nbig, 1u)] = (uint32_t)b;
        return;
    }
    int32_t *p = ent + s;
    for (uint32_t i = 1; i < c; ++i) {
        int32_t v = p[i];
        uint32_t j = i;
        while (j > 0 && p[j - 1] > v) { p[j] = p[j - 1]; --j; }
        p[j] = v;
    }
}

// one CTA per big bucket: bitonic sort of a power-of-two padded copy in global scratch
__global__ void __launch_bounds__(256)
index_sort_big_kernel(const uint32_t *__restrict__ start, const uint32_t *__restrict__ big_list,
                      const unsigned long long *__restrict__ tmp_off, int32_t *__restrict__ ent, int32_t *__restrict__ tmp)
{
    const uint32_t b = big_list[blockIdx.x];
    const uint32_t s = start[b], c = start[b + 1] - s;
    uint32_t P = 1;
    while (P < c) P <<= 1;
    int32_t *t = tmp + tmp_off[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) t[i] = i < c ? ent[s + i] : INT32_MAX;
    __syncthreads();
    for (uint32_t k = 2; k <= P; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
                uint32_t l = i ^ j;
                if (l > i) {
                    int32_t x = t[i], y = t[l];
                    bool up = (i & k) == 0;
                    if ((x > y) == up) { t[i] = y; t[l] = x; }
                }
            }
            __syncthreads();
        }
    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) ent[s + i] = t[i];
}

__global__ void big_sizes_kernel(const uint32_t *__restrict__ start, const uint32_t *__restrict__ big_list, uint32_t nbig,
                                 uint32_t *__restrict__ sizes)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbig) return;
    const uint32_t b = big_list[i], c = start[b + 1] - start[b];
    uint32_t P = 1;
    while (P < c) P <<= 1;
    sizes[i] = P;
}

__global__ void index_finalize_kernel(const uint32_t *__restrict__ keys_all, EntryMap em, int64_t nentries,
                                      int32_t *__restrict__ ent, uint32_t *__restrict__ ekey)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nentries) return;
    const int64_t e = ent[i];
    if (ekey) ekey[i] = em.key(keys_all, e);
    ent[i] = (int32_t)em.pos(e);
}

__global__ void index_pack_kernel(const uint32_t *__restrict__ start, int64_t nbuckets, int shift, uint32_t esc, uint32_t *__restrict__ pk)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbuckets) return;
    const uint32_t s = start[b], c = start[b + 1] - s;
    pk[b] = s | (min(c, esc) << shift);
}

__global__ void index_nkeys_kernel(const uint32_t *__restrict__ start, int64_t nbuckets, const uint32_t *__restrict__ ekey,
                                   unsigned long long *__restrict__ nkeys)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mine = 0;
    if (b < nbuckets) {
        const uint32_t s = start[b], c = start[b + 1] - s;
        if (!ekey) mine = c ? 1 : 0;
        else
            for (uint32_t i = 0; i < c; ++i) { // distinct keys in the bucket (hashed buckets are tiny)
                bool first = true;
                for (uint32_t j = 0; j < i && first; ++j) first = ekey[s + j] != ekey[s + i];
                mine += first;
            }
    }
    for (int d = 16; d; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(nkeys, mine);
}

// seq >= 0: one sequence under `policy`; seq < 0: every sequence of the set (get_seedmap's head pass per sequence)
// ref == NULL: explicit pairs (d_xkeys / d_xpos, npairs of them, insertion order)
static int index_build_impl(pb_ctx *ctx, const pb_seqset *ref, int64_t seq, uint32_t mask, int policy, pb_index **out,
                            const uint32_t *d_xkeys = nullptr, const int32_t *d_xpos = nullptr, int64_t npairs = 0)
{
    *out = nullptr;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    pb_timer_reset(ctx);
    const bool pairs = ref == nullptr;
    const bool whole_set = !pairs && seq < 0;
    const int64_t len = pairs ? npairs : (whole_set ? ref->base[ref->n] : ref->len[seq]);
    EntryMap em;
    em.len = len;
    int64_t nscan;
    if (pairs) {
        em.nhead = len;
        em.xkeys = d_xkeys;
        em.xpos = d_xpos;
        nscan = len;
    } else if (whole_set) {
        em.nhead = len; // entries are line positions already
        nscan = len;
    } else if (policy == PB_POLICY_LOCATOR) {
        em.nhead = len;
        nscan = len;
    } else { // ref_seq.h:291-311
        const int64_t nmax = len - 16;
        em.nhead = std::max<int64_t>(0, std::min<int64_t>(nmax, 20000));
        int64_t ntail = std::min<int64_t>(len - 20000 - 16, 20000);
        nscan = em.nhead + std::max<int64_t>(ntail, 0);
    }
    pb_index *ix = new pb_index();
    ix->ctx = ctx;
    ix->mask = mask;
    ix->policy = policy;
    ix->ref_len = (whole_set || pairs) ? -1 : len;
    ix->whole_set = whole_set;
    ix->nscanned = nscan;
    ix->fn = make_bucket_fn(mask);
    ix->nbuckets = (int64_t)1 << ix->fn.bits;
    const int64_t nb = ix->nbuckets;
    int r;
#define TRYI(x) do { r = (x); if (r != PB_OK) { delete ix; return r; } } while (0)
    DevBuf d_keys, d_count, d_cursor, tmp, d_big, d_nbig, d_nkeys;
    if (!pairs) {
        TRYI(d_keys.alloc(ctx, (size_t)std::max<int64_t>(len, 1) * 4 + 64));
        pb_timer_begin(ctx, PB_T_SEED);
        TRYI(pb_seed_bulk_device(ctx, ref, whole_set ? 0 : ref->base[seq], len, mask, d_keys.as<uint32_t>()));
        pb_timer_end(ctx, PB_T_SEED);
    }
    pb_timer_begin(ctx, PB_T_INDEX);
    TRYI(d_count.alloc_zero(ctx, (size_t)(nb + 1) * 4));
    TRYI(ix->d_start.alloc(ctx, (size_t)(nb + 2) * 4));
    const unsigned gscan = (unsigned)std::max<int64_t>(1, (nscan + 255) / 256);
    const unsigned gset = (unsigned)std::max<int64_t>(1, std::min<int64_t>(pairs ? 1 : ref->n, (int64_t)ctx->sm_count * 8));
    if (whole_set) {
        int64_t ns = 0;
        for (int64_t i = 0; i < ref->n; ++i) ns += std::max(0, ref->len[i] - 16);
        ix->nscanned = ns;
        set_index_count_kernel<<<gset, 256, 0, ctx->stream>>>(d_keys.as<uint32_t>(), ref->d_base.as<int64_t>(), ref->d_len.as<int32_t>(), ref->n, ix->fn, d_count.as<uint32_t>());
        ctx->launches++;
    } else if (nscan > 0) {
        index_count_kernel<<<gscan, 256, 0, ctx->stream>>>(d_keys.as<uint32_t>(), em, nscan, ix->fn, d_count.as<uint32_t>());
        ctx->launches++;
    }
    TRYI(pb_scan_u32(ctx, d_count.as<uint32_t>(), ix->d_start.as<uint32_t>(), nb, tmp));
    uint32_t nent32 = 0;
    TRYI(pb_d2h(ctx, &nent32, ix->d_start.as<uint32_t>() + nb, 4));
    TRYI(pb_sync(ctx));
    ix->nentries = nent32;
    TRYI(ix->d_pos.alloc(ctx, (size_t)std::max<int64_t>(ix->nentries, 1) * 4));
    if (!ix->fn.exact) TRYI(ix->d_key.alloc(ctx, (size_t)std::max<int64_t>(ix->nentries, 1) * 4));
    if (ix->nentries > 0) {
        TRYI(d_cursor.alloc(ctx, (size_t)nb * 4));
        if (cudaMemcpyAsync(d_cursor.p, ix->d_start.p, (size_t)nb * 4, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) {
            delete ix;
            return pb_fail(ctx, PB_ERR_CUDA, "cursor copy failed");
        }
        if (whole_set)
            set_index_scatter_kernel<<<gset, 256, 0, ctx->stream>>>(d_keys.as<uint32_t>(), ref->d_base.as<int64_t>(), ref->d_len.as<int32_t>(), ref->n, ix->fn, d_cursor.as<uint32_t>(), ix->d_pos.as<int32_t>());
        else
            index_scatter_kernel<<<gscan, 256, 0, ctx->stream>>>(d_keys.as<uint32_t>(), em, nscan, ix->fn, d_cursor.as<uint32_t>(), ix->d_pos.as<int32_t>());
        ctx->launches++;
        // big buckets are rare (degenerate repeats); the list can hold at most nentries/SMALL_BUCKET of them
        const int64_t maxbig = ix->nentries / SMALL_BUCKET + 1;
        TRYI(d_big.alloc(ctx, (size_t)maxbig * 4));
        TRYI(d_nbig.alloc_zero(ctx, 16));
        index_sort_small_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, ctx->stream>>>(ix->d_start.as<uint32_t>(), nb, ix->d_pos.as<int32_t>(), d_big.as<uint32_t>(), d_nbig.as<uint32_t>());
        ctx->launches++;
        uint32_t nbig = 0;
        TRYI(pb_d2h(ctx, &nbig, d_nbig.p, 4));
        TRYI(pb_sync(ctx));
        if (nbig) {
            DevBuf d_sizes, d_off, d_tmp, tmp2;
            TRYI(d_sizes.alloc(ctx, (size_t)(nbig + 1) * 4));
            TRYI(d_off.alloc(ctx, (size_t)(nbig + 2) * 8));
            big_sizes_kernel<<<(nbig + 255) / 256, 256, 0, ctx->stream>>>(ix->d_start.as<uint32_t>(), d_big.as<uint32_t>(), nbig, d_sizes.as<uint32_t>());
            ctx->launches++;
            TRYI(pb_scan_i64(ctx, d_sizes.as<uint32_t>(), d_off.as<int64_t>(), nbig, tmp2));
            int64_t tot = 0;
            TRYI(pb_d2h(ctx, &tot, d_off.as<int64_t>() + nbig, 8));
            TRYI(pb_sync(ctx));
            TRYI(d_tmp.alloc(ctx, (size_t)tot * 4 + 16));
            index_sort_big_kernel<<<nbig, 256, 0, ctx->stream>>>(ix->d_start.as<uint32_t>(), d_big.as<uint32_t>(), d_off.as<unsigned long long>(), ix->d_pos.as<int32_t>(), d_tmp.as<int32_t>());
            ctx->launches++;
            TRYI(pb_sync(ctx));
        }
        index_finalize_kernel<<<(unsigned)((ix->nentries + 255) / 256), 256, 0, ctx->stream>>>(d_keys.as<uint32_t>(), em, ix->nentries, ix->d_pos.as<int32_t>(), ix->fn.exact ? nullptr : ix->d_key.as<uint32_t>());
        ctx->launches++;
    }
    if (ix->fn.exact) { // packed headers need at least 2 count bits above the start offset
        int shift = 1;
        while (shift < 32 && ((uint64_t)1 << shift) <= (uint64_t)ix->nentries) ++shift;
        const char *force = getenv("PB_PK_SHIFT"); // test knob: fewer count bits, so that short lists already take the escape path
        if (force && atoi(force) >= shift && atoi(force) <= 30) shift = atoi(force);
        if (shift <= 30) {
            ix->pk_shift = shift;
            ix->pk_esc = (1u << (32 - shift)) - 1u;
            TRYI(ix->d_pk.alloc(ctx, (size_t)nb * 4));
            index_pack_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, ctx->stream>>>(ix->d_start.as<uint32_t>(), nb, shift, ix->pk_esc, ix->d_pk.as<uint32_t>());
            ctx->launches++;
        }
    }
    TRYI(d_nkeys.alloc_zero(ctx, 16));
    index_nkeys_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, ctx->stream>>>(ix->d_start.as<uint32_t>(), nb, ix->fn.exact ? nullptr : ix->d_key.as<uint32_t>(), d_nkeys.as<unsigned long long>());
    ctx->launches++;
    pb_timer_end(ctx, PB_T_INDEX);
    unsigned long long nk = 0;
    TRYI(pb_d2h(ctx, &nk, d_nkeys.p, 8));
    TRYI(pb_sync(ctx));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { delete ix; return pb_fail(ctx, PB_ERR_CUDA, "index build failed: %s", cudaGetErrorString(e)); }
    ix->nkeys = (int64_t)nk;
    pb_timer_collect(ctx);
#undef TRYI
    *out = ix;
    return PB_OK;
}

extern "C" int pb_index_build(pb_ctx *ctx, const pb_seqset *ref, int64_t seq, uint32_t mask, int policy, pb_index **out)
{
    if (!ctx || !ref || !out || seq < 0 || seq >= ref->n) return pb_fail(ctx, PB_ERR_ARG, "pb_index_build: bad argument");
    if (policy != PB_POLICY_LOCATOR && policy != PB_POLICY_REFSEQ) return pb_fail(ctx, PB_ERR_ARG, "unknown index policy %d", policy);
    return index_build_impl(ctx, ref, seq, mask, policy, out);
}

// The seed map as the reference's drivers fill it (locator.cpp:65, ref_seq.h:299,307): seedmap[key].push_back(pos), one pair
// at a time.  hash_table::operator[] collects the pairs and hands them over here in insertion order; per key the positions
// come back in that order.  Keys are looked up exactly (hashed buckets + stored key).  A zero key is not stored (the
// reference never inserts one, locator.cpp:64 -- the host class keeps such an entry itself).
extern "C" int pb_index_build_pairs(pb_ctx *ctx, const uint32_t *keys, const int32_t *pos, int64_t n, pb_index **out)
{
    if (!ctx || !out || n < 0 || (n && (!keys || !pos))) return pb_fail(ctx, PB_ERR_ARG, "pb_index_build_pairs: bad argument");
    if (n > (int64_t)INT32_MAX) return pb_fail(ctx, PB_ERR_DOMAIN, "too many pairs for one index");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf d_k, d_p;
    PB_TRY(d_k.alloc(ctx, (size_t)std::max<int64_t>(n, 1) * 4));
    PB_TRY(d_p.alloc(ctx, (size_t)std::max<int64_t>(n, 1) * 4));
    PB_TRY(pb_h2d(ctx, d_k.p, keys, (size_t)n * 4));
    PB_TRY(pb_h2d(ctx, d_p.p, pos, (size_t)n * 4));
    return index_build_impl(ctx, nullptr, 0, 0xFFFFFFFFu, PB_POLICY_LOCATOR, out, d_k.as<uint32_t>(), d_p.as<int32_t>(), n);
}

extern "C" int pb_index_build_set(pb_ctx *ctx, const pb_seqset *set, uint32_t mask, pb_index **out)
{
    if (!ctx || !set || !out || set->n < 1) return pb_fail(ctx, PB_ERR_ARG, "pb_index_build_set: bad argument");
    if (set->base[set->n] > (int64_t)INT32_MAX) return pb_fail(ctx, PB_ERR_DOMAIN, "set of %lld padded bases: index positions are 32-bit; split the set", (long long)set->base[set->n]);
    for (int64_t i = 0; i < set->n; ++i)
        if (set->len[i] > 20016)
            return pb_fail(ctx, PB_ERR_DOMAIN, "sequence %lld has %d bases: a reference longer than MAX_READ_LEN+16 gets get_seedmap's tail pass "
                           "(ref_seq.h:301-308); index it on its own with pb_index_build(PB_POLICY_REFSEQ)", (long long)i, set->len[i]);
    return index_build_impl(ctx, set, -1, mask, PB_POLICY_REFSEQ, out);
}

extern "C" void pb_index_free(pb_index *ix)
{
    if (!ix) return;
    cudaSetDevice(ix->ctx->device);
    delete ix;
}
extern "C" int64_t pb_index_nkeys(const pb_index *ix) { return ix ? ix->nkeys : 0; }
extern "C" int64_t pb_index_nentries(const pb_index *ix) { return ix ? ix->nentries : 0; }
extern "C" int64_t pb_index_nscanned(const pb_index *ix) { return ix ? ix->nscanned : 0; }
extern "C" uint32_t pb_index_mask(const pb_index *ix) { return ix ? ix->mask : 0; }

// ---------------------------------------------------------------------------------------------
// K2: probe + gather
// ---------------------------------------------------------------------------------------------

struct IndexView {
    const uint32_t *start;
    const int32_t *pos;
    const uint32_t *key; // NULL when exact
    const uint32_t *pk;  // packed bucket headers, NULL when absent
    int pk_shift;
    uint32_t pk_esc;
    BucketFn fn;
};

// bucket range of a query and the number of entries that really carry its key
__device__ __forceinline__ void probe_one(const IndexView &iv, uint32_t key, uint32_t *s, uint32_t *blen, uint32_t *cnt)
{
    if (!key) { *s = 0; *blen = 0; *cnt = 0; return; } // keys with (sd & mask) == 0 are never inserted (locator.cpp:64)
    const uint32_t b = bucket_of(iv.fn, key);
    const uint32_t s0 = __ldg(iv.start + b), c = __ldg(iv.start + b + 1) - s0;
    uint32_t m = c;
    if (iv.key) {
        m = 0;
        for (uint32_t t = 0; t < c; ++t) m += __ldg(iv.key + s0 + t) == key;
    }
    *s = s0; *blen = c; *cnt = m;
}

__global__ void probe_count_kernel(IndexView iv, const uint32_t *__restrict__ keys, int64_t nq, uint32_t *__restrict__ cnt)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint32_t s, bl, c;
    probe_one(iv, keys[q], &s, &bl, &c);
    cnt[q] = c;
}

__global__ void probe_gather_tail_kernel(IndexView iv, const uint32_t *__restrict__ keys, int64_t q0, int64_t nq, const int64_t *__restrict__ qoff,
                                         int32_t *__restrict__ cand_pos, int32_t *__restrict__ cand_q)
{
    const int64_t q = q0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; // queries [q0, nq)
    if (q >= nq) return;
    const int64_t o0 = qoff[q];
    if (qoff[q + 1] == o0) return;
    const uint32_t key = keys[q];
    uint32_t s, bl, c;
    probe_one(iv, key, &s, &bl, &c);
    int64_t o = o0;
    for (uint32_t t = 0; t < bl; ++t) {
        if (iv.key && __ldg(iv.key + s + t) != key) continue;
        cand_pos[o] = __ldg(iv.pos + s + t);
        cand_q[o] = (int32_t)q;
        ++o;
    }
}

// Four queries per thread (exact, direct-address indexes).  One query per thread leaves the kernel latency-bound: a key load
// from HBM, then two dependent random reads of the bucket table, with nothing else in flight.  Here a thread loads four
// consecutive keys as one uint4 and issues the eight bucket-header loads back to back before it uses any of them, so four times
// as many random reads are in flight per resident thread; counts leave as one uint4 store.
// CS: keys and counts are streams that pass through once -- loaded / stored evict-first (ld.global.cs / st.global.cs) so that
// they do not push the bucket table, which every probe hits at random, out of L2
template <int V, bool CS> // V uint4 groups (4*V queries) per thread
__global__ void __launch_bounds__(256)
probe_count4_kernel(IndexView iv, const uint4 *__restrict__ keys4, int64_t nq4, uint4 *__restrict__ cnt4)
{
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (g0 >= nq4) return;
    uint32_t k[4 * V], s[4 * V], e[4 * V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const uint4 x = CS ? __ldcs(keys4 + min(g0 + v, nq4 - 1)) : keys4[min(g0 + v, nq4 - 1)];
        k[4 * v] = x.x; k[4 * v + 1] = x.y; k[4 * v + 2] = x.z; k[4 * v + 3] = x.w;
    }
    if (iv.pk) { // one random read per query: start and (capped) count in one word
#pragma unroll
        for (int t = 0; t < 4 * V; ++t) s[t] = __ldg(iv.pk + bucket_of(iv.fn, k[t]));
#pragma unroll
        for (int t = 0; t < 4 * V; ++t) {
            uint32_t c = s[t] >> iv.pk_shift;
            if (c == iv.pk_esc) { // long list: exact count from the offsets
                const uint32_t b = bucket_of(iv.fn, k[t]);
                c = __ldg(iv.start + b + 1) - __ldg(iv.start + b);
            }
            e[t] = c;
        }
    } else {
#pragma unroll
        for (int t = 0; t < 4 * V; ++t) {
            const uint32_t b = bucket_of(iv.fn, k[t]);
            s[t] = __ldg(iv.start + b);
            e[t] = __ldg(iv.start + b + 1) - s[t];
        }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
        if (g0 + v >= nq4) break;
        uint4 c; // keys with (sd & mask) == 0 are never inserted (locator.cpp:64)
        c.x = k[4 * v] ? e[4 * v] : 0u;
        c.y = k[4 * v + 1] ? e[4 * v + 1] : 0u;
        c.z = k[4 * v + 2] ? e[4 * v + 2] : 0u;
        c.w = k[4 * v + 3] ? e[4 * v + 3] : 0u;
        if (CS) __stcs(cnt4 + g0 + v, c); else cnt4[g0 + v] = c;
    }
}

template <bool CS> // CS: offsets, keys and the candidate arrays are streams (evict-first), see probe_count4_kernel
__global__ void __launch_bounds__(256)
probe_gather4_kernel(IndexView iv, const uint4 *__restrict__ keys4, int64_t nq4, const int64_t *__restrict__ qoff,
                     int32_t *__restrict__ cand_pos, int32_t *__restrict__ cand_q)
{
    const int64_t q4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q4 >= nq4) return;
    const int64_t q = q4 * 4;
    int64_t o[5];
#pragma unroll
    for (int t = 0; t < 5; ++t) o[t] = CS ? __ldcs(qoff + q + t) : qoff[q + t];
    if (o[4] == o[0]) return; // none of the four has a candidate
    const uint4 k4 = CS ? __ldcs(keys4 + q4) : keys4[q4];
    const uint32_t k[4] = {k4.x, k4.y, k4.z, k4.w};
    uint32_t s[4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
        s[t] = o[t + 1] > o[t] ? (iv.pk ? __ldg(iv.pk + bucket_of(iv.fn, k[t])) & ((1u << iv.pk_shift) - 1u) : __ldg(iv.start + bucket_of(iv.fn, k[t]))) : 0u;
    int32_t first[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) first[t] = o[t + 1] > o[t] ? __ldg(iv.pos + s[t]) : 0; // most lists hold one position
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int64_t n = o[t + 1] - o[t];
        if (n <= 0) continue;
        if (CS) { __stcs(cand_pos + o[t], first[t]); __stcs(cand_q + o[t], (int32_t)(q + t)); }
        else { cand_pos[o[t]] = first[t]; cand_q[o[t]] = (int32_t)(q + t); }
        for (int64_t i = 1; i < n; ++i) {
            cand_pos[o[t] + i] = __ldg(iv.pos + s[t] + i);
            cand_q[o[t] + i] = (int32_t)(q + t);
        }
    }
}

// Diagonal-bin voting: one warp per read tallies its candidates' diagonals (pos - j) in 256-base bins with a
// shared-memory histogram; __match_any_sync groups the lanes that hit the same bin so each group issues one shared-memory
// atomic, then a shuffle reduction picks the fullest bin.  The tally is reported per read (votes / best diagonal); it is
// NOT used to reorder or prune candidates -- the reference takes the first success in list order, so
// pruning would change results.  It tells a caller how concentrated a read's seed hits are.
#define VOTE_BINS 256
__global__ void __launch_bounds__(128)
vote_kernel(const int64_t *__restrict__ qoff, const int32_t *__restrict__ cand_pos, const int32_t *__restrict__ cand_q, int64_t nkept,
            int ntrial, int32_t *__restrict__ votes, int32_t *__restrict__ best_diag)
{
    __shared__ int hist[4][VOTE_BINS];
    __shared__ int tag[4][VOTE_BINS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t k = (int64_t)blockIdx.x * 4 + warp;
    if (k >= nkept) return;
    for (int b = lane; b < VOTE_BINS; b += 32) { hist[warp][b] = 0; tag[warp][b] = -1; }
    __syncwarp();
    const int64_t c0 = qoff[k * ntrial], c1 = qoff[(k + 1) * ntrial];
    for (int64_t cb = c0; cb < c1; cb += 32) {
        const int64_t c = cb + lane;
        const bool valid = c < c1;
        int bin = -1;
        if (valid) {
            const int j = cand_q[c] - (int)(k * ntrial);
            bin = (cand_pos[c] - j) >> 8; // 256-base diagonal bins (arithmetic shift: negative diagonals have their own bins)
        }
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned peers = __match_any_sync(act, bin);
            if (lane == __ffs(peers) - 1) { // one shared-memory atomic per distinct bin in this batch of 32
                const int slot = (unsigned)bin % VOTE_BINS;
                atomicAdd(&hist[warp][slot], __popc(peers));
                tag[warp][slot] = bin; // bins that collide in the table share a slot: an upper bound, fine for a tally
            }
        }
        __syncwarp();
    }
    int best = 0, bdiag = 0;
    for (int b = lane; b < VOTE_BINS; b += 32)
        if (hist[warp][b] > best) { best = hist[warp][b]; bdiag = tag[warp][b]; }
    for (int d = 16; d; d >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, best, d), od = __shfl_xor_sync(0xffffffffu, bdiag, d);
        if (ob > best || (ob == best && od < bdiag)) { best = ob; bdiag = od; }
    }
    if (lane == 0) { votes[k] = best; best_diag[k] = bdiag << 8; }
}

int pb_vote(pb_ctx *ctx, const ProbeOut *po, int64_t nkept, int ntrial, int32_t *d_votes, int32_t *d_best_diag)
{
    if (nkept <= 0) return PB_OK;
    vote_kernel<<<(unsigned)((nkept + 3) / 4), 128, 0, ctx->stream>>>(po->d_qoff.as<int64_t>(), po->d_cand_pos.as<int32_t>(),
                                                                     po->d_cand_q.as<int32_t>(), nkept, ntrial, d_votes, d_best_diag);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}

// seeds at the first ntrial offsets of every kept read: key = encode(read + j) & mask (locator.cpp:75)
__global__ void locate_seed_kernel(const uint32_t *__restrict__ pw, const int64_t *__restrict__ base, const int32_t *__restrict__ kept,
                                   int64_t nkept, int ntrial, uint32_t mask, uint32_t *__restrict__ keys)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nkept * ntrial) return;
    const int64_t k = q / ntrial;
    const int j = (int)(q - k * ntrial);
    const int64_t g = base[kept[k]] + j;
    const uint32_t h0 = bswap32(__ldg(pw + (g >> 4))), h1 = bswap32(__ldg(pw + (g >> 4) + 1));
    keys[q] = seed_from_be(h0, h1, (int)(g & 15)) & mask;
}

// assembler-side trials (spaced_seed.cpp:424-426, try_align :261-281): query t of read k is trial j = t/2, forward from
// pos = j (t even) or backward from pos = len-j-16 (t odd); key = seed_at(read, pos) & mask.  quirk: the reference's
// seed_at returns the u32 at BYTE offset pos of the record body when pos%4==0 (Q-S1) -- read from the kept .bin image,
// zero past its end.  A segment shorter than min_overlap is not probed (:280): its key is forced to 0 (= never found).
__global__ void overlap_seed_kernel(const uint32_t *__restrict__ pw, const int64_t *__restrict__ base, const int32_t *__restrict__ len,
                                    const int32_t *__restrict__ kept, int64_t nkept, int max_trial, int min_overlap, uint32_t mask,
                                    const uint8_t *__restrict__ image, int64_t image_bytes, const int64_t *__restrict__ recoff,
                                    int quirk, uint32_t *__restrict__ keys)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ntr = 2 * max_trial;
    if (q >= nkept * ntr) return;
    const int64_t k = q / ntr;
    const int t = (int)(q - k * ntr), j = t >> 1;
    const bool forward = (t & 1) == 0;
    const int r = kept[k];
    const int slen = len[r];
    const int pos = forward ? j : slen - j - 16;
    const int s_len = forward ? slen - pos : pos + 16;
    uint32_t key = 0u;
    if (pos >= 0 && s_len >= min_overlap) {
        if (quirk && (pos & 3) == 0 && image) {
            const int64_t o = recoff[r] + pos;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (o + b < image_bytes) key |= (uint32_t)image[o + b] << (8 * b);
        } else {
            const int64_t g = base[r] + pos;
            const uint32_t h0 = bswap32(__ldg(pw + (g >> 4))), h1 = bswap32(__ldg(pw + (g >> 4) + 1));
            key = seed_from_be(h0, h1, (int)(g & 15));
        }
        key &= mask;
    }
    keys[q] = key;
}

static IndexView view_of(const pb_index *ix)
{
    IndexView iv;
    iv.start = ix->d_start.as<uint32_t>();
    iv.pos = ix->d_pos.as<int32_t>();
    iv.key = ix->fn.exact ? nullptr : ix->d_key.as<uint32_t>();
    iv.pk = ix->pk_esc ? ix->d_pk.as<uint32_t>() : nullptr;
    iv.pk_shift = ix->pk_shift;
    iv.pk_esc = ix->pk_esc;
    iv.fn = ix->fn;
    return iv;
}

// count / gather launches: four queries per thread where the index is direct-address and the arrays are 16-byte aligned, the
// remainder (and hashed indexes) one query per thread
static int launch_probe_count(pb_ctx *ctx, const IndexView &iv, const uint32_t *d_keys, int64_t nq, uint32_t *d_cnt)
{
    int64_t done = 0;
    if (!iv.key && nq >= 4 && (((uintptr_t)d_keys | (uintptr_t)d_cnt) & 15) == 0 && !getenv("PB_PROBE1")) {
        const int64_t nq4 = nq / 4;
        const char *pv = getenv("PB_PROBE_V");
        const int V = pv ? atoi(pv) : 2;
        const bool cs = !(getenv("PB_PROBE_CS") && atoi(getenv("PB_PROBE_CS")) == 0);
        const uint4 *k4 = reinterpret_cast<const uint4 *>(d_keys);
        uint4 *c4 = reinterpret_cast<uint4 *>(d_cnt);
        if (V >= 2) {
            const unsigned g = (unsigned)(((nq4 + 1) / 2 + 255) / 256);
            if (cs) probe_count4_kernel<2, true><<<g, 256, 0, ctx->stream>>>(iv, k4, nq4, c4);
            else probe_count4_kernel<2, false><<<g, 256, 0, ctx->stream>>>(iv, k4, nq4, c4);
        } else {
            const unsigned g = (unsigned)((nq4 + 255) / 256);
            if (cs) probe_count4_kernel<1, true><<<g, 256, 0, ctx->stream>>>(iv, k4, nq4, c4);
            else probe_count4_kernel<1, false><<<g, 256, 0, ctx->stream>>>(iv, k4, nq4, c4);
        }
        PB_LAUNCH_CHECK(ctx);
        done = nq4 * 4;
    }
    if (done < nq) {
        probe_count_kernel<<<(unsigned)((nq - done + 255) / 256), 256, 0, ctx->stream>>>(iv, d_keys + done, nq - done, d_cnt + done);
        PB_LAUNCH_CHECK(ctx);
    }
    return PB_OK;
}

static int launch_probe_gather(pb_ctx *ctx, const IndexView &iv, const uint32_t *d_keys, int64_t nq, const int64_t *d_qoff,
                               int32_t *d_cand_pos, int32_t *d_cand_q)
{
    int64_t done = 0;
    if (!iv.key && nq >= 4 && ((uintptr_t)d_keys & 15) == 0 && !getenv("PB_PROBE1")) {
        const int64_t nq4 = nq / 4;
        const bool cs = !(getenv("PB_PROBE_CS") && atoi(getenv("PB_PROBE_CS")) == 0);
        const unsigned g = (unsigned)((nq4 + 255) / 256);
        if (cs) probe_gather4_kernel<true><<<g, 256, 0, ctx->stream>>>(iv, reinterpret_cast<const uint4 *>(d_keys), nq4, d_qoff, d_cand_pos, d_cand_q);
        else probe_gather4_kernel<false><<<g, 256, 0, ctx->stream>>>(iv, reinterpret_cast<const uint4 *>(d_keys), nq4, d_qoff, d_cand_pos, d_cand_q);
        PB_LAUNCH_CHECK(ctx);
        done = nq4 * 4;
    }
    if (done < nq) {
        probe_gather_tail_kernel<<<(unsigned)((nq - done + 255) / 256), 256, 0, ctx->stream>>>(iv, d_keys, done, nq, d_qoff, d_cand_pos, d_cand_q);
        PB_LAUNCH_CHECK(ctx);
    }
    return PB_OK;
}

// keys (one per query, already on the device) -> counts -> exclusive offsets -> candidate arrays
static int probe_and_gather(pb_ctx *ctx, const pb_index *ix, const uint32_t *d_keys, int64_t nq, ProbeOut *po)
{
    DevBuf d_cnt, tmp;
    PB_TRY(d_cnt.alloc(ctx, (size_t)nq * 4));
    pb_timer_begin(ctx, PB_T_PROBE);
    IndexView iv = view_of(ix);
    PB_TRY(launch_probe_count(ctx, iv, d_keys, nq, d_cnt.as<uint32_t>()));
    PB_TRY(pb_scan_i64(ctx, d_cnt.as<uint32_t>(), po->d_qoff.as<int64_t>(), nq, tmp));
    int64_t ncand = 0;
    PB_TRY(pb_d2h(ctx, &ncand, po->d_qoff.as<int64_t>() + nq, 8));
    PB_TRY(pb_sync(ctx));
    if (ncand > (int64_t)INT32_MAX) return pb_fail(ctx, PB_ERR_DOMAIN, "%lld candidates in one batch; split the read batch", (long long)ncand);
    po->ncand = ncand;
    PB_TRY(po->d_cand_pos.alloc(ctx, (size_t)std::max<int64_t>(ncand, 1) * 4));
    PB_TRY(po->d_cand_q.alloc(ctx, (size_t)std::max<int64_t>(ncand, 1) * 4));
    if (ncand) {
        PB_TRY(launch_probe_gather(ctx, iv, d_keys, nq, po->d_qoff.as<int64_t>(), po->d_cand_pos.as<int32_t>(), po->d_cand_q.as<int32_t>()));
    }
    pb_timer_end(ctx, PB_T_PROBE);
    return PB_OK;
}

static int empty_probe(pb_ctx *ctx, ProbeOut *po)
{
    PB_CUDA(ctx, cudaMemsetAsync(po->d_qoff.p, 0, 16, ctx->stream));
    PB_TRY(po->d_cand_pos.alloc(ctx, 16));
    PB_TRY(po->d_cand_q.alloc(ctx, 16));
    return PB_OK;
}

int pb_locate_seed_probe(pb_ctx *ctx, const pb_index *ix, const pb_seqset *reads, const int32_t *d_kept, int64_t nkept,
                         int ntrial, ProbeOut *po)
{
    const int64_t nq = nkept * ntrial;
    po->ncand = 0;
    PB_TRY(po->d_qoff.alloc(ctx, (size_t)(nq + 2) * 8));
    if (nq == 0) return empty_probe(ctx, po);
    DevBuf d_keys;
    PB_TRY(d_keys.alloc(ctx, (size_t)nq * 4));
    pb_timer_begin(ctx, PB_T_SEED);
    locate_seed_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, ctx->stream>>>(reads->d_packed.as<uint32_t>(), reads->d_base.as<int64_t>(), d_kept, nkept, ntrial, ix->mask, d_keys.as<uint32_t>());
    PB_LAUNCH_CHECK(ctx);
    pb_timer_end(ctx, PB_T_SEED);
    return probe_and_gather(ctx, ix, d_keys.as<uint32_t>(), nq, po);
}

int pb_overlap_seed_probe(pb_ctx *ctx, const pb_index *ix, const pb_seqset *reads, const int32_t *d_kept, int64_t nkept,
                          int max_trial, int min_overlap, int quirk, ProbeOut *po)
{
    const int64_t nq = nkept * 2 * max_trial;
    po->ncand = 0;
    PB_TRY(po->d_qoff.alloc(ctx, (size_t)(nq + 2) * 8));
    if (nq == 0) return empty_probe(ctx, po);
    DevBuf d_keys;
    PB_TRY(d_keys.alloc(ctx, (size_t)nq * 4));
    pb_timer_begin(ctx, PB_T_SEED);
    overlap_seed_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, ctx->stream>>>(
        reads->d_packed.as<uint32_t>(), reads->d_base.as<int64_t>(), reads->d_len.as<int32_t>(), d_kept, nkept, max_trial, min_overlap,
        ix->mask, reads->image_bytes ? reads->d_image.as<uint8_t>() : nullptr, reads->image_bytes,
        reads->image_bytes ? reads->d_recoff.as<int64_t>() : nullptr, quirk, d_keys.as<uint32_t>());
    PB_LAUNCH_CHECK(ctx);
    pb_timer_end(ctx, PB_T_SEED);
    return probe_and_gather(ctx, ix, d_keys.as<uint32_t>(), nq, po);
}

// Bulk probe (bandwidth measurement): every position of the set is a query -- K1 bulk keys, then probe_count, scan,
// probe_gather, all on the device.  Reports queries, candidates and the times of the three stages.
extern "C" int pb_probe_bulk_device(pb_ctx *ctx, const pb_index *ix, const pb_seqset *s, int64_t *nqueries, int64_t *ncand,
                                    float *ms_seed, float *ms_count, float *ms_gather)
{
    if (!ctx || !ix || !s) return pb_fail(ctx, PB_ERR_ARG, "pb_probe_bulk_device: bad argument");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t nq = s->base[s->n];
    if (nq <= 0) return pb_fail(ctx, PB_ERR_ARG, "empty set");
    DevBuf d_keys, d_cnt, d_qoff, d_pos, d_q, tmp;
    PB_TRY(d_keys.alloc(ctx, (size_t)nq * 4 + 64));
    PB_TRY(d_cnt.alloc(ctx, (size_t)nq * 4));
    PB_TRY(d_qoff.alloc(ctx, (size_t)(nq + 2) * 8));
    cudaEvent_t ev[6];
    for (auto &e : ev) PB_CUDA(ctx, cudaEventCreate(&e));
    IndexView iv = view_of(ix);
    // PB_PROBE_L2PERSIST=1 (opt-in: the set-aside is a device-wide limit): pin the packed bucket table in L2 for the bulk passes
    const bool persist = getenv("PB_PROBE_L2PERSIST") && atoi(getenv("PB_PROBE_L2PERSIST")) && iv.pk;
    if (persist) {
        const size_t bytes = (size_t)ix->nbuckets * 4;
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes + (bytes >> 3));
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof av);
        av.accessPolicyWindow.base_ptr = (void *)iv.pk;
        av.accessPolicyWindow.num_bytes = bytes;
        av.accessPolicyWindow.hitRatio = 1.0f;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
    }
    cudaEventRecord(ev[0], ctx->stream);
    PB_TRY(pb_seed_bulk_device(ctx, s, 0, nq, ix->mask, d_keys.as<uint32_t>()));
    cudaEventRecord(ev[1], ctx->stream);
    cudaEventRecord(ev[2], ctx->stream);
    PB_TRY(launch_probe_count(ctx, iv, d_keys.as<uint32_t>(), nq, d_cnt.as<uint32_t>()));
    cudaEventRecord(ev[3], ctx->stream);
    PB_TRY(pb_scan_i64(ctx, d_cnt.as<uint32_t>(), d_qoff.as<int64_t>(), nq, tmp));
    int64_t nc = 0;
    PB_TRY(pb_d2h(ctx, &nc, d_qoff.as<int64_t>() + nq, 8));
    PB_TRY(pb_sync(ctx));
    PB_TRY(d_pos.alloc(ctx, (size_t)std::max<int64_t>(nc, 1) * 4));
    PB_TRY(d_q.alloc(ctx, (size_t)std::max<int64_t>(nc, 1) * 4));
    cudaEventRecord(ev[4], ctx->stream);
    PB_TRY(launch_probe_gather(ctx, iv, d_keys.as<uint32_t>(), nq, d_qoff.as<int64_t>(), d_pos.as<int32_t>(), d_q.as<int32_t>()));
    cudaEventRecord(ev[5], ctx->stream);
    PB_TRY(pb_sync(ctx));
    if (persist) {
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof av);
        cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
        cudaGetLastError();
    }
    float a = 0, b = 0, c = 0;
    cudaEventElapsedTime(&a, ev[0], ev[1]);
    cudaEventElapsedTime(&b, ev[2], ev[3]);
    cudaEventElapsedTime(&c, ev[4], ev[5]);
    for (auto &e : ev) cudaEventDestroy(e);
    if (nqueries) *nqueries = nq;
    if (ncand) *ncand = nc;
    if (ms_seed) *ms_seed = a;
    if (ms_count) *ms_count = b;
    if (ms_gather) *ms_gather = c;
    return PB_OK;
}

// ---------------------------------------------------------------------------------------------
// random-gather peak: the measured ceiling of the bulk probe (K2)
// ---------------------------------------------------------------------------------------------
// Nothing but independent random 4-byte reads of a table (default: 64 MB, the bucket table of a weight-12 mask), eight in
// flight per thread, indices from a register LCG: no key stream in, no count stream out, no bucket function.  What this
// kernel reaches is what the SMs' request path to L2 delivers for one divergent 32-byte sector per read -- the rate the
// probe passes are measured against.
__global__ void __launch_bounds__(256) random_gather_kernel(const uint32_t *__restrict__ table, uint32_t idx_mask, int iters, uint32_t *out)
{
    uint32_t x = ((uint32_t)blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t acc = 0u;
    for (int it = 0; it < iters; ++it) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x = x * 1664525u + 1013904223u;
            v[k] = __ldg(table + ((x >> 4) & idx_mask));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc ^= v[k];
    }
    if (acc == 0x9e3779b9u) out[0] = acc; // keeps the loads alive
}

extern "C" int pb_random_gather_peak(pb_ctx *ctx, size_t table_bytes, double *reads_per_s)
{
    if (!ctx || !reads_per_s) return pb_fail(ctx, PB_ERR_ARG, "pb_random_gather_peak: bad argument");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!table_bytes) table_bytes = (size_t)64 << 20;
    size_t words = 1024;
    while (words * 2 * 4 <= table_bytes) words *= 2; // a power of two: the index is a mask
    DevBuf d_tab, d_out;
    PB_TRY(d_tab.alloc_zero(ctx, words * 4));
    PB_TRY(d_out.alloc_zero(ctx, 64));
    const int iters = 64, threads = 256, use_blocks = 1 << 15; // 2^15 CTAs * 256 threads * 512 reads = 2^32 reads per launch
    cudaEvent_t e0, e1;
    PB_CUDA(ctx, cudaEventCreate(&e0));
    PB_CUDA(ctx, cudaEventCreate(&e1));
    float best = 0.f;
    for (int rep = 0; rep < 4; ++rep) { // the first launch warms up (and pulls the table into L2); keep the fastest
        cudaEventRecord(e0, ctx->stream);
        random_gather_kernel<<<use_blocks, threads, 0, ctx->stream>>>(d_tab.as<uint32_t>(), (uint32_t)(words - 1), iters, d_out.as<uint32_t>());
        cudaEventRecord(e1, ctx->stream);
        ctx->launches++;
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && (best == 0.f || ms < best)) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *reads_per_s = (double)use_blocks * threads * (double)iters * 8.0 / (best * 1e-3);
    return PB_OK;
}

__global__ void find_write_kernel(IndexView iv, const uint32_t *__restrict__ keys, int64_t n, const int64_t *__restrict__ pos_off,
                                  int64_t cap_each, int32_t *__restrict__ out)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const uint32_t key = keys[q];
    uint32_t s, bl, c;
    probe_one(iv, key, &s, &bl, &c);
    int64_t w = 0;
    for (uint32_t t = 0; t < bl && w < cap_each; ++t) {
        if (iv.key && __ldg(iv.key + s + t) != key) continue;
        out[pos_off[q] + w] = __ldg(iv.pos + s + t);
        ++w;
    }
}

extern "C" int pb_index_find_batch(pb_ctx *ctx, const pb_index *ix, const uint32_t *keys, int64_t n, int64_t *count, int32_t *pos,
                                   const int64_t *pos_off, int64_t cap_each)
{
    if (!ctx || !ix || n < 0 || (n && (!keys || !count))) return pb_fail(ctx, PB_ERR_ARG, "pb_index_find_batch: bad argument");
    if (!n) return PB_OK;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf d_keys, d_cnt;
    PB_TRY(d_keys.alloc(ctx, (size_t)n * 4));
    PB_TRY(d_cnt.alloc(ctx, (size_t)n * 4));
    PB_TRY(pb_h2d(ctx, d_keys.p, keys, (size_t)n * 4));
    IndexView iv = view_of(ix);
    const unsigned grid = (unsigned)((n + 255) / 256);
    probe_count_kernel<<<grid, 256, 0, ctx->stream>>>(iv, d_keys.as<uint32_t>(), n, d_cnt.as<uint32_t>());
    PB_LAUNCH_CHECK(ctx);
    std::vector<uint32_t> cnt((size_t)n);
    PB_TRY(pb_d2h(ctx, cnt.data(), d_cnt.p, (size_t)n * 4));
    PB_TRY(pb_sync(ctx));
    int64_t maxend = 0;
    for (int64_t i = 0; i < n; ++i) {
        count[i] = cnt[i];
        if (pos && pos_off) maxend = std::max<int64_t>(maxend, pos_off[i] + std::min<int64_t>(cnt[i], cap_each));
    }
    if (pos && pos_off && maxend > 0) {
        DevBuf d_off, d_out;
        PB_TRY(d_off.alloc(ctx, (size_t)n * 8));
        PB_TRY(d_out.alloc_zero(ctx, (size_t)maxend * 4));
        PB_TRY(pb_h2d(ctx, d_off.p, pos_off, (size_t)n * 8));
        find_write_kernel<<<grid, 256, 0, ctx->stream>>>(iv, d_keys.as<uint32_t>(), n, d_off.as<int64_t>(), cap_each, d_out.as<int32_t>());
        PB_LAUNCH_CHECK(ctx);
        std::vector<int32_t> host((size_t)maxend);
        PB_TRY(pb_d2h(ctx, host.data(), d_out.p, (size_t)maxend * 4));
        PB_TRY(pb_sync(ctx));
        for (int64_t i = 0; i < n; ++i) {
            int64_t m = std::min<int64_t>(cnt[i], cap_each);
            if (m > 0) memcpy(pos + pos_off[i], host.data() + pos_off[i], (size_t)m * 4);
        }
    }
    return PB_OK;
}
