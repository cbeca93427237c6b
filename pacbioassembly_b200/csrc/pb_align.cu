// pb_align.cu -- L2: banded edit-distance aligner on the GPU (replaces src/seq_aligner.h).
//
// The reference fills an 8-byte {cost,parent} cell at a time (seq_aligner.h:151-190).  Here one warp owns
// one alignment and advances a whole band row per step with a bit-parallel recurrence (Myers 1999 /
// Hyyro 2003, transposed: the state is the row's horizontal deltas, the band slides one bit per row).
// 32 lanes x S 32-bit words hold the 2*max_dst+1 band bits; the only cross-lane traffic per row is two
// 2-bit shuffles, two ballots (carry generate / propagate of the multi-word add) and one broadcast.
// Costs AND parents come out identical to the reference, including its tie-breaking
// (diag first, then left if strictly smaller, then up if strictly smaller; seq_aligner.h:164-173):
//   D0 = cost(i,j)==cost(i-1,j-1);  parent = MATCH if Eq | ~D0, else INSERT if h(i,j)=+1, else DELETE.
// tools/bitpar_model.c is the same arithmetic on the CPU, checked against the oracle cell by cell.
//
// Per cell the kernel touches 2 bits of HBM (the two parent planes, written once, coalesced 128 B per
// store instruction) -- versus 8 bytes in the reference.  No tensor cores: this is integer/bit work.
#include <limits.h>
#include <stdlib.h>

#include <algorithm>
#include <map>

#include "pb_internal.cuh"
#include "pb_carry_chain.cuh"

#define FULL 0xffffffffu
// -DPB_BOUNDS_CHECK (tools/ab_build.sh bounds -DPB_BOUNDS_CHECK): every parent-plane address the aligners store to or prefetch
// from is checked against its warp's scratch slot, every ring slot against the ring; a violation prints and traps.  The pool's
// compute-sanitizer is closed, so this build is the memory-safety check (tests/test_gpu_parity.py::test_traceback_prefetch_stress
// run with PB_LIB=build/exp/libpb_bounds.so; profiles/r02_bounds_check.log).
#ifdef PB_BOUNDS_CHECK
#define PB_CHECK_RANGE(what, p, bytes, lo, hi)                                                                                 \
    do {                                                                                                                     \
        const char *_p = reinterpret_cast<const char *>(p);                                                                  \
        if (_p < reinterpret_cast<const char *>(lo) || _p + (bytes) > reinterpret_cast<const char *>(hi) ||                   \
            (reinterpret_cast<uintptr_t>(_p) & ((bytes) - 1))) {                                                             \
            printf("PB_BOUNDS_CHECK %s: %p + %d outside [%p, %p) or misaligned (block %d thread %d)\n", what, (const void *)_p,  \
                   (int)(bytes), (const void *)(lo), (const void *)(hi), (int)blockIdx.x, (int)threadIdx.x);                  \
            __trap();                                                                                                        \
        }                                                                                                                    \
    } while (0)
#else
#define PB_CHECK_RANGE(what, p, bytes, lo, hi) ((void)0)
#endif
#define PB_STAGE_WORDS 128 // staging buffer per plane (words): 4 kbp of seg_b per TMA chunk
#ifndef PB_TB_WINDOWS
// traceback: 32-row parent windows in flight per warp (the one being walked + prefetched ones).  The backward walk is 19 % of K3
// (measured by skipping it: 91.1 -> 74.2 ms).  It is neither its instruction count (a loop with a third fewer instructions per
// step ran in the same time) nor latency (deeper prefetch is slower: 3 windows 89.9 ms, 5 windows 91.7, 8 windows 98.9): every
// 8-byte pair it reads sits in its own 128-byte line of a row written long ago, and those scattered reads land in a stream of
// writes that already runs at ~75 % of the HBM peak.
#define PB_TB_WINDOWS 3
#endif
// Parent layout and traceback prefetch ("Where K3's time goes"): parents are stored as 16-byte units of two
// adjacent band words per lane (one STG.128 per slot pair, one DRAM line per row for the walk while the path sits inside a
// unit), and the traceback prefetches its windows into a shared-memory ring with 16-byte L2-only cp.async.cg copies.  Rows above
// the matrix are zeroed with plain shared-memory stores, never with zero-fill copies: a copy whose source is ignored may still
// be handed a meaningless address by the assembler (found in round 1: only at ptxas -O3, only where windows reach above row 1).
#ifndef PB_TB_RING
#define PB_TB_RING 4 // traceback windows in the shared-memory ring (the one being walked + asynchronous prefetches); 5, 6: same or slightly slower, 8: slower
#endif
#define PB_TB_RING_WORDS (PB_TB_RING * 256 + 8 * PB_TB_RING + 8)
#ifndef PB_PAD_MOD
// Band classes with S % PB_PAD_MOD == 0 keep their Eq planes padded (one word per S words) so that the lane stride S+1 is free of
// shared-memory bank conflicts.  The padding costs ~4 ALU instructions per band word and row (the window of a lane crosses one
// pad word at a position that changes every 32 rows), and the kernel is bound by the ALU pipe, not by shared memory: measured on
// config 2 (A/B builds, same box), padding every even class 100.3 ms, only S=8,16 98.1 ms, only S=16 94.9 ms per step of K3.
#define PB_PAD_MOD 16
#endif
#ifndef ALIGN_WPB
#define ALIGN_WPB 8 // warps (alignments in flight) per CTA, at most: launches that need more shared memory use fewer
#endif
#ifndef PB_MINB3
#define PB_MINB3 6 // resident CTAs per SM asked of ptxas for the narrow-band classes
#endif

struct SeqView {
    const uint32_t *hi, *lo;
    const int64_t *base;
    const int32_t *len;
    int64_t nwords; // addressable words (line + guard)
    // bytes outside {A,C,G,T} (byte-exact DP path): position plane, sorted exception list, per-sequence value tables
    const uint32_t *irr;
    const int64_t *exc_pos;
    const uint8_t *exc_val;
    int64_t nexc;
    const uint32_t *tab;
};

static SeqView seq_view(const pb_seqset *s)
{
    SeqView v;
    v.hi = s->d_hi.as<uint32_t>();
    v.lo = s->d_lo.as<uint32_t>();
    v.base = s->d_base.as<int64_t>();
    v.len = s->d_len.as<int32_t>();
    v.nwords = s->nwords() + 4;
    v.irr = s->d_irr.as<uint32_t>();
    v.exc_pos = s->d_exc_pos.as<int64_t>();
    v.exc_val = s->d_exc_val.as<uint8_t>();
    v.nexc = s->nexc;
    v.tab = s->d_tab.as<uint32_t>();
    return v;
}

// 32 bits of a plane starting at bit `bit`; bits outside the array read 0
__device__ __forceinline__ uint32_t load_window(const uint32_t *__restrict__ arr, int64_t nwords, int64_t bit)
{
    const int64_t wi = bit >> 5;
    const unsigned sh = (unsigned)(bit & 31);
    const uint32_t w0 = (wi >= 0 && wi < nwords) ? __ldg(arr + wi) : 0u;
    const uint32_t w1 = (wi + 1 >= 0 && wi + 1 < nwords) ? __ldg(arr + wi + 1) : 0u;
    return __funnelshift_r(w0, w1, sh);
}

// first exception with line position >= g
__device__ __forceinline__ int64_t exc_lower_bound(const SeqView &V, int64_t g)
{
    int64_t a = 0, b = V.nexc;
    while (a < b) {
        const int64_t m = (a + b) >> 1;
        if (__ldg(V.exc_pos + m) < g) a = m + 1; else b = m;
    }
    return a;
}
// slot (0..3) of byte v in a sequence's table of non-ACGT values, or -1
__device__ __forceinline__ int tab_slot(uint32_t tab, uint32_t v)
{
#pragma unroll
    for (int t = 0; t < 4; ++t)
        if (((tab >> (8 * t)) & 0xFFu) == v) return t;
    return -1;
}

// Eq plane (4..7) for a row whose seg_a element is a non-ACGT byte: look the byte up in the exception list
__device__ __forceinline__ int irr_plane(const SeqView &A, int64_t g, uint32_t a_tab)
{
    const int64_t e = exc_lower_bound(A, g);
    const int slot = (e < A.nexc && __ldg(A.exc_pos + e) == g) ? tab_slot(a_tab, A.exc_val[e]) : -1;
    return 4 + (slot < 0 ? 0 : slot); // slot < 0 cannot happen for a sequence the host admitted (<= 4 distinct values)
}

// The four sequence views an alignment can draw from: reads / ref and, in overlap mode, their reversed copies.
struct PairViews { SeqView A, B, A2, B2; }; // A = reads, B = ref, A2 = reads reversed, B2 = ref reversed

struct CandView {
    const SeqView *a, *b; // seg_a, seg_b of seq_aligner::align
    int64_t a_bit, b_bit;
    int a_len, b_len;
    uint32_t a_tab;
    int j, read_pos, dir;
};

// candidate = (read r of length rlen at line offset rbase, query q of kept read k, seed-map position refpos)
// t = trial (query number within the read); the reference sequence sits at line offset ref_base with ref_len elements
__device__ __forceinline__ void derive_views(const PairViews &pv, const LocateView &lv, int t, int r, int rlen, int64_t rbase,
                                             int64_t ref_base, int ref_len, int refpos, bool irr, CandView &cv)
{
    if (lv.mode == PB_MODE_LOCATE) { // locator.cpp:78-82: align(read[j:], ref[pos:])
        cv.j = t; cv.read_pos = t; cv.dir = 1;
        cv.a = &pv.A; cv.a_bit = rbase + t; cv.a_len = rlen - t;
        cv.b = &pv.B; cv.b_bit = ref_base + refpos; cv.b_len = ref_len - refpos;
        cv.a_tab = irr ? pv.A.tab[r] : 0u;
        return;
    }
    // spaced_seed.cpp:274-285 + ref_seq.h:264,282-286: align(ref view, read view), both forward or both backward
    refpos += lv.ref_shift; // get_accessor(pos): txt_buf + beg + pos, with pre <= beg once the reference has grown in front
    const bool forward = (t & 1) == 0;
    cv.j = t >> 1; cv.dir = forward ? 1 : -1;
    cv.read_pos = forward ? cv.j : rlen - cv.j - 16;
    cv.a_tab = 0u;
    if (forward) {
        cv.a = &pv.B; cv.a_bit = ref_base + refpos; cv.a_len = ref_len - refpos;
        cv.b = &pv.A; cv.b_bit = rbase + cv.read_pos; cv.b_len = rlen - cv.read_pos;
    } else { // element k of a backward accessor at offset o is text[o - k]: position len-1-o of the reversed copy
        const int r_off = refpos + 15, s_off = cv.read_pos + 15;
        cv.a = &pv.B2; cv.a_bit = ref_base + (ref_len - 1 - r_off); cv.a_len = r_off + 1;
        cv.b = &pv.A2; cv.b_bit = rbase + (rlen - 1 - s_off); cv.b_len = s_off + 1;
    }
}

// seq_aligner.h:94-102
__device__ __forceinline__ void derive_params(int a_len, int b_len, double R, int &len_a, int &len_b, int &D)
{
    if (b_len >= a_len) {
        len_a = a_len;
        D = 1 + (int)(len_a * R);
        len_b = min(b_len, len_a + D);
    } else {
        len_b = b_len;
        D = 1 + (int)(len_b * R);
        len_a = min(a_len, len_b + D);
    }
}

// DP cells the reference evaluates in rows 1..n: sum of min(len_b,i+D) - max(1,i-D) + 1 (seq_aligner.h:158-159)
__device__ __forceinline__ long long cells_upto(int n, int D, int len_b)
{
    long long t = max(len_b - D, 0), m = min((long long)n, t);
    long long f = m * (m + 1) / 2 + m * D + ((long long)n - m) * len_b;
    long long m2 = min(n, D + 1);
    long long g = m2;
    if (n > D + 1) { long long x = n - D; g += x * (x + 1) / 2 - 1; }
    return f - g + n;
}

// ---------------------------------------------------------------------------------------------
// K3a: prefix filter.  One thread replays rows 1..min(32,max_dst,len) of a candidate with a single
// 32-bit word and applies the reference's early-failure test (seq_aligner.h:185) exactly.  A candidate
// that fails here would have failed identically in the reference, so nothing is pruned that the
// reference would have kept; everything that survives goes to the full aligner in list order.
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
prefilter_kernel(const __grid_constant__ PairViews pv, const __grid_constant__ LocateView lv, int64_t ncand, double R, int maxn, int maxm,
                 uint8_t *__restrict__ survive, int32_t *__restrict__ rej_cells)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand) return;
    const int q = lv.d_cand_q[c];
    int k, t;
    int64_t ref_base = lv.ref_base;
    int ref_len = lv.ref_len;
    if (lv.d_cand_item) { // all-vs-all: candidates carry their (reference, read) pair and their trial number
        k = lv.d_cand_item[c];
        if (k < 0) { survive[c] = 0; rej_cells[c] = 0; return; } // slot of a dropped self hit
        t = q;
        const int T = lv.d_item_ref[k];
        ref_base = pv.B.base[T];
        ref_len = pv.B.len[T];
    } else {
        k = q / lv.ntrial;
        t = q - k * lv.ntrial;
    }
    const int r = lv.d_kept[k];
    CandView cv;
    derive_views(pv, lv, t, r, pv.A.len[r], pv.A.base[r], ref_base, ref_len, lv.d_cand_pos[c], false, cv);
    const SeqView &A = *cv.a, &B = *cv.b;
    const int a_len = cv.a_len, b_len = cv.b_len;
    const int64_t a_bit = cv.a_bit, b_bit = cv.b_bit;
    int len_a, len_b, D;
    derive_params(a_len, b_len, R, len_a, len_b, D);
    if (len_a >= maxn || D >= maxm) { // seq_aligner.h:104-107 (domain per SURVEY Q-D3)
        survive[c] = 0;
        rej_cells[c] = 0;
        return;
    }
    const int K = min(min(32, D), min(len_a, len_b));
    if (load_window(A.irr, A.nwords, a_bit) | load_window(B.irr, B.nwords, b_bit)) { // raw-byte compare needed: not here
        survive[c] = 1;
        rej_cells[c] = 0;
        return;
    }
    const uint32_t ahi = load_window(A.hi, A.nwords, a_bit), alo = load_window(A.lo, A.nwords, a_bit);
    const uint32_t bhi = load_window(B.hi, B.nwords, b_bit), blo = load_window(B.lo, B.nwords, b_bit);
    const uint32_t peq0 = ~bhi & ~blo, peq1 = ~bhi & blo, peq2 = bhi & ~blo, peq3 = bhi & blo;
    uint32_t Hp = 0xffffffffu, Hn = 0u;
    int cii = 0, fail = 0;
    for (int i = 1; i <= K; ++i) {
        const uint32_t h = (ahi >> (i - 1)) & 1u, l = (alo >> (i - 1)) & 1u;
        const uint32_t Eq = h ? (l ? peq3 : peq2) : (l ? peq1 : peq0);
        const uint32_t Xv = (((Eq & Hp) + Hp) ^ Hp) | Eq;
        const uint32_t Vp = Hn | ~(Xv | Hp), Vn = Hp & Xv;
        const uint32_t D0 = Xv | Hn;
        const uint32_t Xh = Eq | Hn;
        const uint32_t vps = (Vp << 1) | 1u, vns = Vn << 1;
        Hp = vns | ~(Xh | vps);
        Hn = vps & Xh;
        cii += 1 - (int)((D0 >> (i - 1)) & 1u);
        if (i > 10 && (double)cii > i * R) { fail = i; break; }
    }
    survive[c] = fail ? 0 : 1;
    rej_cells[c] = fail ? (int32_t)cells_upto(fail, D, len_b) : 0;
}

static PairViews pair_views(const SeqSets &ss)
{
    PairViews pv;
    pv.A = seq_view(ss.reads);
    pv.B = seq_view(ss.ref);
    pv.A2 = seq_view(ss.reads_rev ? ss.reads_rev : ss.reads);
    pv.B2 = seq_view(ss.ref_rev ? ss.ref_rev : ss.ref);
    return pv;
}

int pb_prefilter(pb_ctx *ctx, const SeqSets &ss, const LocateView &lv, int64_t ncand, double R,
                 int maxn, int maxm, uint8_t *d_survive, int32_t *d_rej_cells)
{
    if (ncand <= 0) return PB_OK;
    prefilter_kernel<<<(unsigned)((ncand + 255) / 256), 256, 0, ctx->stream>>>(pair_views(ss), lv, ncand, R, maxn, maxm, d_survive, d_rej_cells);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
}

// ---------------------------------------------------------------------------------------------
// K3: the banded aligner.  One warp = one alignment; S = band words per lane (band <= 1024*S bits).
// ---------------------------------------------------------------------------------------------

struct AlnRes {
    int ret, len_a, len_b, D, matlen_a, matlen_b, cost, diag_cost, nedit, fail_row;
    long long cells;
};

// ---- 1-D bulk TMA: global -> shared, completion on an mbarrier (cp.async.bulk, SASS UBLKCP) -----------------------
// Each warp stages the bit-plane words of its seg_b window with one bulk copy per plane instead of per-lane loads;
// the Eq planes are then built from shared memory.  16-byte aligned source / destination / size.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t phase)
{
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(phase)
                 : "memory");
    return ok != 0;
}

// ---- cp.async (LDGSTS): global -> shared without a register in between.  The traceback prefetches its parent windows into a
// shared-memory ring with it, because a prefetch that lands in registers has to be complete before the registers can be rotated
// to the next window (ncu: 64 % of the traceback's stalls were long-scoreboard waits on the first use of a window).  With 8-byte
// pairs and .ca copies this bought nothing (4 windows 92.2 ms, 6 windows 94.3, 10 windows 100.9 against 92.4 ms for the register
// walk on the same box: the path leaves its two predicted band words every 10-20 windows, everything fetched ahead is then thrown
// away, and each 8-byte pair cost a 128-byte line); with 16-byte units and L2-only .cg copies it does (see PB_UNIT16).
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src, int src_bytes) // src_bytes 0: zero-fill
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src, int src_bytes) // L2 only (.cg), src_bytes 0: zero-fill
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// a * b + c on the FMA pipe (IMAD) where the integer pipe is the bound
__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// multi-word adds: pb_carry_chain.cuh (generated), one asm statement per carry chain

// One band row for the S words of this lane.  pl points at this lane's first Eq word of the row's plane, sh is the
// row's bit offset inside those words; prow is this lane's pair column (row base + 2*lane) of the row's parent block.  Returns the D0 word of
// slot sd (the main diagonal lives there in one lane); leaves the row's vertical deltas in Vp/Vn.
template <int S>
__device__ __forceinline__ uint32_t row_step(uint32_t (&Hp)[S], uint32_t (&Hn)[S], const uint32_t (&keep)[S], uint32_t (&Vp)[S],
                                             uint32_t (&Vn)[S], const uint32_t *__restrict__ pl, int thrs, unsigned sh, int lane,
                                             int sd, bool stores, uint32_t *__restrict__ prow, int LN = 32,
                                             int tail_off = 0)
{
    // phase A: slide the band one bit (across words and lanes), fetch Eq, block add with carry-in 0
    uint32_t nx = __shfl_down_sync(FULL, (Hp[0] & 1u) | ((Hn[0] & 1u) << 1), 1);
    if (lane == 31) nx = 1u;
    uint32_t Eq[S], x[S], sum[S];
    // Eq words: logical words x0..x0+S of the plane; for even S the plane is stored with one pad word per S words
    // (bank-conflict-free for the lane stride S), which shows up here as a +1 from slot `thrs` on
    constexpr bool PAD = (S % PB_PAD_MOD) == 0;
    uint32_t plw = pl[0];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const uint32_t hp_hi = (s + 1 < S) ? Hp[s + 1] : (nx & 1u);
        const uint32_t hn_hi = (s + 1 < S) ? Hn[s + 1] : (nx >> 1);
        Hp[s] = __funnelshift_r(Hp[s], hp_hi, 1);
        Hn[s] = __funnelshift_r(Hn[s], hn_hi, 1);
        const uint32_t nxt = PAD ? pl[s + 1 + ((s + 1 >= thrs) ? 1 : 0)] : pl[s + 1];
        Eq[s] = __funnelshift_r(plw, nxt, sh) & keep[s]; // no matches beyond the band's upper edge (see align_one)
        plw = nxt;
        x[s] = Eq[s] & Hp[s];
    }
    const uint32_t carry = CarryChain<S>::add(sum, x, Hp);
    uint32_t ones = sum[0];
#pragma unroll
    for (int s = 1; s < S; ++s) ones &= sum[s];
    const uint32_t G = __ballot_sync(FULL, carry);
    const uint32_t P = __ballot_sync(FULL, ones == 0xffffffffu);
    const uint32_t cin = ((((G | P) + G) ^ P) >> lane) & 1u; // carry into this lane's block
    CarryChain<S>::inc(sum, cin);

    // phase B: vertical deltas, D0, MATCH plane
    uint32_t d0w = 0u;
    uint32_t Mw[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const uint32_t Xv = (sum[s] ^ Hp[s]) | Eq[s];
        Vp[s] = Hn[s] | ~(Xv | Hp[s]);
        Vn[s] = Hp[s] & Xv;
        const uint32_t D0 = Xv | Hn[s];
        Mw[s] = Eq[s] | ~D0;
        if (s == sd) d0w = D0;
    }
    uint32_t pv = __shfl_up_sync(FULL, (Vp[S - 1] >> 31) | ((Vn[S - 1] >> 31) << 1), 1);
    if (lane == 0) pv = 1u; // vin = +1 at the band's left edge (and at column 0)

    // phase C: new horizontal deltas, INSERT plane
    uint32_t pprev = pv << 31, nprev = (pv >> 1) << 31; // bit 31 = delta entering this word from the left
    uint32_t heldM = 0u, heldI = 0u;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const uint32_t vps = __funnelshift_l(pprev, Vp[s], 1), vns = __funnelshift_l(nprev, Vn[s], 1);
        pprev = Vp[s];
        nprev = Vn[s];
        const uint32_t Xh = Eq[s] | Hn[s];
        Hp[s] = vns | ~(Xh | vps);
        Hn[s] = vps & Xh;
        // parents of band word w = lane*S+s: {MATCH plane, INSERT plane} as one 8-byte pair at pair index s*32+lane
        // lanes whose words all lie past the band skip the store; a partly used lane writes its S words (row padding)
        // parents as 16-byte units of two adjacent band words per lane: {M[2p], I[2p], M[2p+1], I[2p+1]} at unit p*LN + lane
        // (one STG.128 per slot pair); the last slot of an odd S follows as 8-byte pairs at tail_off.  prow = row base + 4*lane.
        if ((s & 1) == 0 && s + 1 < S) {
            heldM = Mw[s]; heldI = Hp[s];
        } else if (s & 1) {
            if (stores) reinterpret_cast<uint4 *>(prow)[(s >> 1) * LN] = make_uint4(heldM, heldI, Mw[s], Hp[s]);
        } else {
            if (stores) *reinterpret_cast<uint2 *>(prow + tail_off) = make_uint2(Mw[s], Hp[s]);
        }
    }
    return d0w;
}

// goal_cell + coverage test + find_path for one finished forward pass (seq_aligner.h:111-116,191-233); called by the
// whole warp with warp-uniform arguments.  hp_words / hn_words: the final row's horizontal deltas in shared memory (band
// word w at [w]); par_pair(row, w): the {MATCH word, INSERT word} pair of band word w of DP row `row` (zero outside).
// par_addr(row, w): the pair's address in global memory (NULL outside) and ring: PB_TB_RING_WORDS words of this warp's shared
// memory for the asynchronous window prefetch; ring == NULL walks with register prefetch through par_pair instead.
struct NoUnit { __device__ void operator()(int, int &, int &) const {} };
struct NoUnitLoad { __device__ uint4 operator()(int, int, int) const { return make_uint4(0u, 0u, 0u, 0u); } };
template <bool UNITS, class PairAt, class PairAddr, class UnitOf = NoUnit, class UnitLoad = NoUnitLoad>
__device__ __forceinline__ void finish_alignment(int len_a, int len_b, int D, int a_len, double R, int cii, int colbest, int col_i,
                                                 const uint32_t *hp_words, const uint32_t *hn_words, PairAt par_pair, PairAddr par_addr,
                                                 uint32_t *ring, const void *gbase,
                                                 uint8_t *__restrict__ opsrev, uint8_t *__restrict__ ops_out, AlnRes &res,
                                                 UnitOf unit_of = UnitOf(), UnitLoad unit_load = UnitLoad())
{
    const int lane = threadIdx.x & 31;
    // ---- goal_cell, seq_aligner.h:191-213
    int matlen_a, matlen_b, cost;
    if (len_a > len_b) {
        matlen_a = col_i; matlen_b = len_b; cost = colbest;
    } else {
        // last row: cost(len_a, j) for j in (len_a, len_b] from the final horizontal deltas; earliest strict minimum
        matlen_a = len_a; matlen_b = len_a; cost = cii;
        if (lane == 0) {
            int c = cii;
            for (int j = len_a + 1; j <= len_b; ++j) {
                const int k = j - len_a + D;
                c += (int)((hp_words[k >> 5] >> (k & 31)) & 1u) - (int)((hn_words[k >> 5] >> (k & 31)) & 1u);
                if (c < cost) { cost = c; matlen_b = j; }
            }
        }
        cost = __shfl_sync(FULL, cost, 0);
        matlen_b = __shfl_sync(FULL, matlen_b, 0);
    }
    res.matlen_a = matlen_a; res.matlen_b = matlen_b; res.cost = cost;
    res.diag_cost = (a_len <= len_a && a_len <= len_b) ? cii : 0; // locator.cpp:86 (Q-L2)
    if ((double)matlen_b < len_b * (1 - R)) return; // seq_aligner.h:114

    // ---- find_path, seq_aligner.h:214-233: walk the parent planes back from the goal cell.
    // Warp-cooperative: lane r holds the parent pairs of row i0-r around the path's band position (2 band words,
    // 16 bytes), the next 64 rows are prefetched while the current ones are walked, and runs of MATCH along a
    // diagonal (same band bit, consecutive rows) are found with one ballot and written by as many lanes.
    __syncwarp();
    // window = the band word under the path plus the neighbour the path is closer to
    auto window_base = [](int k) -> int { return (k >> 5) - ((k & 31) < 16 ? 1 : 0); };
    int n = 0;
    if (UNITS && ring) {
        // ---- units + asynchronous prefetch: slot q of the ring holds, per lane, the primary unit (16 bytes at ring[q*256 + 4*lane])
        // and the secondary one (ring[q*256 + 128 + 4*lane]); meta[8q..] = i0, b0, n0, b1, n1
        int i = matlen_a, j = matlen_b;
        const int guard = len_a + len_b + 1;
        int *meta = reinterpret_cast<int *>(ring + PB_TB_RING * 256);
        auto fetch = [&](int slot, int i0w, int pb, int pn, int sb, int sn) {
            const int row = i0w - lane;
            const uint2 *q0 = par_addr(row, pb);
            uint32_t *dst = ring + slot * 256 + 4 * lane;
            // rows above the matrix get plain zero stores, not zero-fill copies: a copy whose source is ignored may still be
            // handed a meaningless address by the assembler, and the hardware faults on it
            PB_CHECK_RANGE("ring slot", dst, 16, ring, ring + PB_TB_RING * 256);
            if (q0) PB_CHECK_RANGE("traceback prefetch", q0, pn == 2 ? 16 : 8, gbase, opsrev);
            if (!q0) *reinterpret_cast<uint4 *>(dst) = make_uint4(0u, 0u, 0u, 0u);
            else if (pn == 2) cp_async16(dst, q0, 16);
            else cp_async8(dst, q0, 8);
            if (sn) {
                const uint2 *q1 = par_addr(row, sb);
                if (q1) PB_CHECK_RANGE("traceback prefetch (2nd unit)", q1, sn == 2 ? 16 : 8, gbase, opsrev);
                if (!q1) *reinterpret_cast<uint4 *>(dst + 128) = make_uint4(0u, 0u, 0u, 0u);
                else if (sn == 2) cp_async16(dst + 128, q1, 16);
                else cp_async8(dst + 128, q1, 8);
            }
            if (lane == 0) { int *m = meta + 8 * slot; m[0] = i0w; m[1] = pb; m[2] = pn; m[3] = sb; m[4] = sn; }
            cp_async_commit();
        };
        int cur_slot = 0, cur_i0 = 0, wend = 0, cb0 = 0, cn0 = 0, cb1 = 0, cn1 = 0;
        uint4 U0 = make_uint4(0u, 0u, 0u, 0u), U1 = U0;
        bool have = false;
        while (i > 0 && j > 0 && n < guard) {
            const int k = j - i + D, w = k >> 5, kb = k & 31;
            if (!have || i <= wend || !((unsigned)(w - cb0) < (unsigned)cn0 || (unsigned)(w - cb1) < (unsigned)cn1)) {
                const int nslot = cur_slot + 1 == PB_TB_RING ? 0 : cur_slot + 1;
                bool usual = have;
                if (usual) {
                    const int *m = meta + 8 * nslot;
                    usual = m[0] == i && ((unsigned)(w - m[1]) < (unsigned)m[2] || (unsigned)(w - m[3]) < (unsigned)m[4]);
                }
                __syncwarp();
                if (usual) { // refill the slot just walked with the window PB_TB_RING-1 ahead, same units as the newest one
                    const int last = cur_slot == 0 ? PB_TB_RING - 1 : cur_slot - 1;
                    const int *m = meta + 8 * last;
                    const int pb = m[1], pn = m[2], sb = m[3], sn = m[4];
                    __syncwarp();
                    fetch(cur_slot, i - 32 * (PB_TB_RING - 1), pb, pn, sb, sn);
                    cur_slot = nslot;
                } else { // cold start, or the path left the predicted units
                    cp_async_wait<0>();
                    int pb, pn, sb = 0, sn = 0;
                    unit_of(w, pb, pn);
                    const int pos = (w - pb) * 32 + kb;
                    if (pos < 8 && pb > 0) unit_of(pb - 1, sb, sn);
                    else if (pos >= 32 * pn - 8 && 32 * (pb + pn) <= 2 * D) unit_of(pb + pn, sb, sn);
#pragma unroll
                    for (int t = 0; t < PB_TB_RING; ++t) fetch(t, i - 32 * t, pb, pn, sb, sn);
                    cur_slot = 0;
                }
                cp_async_wait<PB_TB_RING - 1>();
                __syncwarp();
                const int *m = meta + 8 * cur_slot;
                cur_i0 = m[0]; cb0 = m[1]; cn0 = m[2]; cb1 = m[3]; cn1 = m[4];
                U0 = *reinterpret_cast<const uint4 *>(ring + cur_slot * 256 + 4 * lane);
                U1 = cn1 ? *reinterpret_cast<const uint4 *>(ring + cur_slot * 256 + 128 + 4 * lane) : make_uint4(0u, 0u, 0u, 0u);
                wend = cur_i0 - 32;
                have = true;
            }
            const int r0 = cur_i0 - i;
            const int d0 = w - cb0;
            const bool prim = (unsigned)d0 < (unsigned)cn0;
            const int d = prim ? d0 : w - cb1;
            const uint4 u = prim ? U0 : U1;
            const uint32_t mword = d ? u.z : u.x;
            const uint32_t B = __ballot_sync(FULL, (mword >> kb) & 1u) >> r0;
            int run = (~B) ? __ffs(~B) - 1 : 32;
            const int lim = min(32 - r0, min(i, j));
            const bool indel = run < lim;
            run = min(run, lim);
            if (lane < run) opsrev[n + lane] = (uint8_t)PB_MATCH;
            n += run; i -= run; j -= run;
            if (indel) {
                const uint32_t iword = d ? u.w : u.y;
                const uint32_t hb = (__shfl_sync(FULL, iword, r0 + run) >> kb) & 1u;
                if (lane == 0) opsrev[n] = (uint8_t)(hb ? PB_INSERT : PB_DELETE);
                ++n;
                if (hb) --j; else --i;
            }
        }
        cp_async_wait<0>();
        __syncwarp();
        if (n < guard) {
            if (i == 0 && j > 0) {
                for (int t = lane; t < j; t += 32) opsrev[n + t] = (uint8_t)PB_INSERT;
                n += j;
            } else if (j == 0 && i > 0) {
                for (int t = lane; t < i; t += 32) opsrev[n + t] = (uint8_t)PB_DELETE;
                n += i;
            }
        }
    } else {
        int i = matlen_a, j = matlen_b;
        const int guard = len_a + len_b + 1; // a path can never be longer; keeps a corrupted plane from hanging the GPU
        // windows in flight: [0] is the one being walked, [1..] are fetched ahead (DRAM latency under load is several windows long)
        int wi0[PB_TB_WINDOWS], wwb[PB_TB_WINDOWS];
        uint2 wp0[PB_TB_WINDOWS], wp1[PB_TB_WINDOWS];
#pragma unroll
        for (int t = 0; t < PB_TB_WINDOWS; ++t) { wi0[t] = -1; wwb[t] = 0; wp0[t] = make_uint2(0u, 0u); wp1[t] = wp0[t]; }
        int kbase = 0, wend = 0; // the current window covers band bits [kbase, kbase+64) of rows (wend, wend+32]
        bool have = false;
        while (i > 0 && j > 0 && n < guard) {
            const int k = j - i + D;
            if (!have || i <= wend || (unsigned)(k - kbase) >= 64u) {
                const int w = k >> 5, wb = window_base(k);
                if (have && wi0[1] == i && w >= wwb[1] && w <= wwb[1] + 1) { // the usual case: 32 rows consumed, prediction held
#pragma unroll
                    for (int t = 0; t + 1 < PB_TB_WINDOWS; ++t) { wi0[t] = wi0[t + 1]; wwb[t] = wwb[t + 1]; wp0[t] = wp0[t + 1]; wp1[t] = wp1[t + 1]; }
                    wi0[PB_TB_WINDOWS - 1] = wi0[0] - 32 * (PB_TB_WINDOWS - 1);
                    wwb[PB_TB_WINDOWS - 1] = wb;
                    wp0[PB_TB_WINDOWS - 1] = par_pair(wi0[PB_TB_WINDOWS - 1] - lane, wb);
                    wp1[PB_TB_WINDOWS - 1] = par_pair(wi0[PB_TB_WINDOWS - 1] - lane, wb + 1);
                } else { // cold start or the path left the predicted words: fetch now
#pragma unroll
                    for (int t = 0; t < PB_TB_WINDOWS; ++t) {
                        wi0[t] = i - 32 * t; wwb[t] = wb;
                        wp0[t] = par_pair(wi0[t] - lane, wb);
                        wp1[t] = par_pair(wi0[t] - lane, wb + 1);
                    }
                }
                have = true;
                kbase = 32 * wwb[0];
                wend = wi0[0] - 32;
            }
            const int r0 = wi0[0] - i;   // lane that holds the current row
            const int kb = k - kbase;    // 0..63: bit inside the two-word window
            const uint32_t mword = kb < 32 ? wp0[0].x : wp1[0].x;
            const uint32_t B = __ballot_sync(FULL, (mword >> (kb & 31)) & 1u) >> r0; // bit t: cell (i-t, j-t) is MATCH
            int run = (~B) ? __ffs(~B) - 1 : 32;
            const int lim = min(32 - r0, min(i, j)); // rows left in this window / cells left on this diagonal
            const bool indel = run < lim;            // the run ends on a non-MATCH cell that this window still holds
            run = min(run, lim);
            if (lane < run) opsrev[n + lane] = (uint8_t)PB_MATCH;
            n += run; i -= run; j -= run;
            if (indel) { // same diagonal, so same band word and bit: the cell's pair sits in lane r0 + run
                const uint32_t iword = kb < 32 ? wp0[0].y : wp1[0].y;
                const uint32_t hb = (__shfl_sync(FULL, iword, r0 + run) >> (kb & 31)) & 1u;
                if (lane == 0) opsrev[n] = (uint8_t)(hb ? PB_INSERT : PB_DELETE);
                ++n;
                if (hb) --j; else --i;
            }
        }
        if (n < guard) {
            if (i == 0 && j > 0) { // init_cell row 0: INSERT all the way
                for (int t = lane; t < j; t += 32) opsrev[n + t] = (uint8_t)PB_INSERT;
                n += j;
            } else if (j == 0 && i > 0) { // init_cell column 0: DELETE all the way
                for (int t = lane; t < i; t += 32) opsrev[n + t] = (uint8_t)PB_DELETE;
                n += i;
            }
        }
    }
    __syncwarp();
    if (ops_out)
        for (int k = lane; k < n; k += 32) ops_out[k] = __ldcg(opsrev + (n - 1 - k));
    res.nedit = n;
    res.ret = matlen_b;
}

// Inlined into its one call site per kernel: as a separate function the row loop re-materialised the global-store descriptor
// from vector registers before every store (two R2UR per parent pair); measured 95.0 -> 91.2 ms of K3 per config-2 step.
template <int S, bool IRR>
__device__ __forceinline__ void align_one(const SeqView &A, int64_t a_bit, int a_len, uint32_t a_tab, const SeqView &B, int64_t b_bit, int b_len,
                                       double R, int maxn, int maxm, uint32_t *__restrict__ planes, int PW,
                                       uint32_t *__restrict__ par, uint8_t *__restrict__ opsrev, uint8_t *__restrict__ ops_out,
                                       uint32_t *__restrict__ raw, int RW, uint64_t *bar, uint32_t &phase,
                                       AlnRes &res)
{
    constexpr int T = 32 * S;
    const int lane = threadIdx.x & 31;
    int len_a, len_b, D;
    derive_params(a_len, b_len, R, len_a, len_b, D);
    res.ret = -1; res.len_a = len_a; res.len_b = len_b; res.D = D;
    res.matlen_a = res.matlen_b = res.cost = res.diag_cost = res.nedit = res.fail_row = 0;
    res.cells = 0;
    if (len_a >= maxn || D >= maxm) return; // seq_aligner.h:104-107


    // ---- Eq planes of seg_b in shared memory: plane c, bit t <-> (b[t - D] == c), zero outside [0,len_b)
    constexpr bool PAD = (S % PB_PAD_MOD) == 0; // see row_step: physical index of logical word x is x + x/S in the padded classes
    const int PWn = ((len_a + 31) >> 5) + T + 1;
    // The plane words that cover seg_b = line bits [b_bit, b_bit + len_b) come in through a small staging buffer filled by
    // bulk TMA copies (16-byte granules), PB_STAGE_WORDS words of each plane at a time; consecutive chunks overlap by four
    // words because a 32-bit window straddles two words.
    const int64_t g0 = b_bit - D;             // line bit of bit 0 of plane word 0 (may be negative)
    const int64_t w_first = max((int64_t)0, g0 >> 5) & ~(int64_t)3; // first staged line word
    const int64_t w_end = min(B.nwords, (((b_bit + len_b + 31) >> 5) + 1 + 3) & ~(int64_t)3);
    const int k0 = (int)((g0 >> 5) - w_first);                   // raw word index (relative to w_first) under plane word 0
    for (int64_t cw = w_first; cw < w_end; cw += PB_STAGE_WORDS - 4) {
        const int n_raw = (int)min((int64_t)PB_STAGE_WORDS, w_end - cw);
        __syncwarp();
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // earlier generic reads of `raw` before the async writes
            const uint32_t bytes = (uint32_t)n_raw * 4u;
            mbar_expect_tx(bar, bytes * (IRR ? 3u : 2u));
            tma_load_1d(raw, B.hi + cw, bytes, bar);
            tma_load_1d(raw + RW, B.lo + cw, bytes, bar);
            if (IRR) tma_load_1d(raw + 2 * RW, B.irr + cw, bytes, bar);
        }
        int spins = 0;
        while (!mbar_try_wait(bar, phase)) {
            if (++spins > (1 << 22)) __trap(); // a copy that never lands must not hang the GPU
        }
        phase ^= 1u;
        // plane words whose first raw word lies in this chunk (the last chunk also takes what is left)
        const int c_lo = (int)(cw - w_first), c_hi = (cw + PB_STAGE_WORDS - 4 >= w_end) ? INT_MAX : c_lo + PB_STAGE_WORDS - 4;
        const int x_lo = (cw == w_first) ? 0 : max(0, c_lo - k0); // the first chunk also takes the words in front of the line
        const int x_hi = (c_hi == INT_MAX) ? PWn : min(PWn, max(0, c_hi - k0));
        for (int x = x_lo + lane; x < x_hi; x += 32) {
            const int bidx0 = 32 * x - D; // b index of bit 0 of this word
            uint32_t valid;
            if (bidx0 >= len_b || bidx0 + 31 < 0) valid = 0u;
            else {
                valid = 0xffffffffu;
                if (bidx0 < 0) valid &= 0xffffffffu << (-bidx0);
                if (bidx0 + 31 >= len_b) valid &= 0xffffffffu >> (bidx0 + 32 - len_b);
            }
            uint32_t hi = 0u, lo = 0u;
            if (valid) {
                const int wi = x + k0 - c_lo; // index into the staged chunk
                const unsigned sh = (unsigned)(g0 & 31);
                auto win = [&](const uint32_t *pl) -> uint32_t {
                    const uint32_t w0 = (wi >= 0 && wi < n_raw) ? pl[wi] : 0u;
                    const uint32_t w1 = (wi + 1 >= 0 && wi + 1 < n_raw) ? pl[wi + 1] : 0u;
                    return __funnelshift_r(w0, w1, sh);
                };
                hi = win(raw);
                lo = win(raw + RW);
                if (IRR) valid &= ~win(raw + 2 * RW); // a non-ACGT byte equals none of A,C,G,T
            }
        const int px = PAD ? x + x / S : x;
        planes[0 * PW + px] = ~hi & ~lo & valid;
        planes[1 * PW + px] = ~hi & lo & valid;
        planes[2 * PW + px] = hi & ~lo & valid;
        planes[3 * PW + px] = hi & lo & valid;
        if (IRR) {
            planes[4 * PW + px] = 0u; planes[5 * PW + px] = 0u; planes[6 * PW + px] = 0u; planes[7 * PW + px] = 0u;
        }
        }
    }
    if (IRR) { // planes 4..7: seg_b equals the k-th non-ACGT byte value of seg_a's sequence (raw compare, seq_aligner.h:136)
        __syncwarp();
        const int64_t e0 = exc_lower_bound(B, b_bit), e1 = exc_lower_bound(B, b_bit + len_b);
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const int slot = tab_slot(a_tab, B.exc_val[e]);
            if (slot >= 0) { // unused table slots hold 'A', which never shows up in the exception list
                const int t = (int)(B.exc_pos[e] - b_bit) + D;
                const int x = t >> 5, px = PAD ? x + x / S : x;
                atomicOr(&planes[(4 + slot) * PW + px], 1u << (t & 31));
            }
        }
    }
    __syncwarp();

    // ---- row 0: h = -1 for (fake) columns j <= 0, +1 for j >= 1.
    // Upper band edge: cell (i, i+D) must not see its "up" neighbour (seq_aligner.h:170).  Band bits k > 2D are
    // computed like ordinary cells but with Eq forced to 0; by induction over rows their cost is then
    // cost(i, i+D) + (k - 2D), i.e. h = +1 all the way up, so the bit that slides into k = 2D every row is the
    // +1 a pinned edge would hold and U+1 = Dg+2 never beats Dg+m.  (tools/bitpar_model.c checks this.)
    uint32_t Hp[S], Hn[S], keep[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int k0 = 32 * (lane * S + s);
        uint32_t hn;
        if (k0 + 31 <= D) hn = 0xffffffffu;
        else if (k0 > D) hn = 0u;
        else hn = 0xffffffffu >> (31 - (D - k0));
        Hn[s] = hn;
        Hp[s] = ~hn;
        uint32_t kp; // bits k <= 2D
        if (k0 > 2 * D) kp = 0u;
        else if (k0 + 31 <= 2 * D) kp = 0xffffffffu;
        else kp = 0xffffffffu >> (31 - (2 * D - k0));
        keep[s] = kp;
    }
    const int wd = D >> 5, Ld = wd / S, sd = wd % S; // owner of the main-diagonal bit k = D
    const int LN = 32;
    const bool stores = true; // measured: predicating the pair stores costs 40 % (A/B on B200), the padding writes are cheaper
    const size_t rstride = (size_t)2 * LN * S; // words per parent row
    const int lane_off = 4 * lane, tail_off = (S / 2) * LN * 4 - 2 * lane; // units of 4 words per lane, then the odd slot's pairs

    int cii = 0;                          // cost(i,i), warp-uniform, advanced once per 32-row block
    int colc = 0, colbest = 0, col_i = 0; // cost(i,len_b) tracking when len_a > len_b
    int fail_row = 0;
    int rows_done = 0;
    uint32_t Vp[S], Vn[S];

    // ---- rows 1..min(len_a,len_b) in blocks of 32: the early-failure test (seq_aligner.h:185) is evaluated once
    // per block from the block's 32 diagonal D0 bits; a failing block reports its first failing row exactly.
    const int nfast = min(len_a, len_b);
    uint32_t awh_n = load_window(A.hi, A.nwords, a_bit), awl_n = load_window(A.lo, A.nwords, a_bit);
    for (int i0 = 1; i0 <= nfast; i0 += 32) {
        const int tmax = min(32, nfast - i0 + 1);
        const uint32_t awh = awh_n, awl = awl_n; // this block's 32 bases of seg_a (warp-uniform); fetch the next block's now
        awh_n = load_window(A.hi, A.nwords, a_bit + i0 + 31);
        awl_n = load_window(A.lo, A.nwords, a_bit + i0 + 31);
        const uint32_t awi = IRR ? load_window(A.irr, A.nwords, a_bit + i0 - 1) : 0u;
        const int thr = (int)((i0 + lane) * R); // cost > i*R  <=>  cost > floor(i*R) for an integer cost
        const int q = (i0 - 1) >> 5; // first plane word of the block's rows (logical index)
        const uint32_t *plq = planes + q + lane * S + (PAD ? lane + q / S : 0);
        const int thrs = S - q % S;
        uint32_t *prow = par + (size_t)(i0 - 1) * rstride + lane_off;
        PB_CHECK_RANGE("parent rows of a block", par + (size_t)(i0 - 1) * rstride, 16, par, opsrev);
        PB_CHECK_RANGE("parent rows of a block (end)", par + (size_t)(i0 - 1 + tmax) * rstride - 4, 4, par, opsrev);
        uint32_t hist = 0u;
        for (int t = 0; t < tmax; ++t) {
            int ca = (int)(((awh >> t) & 1u) * 2u + ((awl >> t) & 1u));
            if (IRR && ((awi >> t) & 1u)) ca = irr_plane(A, a_bit + i0 - 1 + t, a_tab);
            const uint32_t d0w = row_step<S>(Hp, Hn, keep, Vp, Vn, plq + ca * PW, thrs, (unsigned)t, lane, sd, stores, prow, LN, tail_off);
            hist = __funnelshift_r(hist, d0w >> (D & 31), 1); // row t's diagonal D0 bit enters at bit 31 (meaningful in the diagonal's owner lane)
            prow += rstride;
        }
        hist >>= 32 - tmax; // row t of the block now sits at bit t
        hist = __shfl_sync(FULL, hist, Ld);
        const int cdiag = cii + (lane + 1) - __popc(hist & (0xffffffffu >> (31 - lane))); // cost(i0+lane, i0+lane)
        const uint32_t badm = __ballot_sync(FULL, lane < tmax && i0 + lane > 10 && cdiag > thr);
        if (badm) {
            fail_row = i0 + __ffs(badm) - 1;
            rows_done = fail_row;
            break;
        }
        cii += tmax - __popc(hist & (0xffffffffu >> (32 - tmax)));
        rows_done = i0 + tmax - 1;
    }

    // ---- rows len_b+1..len_a (only when seg_a outruns seg_b): no failure test there (the cell is never written,
    // Q-D2); follow cost(i, len_b) down the last column through the vertical deltas
    if (!fail_row && len_a > len_b) {
        colc = colbest = cii;
        col_i = len_b;
        uint32_t awh = 0u, awl = 0u, awi = 0u;
        for (int i = len_b + 1; i <= len_a; ++i) {
            const int t = (i - 1) & 31;
            if (t == 0 || i == len_b + 1) {
                awh = load_window(A.hi, A.nwords, a_bit + (i - 1 - t));
                awl = load_window(A.lo, A.nwords, a_bit + (i - 1 - t));
                if (IRR) awi = load_window(A.irr, A.nwords, a_bit + (i - 1 - t));
            }
            int ca = (int)(((awh >> t) & 1u) * 2u + ((awl >> t) & 1u));
            if (IRR && ((awi >> t) & 1u)) ca = irr_plane(A, a_bit + i - 1, a_tab);
            const int q = (i - 1) >> 5;
            row_step<S>(Hp, Hn, keep, Vp, Vn, planes + ca * PW + q + lane * S + (PAD ? lane + q / S : 0), S - q % S, (unsigned)t,
                        lane, sd, stores, par + (size_t)(i - 1) * rstride + lane_off, LN, tail_off);
            const int k = len_b - i + D, wk = k >> 5, Lk = wk / S, sk = wk % S;
            uint32_t vpw = 0u, vnw = 0u;
#pragma unroll
            for (int s = 0; s < S; ++s)
                if (s == sk) { vpw = Vp[s]; vnw = Vn[s]; }
            vpw = __shfl_sync(FULL, vpw, Lk);
            vnw = __shfl_sync(FULL, vnw, Lk);
            colc += (int)((vpw >> (k & 31)) & 1u) - (int)((vnw >> (k & 31)) & 1u);
            if (colc < colbest) { colbest = colc; col_i = i; }
            rows_done = i;
        }
    }
    res.cells = cells_upto(rows_done, D, len_b);
    if (fail_row) { res.fail_row = fail_row; return; }

    // ---- goal_cell, coverage test, find_path
    __syncwarp();
#pragma unroll
    for (int s = 0; s < S; ++s) {
        planes[lane * S + s] = Hp[s];
        planes[T + lane * S + s] = Hn[s];
    }
    __syncwarp();
    auto par_addr = [&](int row, int w) -> const uint2 * { // {MATCH word, INSERT word} of band word w of DP row `row`
        if (row < 1 || w < 0 || 32 * w > 2 * D) return nullptr;
        const int L = w / S, s = w - L * S;
        const uint32_t *rb = par + (size_t)(row - 1) * rstride;
        if (s < (S & ~1)) return reinterpret_cast<const uint2 *>(rb + ((s >> 1) * LN + L) * 4 + (s & 1) * 2);
        return reinterpret_cast<const uint2 *>(rb + (S / 2) * LN * 4 + L * 2);
    };
    // the unit that holds band word w: its first word and its width in words (2: a 16-byte unit, 1: a lone pair)
    auto unit_of = [&](int w, int &b, int &n) {
        const int s = w % S;
        if (s < (S & ~1)) { b = w - (s & 1); n = 2; } else { b = w; n = 1; }
    };
    auto unit_load = [&](int row, int b, int n) -> uint4 { // one load per row: 16 bytes for a two-word unit, 8 for a lone pair
        const uint2 *q = par_addr(row, b);
        if (!q) return make_uint4(0u, 0u, 0u, 0u);
        if (n == 2) return __ldcg(reinterpret_cast<const uint4 *>(q));
        const uint2 v = __ldcg(q);
        return make_uint4(v.x, v.y, 0u, 0u);
    };
    auto par_pair = [&](int row, int w) -> uint2 {
        const uint2 *q = par_addr(row, w);
        return q ? __ldcg(q) : make_uint2(0u, 0u);
    };
    uint32_t *ring = planes + 2 * T; // behind the final deltas; the Eq planes and the staging area are dead by now
    finish_alignment<true>(len_a, len_b, D, a_len, R, cii, colbest, col_i, planes, planes + T, par_pair, par_addr, ring, par, opsrev, ops_out, res,
                           unit_of, unit_load);
}

#include "pb_align_nb.cuh"

struct AlignLaunch {
    SeqView A, B;   // pairs mode: seg_a / seg_b sets; locate modes: reads / ref
    SeqView A2, B2; // overlap mode: reversed reads / ref
    double R;
    int maxn, maxm;
    int PW;              // plane stride (words) for this launch
    int RW;              // staging stride (words) per raw plane
    int warp_words;      // shared-memory words per warp (planes + staging)
    int wpb;             // warps per CTA of this launch
    size_t slot_words;   // scratch words per warp slot
    size_t par_words;    // of which parent planes
    uint32_t *scratch;
    int *queue;          // work counter
    const int32_t *order; // item ids, longest first
    int nitems;
    uint8_t *ops;
    const int64_t *ops_off;
    unsigned long long *stats; // [0] reference-equivalent DP cells of the alignments K3 ran, [1] alignments run by K3, [2] band
                               // cells actually computed (strip width x rows in the first pass), [3] items redone (may be NULL)
    uint8_t *redo;             // two-pass locate: per item, 1 = the full-band kernel must (still) run it.  The first pass (strip)
                               // clears or keeps it; the second pass skips items whose flag is 0.  NULL: single pass
    int g256;                  // first pass: goal-side width of the strip in 1/256 of max_dst (see nb_target)
};

// resident CTAs per SM the register allocation is held to (more warps hide the shuffle / ballot latencies of a row)
template <int S> struct MinBlocks { static constexpr int v = (S <= 3 ? PB_MINB3 : (S <= 5 ? 5 : (S <= 12 ? 4 : 3))) * 4 / ALIGN_WPB; };

// PAIRS = all-vs-all items (pb_overlap_all_run): a separate instantiation, so the single-reference kernels keep their
// register allocation
template <int S, bool IRR, bool PAIRS>
__global__ void __launch_bounds__(ALIGN_WPB * 32, MinBlocks<S>::v)
align_locate_kernel(const __grid_constant__ AlignLaunch p, const __grid_constant__ LocateView lv, const uint8_t *__restrict__ survive,
                    const int32_t *__restrict__ rej_cells,
                    pb_locate_rec *__restrict__ recs)
{
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ __align__(8) uint64_t bars[ALIGN_WPB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per warp: Eq planes, then the TMA staging area for seg_b's raw plane words
    uint32_t *planes = smem + (size_t)warp * p.warp_words;
    uint32_t *raw = planes + (size_t)(IRR ? 8 : 4) * p.PW;
    uint64_t *bar = &bars[warp];
    uint32_t phase = 0u;
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    const size_t slot = (size_t)blockIdx.x * p.wpb + warp;
    uint32_t *par = p.scratch + slot * p.slot_words;
    uint8_t *opsrev = reinterpret_cast<uint8_t *>(par + p.par_words);
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(p.queue, 1);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= p.nitems) break;
        const int k = p.order[idx];
        if (p.redo && !p.redo[k]) continue; // certified by the first pass
        const int r = lv.d_kept[k];
        const int rlen = p.A.len[r];
        const int64_t rbase = p.A.base[r];
        // all-vs-all: item = (reference T, read r) with its own candidate range
        const int64_t c0 = PAIRS ? lv.d_item_beg[k] : lv.d_qoff[(int64_t)k * lv.ntrial];
        const int64_t c1 = PAIRS ? lv.d_item_end[k] : lv.d_qoff[(int64_t)(k + 1) * lv.ntrial];
        const int T = PAIRS ? lv.d_item_ref[k] : 0;
        const int64_t ref_base = PAIRS ? p.B.base[T] : lv.ref_base;
        const int ref_len = PAIRS ? p.B.len[T] : lv.ref_len;
        long long cells = 0;
        int ncand = 0;
        bool found = false;
        AlnRes res;
        int win_j = 0, win_pos = 0, win_rpos = 0, win_dir = 0;
        for (int64_t cb = c0; cb < c1 && !found; cb += 32) {
            const int64_t c = cb + lane;
            const int nvalid = (int)min((int64_t)32, c1 - cb);
            int sv = 0, rc = 0;
            if (lane < nvalid) { sv = survive[c]; rc = rej_cells[c]; }
            uint32_t mask = __ballot_sync(FULL, sv);
            int done = 0; // lanes [0,done) already accounted
            while (mask) {
                const int f = __ffs(mask) - 1;
                cells += __reduce_add_sync(FULL, (lane >= done && lane < f) ? rc : 0);
                ncand += f - done + 1;
                const int q = lv.d_cand_q[cb + f];
                const int pos = lv.d_cand_pos[cb + f];
                CandView cv;
                derive_views(*reinterpret_cast<const PairViews *>(&p.A), lv, PAIRS ? q : q - k * lv.ntrial, r, rlen, rbase, ref_base,
                             ref_len, pos, IRR, cv);
                align_one<S, IRR>(*cv.a, cv.a_bit, cv.a_len, cv.a_tab, *cv.b, cv.b_bit, cv.b_len, p.R, p.maxn, p.maxm, planes,
                             p.PW, par, opsrev, p.ops ? p.ops + p.ops_off[k] : nullptr, raw, p.RW, bar, phase, res);
                cells += res.cells;
                if (lane == 0 && p.stats) {
                    atomicAdd(p.stats, (unsigned long long)res.cells); atomicAdd(p.stats + 1, 1ull);
                    atomicAdd(p.stats + 2, (unsigned long long)res.cells);
                    // integer-ALU warp instructions of the row loop: 19 per band word + 23 per row (SASS of align_locate_kernel<S>)
                    atomicAdd(p.stats + 4, (unsigned long long)(res.fail_row ? res.fail_row : res.len_a) * (19ull * S + 23ull));
                }
                done = f + 1;
                mask &= mask - 1;
                const bool ok = lv.mode == PB_MODE_LOCATE ? res.ret > 0                                      // locator.cpp:82
                                                          : (res.ret >= 0 && res.matlen_a >= lv.min_overlap); // ref_seq.h:264-265
                if (ok) { found = true; win_j = cv.j; win_pos = pos; win_rpos = cv.read_pos; win_dir = cv.dir; break; }
            }
            if (!found) {
                cells += __reduce_add_sync(FULL, (lane >= done && lane < nvalid) ? rc : 0);
                ncand += nvalid - done;
            }
        }
        if (lane == 0) {
            pb_locate_rec rec;
            rec.nseq = PAIRS ? r : k; rec.found = found ? 1 : 0;
            rec.j = found ? win_j : 0; rec.pos = found ? win_pos : 0;
            rec.cost = found ? res.cost : 0;
            // locator: len - j and cost(len-j,len-j) (locator.cpp:85-86); overlap: try_align's pos and direction
            rec.seg_len = !found ? 0 : (lv.mode == PB_MODE_LOCATE ? rlen - win_j : win_rpos);
            rec.diag_cost = !found ? 0 : (lv.mode == PB_MODE_LOCATE ? res.diag_cost : win_dir);
            rec.matlen_a = found ? res.matlen_a : 0; rec.matlen_b = found ? res.matlen_b : 0;
            rec.nedit = found ? res.nedit : 0;
            rec.ncand = ncand; rec._pad = T; rec.cells = cells; // all-vs-all: pb_pair_rec::ref_id
            recs[k] = rec;
        }
    }
}

// First pass of the two-pass locate (K3n): the same item loop over align_one_nb.  An item whose result is certified gets its
// record and redo[k] = 0; one with a candidate the strip cannot decide keeps redo[k] = 1 and is run again, from its first
// candidate, by align_locate_kernel (full band) in the second pass.
template <int S>
__global__ void __launch_bounds__(ALIGN_WPB * 32, MinBlocks<S>::v)
align_locate_nb_kernel(const __grid_constant__ AlignLaunch p, const __grid_constant__ LocateView lv, const uint8_t *__restrict__ survive,
                       const int32_t *__restrict__ rej_cells, pb_locate_rec *__restrict__ recs)
{
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ __align__(8) uint64_t bars[ALIGN_WPB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *planes = smem + (size_t)warp * p.warp_words;
    uint32_t *raw = planes + (size_t)4 * p.PW;
    uint64_t *bar = &bars[warp];
    uint32_t phase = 0u;
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    const size_t slot = (size_t)blockIdx.x * p.wpb + warp;
    uint32_t *par = p.scratch + slot * p.slot_words;
    uint8_t *opsrev = reinterpret_cast<uint8_t *>(par + p.par_words);
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(p.queue, 1);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= p.nitems) break;
        const int k = p.order[idx];
        const int r = lv.d_kept[k];
        const int rlen = p.A.len[r];
        const int64_t rbase = p.A.base[r];
        const int64_t c0 = lv.d_qoff[(int64_t)k * lv.ntrial];
        const int64_t c1 = lv.d_qoff[(int64_t)(k + 1) * lv.ntrial];
        const int64_t ref_base = lv.ref_base;
        const int ref_len = lv.ref_len;
        long long cells = 0, k3_cells = 0, band_cells = 0, alu_rows = 0;
        int ncand = 0, nrun = 0, redo = 0;
        int tbc[2] = {0, 0}; // traceback: rounds recomputed, cold starts of the window ring
        bool found = false;
        AlnRes res;
        int win_j = 0, win_pos = 0, win_rpos = 0, win_dir = 0;
        for (int64_t cb = c0; cb < c1 && !found && !redo; cb += 32) {
            const int64_t c = cb + lane;
            const int nvalid = (int)min((int64_t)32, c1 - cb);
            int sv = 0, rc = 0;
            if (lane < nvalid) { sv = survive[c]; rc = rej_cells[c]; }
            uint32_t mask = __ballot_sync(FULL, sv);
            int done = 0;
            while (mask) {
                const int f = __ffs(mask) - 1;
                cells += __reduce_add_sync(FULL, (lane >= done && lane < f) ? rc : 0);
                ncand += f - done + 1;
                const int q = lv.d_cand_q[cb + f];
                const int pos = lv.d_cand_pos[cb + f];
                CandView cv;
                derive_views(*reinterpret_cast<const PairViews *>(&p.A), lv, q - k * lv.ntrial, r, rlen, rbase, ref_base,
                             ref_len, pos, false, cv);
                align_one_nb<S>(*cv.a, cv.a_bit, cv.a_len, *cv.b, cv.b_bit, cv.b_len, p.R, p.maxn, p.maxm, p.g256, planes, p.PW, par,
                                p.par_words, opsrev, (int)((p.slot_words - p.par_words) & ~(size_t)31), p.ops ? p.ops + p.ops_off[k] : nullptr, raw, p.RW,
                                raw + 2 * p.RW, bar, phase, res, redo,
                                band_cells, tbc);
                alu_rows += res.fail_row ? res.fail_row : (redo ? 0 : res.len_a);
                if (redo) break;
                cells += res.cells; k3_cells += res.cells; ++nrun;
                done = f + 1;
                mask &= mask - 1;
                const bool ok = lv.mode == PB_MODE_LOCATE ? res.ret > 0                                      // locator.cpp:82
                                                          : (res.ret >= 0 && res.matlen_a >= lv.min_overlap); // ref_seq.h:264-265
                if (ok) { found = true; win_j = cv.j; win_pos = pos; win_rpos = cv.read_pos; win_dir = cv.dir; break; }
            }
            if (!found && !redo) {
                cells += __reduce_add_sync(FULL, (lane >= done && lane < nvalid) ? rc : 0);
                ncand += nvalid - done;
            }
        }
        if (lane == 0) {
            if (p.stats) { // what the strip computed counts even when the item has to be redone
                atomicAdd(p.stats + 2, (unsigned long long)band_cells);
                atomicAdd(p.stats + 4, (unsigned long long)alu_rows * (11ull * S + 12ull)); // SASS of align_locate_nb_kernel<S>: 11 per word + 12 per row
                if (redo) atomicAdd(p.stats + 3, 1ull);
                else { atomicAdd(p.stats, (unsigned long long)k3_cells); atomicAdd(p.stats + 1, (unsigned long long)nrun); }
                if (tbc[0]) { atomicAdd(p.stats + 5, (unsigned long long)tbc[0]); atomicAdd(p.stats + 6, (unsigned long long)tbc[1]); }
            }
            p.redo[k] = (uint8_t)redo;
            if (!redo) {
                pb_locate_rec rec;
                rec.nseq = k; rec.found = found ? 1 : 0;
                rec.j = found ? win_j : 0; rec.pos = found ? win_pos : 0;
                rec.cost = found ? res.cost : 0;
                rec.seg_len = !found ? 0 : (lv.mode == PB_MODE_LOCATE ? rlen - win_j : win_rpos);
                rec.diag_cost = !found ? 0 : (lv.mode == PB_MODE_LOCATE ? res.diag_cost : win_dir);
                rec.matlen_a = found ? res.matlen_a : 0; rec.matlen_b = found ? res.matlen_b : 0;
                rec.nedit = found ? res.nedit : 0;
                rec.ncand = ncand; rec._pad = 0; rec.cells = cells;
                recs[k] = rec;
            }
        }
    }
}

template <int S, bool IRR>
__global__ void __launch_bounds__(ALIGN_WPB * 32, MinBlocks<S>::v)
align_pairs_kernel(const __grid_constant__ AlignLaunch p, pb_align_out *__restrict__ out)
{
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ __align__(8) uint64_t bars[ALIGN_WPB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per warp: Eq planes, then the TMA staging area for seg_b's raw plane words
    uint32_t *planes = smem + (size_t)warp * p.warp_words;
    uint32_t *raw = planes + (size_t)(IRR ? 8 : 4) * p.PW;
    uint64_t *bar = &bars[warp];
    uint32_t phase = 0u;
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    const size_t slot = (size_t)blockIdx.x * p.wpb + warp;
    uint32_t *par = p.scratch + slot * p.slot_words;
    uint8_t *opsrev = reinterpret_cast<uint8_t *>(par + p.par_words);
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(p.queue, 1);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= p.nitems) break;
        const int k = p.order[idx];
        if (p.redo && !p.redo[k]) continue; // certified by the strip pass
        AlnRes res;
        align_one<S, IRR>(p.A, p.A.base[k], p.A.len[k], IRR ? p.A.tab[k] : 0u, p.B, p.B.base[k], p.B.len[k], p.R, p.maxn, p.maxm, planes, p.PW, par, opsrev,
                     p.ops ? p.ops + p.ops_off[k] : nullptr, raw, p.RW, bar, phase, res);
        if (lane == 0 && p.stats) { atomicAdd(p.stats, (unsigned long long)res.cells); atomicAdd(p.stats + 1, 1ull); }
        if (lane == 0) {
            pb_align_out o;
            o.ret = res.ret; o.len_a = res.len_a; o.len_b = res.len_b; o.max_dst = res.D;
            o.matlen_a = res.matlen_a; o.matlen_b = res.matlen_b; o.cost = res.cost; o.diag_cost = res.diag_cost;
            o.nedit = res.nedit; o.fail_row = res.fail_row; o.cells = res.cells;
            out[k] = o;
        }
    }
}

// First pass of pb_align_batch for the wide bands: the certified strip (align_one_nb) on explicit pairs.  A pair the strip cannot
// certify keeps redo[k] = 1 and is run by align_pairs_kernel afterwards, which skips the others.
template <int S>
__global__ void __launch_bounds__(ALIGN_WPB * 32, MinBlocks<S>::v)
align_pairs_nb_kernel(const __grid_constant__ AlignLaunch p, pb_align_out *__restrict__ out)
{
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ __align__(8) uint64_t bars[ALIGN_WPB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *planes = smem + (size_t)warp * p.warp_words;
    uint32_t *raw = planes + (size_t)4 * p.PW;
    uint64_t *bar = &bars[warp];
    uint32_t phase = 0u;
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    const size_t slot = (size_t)blockIdx.x * p.wpb + warp;
    uint32_t *par = p.scratch + slot * p.slot_words;
    uint8_t *opsrev = reinterpret_cast<uint8_t *>(par + p.par_words);
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(p.queue, 1);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= p.nitems) break;
        const int k = p.order[idx];
        AlnRes res;
        int redo = 0;
        long long band_cells = 0;
        int tbc[2] = {0, 0};
        align_one_nb<S>(p.A, p.A.base[k], p.A.len[k], p.B, p.B.base[k], p.B.len[k], p.R, p.maxn, p.maxm, p.g256, planes, p.PW, par,
                        p.par_words, opsrev, (int)((p.slot_words - p.par_words) & ~(size_t)31), p.ops ? p.ops + p.ops_off[k] : nullptr, raw,
                        p.RW, raw + 2 * p.RW, bar, phase, res, redo, band_cells, tbc);
        if (lane == 0) {
            p.redo[k] = (uint8_t)redo;
            if (p.stats) {
                atomicAdd(p.stats + 2, (unsigned long long)band_cells);
                if (redo) atomicAdd(p.stats + 3, 1ull);
                else { atomicAdd(p.stats, (unsigned long long)res.cells); atomicAdd(p.stats + 1, 1ull); }
            }
            if (!redo) {
                pb_align_out o;
                o.ret = res.ret; o.len_a = res.len_a; o.len_b = res.len_b; o.max_dst = res.D;
                o.matlen_a = res.matlen_a; o.matlen_b = res.matlen_b; o.cost = res.cost; o.diag_cost = res.diag_cost;
                o.nedit = res.nedit; o.fail_row = res.fail_row; o.cells = res.cells;
                out[k] = o;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Narrow bands: several alignments per warp.  A band of at most 32*LANES bits (LANES = 4, 8 or 16 lanes, one word per lane)
// leaves most of a warp idle in align_one, so here a warp advances 32/LANES alignments in lockstep: the shuffles and ballots
// of a row are shared by all groups (each group cuts its own bits out of the ballot, band edges sit at group edges), then
// the groups are finished (goal cell, traceback) one after the other by the whole warp.  Same arithmetic as align_one with
// S = 1; the early-failure test is evaluated per row.  Used by pb_align_batch (config 3's band 32..128 sweep).
// ---------------------------------------------------------------------------------------------
template <int LANES>
__global__ void __launch_bounds__(ALIGN_WPB * 32)
align_pairs_packed_kernel(const __grid_constant__ AlignLaunch p, pb_align_out *__restrict__ out)
{
    constexpr int G = 32 / LANES;
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / LANES, gl = lane % LANES, gbase = g * LANES;
    constexpr uint32_t gmask = (1u << LANES) - 1u;
    uint32_t *planes = smem + (size_t)warp * p.warp_words + (size_t)g * 4 * p.PW;
    const size_t slot = ((size_t)blockIdx.x * p.wpb + warp) * G + g;
    uint32_t *par = p.scratch + slot * p.slot_words; // rows x LANES pairs, the last row is a dump for finished groups
    const int dump_row = (int)(p.par_words / (2 * LANES)) - 1;
    const SeqView &A = p.A, &B = p.B;
    for (;;) {
        int idx0 = 0;
        if (lane == 0) idx0 = atomicAdd(p.queue, G);
        idx0 = __shfl_sync(FULL, idx0, 0);
        if (idx0 >= p.nitems) break;
        const bool has = idx0 + g < p.nitems;
        const int k = has ? p.order[idx0 + g] : 0;
        int a_len = 0, b_len = 0;
        int64_t a_bit = 0, b_bit = 0;
        if (has) { a_len = A.len[k]; a_bit = A.base[k]; b_len = B.len[k]; b_bit = B.base[k]; }
        int len_a = 0, len_b = 0, D = 0;
        derive_params(a_len, b_len, p.R, len_a, len_b, D);
        const bool dom = has && !(len_a >= p.maxn || D >= p.maxm); // seq_aligner.h:104-107
        const int rows = dom ? len_a : 0;

        // Eq planes of this group's seg_b (as in align_one, plain loads)
        const int PWn = ((rows + 31) >> 5) + LANES + 1;
        if (dom)
            for (int x = gl; x < PWn; x += LANES) {
                const int bidx0 = 32 * x - D;
                uint32_t valid;
                if (bidx0 >= len_b || bidx0 + 31 < 0) valid = 0u;
                else {
                    valid = 0xffffffffu;
                    if (bidx0 < 0) valid &= 0xffffffffu << (-bidx0);
                    if (bidx0 + 31 >= len_b) valid &= 0xffffffffu >> (bidx0 + 32 - len_b);
                }
                uint32_t hi = 0u, lo = 0u;
                if (valid) {
                    hi = load_window(B.hi, B.nwords, b_bit + bidx0);
                    lo = load_window(B.lo, B.nwords, b_bit + bidx0);
                }
                planes[0 * p.PW + x] = ~hi & ~lo & valid;
                planes[1 * p.PW + x] = ~hi & lo & valid;
                planes[2 * p.PW + x] = hi & ~lo & valid;
                planes[3 * p.PW + x] = hi & lo & valid;
            }
        __syncwarp();

        // row 0 (see align_one): this lane's single band word holds bits k0 .. k0+31
        const int k0 = 32 * gl;
        uint32_t Hn = (k0 + 31 <= D) ? 0xffffffffu : (k0 > D ? 0u : 0xffffffffu >> (31 - (D - k0)));
        uint32_t Hp = ~Hn;
        const uint32_t keep = (k0 > 2 * D) ? 0u : (k0 + 31 <= 2 * D ? 0xffffffffu : 0xffffffffu >> (31 - (2 * D - k0)));
        const int diag_lane = gbase + min(D >> 5, LANES - 1);
        int cii = 0, colc = 0, colbest = 0, col_i = 0, fail_row = 0, rows_done = 0;
        const int maxrows = __reduce_max_sync(FULL, rows);
        const int min_lb = __reduce_min_sync(FULL, dom ? len_b : INT_MAX); // first row at which some group follows its last column
        uint32_t awh = 0u, awl = 0u, Hp_fin = 0u, Hn_fin = 0u;
        for (int i = 1; i <= maxrows; ++i) {
            const int t = (i - 1) & 31;
            const bool live = i <= rows && !fail_row;
            if (t == 0 && i <= rows) {
                awh = load_window(A.hi, A.nwords, a_bit + i - 1);
                awl = load_window(A.lo, A.nwords, a_bit + i - 1);
            }
            const int ca = (int)(((awh >> t) & 1u) * 2u + ((awl >> t) & 1u));
            const uint32_t *pl = planes + ca * p.PW + ((i - 1) >> 5) + gl;
            // slide the band: bit 0 of the next lane's word comes in at the top; the group's last lane takes the edge value
            uint32_t nx = __shfl_down_sync(FULL, (Hp & 1u) | ((Hn & 1u) << 1), 1);
            if (gl == LANES - 1) nx = 1u;
            Hp = __funnelshift_r(Hp, nx & 1u, 1);
            Hn = __funnelshift_r(Hn, nx >> 1, 1);
            uint32_t Eq = 0u;
            if (i <= rows) Eq = __funnelshift_r(pl[0], pl[1], t) & keep;
            const uint32_t x = Eq & Hp;
            uint32_t sum = x + Hp;
            const uint32_t Gb = (__ballot_sync(FULL, sum < x) >> gbase) & gmask;
            const uint32_t Pb = (__ballot_sync(FULL, sum == 0xffffffffu) >> gbase) & gmask;
            sum += ((((Gb | Pb) + Gb) ^ Pb) >> gl) & 1u; // carries never leave the group
            const uint32_t Xv = (sum ^ Hp) | Eq;
            const uint32_t Vp = Hn | ~(Xv | Hp), Vn = Hp & Xv;
            const uint32_t D0 = Xv | Hn;
            const uint32_t Mw = Eq | ~D0;
            uint32_t pv = __shfl_up_sync(FULL, (Vp >> 31) | ((Vn >> 31) << 1), 1);
            if (gl == 0) pv = 1u;
            const uint32_t vps = (Vp << 1) | (pv & 1u), vns = (Vn << 1) | (pv >> 1);
            const uint32_t Xh = Eq | Hn;
            Hp = vns | ~(Xh | vps);
            Hn = vps & Xh;
            if (i == rows) { Hp_fin = Hp; Hn_fin = Hn; } // this group's last row: the goal scan reads these deltas
            // parents: finished groups write into the dump row (a pointer select; predicated stores are slow here)
            reinterpret_cast<uint2 *>(par + (size_t)(live ? i - 1 : dump_row) * (2 * LANES))[gl] = make_uint2(Mw, Hp);
            const uint32_t d0bit = (__shfl_sync(FULL, D0, diag_lane) >> (D & 31)) & 1u;
            // cost(i, len_b) for rows below seg_b's end follows the vertical delta at column len_b
            const int kcol = min(max(len_b - i + D, 0), 32 * LANES - 1);
            uint32_t vpw = 0u, vnw = 0u;
            if (i > min_lb) { // warp-uniform: only rows past some group's seg_b pay for these
                vpw = __shfl_sync(FULL, Vp, gbase + (kcol >> 5));
                vnw = __shfl_sync(FULL, Vn, gbase + (kcol >> 5));
            }
            if (live) {
                rows_done = i;
                if (i <= len_b) {
                    cii += 1 - (int)d0bit;
                    if (i > 10 && (double)cii > i * p.R) fail_row = i; // seq_aligner.h:185
                    if (i == len_b) { colc = colbest = cii; col_i = i; }
                } else {
                    colc += (int)((vpw >> (kcol & 31)) & 1u) - (int)((vnw >> (kcol & 31)) & 1u);
                    if (colc < colbest) { colbest = colc; col_i = i; }
                }
            }
        }
        // final horizontal deltas of every group into its own plane area, then finish the groups one by one
        __syncwarp();
        planes[gl] = Hp_fin;
        planes[LANES + gl] = Hn_fin;
        __syncwarp();
        for (int gg = 0; gg < G; ++gg) {
            const int src = gg * LANES;
            const int f_has = __shfl_sync(FULL, (int)has, src), f_dom = __shfl_sync(FULL, (int)dom, src);
            if (!f_has) continue;
            const int f_k = __shfl_sync(FULL, k, src), f_la = __shfl_sync(FULL, len_a, src), f_lb = __shfl_sync(FULL, len_b, src);
            const int f_D = __shfl_sync(FULL, D, src), f_alen = __shfl_sync(FULL, a_len, src), f_cii = __shfl_sync(FULL, cii, src);
            const int f_cb = __shfl_sync(FULL, colbest, src), f_ci = __shfl_sync(FULL, col_i, src);
            const int f_fail = __shfl_sync(FULL, fail_row, src), f_rows = __shfl_sync(FULL, rows_done, src);
            const uint32_t *f_planes = smem + (size_t)warp * p.warp_words + (size_t)gg * 4 * p.PW;
            uint32_t *f_par = p.scratch + (((size_t)blockIdx.x * p.wpb + warp) * G + gg) * p.slot_words;
            uint8_t *f_opsrev = reinterpret_cast<uint8_t *>(f_par + p.par_words);
            AlnRes res;
            res.ret = -1; res.len_a = f_la; res.len_b = f_lb; res.D = f_D;
            res.matlen_a = res.matlen_b = res.cost = res.diag_cost = res.nedit = 0;
            res.fail_row = f_fail;
            res.cells = f_dom ? cells_upto(f_rows, f_D, f_lb) : 0;
            if (f_dom && !f_fail) {
                auto par_pair = [&](int row, int w) -> uint2 {
                    if (row < 1 || w < 0 || 32 * w > 2 * f_D) return make_uint2(0u, 0u);
                    return __ldcg(reinterpret_cast<const uint2 *>(f_par + (size_t)(row - 1) * (2 * LANES)) + w);
                };
                auto no_addr = [](int, int) -> const uint2 * { return nullptr; }; // narrow bands: synchronous window loads
                finish_alignment<false>(f_la, f_lb, f_D, f_alen, p.R, f_cii, f_cb, f_ci, f_planes, f_planes + LANES, par_pair, no_addr,
                                 (uint32_t *)nullptr, (const void *)nullptr, f_opsrev, p.ops ? p.ops + p.ops_off[f_k] : nullptr, res);
            }
            if (lane == 0) {
                if (p.stats) { atomicAdd(p.stats, (unsigned long long)res.cells); atomicAdd(p.stats + 1, 1ull); }
                pb_align_out o;
                o.ret = res.ret; o.len_a = res.len_a; o.len_b = res.len_b; o.max_dst = res.D;
                o.matlen_a = res.matlen_a; o.matlen_b = res.matlen_b; o.cost = res.cost; o.diag_cost = res.diag_cost;
                o.nedit = res.nedit; o.fail_row = res.fail_row; o.cells = res.cells;
                out[f_k] = o;
            }
            __syncwarp();
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Very narrow bands: one alignment per THREAD.  A band of at most 32*W bits (W = 3, 5 or 9 words: max_dst <= 47, 79, 143 --
// config 3's band 32 / 64 / 128 points) fits the registers of one thread, so the row needs no shuffle, no ballot and no second
// carry pass: the W-word add is one carry chain inside the thread, the vertical delta entering word s is the top bit of word
// s-1.  Same arithmetic and the same band-edge rules as align_one (bit 0 = the band's left edge, Eq forced to 0 above bit
// 2*max_dst); a warp advances 32 alignments in lockstep, parents go out as [row][word][lane] pairs -- every store instruction
// writes 256 contiguous bytes.
// A batch of a few thousand pairs is a few hundred warps, at most one per scheduler: the time of a point is rows x the latency
// of ONE row, so the row is written for latency.  Whole 32-row blocks in which nothing special can happen to any alignment of
// the warp (no last row, no end of seg_b) run an unrolled, branch-free loop -- the Eq fetches, stores and the failure test of
// neighbouring rows overlap the carry chain; the rows after that take the general loop.  seg_b's Eq words live in the thread's
// own column of shared memory ([plane][word][thread], constant stride: every offset an immediate); a block needs ONE new plane
// word (two raw loads, issued a block ahead).  The traceback fetches, for TBR rows at once, the 32 parent bits around the
// lane's diagonal -- a lane can then take TBR steps whatever they are, with one vote per TBR steps; transcripts are written
// [step][lane], so the byte stores of a step coalesce.
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(ALIGN_WPB * 32)
align_pairs_thread_kernel(const __grid_constant__ AlignLaunch p, pb_align_out *__restrict__ out)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NT = ALIGN_WPB * 32;         // column stride, whatever the CTA's size (smaller CTAs leave columns unused)
    constexpr int PWT = W + 1;                 // plane words a block of 32 rows can touch
    uint32_t *pl = smem + tid;                 // plane c, word s of this thread: pl[(c * PWT + s) * NT]
    uint32_t *vv = smem + 4 * PWT * NT + tid;  // Vp / Vn (or the final Hp / Hn) of this thread: vv[s * NT], vv[(W + s) * NT]
    const size_t wslot = ((size_t)blockIdx.x * p.wpb + warp) * 32; // this warp's 32 alignment slots are contiguous
    uint32_t *par = p.scratch + wslot * p.slot_words;               // [row][word][lane] {MATCH word, INSERT word}
    uint8_t *ops_w = reinterpret_cast<uint8_t *>(par + 32 * p.par_words) + lane; // reversed transcripts, [step][lane]
    const SeqView &A = p.A, &B = p.B;
    for (;;) {
        int idx0 = 0;
        if (lane == 0) idx0 = atomicAdd(p.queue, 32);
        idx0 = __shfl_sync(FULL, idx0, 0);
        if (idx0 >= p.nitems) break;
        const bool has = idx0 + lane < p.nitems;
        const int k = has ? p.order[idx0 + lane] : 0;
        int a_len = 0, b_len = 0;
        int64_t a_bit = 0, b_bit = 0;
        if (has) { a_len = A.len[k]; a_bit = A.base[k]; b_len = B.len[k]; b_bit = B.base[k]; }
        int len_a = 0, len_b = 0, D = 0;
        derive_params(a_len, b_len, p.R, len_a, len_b, D);
        const bool dom = has && !(len_a >= p.maxn || D >= p.maxm) && 2 * D + 1 <= 32 * W; // seq_aligner.h:104-107 (the host routes by band)
        const int rows = dom ? len_a : 0;
        uint32_t Hp[W], Hn[W], keep[W];
#pragma unroll
        for (int s = 0; s < W; ++s) {
            const int k0 = 32 * s;
            Hn[s] = (k0 + 31 <= D) ? 0xffffffffu : (k0 > D ? 0u : 0xffffffffu >> (31 - (D - k0)));
            Hp[s] = ~Hn[s];
            keep[s] = (k0 > 2 * D) ? 0u : (k0 + 31 <= 2 * D ? 0xffffffffu : 0xffffffffu >> (31 - (2 * D - k0)));
        }
        const int sd = D >> 5, dbit = D & 31;
        int cii = 0, colc = 0, colbest = 0, col_i = 0, fail_row = 0;
        int matlen_a = 0, matlen_b = 0, cost = 0;
        const int maxrows = __reduce_max_sync(FULL, rows);
        const int min_lb = __reduce_min_sync(FULL, dom ? len_b : INT_MAX);

        // seg_b's bit planes, plane word x: bit y <-> b[32x - D + y] = raw line bits from b_bit + 32x - D on, i.e. raw words
        // wb0 + x and wb0 + x + 1 funnelled by shb.  hiw / low hold the plane words of the current block, rawh / rawl the last
        // raw words read, nrh / nrl / nawh / nawl what the NEXT block needs (loaded a block ahead).
        const int64_t wb0 = (b_bit - D) >> 5;
        const unsigned shb = (unsigned)((b_bit - D) & 31);
        auto raw_b = [&](const uint32_t *__restrict__ arr, int64_t wi) -> uint32_t { return (dom && wi >= 0 && wi < B.nwords) ? __ldg(arr + wi) : 0u; };
        uint32_t hiw[PWT], low[PWT], rawh, rawl, nrh, nrl, awh = 0u, awl = 0u, nawh = 0u, nawl = 0u;
        rawh = raw_b(B.hi, wb0);
        rawl = raw_b(B.lo, wb0);
#pragma unroll
        for (int s = 0; s < PWT; ++s) {
            const uint32_t h = raw_b(B.hi, wb0 + s + 1), l = raw_b(B.lo, wb0 + s + 1);
            hiw[s] = __funnelshift_r(rawh, h, shb);
            low[s] = __funnelshift_r(rawl, l, shb);
            rawh = h; rawl = l;
        }
        nrh = nrl = 0u;
        if (dom) { nawh = load_window(A.hi, A.nwords, a_bit); nawl = load_window(A.lo, A.nwords, a_bit); }
        // block q = rows 32q+1 .. 32q+32; called for q = 0, 1, 2, ... in turn
        auto refill = [&](int q) {
            if (q > 0) {
#pragma unroll
                for (int s = 0; s + 1 < PWT; ++s) { hiw[s] = hiw[s + 1]; low[s] = low[s + 1]; }
                hiw[PWT - 1] = __funnelshift_r(rawh, nrh, shb);
                low[PWT - 1] = __funnelshift_r(rawl, nrl, shb);
                rawh = nrh; rawl = nrl;
            }
#pragma unroll
            for (int s = 0; s < PWT; ++s) {
                const int bidx0 = 32 * (q + s) - D; // b index of bit 0 of plane word q + s
                uint32_t valid;
                if (bidx0 >= len_b || bidx0 + 31 < 0) valid = 0u;
                else {
                    valid = 0xffffffffu;
                    if (bidx0 < 0) valid &= 0xffffffffu << (-bidx0);
                    if (bidx0 + 31 >= len_b) valid &= 0xffffffffu >> (bidx0 + 32 - len_b);
                }
                const uint32_t hi = hiw[s], lo = low[s];
                pl[(0 * PWT + s) * NT] = ~hi & ~lo & valid;
                pl[(1 * PWT + s) * NT] = ~hi & lo & valid;
                pl[(2 * PWT + s) * NT] = hi & ~lo & valid;
                pl[(3 * PWT + s) * NT] = hi & lo & valid;
            }
            awh = nawh; awl = nawl;
            if (32 * (q + 1) < rows) { // the block after this one
                nrh = raw_b(B.hi, wb0 + q + 1 + PWT);
                nrl = raw_b(B.lo, wb0 + q + 1 + PWT);
                nawh = load_window(A.hi, A.nwords, a_bit + 32 * (q + 1));
                nawl = load_window(A.lo, A.nwords, a_bit + 32 * (q + 1));
            }
        };
        // one row: slides the band, updates Hp / Hn, leaves the vertical deltas in Vp / Vn, stores the parents when st;
        // returns the D0 word the main diagonal sits in
        auto row_core = [&](int i, int t, bool st, uint32_t (&Vp)[W], uint32_t (&Vn)[W]) -> uint32_t {
            const uint32_t ca = ((awh >> t) & 1u) * 2u + ((awl >> t) & 1u);
            const uint32_t *pc = pl + ca * (PWT * NT);
            uint32_t Eq[W], x[W], sum[W];
            uint32_t w0 = pc[0];
#pragma unroll
            for (int s = 0; s < W; ++s) { // slide the band one bit; the bit entering at the top is the +1 of a pinned edge
                Hp[s] = __funnelshift_r(Hp[s], s + 1 < W ? Hp[s + 1] : 1u, 1);
                Hn[s] = __funnelshift_r(Hn[s], s + 1 < W ? Hn[s + 1] : 0u, 1);
                const uint32_t w1 = pc[(s + 1) * NT];
                Eq[s] = __funnelshift_r(w0, w1, t) & keep[s];
                w0 = w1;
                x[s] = Eq[s] & Hp[s];
            }
            CarryChain<W>::add_nc(sum, x, Hp);
            uint32_t d0w = 0u;
            uint32_t pprev = 0x80000000u, nprev = 0u; // vin = +1 at the band's left edge
            uint2 *prow = reinterpret_cast<uint2 *>(par + (size_t)(i - 1) * (W * 64)) + lane;
#pragma unroll
            for (int s = 0; s < W; ++s) {
                const uint32_t Xv = (sum[s] ^ Hp[s]) | Eq[s];
                Vp[s] = Hn[s] | ~(Xv | Hp[s]);
                Vn[s] = Hp[s] & Xv;
                const uint32_t D0 = Xv | Hn[s];
                const uint32_t Mw = Eq[s] | ~D0;
                if (s <= (W - 1) / 2 && s == sd) d0w = D0; // max_dst <= 16*W - 1: the diagonal sits in the lower half
                const uint32_t vps = __funnelshift_l(pprev, Vp[s], 1), vns = __funnelshift_l(nprev, Vn[s], 1);
                pprev = Vp[s]; nprev = Vn[s];
                const uint32_t Xh = Eq[s] | Hn[s];
                Hp[s] = vns | ~(Xh | vps);
                Hn[s] = vps & Xh;
                if (st) prow[s * 32] = make_uint2(Mw, Hp[s]);
            }
            return d0w;
        };

        // ---- rows 1 .. fast_rows: whole blocks before any alignment's last row or the end of its seg_b
        int i = 1;
        {
            const int lim = dom ? min(rows, len_b) - 1 : INT_MAX;
            const int fast_rows = max(0, min(__reduce_min_sync(FULL, lim), maxrows)) & ~31;
            for (; i <= fast_rows; i += 32) {
                if (!__any_sync(FULL, dom && !fail_row)) break; // every alignment of the warp has failed
                refill((i - 1) >> 5);
#pragma unroll 8
                for (int t = 0; t < 32; ++t) {
                    uint32_t Vp[W], Vn[W];
                    const uint32_t d0w = row_core(i + t, t, dom, Vp, Vn); // a failed alignment's parents are never read
                    cii += 1 - (int)((d0w >> dbit) & 1u);
                    if (dom && !fail_row && i + t > 10 && (double)cii > (i + t) * p.R) fail_row = i + t; // seq_aligner.h:185
                }
            }
        }
        // ---- the remaining rows: last rows, rows past seg_b's end, goal cells
        const bool any_left = __any_sync(FULL, dom && !fail_row);
        for (; any_left && i <= maxrows; ++i) {
            const int t = (i - 1) & 31;
            const bool live = i <= rows && !fail_row;
            if (t == 0) refill((i - 1) >> 5);
            uint32_t Vp[W], Vn[W];
            const uint32_t d0w = row_core(i, t, live, Vp, Vn);
            if (i > min_lb) { // warp-uniform: only rows past some alignment's seg_b pay for this
#pragma unroll
                for (int s = 0; s < W; ++s) { vv[s * NT] = Vp[s]; vv[(W + s) * NT] = Vn[s]; }
            }
            if (live) {
                if (i <= len_b) {
                    cii += 1 - (int)((d0w >> dbit) & 1u);
                    if (i > 10 && (double)cii > i * p.R) fail_row = i; // seq_aligner.h:185
                    if (i == len_b) { colc = colbest = cii; col_i = i; }
                } else { // cost(i, len_b) follows the vertical delta at column len_b (no failure test there, Q-D2)
                    const int kc = len_b - i + D;
                    colc += (int)((vv[(kc >> 5) * NT] >> (kc & 31)) & 1u) - (int)((vv[(W + (kc >> 5)) * NT] >> (kc & 31)) & 1u);
                    if (colc < colbest) { colbest = colc; col_i = i; }
                }
                if (i == rows && !fail_row) { // goal_cell, seq_aligner.h:191-213
                    if (len_a > len_b) { matlen_a = col_i; matlen_b = len_b; cost = colbest; }
                    else {
#pragma unroll
                        for (int s = 0; s < W; ++s) { vv[s * NT] = Hp[s]; vv[(W + s) * NT] = Hn[s]; }
                        matlen_a = len_a; matlen_b = len_a; cost = cii;
                        int c = cii;
                        for (int j = len_a + 1; j <= len_b; ++j) {
                            const int kk = j - len_a + D;
                            c += (int)((vv[(kk >> 5) * NT] >> (kk & 31)) & 1u) - (int)((vv[(W + (kk >> 5)) * NT] >> (kk & 31)) & 1u);
                            if (c < cost) { cost = c; matlen_b = j; }
                        }
                    }
                }
            }
        }
        // ---- coverage test, find_path (seq_aligner.h:114, 214-233), record.  Every lane walks its own path, the lanes step
        // together.  A round: every walking lane fetches, for the TBR rows from where it stands, the 32 {MATCH, INSERT} bits
        // around its diagonal (two adjacent band words funnelled; bits outside the band read 0 and are never reached) into
        // its column of shared memory (the Eq planes are dead by now), then takes TBR steps -- a step moves at most one row up
        // and one bit sideways, so the window holds whatever the steps turn out to be.
        pb_align_out o;
        o.ret = -1; o.len_a = len_a; o.len_b = len_b; o.max_dst = D;
        o.matlen_a = o.matlen_b = o.cost = o.diag_cost = o.nedit = 0;
        o.fail_row = fail_row;
        o.cells = dom ? cells_upto(fail_row ? fail_row : len_a, D, len_b) : 0;
        bool walk = false;
        if (dom && !fail_row) {
            o.matlen_a = matlen_a; o.matlen_b = matlen_b; o.cost = cost;
            o.diag_cost = (a_len <= len_a && a_len <= len_b) ? cii : 0;
            walk = !((double)matlen_b < len_b * (1 - p.R));
        }
#ifdef PB_EXP_NOTB // timing experiment only (tools/ab_build.sh): the forward pass without the traceback
        walk = false;
#endif
        {
            int ti = walk ? matlen_a : 0, tj = walk ? matlen_b : 0, n = 0;
            const int guard = len_a + len_b + 1;
            constexpr int TBR = (4 * PWT + 2 * W) / 2 < 15 ? (4 * PWT + 2 * W) / 2 : 15; // rows per round: what fits the thread's column
            uint32_t *win = smem + tid;                                                  // row u of the round: win[(2u) * NT], win[(2u + 1) * NT]
            for (;;) {
                const bool act = ti > 0 && tj > 0 && n < guard;
                if (!__any_sync(FULL, act)) break;
                const int p0 = tj - ti + D - 16; // band bit of window bit 0
                const int wlo = p0 >> 5;
                const unsigned shw = (unsigned)(p0 & 31);
                { // branch-free: clamped addresses, all loads of the round in flight together, out-of-range words read as 0
                    const int wl = min(max(wlo, 0), W - 1), wh = min(max(wlo + 1, 0), W - 1);
                    const uint32_t ml = (wlo >= 0 && wlo < W) ? 0xffffffffu : 0u, mh = (wlo + 1 >= 0 && wlo + 1 < W) ? 0xffffffffu : 0u;
                    uint2 a0[TBR], a1[TBR];
#pragma unroll
                    for (int u = 0; u < TBR; ++u) {
                        const uint2 *base = reinterpret_cast<const uint2 *>(par + (size_t)(max(ti - u, 1) - 1) * (W * 64)) + lane;
                        a0[u] = __ldcg(base + wl * 32);
                        a1[u] = __ldcg(base + wh * 32);
                    }
#pragma unroll
                    for (int u = 0; u < TBR; ++u) {
                        const uint32_t rl = ti - u >= 1 ? ml : 0u, rh = ti - u >= 1 ? mh : 0u;
                        win[(2 * u) * NT] = __funnelshift_r(a0[u].x & rl, a1[u].x & rh, shw);
                        win[(2 * u + 1) * NT] = __funnelshift_r(a0[u].y & rl, a1[u].y & rh, shw);
                    }
                }
                const int w_top = ti;
                int bit = 16;
#pragma unroll
                for (int u = 0; u < TBR; ++u) {
                    const bool ok = ti > 0 && tj > 0 && n < guard;
                    const int idx = w_top - ti; // 0 .. u
                    const uint32_t m = win[(2 * idx) * NT], ins = win[(2 * idx + 1) * NT];
                    const int mb = (int)((m >> bit) & 1u);            // MATCH wins, then INSERT, else DELETE (seq_aligner.h:214-233)
                    const int ib = (int)((ins >> bit) & 1u) & (mb ^ 1);
                    const int db = (mb | ib) ^ 1;
                    if (ok) {
                        ops_w[(size_t)n * 32] = (uint8_t)(mb ? PB_MATCH : (ib ? PB_INSERT : PB_DELETE));
                        ti -= mb | db;
                        tj -= mb | ib;
                        bit += db - ib;
                        ++n;
                    }
                }
            }
            if (walk) {
                if (n < guard) {
                    for (; tj > 0 && ti == 0; --tj) { ops_w[(size_t)n * 32] = (uint8_t)PB_INSERT; ++n; } // init_cell row 0
                    for (; ti > 0 && tj == 0; --ti) { ops_w[(size_t)n * 32] = (uint8_t)PB_DELETE; ++n; } // init_cell column 0
                }
                if (p.ops) {
                    uint8_t *dst = p.ops + p.ops_off[k];
                    for (int q = 0; q < n; ++q) dst[q] = ops_w[(size_t)(n - 1 - q) * 32];
                }
                o.nedit = n;
                o.ret = matlen_b;
            }
        }
        if (has) {
            if (p.stats) { atomicAdd(p.stats, (unsigned long long)o.cells); atomicAdd(p.stats + 1, 1ull); }
            out[k] = o;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// host side: size classes, scratch, launches
// ---------------------------------------------------------------------------------------------

// Band classes.  A class key is S (band words per lane) for plain ACGT input, 1000 + S for the byte-exact variant
// (8 Eq planes) that pairs holding other bytes take; such input is rare, so three sizes are enough there.
static const int kClasses[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16};
static const int kIrrClasses[] = {2, 6, 16};
static inline int key_S(int key) { return key % 1000; }
static int class_for_band_plain(int D)
{
    const int NW = (2 * D + 1 + 31) >> 5;
    for (int S : kClasses)
        if (32 * S >= NW) return S;
    return -1;
}
static inline bool key_irr(int key) { return key >= 1000 && key < 2000; }
// 2000 + LANES: several narrow-band alignments per warp (pairs mode only), LANES lanes x one word each
static inline bool key_packed(int key) { return key >= 2000 && key < 3000; }
static inline int key_lanes(int key) { return key - 2000; }
// 3000 + S: first pass of the two-pass locate over a certified strip (align_locate_nb_kernel<S>)
static const int kNarrowClasses[] = {1, 2, 3, 4, 5, 6, 7, 8};
static inline bool key_narrow(int key) { return key >= 3000 && key < 4000; }
// 4000 + W: one alignment per thread, W band words (pairs mode only; max_dst <= 16*W - 1)
static inline bool key_thread(int key) { return key >= 4000; }

// The strip class of an item whose widest band is D: the smallest S whose strip reaches the wanted goal-side width on the side(s)
// the item's candidates can have their goal on; -1: run the full band only (tiny bands, or no gain).
static int narrow_class_for_band_uncached(int D, int g256, bool both_sides);
static int narrow_class_for_band(int D, int g256, bool both_sides)
{ // called once per read and step: memoised per band half-width (g256 is fixed for the life of the process)
    static thread_local std::vector<int> memo[2];
    std::vector<int> &m = memo[both_sides ? 1 : 0];
    if (D < 0 || D > (1 << 16)) return narrow_class_for_band_uncached(D, g256, both_sides);
    if ((size_t)D >= m.size()) m.resize((size_t)D + 1024, INT_MIN);
    if (m[D] == INT_MIN) m[D] = narrow_class_for_band_uncached(D, g256, both_sides);
    return m[D];
}
static int narrow_class_for_band_uncached(int D, int g256, bool both_sides)
{
    const int full = class_for_band_plain(D);
    for (int S : kNarrowClasses) {
        if (full > 0 && S > full) break;
        int Wl, NBw, Wg;
        const int tgt = nb_target(D, g256);
        if (nb_policy(D, S, 0, tgt, &Wl, &NBw, &Wg) != 2) continue;
        if (both_sides && nb_policy(D, S, 1, tgt, &Wl, &NBw, &Wg) != 2) continue;
        return 3000 + S;
    }
    return -1;
}

static int class_for_band(int D, bool irr)
{ // smallest S with 32*S words >= ceil((2D+1)/32); returns the class key or -1
    const int NW = (2 * D + 1 + 31) >> 5;
    if (irr) {
        for (int S : kIrrClasses)
            if (32 * S >= NW) return 1000 + S;
        return -1;
    }
    for (int S : kClasses)
        if (32 * S >= NW) return S;
    return -1;
}

struct ClassPlan {
    std::vector<int32_t> items;
    int max_rows = 0, max_D = 0;
    double work = 0; // estimated instruction count of the class: its share of the machine
};

template <int S, bool IRR> struct KernelSel {
    static const void *locate() { return (const void *)align_locate_kernel<S, IRR, false>; }
    static const void *locate_pairs() { return (const void *)align_locate_kernel<S, false, true>; }
    static const void *pairs() { return (const void *)align_pairs_kernel<S, IRR>; }
};
template <int S> struct NarrowSel {
    static const void *locate() { return (const void *)align_locate_nb_kernel<S>; }
    static const void *pairs() { return (const void *)align_pairs_nb_kernel<S>; }
};

// locate: 0 = pairs of sequences (pb_align_batch), 1 = locate / overlap items, 2 = all-vs-all items
static const void *kernel_ptr(int key, int locate)
{
    if (key_thread(key)) {
        switch (key_S(key)) {
            case 3: return (const void *)align_pairs_thread_kernel<3>;
            case 5: return (const void *)align_pairs_thread_kernel<5>;
            case 9: return (const void *)align_pairs_thread_kernel<9>;
        }
        return nullptr;
    }
    if (key_narrow(key)) {
        switch (key_S(key)) {
#define CASE(s) case s: return locate ? NarrowSel<s>::locate() : NarrowSel<s>::pairs();
            CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
        }
        return nullptr;
    }
    if (key_packed(key)) {
        switch (key_lanes(key)) {
            case 4: return (const void *)align_pairs_packed_kernel<4>;
            case 8: return (const void *)align_pairs_packed_kernel<8>;
            case 16: return (const void *)align_pairs_packed_kernel<16>;
        }
        return nullptr;
    }
    const int S = key_S(key);
    if (key_irr(key)) {
        switch (S) {
#define CASE(s) case s: return locate ? KernelSel<s, true>::locate() : KernelSel<s, true>::pairs();
            CASE(2) CASE(6) CASE(16)
#undef CASE
        }
        return nullptr;
    }
    switch (S) {
#define CASE(s) case s: return locate == 2 ? KernelSel<s, false>::locate_pairs() : (locate ? KernelSel<s, false>::locate() : KernelSel<s, false>::pairs());
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(14) CASE(16)
#undef CASE
    }
    return nullptr;
}

struct LaunchGeom {
    int PW, RW, warp_words;
    size_t smem_bytes, slot_words, par_words;
    int blocks;     // CTAs to launch
    int max_blocks; // min(full occupancy, one warp per item)
    int groups;     // alignments in flight per warp (1, or 32/LANES for the packed narrow-band kernels)
    int wpb;        // warps per CTA
};

static int plan_launch(pb_ctx *ctx, int key, const ClassPlan &cp, int locate, size_t scratch_budget, LaunchGeom *g)
{
    const size_t ops_bytes = ((size_t)2 * cp.max_rows + cp.max_D + 64 + 127) & ~(size_t)127; // keeps every slot 128 B aligned
    if (key_thread(key)) {
        const int W = key_S(key);
        g->groups = 32;                        // alignments per warp
        g->PW = W + 1;
        g->RW = 0;
        g->warp_words = 32 * (4 * (W + 1) + 2 * W); // per thread: 4 planes x (W+1) Eq words + Vp / Vn
        g->par_words = (size_t)std::max(cp.max_rows, 1) * 2 * W; // per alignment; a warp's 32 slots interleave as [row][word][lane]
    } else if (key_packed(key)) {
        const int LANES = key_lanes(key);
        g->groups = 32 / LANES;
        g->PW = (((cp.max_rows + 31) >> 5) + LANES + 2 + 3) & ~3;
        g->RW = 0;
        g->warp_words = g->groups * 4 * g->PW;
        g->par_words = ((size_t)std::max(cp.max_rows, 1) + 1) * 2 * LANES; // one extra row: the dump for finished groups
        g->par_words = (g->par_words + 31) & ~(size_t)31;
    } else {
        const int S = key_S(key), T = 32 * S;
        g->groups = 1;
        const int logical = ((cp.max_rows + 31) >> 5) + T + 2;
        g->PW = logical + ((S % PB_PAD_MOD) == 0 && !key_narrow(key) ? logical / S + 2 : 0); // padded classes: one pad word per S words (bank conflicts)
        g->PW = (g->PW + 3) & ~3; // 16-byte aligned sub-arrays (TMA destination)
        g->RW = PB_STAGE_WORDS; // TMA staging buffer per raw plane
        g->warp_words = (key_irr(key) ? 8 : 4) * g->PW + (key_irr(key) ? 3 : 2) * g->RW;
        if (key_narrow(key)) { // the strip pass keeps its planes for the traceback: goal scan / window ring behind them
            g->warp_words += nb_tb_words(S);
            g->par_words = nb_tile_words(S) + nb_ck_words(S, std::max(cp.max_rows, 1)); // a round's tiles + one checkpoint per block
        } else {
            g->warp_words = std::max(g->warp_words, (2 * T + PB_TB_RING_WORDS + 3) & ~3); // final deltas + the traceback's window ring
            g->par_words = (size_t)std::max(cp.max_rows, 1) * 2 * T;
        }
    }
    // fewer warps per CTA when the per-warp planes are large (long sequences): the packed kernels stay under 96 KB so that two
    // CTAs fit an SM, the others under the 200 KB a CTA may ask for
    g->wpb = ALIGN_WPB;
    const size_t smem_cap = key_packed(key) ? 96 * 1024 : 200 * 1024;
    while (g->wpb > 1 && (size_t)g->wpb * g->warp_words * sizeof(uint32_t) > smem_cap) g->wpb >>= 1;
    // packed kernels, few items: spread them over more, smaller CTAs rather than leave SMs without work.  (Not for the
    // one-alignment-per-warp kernels: measured on config 2, smaller CTAs for the sparse wide-band classes cost 5 %.)
    while ((key_packed(key) || key_thread(key)) && g->wpb > 2 && (int64_t)cp.items.size() < (int64_t)g->wpb * g->groups * ctx->sm_count) g->wpb >>= 1;
    g->smem_bytes = (size_t)(key_thread(key) ? ALIGN_WPB : g->wpb) * g->warp_words * sizeof(uint32_t); // thread kernels: fixed column stride
    g->slot_words = g->par_words + ops_bytes / 4;
    const void *fn = kernel_ptr(key, locate);
    if (g->smem_bytes > 200 * 1024) return pb_fail(ctx, PB_ERR_DOMAIN, "sequence of %d rows needs %zu bytes of shared memory per CTA", cp.max_rows, g->smem_bytes);
    int occ = 0;
    { // the attribute and the occupancy query cost ~20 us each and are asked for every class of every step: remembered per
      // (kernel, warps per CTA, shared memory); the attribute only ever needs raising
        struct Key {
            const void *fn; int dev, wpb; size_t smem;
            bool operator<(const Key &o) const
            { return fn != o.fn ? fn < o.fn : (dev != o.dev ? dev < o.dev : (wpb != o.wpb ? wpb < o.wpb : smem < o.smem)); }
        };
        static thread_local std::map<Key, int> occ_memo;
        static thread_local std::map<std::pair<const void *, int>, size_t> smem_set;
        const Key key = {fn, ctx->device, g->wpb, g->smem_bytes};
        auto it = occ_memo.find(key);
        if (it == occ_memo.end()) {
            size_t &have = smem_set[std::make_pair(fn, ctx->device)];
            if (g->smem_bytes > have) {
                PB_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem_bytes));
                have = g->smem_bytes;
            }
            PB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, g->wpb * 32, g->smem_bytes));
            occ_memo[key] = occ;
        } else occ = it->second;
    }
    if (occ < 1) occ = 1;
    int64_t blocks = (int64_t)occ * ctx->sm_count;
    const int64_t per_cta = (int64_t)g->wpb * g->groups; // alignments in flight per CTA
    blocks = std::min<int64_t>(blocks, ((int64_t)cp.items.size() + per_cta - 1) / per_cta);
    const size_t slot_bytes = g->slot_words * 4;
    const int64_t by_mem = (int64_t)(scratch_budget / (slot_bytes * per_cta));
    if (by_mem < 1) return pb_fail(ctx, PB_ERR_NOMEM, "scratch budget %zu too small for one CTA (%zu bytes per alignment)", scratch_budget, slot_bytes);
    blocks = std::max<int64_t>(1, std::min(blocks, by_mem));
    g->blocks = (int)blocks;
    g->max_blocks = (int)blocks;
    return PB_OK;
}

static size_t scratch_budget(pb_ctx *ctx)
{
    if (ctx->scratch_budget_cached) { // fixed after the first call: a moving budget would re-plan (and re-allocate) every step
        size_t b = ctx->scratch_budget_cached;
        if (ctx->scratch_limit && ctx->scratch_limit < b) b = ctx->scratch_limit;
        return b;
    }
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); fr = (size_t)8 << 30; }
    size_t b = (size_t)((double)(fr + ctx->scratch_bytes) * 0.5); // what we already hold counts as available
    ctx->scratch_budget_cached = b;
    if (ctx->scratch_limit && ctx->scratch_limit < b) b = ctx->scratch_limit;
    return b;
}

template <class LaunchFn>
static int run_classes(pb_ctx *ctx, std::map<int, ClassPlan> &plans, int locate, AlignLaunch base, LaunchFn &&launch, bool spread = true,
                       double slots_scale = 1.0)
{
    if (plans.empty()) return PB_OK;
    const size_t budget = scratch_budget(ctx);
    // geometry per class; every class gets its own scratch region so that the launches can overlap
    std::map<int, LaunchGeom> geoms;
    size_t need = 0;
    for (auto &kv : plans) {
        LaunchGeom g;
        PB_TRY(plan_launch(ctx, kv.first, kv.second, locate, budget, &g));
        geoms[kv.first] = g;
        need += (size_t)g.blocks * g.wpb * g.groups * g.slot_words * 4;
    }
    // The classes run concurrently and share the SMs, so give each a share of the resident warps (and of the
    // scratch) in proportion to its share of the DP cells: they then drain at about the same time.
    {
        double total_work = 0;
        for (auto &kv : plans) total_work += kv.second.work;
        const char *wenv = getenv("PB_WARPS_PER_SM"); // tuning knob: resident aligner warps per SM shared by all classes
        const double warp_slots = (double)ctx->sm_count * (wenv ? atof(wenv) : 24.0) * slots_scale;
        need = 0;
        for (auto &kv : geoms) {
            const ClassPlan &cp = plans[kv.first];
            const double share = total_work > 0 ? cp.work / total_work : 1.0 / plans.size();
            const int want = (int)(warp_slots * share / kv.second.wpb + 0.999);
            kv.second.blocks = std::max(1, std::min(kv.second.blocks, want));
            need += (size_t)kv.second.blocks * kv.second.wpb * kv.second.groups * kv.second.slot_words * 4;
        }
        if (need > budget) { // still too much scratch: shrink every grid by the same factor (at least one CTA each)
            const double f = (double)budget / (double)need;
            need = 0;
            for (auto &kv : geoms) {
                kv.second.blocks = std::max(1, (int)(kv.second.blocks * f));
                need += (size_t)kv.second.blocks * kv.second.wpb * kv.second.groups * kv.second.slot_words * 4;
            }
        } else if (spread) {
            // Spend what is left of the budget on extra CTAs for the narrow-band classes (cheap slots): they wait behind
            // the wide-band kernels launched before them and move onto SMs as those drain, which balances the tail.
            for (auto &kv : geoms) { // ascending band class
                LaunchGeom &g = kv.second;
                const size_t per_block = (size_t)g.wpb * g.groups * g.slot_words * 4;
                const int64_t room = (int64_t)((budget - need) / per_block);
                const int extra = (int)std::max<int64_t>(0, std::min<int64_t>(room, (int64_t)g.max_blocks - g.blocks));
                g.blocks += extra;
                need += (size_t)extra * per_block;
            }
        }
    }
    DevBuf d_order, d_queue;
    if (need + 256 > ctx->scratch_bytes) { // grow-only, kept across calls
        PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch) cudaFree(ctx->scratch);
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        // some headroom for the next batch, but never past the budget the plan was cut to
        size_t want = std::min(need + need / 8, std::max(budget, need)) + 256;
        cudaError_t e = cudaMalloc(&ctx->scratch, want);
        if (e != cudaSuccess && want > need + 256) { // no room for the headroom: ask for exactly what this plan needs
            cudaGetLastError();
            want = need + 256;
            e = cudaMalloc(&ctx->scratch, want);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            ctx->scratch = nullptr;
            ctx->scratch_budget_cached = 0; // free memory has changed since the budget was taken: the next call measures it again
            return pb_fail(ctx, PB_ERR_NOMEM, "aligner scratch of %zu bytes: %s", want, cudaGetErrorString(e));
        }
        ctx->scratch_bytes = want;
    }
    size_t nitems = 0;
    for (auto &kv : plans) nitems += kv.second.items.size();
    PB_TRY(d_order.alloc(ctx, nitems * 4 + 16));
    PB_TRY(d_queue.alloc_zero(ctx, (size_t)plans.size() * 4 + 16));
    // launch order: widest band first.  Those kernels have few, long-running alignments and must be resident from the
    // start; the narrow-band kernels behind them are oversubscribed and fill the SMs as CTAs retire.  (Measured: launching
    // the heaviest class first starves the wide classes, 134 ms vs 105 ms per config-2 step.)
    std::vector<int> launch_order;
    for (auto it = plans.rbegin(); it != plans.rend(); ++it) launch_order.push_back(it->first);
    {
        std::vector<int32_t> all;
        all.reserve(nitems);
        for (int c : launch_order) all.insert(all.end(), plans[c].items.begin(), plans[c].items.end());
        PB_TRY(pb_h2d(ctx, d_order.p, all.data(), nitems * 4));
    }
    // one stream per band class: a class with few, long alignments no longer leaves the other SMs idle
    while (ctx->aux_streams.size() < plans.size()) {
        cudaStream_t st;
        cudaEvent_t ev;
        PB_CUDA(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        PB_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->aux_streams.push_back(st);
        ctx->aux_events.push_back(ev);
    }
    if (!ctx->fork_event) PB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
    const bool trace = getenv("PB_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    std::vector<std::string> tdesc;
    PB_CUDA(ctx, cudaEventRecord(ctx->fork_event, ctx->stream));
    size_t off = 0, soff = 0;
    int ci = 0;
    for (int cls : launch_order) {
        ClassPlan &cp = plans[cls];
        const LaunchGeom &g = geoms[cls];
        AlignLaunch p = base;
        p.PW = g.PW;
        p.RW = g.RW;
        p.warp_words = g.warp_words;
        p.wpb = g.wpb;
        p.slot_words = g.slot_words;
        p.par_words = g.par_words;
        p.scratch = reinterpret_cast<uint32_t *>(ctx->scratch) + soff;
        p.queue = d_queue.as<int>() + ci;
        p.order = d_order.as<int32_t>() + off;
        p.nitems = (int)cp.items.size();
        cudaStream_t st = ctx->aux_streams[ci];
        PB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->fork_event, 0));
        if (trace) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            tev.push_back(a); tev.push_back(b);
            char buf[256];
            snprintf(buf, sizeof buf, "S=%d%s items=%zu blocks=%d rows<=%d band<=%d slotMB=%.2f work=%.3g", key_S(cls), key_irr(cls) ? "x" : (key_packed(cls) ? "p" : ""), cp.items.size(), g.blocks, cp.max_rows, cp.max_D, g.slot_words * 4 / 1048576.0, cp.work);
            tdesc.push_back(buf);
            cudaEventRecord(a, st);
        }
        PB_TRY(launch(cls, p, g, st));
        if (trace) cudaEventRecord(tev.back(), st);
        PB_CUDA(ctx, cudaEventRecord(ctx->aux_events[ci], st));
        PB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->aux_events[ci], 0));
        off += cp.items.size();
        soff += (size_t)g.blocks * g.wpb * g.groups * g.slot_words;
        ++ci;
    }
    if (trace) { // PB_TRACE=1: per-class device times (classes overlap, so they do not add up)
        cudaStreamSynchronize(ctx->stream);
        for (size_t c = 0; c < tdesc.size(); ++c) {
            float ms = 0, ms0 = 0;
            cudaEventElapsedTime(&ms, tev[2 * c], tev[2 * c + 1]);
            cudaEventElapsedTime(&ms0, tev[0], tev[2 * c]);
            fprintf(stderr, "[pb_trace] %s start=+%.2fms dur=%.2fms\n", tdesc[c].c_str(), ms0, ms);
        }
        for (auto e : tev) cudaEventDestroy(e);
        fprintf(stderr, "[pb_trace] scratch %.2f GB of budget %.2f GB\n", need / 1e9, budget / 1e9);
    }
    return PB_OK;
}

// The host-side plan of a locate step: which band class every item runs in, in both passes.  It depends on the read lengths
// only, so the pipelined entry points build it (pb_align_locate_prepare) before they wait for the previous step's kernels, and
// the planning -- ~2 ms for 100 k reads -- no longer sits between the prefix filter and the aligner with the GPU idle.
struct LocatePlan {
    std::vector<int32_t> lens;
    double R = 0;
    int maxn = 0, maxm = 0, mode = 0;
    bool any_irr = false;
    std::map<int, ClassPlan> plans, narrow_plans;
    int64_t n_narrow = 0;
    int g256 = 192;
};

static int build_locate_plan(pb_ctx *ctx, const std::vector<int32_t> &kept_lens, const std::vector<uint8_t> *kept_irr, double R,
                             int maxn, int maxm, int mode, LocatePlan *lp)
{
    const int64_t nkept = (int64_t)kept_lens.size();
    lp->lens = kept_lens;
    lp->R = R; lp->maxn = maxn; lp->maxm = maxm; lp->mode = mode;
    lp->any_irr = false;
    lp->plans.clear(); lp->narrow_plans.clear(); lp->n_narrow = 0;
    // longest first, ties in input order: a counting sort (all-vs-all batches hold millions of items)
    std::vector<int32_t> order((size_t)nkept);
    {
        int maxlen = 0;
        for (int64_t k = 0; k < nkept; ++k) maxlen = std::max(maxlen, kept_lens[k]);
        std::vector<int64_t> first((size_t)maxlen + 2, 0);
        for (int64_t k = 0; k < nkept; ++k) first[(size_t)(maxlen - kept_lens[k]) + 1]++;
        for (size_t l = 1; l < first.size(); ++l) first[l] += first[l - 1];
        for (int64_t k = 0; k < nkept; ++k) order[(size_t)first[(size_t)(maxlen - kept_lens[k])]++] = (int32_t)k;
    }
    // Two passes.  First pass (K3n, align_locate_nb_kernel): every plain-ACGT item whose band admits a certified strip runs over
    // the strip; what it certifies is final.  Second pass (align_locate_kernel, full band): the items the first pass could not
    // take (byte-exact variant, tiny bands) plus those it flagged -- the kernel skips every item whose redo flag is clear, so the
    // second pass is planned for the items that are its own plus a small allowance.  PB_NARROW=0 runs the full band only.
    static const bool narrow_on = !(getenv("PB_NARROW") && atoi(getenv("PB_NARROW")) == 0);
    // goal-side width of the strip as a share of max_dst: what a certified alignment may cost.  A/B on config 2 (CLR reads, cost
    // 0.15-0.22 of the length at R = 0.3): 0.80 -> 57.9 ms of K3 and 8 reads redone, 0.75 -> 57.0 / 10, 0.70 -> 55.5 / 136,
    // 0.65 -> 6063 redone.  Whatever the value, results are exact: the certificate decides, the full band redoes the rest.
    static const int g256 = getenv("PB_NARROW_G") ? std::max(128, std::min(256, (int)(atof(getenv("PB_NARROW_G")) * 256.0))) : 192;
    lp->g256 = g256;
    // The strip pass runs in locate mode only.  In the overlap modes the goal may lie on either side, so the strip needs the goal
    // width on both (1.5 max_dst + 64 against the band's 2 max_dst + 1: rarely a class lower), three of four candidate alignments
    // fail within a few blocks and the overlaps that succeed are short: measured on config 5 (all-vs-all, 50 k reads) the two
    // passes cost 381 ms of K3 where the full band alone takes 268 ms.
    const bool both_sides = mode != PB_MODE_LOCATE;
    // the class of an item is a function of its length (and of the rare non-ACGT flag): looked up once per distinct length
    struct PerLen { int cls = INT_MIN, ncls = -1, D = 0, rows = 0; double w_full = 0, w_narrow = 0; };
    std::vector<PerLen> by_len[2];
    for (int32_t k : order) {
        const int L = kept_lens[k];
        const int irr = kept_irr && (*kept_irr)[k] ? 1 : 0;
        lp->any_irr = lp->any_irr || irr;
        std::vector<PerLen> &tab = by_len[irr];
        if ((size_t)L >= tab.size()) tab.resize((size_t)L + 1);
        PerLen &pl = tab[L];
        if (pl.cls == INT_MIN) {
            // widest band / longest seg_a any candidate of this read can reach the DP with: len_a <= L, and the domain
            // check (seq_aligner.h:104) turns away len_a >= maxn or max_dst >= maxm before any row is computed
            pl.D = std::min(1 + (int)(L * R), maxm - 1);
            pl.cls = class_for_band(std::max(pl.D, 1), irr != 0); // non-ACGT bytes in the read (or the contig): byte-exact variant
            if (pl.cls < 0) return pb_fail(ctx, PB_ERR_DOMAIN, "band half-width %d exceeds PB_MAX_BAND", pl.D);
            // rows = len_a: the read itself (locator), or the reference view cut to len_b + max_dst (overlap, seq_aligner.h:100)
            pl.rows = std::min(mode == PB_MODE_LOCATE ? L : L + pl.D, std::max(maxn - 1, 1));
            pl.ncls = (narrow_on && !irr && !both_sides) ? narrow_class_for_band(pl.D, g256, false) : -1;
            // ~instructions: rows x (per-word + per-row cost); an item of the first pass comes back only when it could not be certified
            pl.w_full = (double)L * (30.0 * key_S(pl.cls) + 60.0) * (pl.ncls < 0 ? 1.0 : 0.02);
            pl.w_narrow = pl.ncls < 0 ? 0.0 : (double)L * (16.0 * key_S(pl.ncls) + 40.0);
        }
        ClassPlan &cp = lp->plans[pl.cls];
        cp.items.push_back(k);
        cp.max_rows = std::max(cp.max_rows, pl.rows);
        cp.max_D = std::max(cp.max_D, pl.D);
        cp.work += pl.w_full;
        if (pl.ncls >= 0) {
            ClassPlan &np = lp->narrow_plans[pl.ncls];
            np.items.push_back(k);
            np.max_rows = std::max(np.max_rows, pl.rows);
            np.max_D = std::max(np.max_D, pl.D);
            np.work += pl.w_narrow;
            ++lp->n_narrow;
        }
    }
    return PB_OK;
}

void pb_locate_plan_free(LocatePlan *lp) { delete lp; }

// Pipelined callers: plan the step from the kept reads' lengths before anything of it is queued (no non-ACGT reads assumed;
// a batch that turns out to hold some is planned again when it arrives).
int pb_align_locate_prepare(pb_ctx *ctx, const std::vector<int32_t> &kept_lens, double R, int maxn, int maxm, int mode)
{
    LocatePlan *lp = new LocatePlan();
    int r = build_locate_plan(ctx, kept_lens, nullptr, R, maxn, maxm, mode, lp);
    if (r != PB_OK) { delete lp; return r; }
    if (ctx->planned) delete ctx->planned;
    ctx->planned = lp;
    return PB_OK;
}

int pb_align_locate(pb_ctx *ctx, const SeqSets &ss, const LocateView &lv, int64_t nkept,
                    const std::vector<int32_t> &kept_lens, const std::vector<uint8_t> &kept_irr, double R, int maxn, int maxm,
                    const uint8_t *d_survive, const int32_t *d_rej_cells, pb_locate_rec *d_recs, uint8_t *d_ops,
                    const int64_t *d_ops_off, unsigned long long *d_stats)
{
    if (nkept == 0) { pb_timer_begin(ctx, PB_T_ALIGN); return PB_OK; }
    LocatePlan local, *lp = nullptr;
    if (ctx->planned) { // a plan made ahead of time: valid if it was made for exactly these items
        LocatePlan *pp = ctx->planned;
        bool irr = false;
        for (uint8_t f : kept_irr) irr = irr || f;
        if (!irr && pp->R == R && pp->maxn == maxn && pp->maxm == maxm && pp->mode == lv.mode && pp->lens.size() == kept_lens.size() &&
            memcmp(pp->lens.data(), kept_lens.data(), kept_lens.size() * sizeof(int32_t)) == 0)
            lp = pp;
    }
    if (!lp) {
        PB_TRY(build_locate_plan(ctx, kept_lens, &kept_irr, R, maxn, maxm, lv.mode, &local));
        lp = &local;
    }
    std::map<int, ClassPlan> &plans = lp->plans, &narrow_plans = lp->narrow_plans;
    const int64_t n_narrow = lp->n_narrow;
    const int g256 = lp->g256;
    AlignLaunch base;
    memset(&base, 0, sizeof base);
    {
        PairViews pv = pair_views(ss);
        base.A = pv.A; base.B = pv.B; base.A2 = pv.A2; base.B2 = pv.B2;
    }
    base.R = R; base.maxn = maxn; base.maxm = maxm;
    base.ops = d_ops; base.ops_off = d_ops_off;
    base.stats = d_stats;
    base.g256 = g256;
    const int kmode = lv.d_item_ref ? 2 : 1;
    auto launch = [&](int key, const AlignLaunch &p, const LaunchGeom &g, cudaStream_t st) -> int {
        void *args[] = {(void *)&p, (void *)&lv, (void *)&d_survive, (void *)&d_rej_cells, (void *)&d_recs};
        PB_CUDA(ctx, cudaLaunchKernel(kernel_ptr(key, kmode), dim3(g.blocks), dim3(g.wpb * 32), args, g.smem_bytes, st));
        ctx->launches++;
        return PB_OK;
    };
    pb_timer_begin(ctx, PB_T_ALIGN); // the callers close the stage; planning above is host time, not kernel time
    DevBuf d_redo;
    if (n_narrow > 0) {
        PB_TRY(d_redo.alloc(ctx, (size_t)nkept + 16));
        PB_CUDA(ctx, cudaMemsetAsync(d_redo.p, 1, (size_t)nkept, ctx->stream)); // items the first pass never sees stay flagged
        base.redo = d_redo.as<uint8_t>();
        PB_TRY(run_classes(ctx, narrow_plans, kmode, base, launch));
    }
    // a second pass stays as small as its own items need: nearly all of them have been settled by the first, so it gets an eighth
    // of the resident warps (and of the scratch: its slots hold stored parents, 15x the first pass's)
    return run_classes(ctx, plans, kmode, base, launch, n_narrow == 0, n_narrow > 0 && (double)n_narrow >= 0.9 * (double)nkept ? 0.125 : 1.0);
}

int pb_align_pairs(pb_ctx *ctx, const pb_seqset *A, const pb_seqset *B, int64_t n, double R, int maxn, int maxm,
                   pb_align_out *d_out, uint8_t *d_ops, const int64_t *d_ops_off)
{
    if (n == 0) { pb_timer_begin(ctx, PB_T_ALIGN); return PB_OK; }
    std::map<int, ClassPlan> plans, narrow_plans;
    int64_t n_narrow = 0;
    // wide bands (one warp per alignment) go through the certified strip first, as in pb_align_locate: PB_NARROW=0 turns it off
    static const bool narrow_on = !(getenv("PB_NARROW") && atoi(getenv("PB_NARROW")) == 0);
    static const int g256 = getenv("PB_NARROW_G") ? std::max(128, std::min(256, (int)(atof(getenv("PB_NARROW_G")) * 256.0))) : 192;
    std::vector<int32_t> order((size_t)n);
    for (int64_t k = 0; k < n; ++k) order[k] = (int32_t)k;
    std::vector<int> la((size_t)n), D((size_t)n);
    std::vector<uint8_t> goal_left((size_t)n);
    for (int64_t k = 0; k < n; ++k) {
        int lb;
        pb_align_params(A->len[k], B->len[k], R, &la[k], &lb, &D[k]);
        goal_left[k] = la[k] > lb; // the goal cell is searched on the last column (seq_aligner.h:191-213)
    }
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return (int64_t)la[x] * D[x] > (int64_t)la[y] * D[y]; });
    for (int32_t k : order) {
        const bool rejected = la[k] >= maxn || D[k] >= maxm;
        int cls = 1; // rejected by the domain check: any kernel will do, nothing is computed
        if (!rejected) {
            const bool irr = ((A->flags[k] | B->flags[k]) & PB_FLAG_IRREGULAR) != 0;
            const int NW = (2 * D[k] + 1 + 31) >> 5;
            if (!irr && NW <= 9 && !getenv("PB_NO_THREAD")) cls = 4000 + (NW <= 3 ? 3 : (NW <= 5 ? 5 : 9)); // one alignment per thread
            else if (!irr && NW <= 16 && !getenv("PB_NO_PACKED")) cls = 2000 + (NW <= 4 ? 4 : (NW <= 8 ? 8 : 16)); // several alignments per warp
            else cls = class_for_band(D[k], irr);
            if (cls < 0) return pb_fail(ctx, PB_ERR_DOMAIN, "band half-width %d exceeds PB_MAX_BAND", D[k]);
        }
        int ncls = -1;
        if (!rejected && narrow_on && cls < 1000) ncls = narrow_class_for_band(D[k], g256, goal_left[k] != 0);
        ClassPlan &cp = plans[cls];
        cp.items.push_back(k);
        if (!rejected) {
            cp.max_rows = std::max(cp.max_rows, la[k]);
            cp.max_D = std::max(cp.max_D, D[k]);
            cp.work += key_thread(cls) ? (double)la[k] * (19.0 * key_S(cls) + 20.0) / 32.0
                       : (key_packed(cls) ? (double)la[k] * 90.0 * key_lanes(cls) / 32.0
                                          : (double)la[k] * (30.0 * key_S(cls) + 60.0) * (ncls < 0 ? 1.0 : 0.02));
        }
        if (ncls >= 0) { // the full-band class above only sees it again when the strip could not certify it
            ClassPlan &np = narrow_plans[ncls];
            np.items.push_back(k);
            np.max_rows = std::max(np.max_rows, la[k]);
            np.max_D = std::max(np.max_D, D[k]);
            np.work += (double)la[k] * (16.0 * key_S(ncls) + 40.0);
            ++n_narrow;
        }
    }
    pb_timer_begin(ctx, PB_T_ALIGN); // the caller closes the stage; planning above is host time, not kernel time
    AlignLaunch base;
    memset(&base, 0, sizeof base);
    base.A = seq_view(A);
    base.B = seq_view(B);
    base.R = R; base.maxn = maxn; base.maxm = maxm;
    base.ops = d_ops; base.ops_off = d_ops_off;
    base.g256 = g256;
    auto launch = [&](int key, const AlignLaunch &p, const LaunchGeom &g, cudaStream_t st) -> int {
        void *args[] = {(void *)&p, (void *)&d_out};
        PB_CUDA(ctx, cudaLaunchKernel(kernel_ptr(key, 0), dim3(g.blocks), dim3(g.wpb * 32), args, g.smem_bytes, st));
        ctx->launches++;
        return PB_OK;
    };
    DevBuf d_redo;
    if (n_narrow > 0) {
        PB_TRY(d_redo.alloc(ctx, (size_t)n + 16));
        PB_CUDA(ctx, cudaMemsetAsync(d_redo.p, 1, (size_t)n, ctx->stream)); // pairs the first pass never sees stay flagged
        base.redo = d_redo.as<uint8_t>();
        PB_TRY(run_classes(ctx, narrow_plans, 0, base, launch));
    }
    return run_classes(ctx, plans, 0, base, launch, n_narrow == 0, n_narrow > 0 && (double)n_narrow >= 0.9 * (double)n ? 0.125 : 1.0);
}
