// pb_alignw.cu -- quality-weighted banded aligner (BASELINE config 3, "quality-weighted scoring").
//
// A LABELLED EXTENSION: the reference has no quality-aware DP -- its scoring hooks match(c,d) = (c != d) and indel(c) = 1 are
// hard-wired (seq_aligner.h:136-137) and quality.cpp is a stand-alone mean-of-ASCII tool.  This is seq_aligner::align with
// those two hooks replaced by per-element table lookups, everything else kept: parameter derivation (seq_aligner.h:94-107),
// init_cell (:139-150), the band and its edge rules, the tie-breaking order diag / left-if-strictly-smaller /
// up-if-strictly-smaller (:164-173), early failure (:185, threshold scaled by fail_scale), goal_cell (:191-213), the coverage
// test (:114) and find_path (:214-233).
//     match(a_i, b_j) = (a_i != b_j) ? wa[i] : 0        DELETE (a_i skipped) costs wa[i]        INSERT (b_j skipped) costs wb[j]
// With all weights 1 and fail_scale 1 it is the reference's aligner (tested against pb_align_batch and the oracle).
//
// Weights break the +-1 delta structure the bit-parallel kernels (pb_align.cu) live on, so this one is the textbook wavefront:
// one warp per alignment, the cells of an anti-diagonal i + j = d are independent and are spread over the lanes, costs are
// 32-bit integers in shared memory indexed by diagonal k = j - i + max_dst.  Cells of one anti-diagonal all have the same
// parity of k and read only neighbours of the other parity (left = k-1, up = k+1, both finished one step earlier) and their own
// slot (diag, two steps earlier), so the update is in place with one __syncwarp per step and no shuffles.  Parents are packed two
// bits per cell into one word per lane and step ([step][word][lane]: every store instruction writes 128 contiguous bytes); the
// traceback fetches them 32 anti-diagonals at a time.  No tensor cores: integer min/add work.
#include <limits.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "pb_internal.cuh"

#define FULL 0xffffffffu
#define W_INF 0x3fffffff
#define WALIGN_WPB 8

struct WLaunch {
    const uint8_t *a, *wa, *b, *wb;
    const int64_t *a_off, *b_off;
    const int32_t *a_len, *b_len;
    int n;
    double R, fail_scale;
    int maxn, maxm;
    int cst_words;     // shared-memory ints per warp (2*Dmax + 3, padded)
    size_t slot_words; // scratch words per warp: parents + reversed transcript
    size_t par_words;
    uint32_t *scratch;
    int *queue;
    pb_align_out *out;
    uint8_t *ops;
    const int64_t *ops_off;
};

__device__ __forceinline__ long long w_cells_upto(int n, int D, int len_b)
{ // DP cells the reference evaluates in rows 1..n (seq_aligner.h:158-159)
    long long t = max(len_b - D, 0), m = min((long long)n, t);
    long long f = m * (m + 1) / 2 + m * D + ((long long)n - m) * len_b;
    long long m2 = min(n, D + 1);
    long long g = m2;
    if (n > D + 1) { long long x = n - D; g += x * (x + 1) / 2 - 1; }
    return f - g + n;
}

// goal_cell (seq_aligner.h:191-213), coverage test (:114) and find_path (:214-233) of one alignment whose last row / column
// sits in cst[] and whose parents were written as [step][word][lane] with cell t = k >> 1 in lane t / cpl, slot t % cpl
__device__ __forceinline__ void w_finish(const WLaunch &p, int idx, int lane, const int *cst, const uint32_t *par, uint8_t *opsrev,
                                         int len_a, int len_b, int D, int a_len, int cpl, int wpl, int fail_row, pb_align_out &o)
{
    if (fail_row) {
        o.fail_row = fail_row;
        o.cells = w_cells_upto(fail_row, D, len_b);
    } else {
        o.cells = w_cells_upto(len_a, D, len_b);
        // ---- goal_cell, seq_aligner.h:191-213: the last row (or column) still sits in cst[]
        int best = W_INF, bestpos = INT_MAX; // pos = j (last row) or i (last column); earliest strict minimum
        if (len_a > len_b) {
            for (int i = len_b + lane; i <= len_a; i += 32) {
                const int c = cst[len_b - i + D];
                if (c < best) { best = c; bestpos = i; }
            }
        } else {
            for (int j = len_a + lane; j <= len_b; j += 32) {
                const int c = cst[j - len_a + D];
                if (c < best) { best = c; bestpos = j; }
            }
        }
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            const int oc = __shfl_xor_sync(FULL, best, s), op = __shfl_xor_sync(FULL, bestpos, s);
            if (oc < best || (oc == best && op < bestpos)) { best = oc; bestpos = op; }
        }
        const int matlen_a = len_a > len_b ? bestpos : len_a, matlen_b = len_a > len_b ? len_b : bestpos;
        o.matlen_a = matlen_a; o.matlen_b = matlen_b; o.cost = best;
        o.diag_cost = (a_len <= len_a && a_len <= len_b) ? cst[D] : 0;
        if (!((double)matlen_b < len_b * (1 - p.R))) { // seq_aligner.h:114
            // ---- find_path, seq_aligner.h:214-233: lane L keeps the parent word of anti-diagonal dbase - L at the
            // (word, lane) position the path is in; a window lasts until the path leaves it or moves to another word
            int i = matlen_a, j = matlen_b, n = 0;
            int dbase = -1, w_word = -1, w_lane = -1;
            uint32_t pre = 0u;
            const int guard = len_a + len_b + 1;
            while ((i > 0 || j > 0) && n < guard) {
                const int d = i + j, k = j - i + D, t = k >> 1;
                const int ln = t / cpl, u = t - ln * cpl, wd = u >> 4;
                if (dbase < d || dbase - d > 31 || wd != w_word || ln != w_lane) {
                    dbase = d; w_word = wd; w_lane = ln;
                    const int dd = d - lane;
                    pre = dd >= 1 ? __ldcg(par + ((size_t)dd * wpl + wd) * 32 + ln) : 0u;
                }
                const uint32_t wv = __shfl_sync(FULL, pre, dbase - d);
                const uint32_t code = (wv >> (2 * (u & 15))) & 3u;
                if (lane == 0) opsrev[n] = (uint8_t)code;
                ++n;
                if (code == PB_MATCH) { --i; --j; }
                else if (code == PB_INSERT) --j;
                else if (code == PB_DELETE) --i;
                else break; // cannot happen: every cell on a path was computed
            }
            __syncwarp();
            if (p.ops) {
                uint8_t *dst = p.ops + p.ops_off[idx];
                for (int q = lane; q < n; q += 32) dst[q] = opsrev[n - 1 - q];
            }
            o.nedit = n;
            o.ret = matlen_b;
        }
    }
}

__global__ void __launch_bounds__(WALIGN_WPB * 32) alignw_kernel(const __grid_constant__ WLaunch p)
{
    extern __shared__ __align__(16) int wsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *cst = wsm + (size_t)warp * p.cst_words + 1; // cst[-1] and cst[2D+1] are the band's edge sentinels
    const size_t slot = (size_t)blockIdx.x * (blockDim.x >> 5) + warp;
    uint32_t *par = p.scratch + slot * p.slot_words;
    uint8_t *opsrev = reinterpret_cast<uint8_t *>(par + p.par_words);
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(p.queue, 1);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= p.n) break;
        const uint8_t *a = p.a + p.a_off[idx], *wa = p.wa + p.a_off[idx];
        const uint8_t *b = p.b + p.b_off[idx], *wb = p.wb + p.b_off[idx];
        const int a_len = p.a_len[idx], b_len = p.b_len[idx];
        int len_a, len_b, D;
        if (b_len >= a_len) { len_a = a_len; D = 1 + (int)(len_a * p.R); len_b = min(b_len, len_a + D); }
        else { len_b = b_len; D = 1 + (int)(len_b * p.R); len_a = min(a_len, len_b + D); }
        pb_align_out o;
        o.ret = -1; o.len_a = len_a; o.len_b = len_b; o.max_dst = D;
        o.matlen_a = o.matlen_b = o.cost = o.diag_cost = o.nedit = o.fail_row = 0;
        o.cells = 0;
        const bool dom = !(len_a >= p.maxn || D >= p.maxm); // seq_aligner.h:104-107
        const int cpl = (D + 1 + 31) / 32;  // cells per lane and step (cell t = k >> 1 belongs to lane t / cpl)
        const int wpl = (cpl + 15) / 16;     // parent words per lane and step
        int fail_row = 0;
        if (dom) {
            __syncwarp();
            for (int k = lane - 1; k <= 2 * D + 1; k += 32) cst[k] = W_INF;
            __syncwarp();
            if (lane == 0) cst[D] = 0; // cell (0,0)
            __syncwarp();
            const int nsteps = len_a + len_b;
            const int nfast = min(len_a, len_b);
            for (int d = 1; d <= nsteps; ++d) {
                // cells (i, d - i): max(0, d - len_b) <= i <= min(len_a, d), band 0 <= k <= 2D with k = d - 2i + D
                const int i_lo = max(0, d - len_b), i_hi = min(len_a, d);
                const int k_lo = max(0, d + D - 2 * i_hi), k_hi = min(2 * D, d + D - 2 * i_lo);
                const int par_k = (d + D) & 1;
                uint32_t *prow = par + (size_t)d * wpl * 32 + lane;
                uint32_t word = 0u;
                for (int u = 0; u < cpl; ++u) {
                    const int k = 2 * (lane * cpl + u) + par_k;
                    uint32_t code = 0u;
                    if (k >= k_lo && k <= k_hi) {
                        const int i = (d + D - k) >> 1, j = d - i;
                        const int wai = i > 0 ? (int)__ldg(wa + i - 1) : 1, wbj = j > 0 ? (int)__ldg(wb + j - 1) : 1;
                        const int sub = (i > 0 && j > 0 && __ldg(a + i - 1) != __ldg(b + j - 1)) ? wai : 0;
                        int c = cst[k] + sub;
                        code = PB_MATCH;
                        int t = cst[k - 1] + wbj;
                        if (t < c) { c = t; code = PB_INSERT; } // k == 0 reads the sentinel: no INSERT at i - j == max_dst
                        t = cst[k + 1] + wai;
                        if (t < c) { c = t; code = PB_DELETE; } // k == 2D likewise: no DELETE at j - i == max_dst
                        cst[k] = min(c, W_INF);
                    }
                    word |= code << (2 * (u & 15));
                    if ((u & 15) == 15 || u == cpl - 1) { prow[(u >> 4) * 32] = word; word = 0u; }
                }
                __syncwarp();
                if ((d & 1) == 0) { // row i = d/2 is complete up to its diagonal cell: early failure, seq_aligner.h:185
                    const int i = d >> 1;
                    if (i > 10 && i <= nfast && (double)cst[D] > i * p.R * p.fail_scale) { fail_row = i; break; }
                }
            }
            w_finish(p, idx, lane, cst, par, opsrev, len_a, len_b, D, a_len, cpl, wpl, fail_row, o);
        }
        if (lane == 0) p.out[idx] = o;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// The same wavefront with the costs in REGISTERS (bands of up to 64*CPL - 1 diagonals; CPL = 1, 2, 3, 5, 9, 17: config 3's
// points).  Lane L owns the 2*CPL consecutive diagonals k = 2*CPL*L .. 2*CPL*(L+1) - 1, i.e. the cells t = CPL*L + u of either
// parity -- the parents' layout above.  A step of parity P updates the lane's CPL slots 2u + P from their neighbours of the
// other parity, which are registers of the same lane except at the lane's two ends: ONE shuffle per step, no shared memory, no
// __syncwarp.  The CPL cells of a lane sit on consecutive rows and columns (i = iTop - u, j = jBase + u), and from one step to
// the next exactly one of iTop / jBase moves by one: the lane keeps a window of CPL elements (and weights) of each sequence in
// registers, shifts ONE of them per step and loads ONE new element, requested two steps before it is used -- the kernel above
// does four byte loads per cell.  Same arithmetic, same order of the comparisons (diag, left if strictly smaller, up if strictly
// smaller), same clamp; at the end the costs are spilled to cst[] and the goal cell / traceback are the shared code.
// ---------------------------------------------------------------------------------------------
// state of one lane: 2*CPL cost slots, the two element windows ({element, weight << 8} per entry; cell u uses a[i - 1] with
// i = iTop - u and b[j - 1] with j = jBase + u) and what the next shift of either window brings in
template <int CPL> struct WLane {
    int c[2 * CPL];
    int A[CPL], B[CPL];
    int nA, nB;
    int iTop, jBase;
};
struct WSeq { // one alignment's operands
    const uint8_t *a, *wa, *b, *wb;
    int len_a, len_b, D;
};
// elements outside a sequence read as {0, weight 1}: the weights of init_cell's row 0 and column 0
__device__ __forceinline__ int w_elem(const uint8_t *__restrict__ e, const uint8_t *__restrict__ w, int x, int len)
{
    int r = 1 << 8;
    if (x >= 0 && x < len) r = (int)__ldg(e + x) | ((int)__ldg(w + x) << 8);
    return r;
}

// One step (anti-diagonal d, parity P of its diagonals) for the CPL cells of this lane.
template <int CPL, int P>
__device__ __forceinline__ void w_step(WLane<CPL> &L, const WSeq &q, int d, int lane, uint32_t *__restrict__ par)
{
    constexpr int NS = 2 * CPL, WPL = (CPL + 15) / 16;
    if (P == 0) { // iTop moves: the a window shifts up by one element
#pragma unroll
        for (int u = CPL - 1; u > 0; --u) L.A[u] = L.A[u - 1];
        L.A[0] = L.nA;
        ++L.iTop;
        L.nA = w_elem(q.a, q.wa, L.iTop, q.len_a);
    } else {      // jBase moves: the b window
#pragma unroll
        for (int u = 0; u + 1 < CPL; ++u) L.B[u] = L.B[u + 1];
        L.B[CPL - 1] = L.nB;
        ++L.jBase;
        L.nB = w_elem(q.b, q.wb, L.jBase + CPL - 1, q.len_b);
    }
    // cells (i, d - i) of the matrix inside the band: max(0, d - len_b, ceil((d - D) / 2)) <= i <= min(len_a, d)
    const int i_lo = max(max(0, d - q.len_b), (d - q.D + 1) >> 1), i_hi = min(q.len_a, d);
    const unsigned x_lo = (unsigned)(L.iTop - i_lo), i_rng = (unsigned)(i_hi - i_lo); // cell u is one of them iff x_lo - u <= i_rng (unsigned)
    int edge; // the one neighbour that lives in another lane; the band's edges read as W_INF
    if (P == 0) { edge = __shfl_up_sync(FULL, L.c[NS - 1], 1); if (lane == 0) edge = W_INF; }
    else { edge = __shfl_down_sync(FULL, L.c[0], 1); if (lane == 31) edge = W_INF; }
    uint32_t word = 0u;
    uint32_t *prow = par + (size_t)d * (WPL * 32) + lane;
#pragma unroll
    for (int u = 0; u < CPL; ++u) {
        const int s = 2 * u + P;
        const int left = (s == 0) ? edge : L.c[s > 0 ? s - 1 : 0];
        const int up = (s == NS - 1) ? edge : L.c[s < NS - 1 ? s + 1 : NS - 1];
        const int wai = L.A[u] >> 8, wbj = L.B[u] >> 8;
        // (row 0 and column 0 need no test of their own: their diagonal neighbour does not exist, its slot holds W_INF, and
        // whatever is added to that loses against the one finite neighbour and is clamped away)
        const int sub = (((L.A[u] ^ L.B[u]) & 0xff) != 0) ? wai : 0;
        int cc = L.c[s] + sub;
        uint32_t code = PB_MATCH;
        int t = left + wbj;
        if (t < cc) { cc = t; code = PB_INSERT; }
        t = up + wai;
        if (t < cc) { cc = t; code = PB_DELETE; }
        if (x_lo - (unsigned)u <= i_rng) {
            L.c[s] = min(cc, W_INF);
            word |= code << (2 * (u & 15));
        }
        if ((u & 15) == 15 || u == CPL - 1) { prow[(u >> 4) * 32] = word; word = 0u; }
    }
}

template <int CPL>
__global__ void __launch_bounds__(WALIGN_WPB * 32) alignw_reg_kernel(const __grid_constant__ WLaunch p)
{
    extern __shared__ __align__(16) int wsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *cst = wsm + (size_t)warp * p.cst_words + 1;
    const size_t slot = (size_t)blockIdx.x * (blockDim.x >> 5) + warp;
    uint32_t *par = p.scratch + slot * p.slot_words;
    uint8_t *opsrev = reinterpret_cast<uint8_t *>(par + p.par_words);
    constexpr int NS = 2 * CPL;            // cost slots per lane
    constexpr int WPL = (CPL + 15) / 16;   // parent words per lane and step
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(p.queue, 1);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= p.n) break;
        WSeq q;
        q.a = p.a + p.a_off[idx]; q.wa = p.wa + p.a_off[idx];
        q.b = p.b + p.b_off[idx]; q.wb = p.wb + p.b_off[idx];
        const int a_len = p.a_len[idx], b_len = p.b_len[idx];
        int len_a, len_b, D;
        if (b_len >= a_len) { len_a = a_len; D = 1 + (int)(len_a * p.R); len_b = min(b_len, len_a + D); }
        else { len_b = b_len; D = 1 + (int)(len_b * p.R); len_a = min(a_len, len_b + D); }
        q.len_a = len_a; q.len_b = len_b; q.D = D;
        pb_align_out o;
        o.ret = -1; o.len_a = len_a; o.len_b = len_b; o.max_dst = D;
        o.matlen_a = o.matlen_b = o.cost = o.diag_cost = o.nedit = o.fail_row = 0;
        o.cells = 0;
        const bool dom = !(len_a >= p.maxn || D >= p.maxm) && D + 1 <= 32 * CPL; // seq_aligner.h:104-107 (the host picks CPL by the widest band)
        int fail_row = 0;
        if (dom) {
            WLane<CPL> L;
#pragma unroll
            for (int s = 0; s < NS; ++s) L.c[s] = (NS * lane + s == D) ? 0 : W_INF; // cell (0,0) sits on diagonal k = D
            // the state BEFORE step 1: that step (parity P1) moves iTop (P1 == 0) or jBase (P1 == 1) to its own values
            const int P1 = (1 + D) & 1;
            const int iTop1 = ((1 + D - P1) >> 1) - CPL * lane;
            L.iTop = iTop1 - (P1 == 0 ? 1 : 0);
            L.jBase = 1 - iTop1 - (P1 == 1 ? 1 : 0);
#pragma unroll
            for (int u = 0; u < CPL; ++u) {
                L.A[u] = w_elem(q.a, q.wa, L.iTop - 1 - u, len_a);
                L.B[u] = w_elem(q.b, q.wb, L.jBase - 1 + u, len_b);
            }
            L.nA = w_elem(q.a, q.wa, L.iTop, len_a);
            L.nB = w_elem(q.b, q.wb, L.jBase + CPL - 1, len_b);
            const int nsteps = len_a + len_b, nfast = min(len_a, len_b);
            const int pD = D & 1, laneD = (D >> 1) / CPL, sD = D - NS * laneD; // the diagonal's slot: updated by the steps of even d
            // early failure (seq_aligner.h:185) after a step of even d: row i = d/2 is complete up to its diagonal cell
            auto failed_at = [&](auto PC, int d) -> bool {
                constexpr int P = decltype(PC)::value;
                const int i = d >> 1;
                if (P != pD || i <= 10 || i > nfast) return false; // P == pD: d is even
                int dv = 0;
#pragma unroll
                for (int s = P; s < NS; s += 2)
                    if (s == sD) dv = L.c[s];
                dv = __shfl_sync(FULL, dv, laneD);
                if ((double)dv > i * p.R * p.fail_scale) { fail_row = i; return true; }
                return false;
            };
            {
                int d = 1;
                bool failed = false;
                const std::integral_constant<int, 0> P0c;
                const std::integral_constant<int, 1> P1c;
                if (P1 == 1) { w_step<CPL, 1>(L, q, d, lane, par); failed = failed_at(P1c, d); ++d; }
                while (!failed && d <= nsteps) {
                    w_step<CPL, 0>(L, q, d, lane, par); failed = failed_at(P0c, d); ++d;
                    if (failed || d > nsteps) break;
                    w_step<CPL, 1>(L, q, d, lane, par); failed = failed_at(P1c, d); ++d;
                }
            }
            // the last row / column into cst[] for the goal scan
            __syncwarp();
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int k = NS * lane + s;
                if (k <= 2 * D + 1) cst[k] = L.c[s];
            }
            __syncwarp();
            w_finish(p, idx, lane, cst, par, opsrev, len_a, len_b, D, a_len, CPL, WPL, fail_row, o);
        }
        if (lane == 0) p.out[idx] = o;
        __syncwarp();
    }
}

extern "C" int pb_align_weighted_batch(pb_ctx *ctx, const char *a_text, const uint8_t *a_w, const int64_t *a_off,
                                       const int32_t *a_len, const char *b_text, const uint8_t *b_w, const int64_t *b_off,
                                       const int32_t *b_len, int64_t n, double R, double fail_scale, int maxn, int maxm,
                                       pb_align_out *out, uint8_t *ops, const int64_t *ops_off)
{
    if (!ctx || n < 0 || (n && (!a_text || !a_w || !a_off || !a_len || !b_text || !b_w || !b_off || !b_len || !out)) || (ops && !ops_off))
        return pb_fail(ctx, PB_ERR_ARG, "pb_align_weighted_batch: bad argument");
    if (!n) return PB_OK;
    if (n > INT32_MAX) return pb_fail(ctx, PB_ERR_ARG, "pb_align_weighted_batch: too many pairs in one call");
    if (!(fail_scale > 0)) return pb_fail(ctx, PB_ERR_ARG, "fail_scale must be positive");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    pb_timer_reset(ctx);
    pb_timer_begin(ctx, PB_T_TOTAL);
    // extents of the two blobs, widest band, longest path
    int64_t a_hi = 0, b_hi = 0, extent = 0;
    int Dmax = 1;
    int64_t steps_max = 1;
    for (int64_t i = 0; i < n; ++i) {
        if (a_len[i] < 0 || b_len[i] < 0 || a_off[i] < 0 || b_off[i] < 0) return pb_fail(ctx, PB_ERR_ARG, "negative length or offset");
        a_hi = std::max(a_hi, a_off[i] + a_len[i]);
        b_hi = std::max(b_hi, b_off[i] + b_len[i]);
        int la, lb, D;
        pb_align_params(a_len[i], b_len[i], R, &la, &lb, &D);
        if (la < maxn && D < maxm) {
            Dmax = std::max(Dmax, D);
            steps_max = std::max<int64_t>(steps_max, (int64_t)la + lb);
        }
        if (ops) {
            if (ops_off[i] < 0) return pb_fail(ctx, PB_ERR_ARG, "negative ops offset");
            extent = std::max<int64_t>(extent, ops_off[i] + a_len[i] + b_len[i] + 1);
        }
    }
    DevBuf d_a, d_wa, d_b, d_wb, d_aoff, d_boff, d_alen, d_blen, d_out, d_ops, d_ops_off, d_queue;
    PB_TRY(d_a.alloc(ctx, (size_t)a_hi + 16)); PB_TRY(d_wa.alloc(ctx, (size_t)a_hi + 16));
    PB_TRY(d_b.alloc(ctx, (size_t)b_hi + 16)); PB_TRY(d_wb.alloc(ctx, (size_t)b_hi + 16));
    PB_TRY(d_aoff.alloc(ctx, (size_t)n * 8)); PB_TRY(d_boff.alloc(ctx, (size_t)n * 8));
    PB_TRY(d_alen.alloc(ctx, (size_t)n * 4)); PB_TRY(d_blen.alloc(ctx, (size_t)n * 4));
    PB_TRY(d_out.alloc_zero(ctx, (size_t)n * sizeof(pb_align_out)));
    PB_TRY(d_queue.alloc_zero(ctx, 16));
    pb_timer_begin(ctx, PB_T_H2D);
    PB_TRY(pb_h2d(ctx, d_a.p, a_text, (size_t)a_hi)); PB_TRY(pb_h2d(ctx, d_wa.p, a_w, (size_t)a_hi));
    PB_TRY(pb_h2d(ctx, d_b.p, b_text, (size_t)b_hi)); PB_TRY(pb_h2d(ctx, d_wb.p, b_w, (size_t)b_hi));
    PB_TRY(pb_h2d(ctx, d_aoff.p, a_off, (size_t)n * 8)); PB_TRY(pb_h2d(ctx, d_boff.p, b_off, (size_t)n * 8));
    PB_TRY(pb_h2d(ctx, d_alen.p, a_len, (size_t)n * 4)); PB_TRY(pb_h2d(ctx, d_blen.p, b_len, (size_t)n * 4));
    if (ops) {
        PB_TRY(d_ops.alloc_zero(ctx, (size_t)extent + 16));
        PB_TRY(d_ops_off.alloc(ctx, (size_t)n * 8));
        PB_TRY(pb_h2d(ctx, d_ops_off.p, ops_off, (size_t)n * 8));
    }
    pb_timer_end(ctx, PB_T_H2D);

    WLaunch p;
    memset(&p, 0, sizeof p);
    p.a = d_a.as<uint8_t>(); p.wa = d_wa.as<uint8_t>(); p.b = d_b.as<uint8_t>(); p.wb = d_wb.as<uint8_t>();
    p.a_off = d_aoff.as<int64_t>(); p.b_off = d_boff.as<int64_t>(); p.a_len = d_alen.as<int32_t>(); p.b_len = d_blen.as<int32_t>();
    p.n = (int)n; p.R = R; p.fail_scale = fail_scale; p.maxn = maxn; p.maxm = maxm;
    p.cst_words = (2 * Dmax + 3 + 4 + 3) & ~3;
    // bands of config 3's sizes keep their costs in registers (alignw_reg_kernel<CPL>), wider ones in shared memory
    int cpl = (Dmax + 1 + 31) / 32;
    const void *fn = (const void *)alignw_kernel;
    if (!getenv("PB_W_SMEM")) {
#define PICK(n) if (fn == (const void *)alignw_kernel && cpl <= n) { cpl = n; fn = (const void *)alignw_reg_kernel<n>; }
        PICK(1) PICK(2) PICK(3) PICK(5) PICK(9) PICK(17)
#undef PICK
    }
    const int wpl = (cpl + 15) / 16;
    p.par_words = (size_t)(steps_max + 1) * wpl * 32;
    p.slot_words = p.par_words + (((size_t)steps_max + 64 + 127) & ~(size_t)127) / 4;
    int wpb = WALIGN_WPB;
    while (wpb > 1 && (size_t)wpb * p.cst_words * 4 > 96 * 1024) wpb >>= 1;
    const size_t smem = (size_t)wpb * p.cst_words * 4;
    if (smem > 200 * 1024) return pb_fail(ctx, PB_ERR_DOMAIN, "band half-width %d needs %zu bytes of shared memory", Dmax, smem);
    PB_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, wpb * 32, smem));
    if (occ < 1) occ = 1;
    int64_t blocks = std::min<int64_t>((int64_t)occ * ctx->sm_count, (n + wpb - 1) / wpb);
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); fr = (size_t)8 << 30; }
    size_t budget = (size_t)((double)fr * 0.5);
    if (ctx->scratch_limit && ctx->scratch_limit < budget) budget = ctx->scratch_limit;
    const size_t per_block = (size_t)wpb * p.slot_words * 4;
    blocks = std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)(budget / per_block)));
    DevBuf d_scratch;
    PB_TRY(d_scratch.alloc(ctx, (size_t)blocks * per_block + 256));
    p.scratch = d_scratch.as<uint32_t>();
    p.queue = d_queue.as<int>();
    p.out = d_out.as<pb_align_out>();
    p.ops = ops ? d_ops.as<uint8_t>() : nullptr;
    p.ops_off = ops ? d_ops_off.as<int64_t>() : nullptr;
    pb_timer_begin(ctx, PB_T_ALIGN);
    {
        void *args[] = {(void *)&p};
        PB_CUDA(ctx, cudaLaunchKernel(fn, dim3((unsigned)blocks), dim3(wpb * 32), args, smem, ctx->stream));
    }
    PB_LAUNCH_CHECK(ctx);
    pb_timer_end(ctx, PB_T_ALIGN);
    pb_timer_begin(ctx, PB_T_D2H);
    PB_TRY(pb_d2h(ctx, out, d_out.p, (size_t)n * sizeof(pb_align_out)));
    if (ops) {
        int64_t lo = INT64_MAX;
        for (int64_t i = 0; i < n; ++i) lo = std::min(lo, ops_off[i]);
        PB_TRY(pb_d2h(ctx, ops + lo, d_ops.as<uint8_t>() + lo, (size_t)(extent - lo)));
    }
    pb_timer_end(ctx, PB_T_D2H);
    pb_timer_end(ctx, PB_T_TOTAL);
    PB_TRY(pb_sync(ctx));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return pb_fail(ctx, PB_ERR_CUDA, "weighted align kernel failed: %s", cudaGetErrorString(e));
    pb_timer_collect(ctx);
    return PB_OK;
}
