// pb_align_nb.cuh -- K3n: the NARROW, BLOCK-STATIONARY first pass of the banded aligner (included by pb_align.cu).
//
// The reference fills the whole band |i-j| <= max_dst (seq_aligner.h:158-159).  K3n computes a strip of offsets
// [-Wl, +Wr] around the main diagonal instead and proves, per alignment, that what it reports is what the full band
// gives; an alignment it cannot certify is flagged and redone by the full-band kernel (align_one).  The argument, the
// frame and the certificate are written out and tested cell for cell against the oracle in tools/narrow_model.c:
//   * every off-diagonal step costs 1, so a path of cost c to a cell of offset o stays within [-(c-o)/2, (c+o)/2];
//   * the strip only removes paths and treats what lies outside as "neighbour + 1", so computed >= true everywhere,
//     with equality for every cell whose optimal paths fit the strip;
//   * early failure (seq_aligner.h:185) compares cost(i,i) with floor(i*R) <= max_dst-1: exact once both sides of the
//     strip are >= max_dst/2;
//   * goal cell and traceback are exact when the final cost m <= the strip's goal side (right if the goal is searched
//     on the last row, left if on the last column); the other side needs m/2 <= max_dst/2, which always holds.
// Block-stationary frame: for the 32 rows of a block, frame bit c is COLUMN j = i0 - Wl + c.  Nothing slides per row
// (no state shift, no funnel shift of the Eq words, no moving edge mask, no 2-bit exchange with the next lane); between
// blocks the state moves down one whole word and the word entering on the right starts at h = +1.  Wl is a multiple of
// 32*S, so the main diagonal sits in slot 0 of one lane and row t's diagonal bit is bit t.  The frame stays strictly
// inside the reference's band (Wl + 32 <= D, NB - Wl <= D): none of the reference's edge rules is ever in play.
// Per band word and row: 8 LOP3 + 2 SHF + 2 IADD3 (was 11 + 6 + 2), and the strip is ~1.25 D wide instead of 2 D; parents
// are written for ~0.85 D of it (the stored strip, see align_one_nb).
#pragma once

#ifndef PB_NB_RING
#define PB_NB_RING 3 // traceback windows in flight in the strip pass (A/B on config 2: 3 -> 57.0 ms, 4 -> 57.9, 6 -> 61.0 of K3)
#endif

// The strip for band half-width D in band class S.  goal_left: the goal is searched on the last column (len_a > len_b).
// target: wanted width of the goal side.  0: no certified strip in this class; 1: valid, goal side limited by the class
// capacity; 2: target met (or the goal side is as wide as the reference band allows).  == tools/narrow_model.c
__host__ __device__ inline int nb_policy(int D, int S, int goal_left, int target, int *Wl_out, int *NBw_out, int *Wgoal_out)
{
    const int unit = 32 * S, cap = 1024 * S;
    const int Wh = D / 2;
    if (target < Wh) target = Wh;
    int Wl, NB, full = 0;
    if (!goal_left) {
        Wl = (Wh + unit - 1) / unit * unit;
        if (Wl + 32 > D) return 0;
        NB = (Wl + 32 + target + 31) & ~31;
        if (NB >= ((D + Wl) & ~31)) { NB = (D + Wl) & ~31; full = 1; }
        if (NB > cap) { NB = cap; full = 0; }
        if (NB - 32 - Wl < Wh) return 0;
        *Wgoal_out = NB - 32 - Wl;
    } else {
        int wl_max = target + unit - 1;
        if (wl_max >= D - 32) { wl_max = D - 32; full = 1; }
        if (wl_max > cap - 32 - Wh) { wl_max = cap - 32 - Wh; full = 0; }
        if (wl_max < unit) return 0;
        Wl = wl_max / unit * unit;
        if (Wl < Wh) return 0;
        NB = (Wl + 32 + Wh + 31) & ~31;
        if (NB > cap || NB - Wl > D) return 0;
        *Wgoal_out = Wl;
    }
    *Wl_out = Wl;
    *NBw_out = NB / 32;
    return (full || *Wgoal_out >= target) ? 2 : 1;
}
// goal-side width asked for: the cost a certified alignment may have.  num/256 of max_dst (PB_NARROW_G, default 0.80)
__host__ __device__ inline int nb_target(int D, int g256) { return (int)(((long long)D * g256) >> 8) + 1; }

// One row of the stationary frame for the S words of this lane.  pl: this lane's first Eq word of the row's plane (word
// aligned with the frame); prow: this lane's unit column of the row's parent block.  Returns the D0 word of slot 0.
template <int S>
__device__ __forceinline__ uint32_t row_step_nb(uint32_t (&Hp)[S], uint32_t (&Hn)[S], uint32_t (&Vp)[S], uint32_t (&Vn)[S],
                                                const uint32_t *__restrict__ pl, int lane, bool lane0, uint32_t *__restrict__ prow, int tail_off,
                                                bool st_on, int NL)
{
    uint32_t Eq[S], x[S], sum[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { Eq[s] = pl[s]; x[s] = Eq[s] & Hp[s]; }
    sum[0] = add_cc(x[0], Hp[0]);
#pragma unroll
    for (int s = 1; s < S; ++s) sum[s] = addc_cc(x[s], Hp[s]);
    const uint32_t carry = addc(0u, 0u);
    uint32_t ones = sum[0];
#pragma unroll
    for (int s = 1; s < S; ++s) ones &= sum[s];
    const uint32_t G = __ballot_sync(FULL, carry);
    const uint32_t P = __ballot_sync(FULL, ones == 0xffffffffu);
    const uint32_t cin = ((((G | P) + G) ^ P) >> lane) & 1u; // carry into this lane's block of S words
    sum[0] = add_cc(sum[0], cin);
#pragma unroll
    for (int s = 1; s < S; ++s) sum[s] = addc_cc(sum[s], 0u);

    uint32_t Mw[S];
    uint32_t d0w = 0u;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const uint32_t Xv = (sum[s] ^ Hp[s]) | Eq[s];
        Vp[s] = Hn[s] | ~(Xv | Hp[s]);
        Vn[s] = Hp[s] & Xv;
        Mw[s] = Eq[s] | ~(Xv | Hn[s]); // MATCH iff Eq | ~D0
        if (s == 0) d0w = Xv | Hn[s];
    }
    // the vertical deltas of the column left of this lane's block enter at bit 0 of its first word; +1 at the frame's left edge
    uint32_t pprev = __shfl_up_sync(FULL, Vp[S - 1], 1), nprev = __shfl_up_sync(FULL, Vn[S - 1], 1);
    if (lane0) { pprev = 0x80000000u; nprev = 0u; }
    uint32_t heldM = 0u, heldI = 0u;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const uint32_t vps = __funnelshift_l(pprev, Vp[s], 1), vns = __funnelshift_l(nprev, Vn[s], 1);
        pprev = Vp[s];
        nprev = Vn[s];
        const uint32_t Xh = Eq[s] | Hn[s];
        Hp[s] = vns | ~(Xh | vps);
        Hn[s] = vps & Xh;
        // parents as 16-byte units {M[2p], I[2p], M[2p+1], I[2p+1]} at unit p*NL + (lane - first stored lane); an odd S ends in
        // 8-byte pairs at tail_off.  Only the NL lanes of the stored strip write (rows are compact: no holes in a line).
        if ((s & 1) == 0 && s + 1 < S) {
            heldM = Mw[s]; heldI = Hp[s];
        } else if (s & 1) {
            if (st_on) reinterpret_cast<uint4 *>(prow)[(s >> 1) * NL] = make_uint4(heldM, heldI, Mw[s], Hp[s]);
        } else {
            if (st_on) *reinterpret_cast<uint2 *>(prow + tail_off) = make_uint2(Mw[s], Hp[s]);
        }
    }
    return d0w;
}

// seq_aligner::align over the strip.  res.redo = 1: nothing certified, the full-band kernel must run this candidate.
template <int S>
__device__ __forceinline__ void align_one_nb(const SeqView &A, int64_t a_bit, int a_len, const SeqView &B, int64_t b_bit, int b_len,
                                             double R, int maxn, int maxm, int g256, int s256, uint32_t *__restrict__ planes, int PW,
                                             uint32_t *__restrict__ par, size_t par_words, uint8_t *__restrict__ opsrev,
                                             uint8_t *__restrict__ ops_out, uint32_t *__restrict__ raw, int RW, uint64_t *bar,
                                             uint32_t &phase, AlnRes &res, int &redo, long long &band_cells)
{
    constexpr int T = 32 * S;
    const int lane = threadIdx.x & 31;
    const bool lane0 = lane == 0;
    int len_a, len_b, D;
    derive_params(a_len, b_len, R, len_a, len_b, D);
    res.ret = -1; res.len_a = len_a; res.len_b = len_b; res.D = D;
    res.matlen_a = res.matlen_b = res.cost = res.diag_cost = res.nedit = res.fail_row = 0;
    res.cells = 0;
    redo = 0;
    if (len_a >= maxn || D >= maxm) return; // seq_aligner.h:104-107
    const bool goal_left = len_a > len_b;
    int Wl = 0, NBw = 0, Wgoal = 0;
    if (!nb_policy(D, S, goal_left ? 1 : 0, nb_target(D, g256), &Wl, &NBw, &Wgoal)) { redo = 1; return; }
    // rows: past len_b + Wl the last column has left the frame (its cells cost more than Wl >= any certified minimum)
    const int rows_max = goal_left ? min(len_a, len_b + Wl) : len_a;
    // The STORED strip is narrower still.  A path of cost m ending at offset o_g stays within [-(m - o_g)/2, (m + o_g)/2], so
    // parents are needed for offsets [-Wgoal/2, +Sg] only (mirrored when the goal is on the last column): the non-goal side is
    // safe for every certified cost (m <= Wgoal), the goal side Sg = s256/256 of max_dst is checked against (m + o_g)/2 once
    // the goal is known.  Parents go out for the lanes that own those frame bits, rows compacted to them.
    const int Sh = (Wgoal + 1) / 2, Sg = max(Sh, (int)(((long long)D * s256) >> 8) + 1);
    const int lo_off = goal_left ? Sg : Sh, hi_off = goal_left ? Sh : Sg;
    int L_lo = max(0, Wl - lo_off) / T, L_hi = min(31, min((Wl + hi_off + 31) / T, (32 * NBw - 1) / T));
    if (((L_hi - L_lo + 1) & 1) && (S & 1)) { // rows of 16-byte units: an odd S needs an even lane count to stay 16-byte aligned
        if (L_hi < 31) ++L_hi; else --L_lo;
    }
    const int NL = L_hi - L_lo + 1;
    const int lo_eff = Wl - L_lo * T, hi_eff = (L_hi + 1) * T - 1 - Wl - 31; // offsets stored in EVERY row of a block
#ifdef PB_EXP_NO_ST // timing experiment only: no parent stores (results are wrong)
    const bool st_on = false;
#else
    const bool st_on = lane >= L_lo && lane <= L_hi;
#endif
    const size_t rstride = (size_t)2 * S * NL; // words per parent row
    if ((size_t)rows_max * rstride > par_words) { redo = 1; return; } // the slot was sized for the item's own length

    // ---- Eq planes of seg_b in shared memory: plane c, bit t <-> (b[t - Wl] == c), zero outside [0,len_b)
    const int PWn = min(PW, ((rows_max + 31) >> 5) + T + 1);
    const int64_t g0 = b_bit - Wl;
    const int64_t w_first = max((int64_t)0, g0 >> 5) & ~(int64_t)3;
    const int64_t w_end = min(B.nwords, (((b_bit + len_b + 31) >> 5) + 1 + 3) & ~(int64_t)3);
    const int k0 = (int)((g0 >> 5) - w_first);
    for (int64_t cw = w_first; cw < w_end; cw += PB_STAGE_WORDS - 4) {
        const int c_lo = (int)(cw - w_first);
        if (max(0, c_lo - k0) >= PWn && cw != w_first) break; // the rest of seg_b lies right of every frame
        const int n_raw = (int)min((int64_t)PB_STAGE_WORDS, w_end - cw);
        __syncwarp();
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t bytes = (uint32_t)n_raw * 4u;
            mbar_expect_tx(bar, bytes * 2u);
            tma_load_1d(raw, B.hi + cw, bytes, bar);
            tma_load_1d(raw + RW, B.lo + cw, bytes, bar);
        }
        int spins = 0;
        while (!mbar_try_wait(bar, phase)) {
            if (++spins > (1 << 22)) __trap();
        }
        phase ^= 1u;
        const int c_hi = (cw + PB_STAGE_WORDS - 4 >= w_end) ? INT_MAX : c_lo + PB_STAGE_WORDS - 4;
        const int x_lo = (cw == w_first) ? 0 : max(0, c_lo - k0);
        const int x_hi = (c_hi == INT_MAX) ? PWn : min(PWn, max(0, c_hi - k0));
        for (int x = x_lo + lane; x < x_hi; x += 32) {
            const int bidx0 = 32 * x - Wl;
            uint32_t valid;
            if (bidx0 >= len_b || bidx0 + 31 < 0) valid = 0u;
            else {
                valid = 0xffffffffu;
                if (bidx0 < 0) valid &= 0xffffffffu << (-bidx0);
                if (bidx0 + 31 >= len_b) valid &= 0xffffffffu >> (bidx0 + 32 - len_b);
            }
            uint32_t hi = 0u, lo = 0u;
            if (valid) {
                const int wi = x + k0 - c_lo;
                const unsigned sh = (unsigned)(g0 & 31);
                auto win = [&](const uint32_t *pl) -> uint32_t {
                    const uint32_t w0 = (wi >= 0 && wi < n_raw) ? pl[wi] : 0u;
                    const uint32_t w1 = (wi + 1 >= 0 && wi + 1 < n_raw) ? pl[wi + 1] : 0u;
                    return __funnelshift_r(w0, w1, sh);
                };
                hi = win(raw);
                lo = win(raw + RW);
            }
            planes[0 * PW + x] = ~hi & ~lo & valid;
            planes[1 * PW + x] = ~hi & lo & valid;
            planes[2 * PW + x] = hi & ~lo & valid;
            planes[3 * PW + x] = hi & lo & valid;
        }
    }
    __syncwarp();

    // ---- row 0 in the frame of block 0: columns j <= 0 (c < Wl) are the fake cells cost(i,j) = i + |j|, h = -1
    const int Ld = Wl / T; // lane that owns the main diagonal (slot 0, bit t in row t of a block)
    uint32_t Hp[S], Hn[S], Vp[S], Vn[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { Hn[s] = lane < Ld ? 0xffffffffu : 0u; Hp[s] = ~Hn[s]; }
    const int wt = NBw - 1, Lt = wt / S, st = wt - Lt * S; // the frame's last valid word: it starts every block at h = +1
    const int lane_off = 4 * (lane - L_lo), tail_off = (S / 2) * NL * 4 - 2 * (lane - L_lo);

    int cii = 0;
    int colc = 0, colbest = 0, col_i = 0;
    int fail_row = 0;
    const int nfast = min(len_a, len_b);
    for (int i0 = 1; i0 <= rows_max; i0 += 32) {
        const int q = (i0 - 1) >> 5;
        if (q > 0) { // the frame moves right by one word
            const uint32_t hp0 = __shfl_down_sync(FULL, Hp[0], 1), hn0 = __shfl_down_sync(FULL, Hn[0], 1);
#pragma unroll
            for (int s = 0; s + 1 < S; ++s) { Hp[s] = Hp[s + 1]; Hn[s] = Hn[s + 1]; }
            Hp[S - 1] = hp0; Hn[S - 1] = hn0;
#pragma unroll
            for (int s = 0; s < S; ++s)
                if (lane == Lt && s == st) { Hp[s] = 0xffffffffu; Hn[s] = 0u; }
        }
        // this block's 32 elements of seg_a: lane t keeps the plane offset of row t
        const uint32_t awh = load_window(A.hi, A.nwords, a_bit + i0 - 1), awl = load_window(A.lo, A.nwords, a_bit + i0 - 1);
        const int my_off = (int)(((awh >> lane) & 1u) * 2u + ((awl >> lane) & 1u)) * PW;
        const uint32_t *plq = planes + q + lane * S;
        uint32_t *prow = par + (size_t)(i0 - 1) * rstride + lane_off;
        PB_CHECK_RANGE("parent rows of a block", par + (size_t)(i0 - 1) * rstride, 16, par, par + par_words);
        PB_CHECK_RANGE("parent rows of a block (end)", par + (size_t)(i0 - 1 + min(32, rows_max - i0 + 1)) * rstride - 4, 4, par, par + par_words);
        const int tfast = max(0, min(32, nfast - i0 + 1)); // rows of this block with an early-failure test
        const int tall = min(32, rows_max - i0 + 1);
        if (tfast > 0) {
            uint32_t hist = 0u, tb = 1u;
            for (int t = 0; t < tfast; ++t) {
                const int off = __shfl_sync(FULL, my_off, t);
                const uint32_t d0w = row_step_nb<S>(Hp, Hn, Vp, Vn, plq + off, lane, lane0, prow, tail_off, st_on, NL);
                hist |= d0w & tb; // row t's diagonal D0 bit is bit t of slot 0 in the diagonal's owner lane
                tb <<= 1;
                prow += rstride;
            }
            hist = __shfl_sync(FULL, hist, Ld);
            const int thr = (int)((i0 + lane) * R); // cost > i*R  <=>  cost > floor(i*R)
            const int cdiag = cii + (lane + 1) - __popc(hist & (0xffffffffu >> (31 - lane)));
            const uint32_t badm = __ballot_sync(FULL, lane < tfast && i0 + lane > 10 && cdiag > thr);
            if (badm) { fail_row = i0 + __ffs(badm) - 1; break; }
            cii += tfast - __popc(hist & (0xffffffffu >> (32 - tfast)));
        }
        if (tall > tfast) { // rows below seg_b's end (len_a > len_b): follow cost(i, len_b) down the last column (Q-D2: no test)
            if (i0 + tfast - 1 == len_b || (tfast == 0 && i0 - 1 == len_b)) { colc = colbest = cii; col_i = len_b; }
            for (int t = tfast; t < tall; ++t) {
                const int i = i0 + t;
                const int off = __shfl_sync(FULL, my_off, t);
                row_step_nb<S>(Hp, Hn, Vp, Vn, plq + off, lane, lane0, prow, tail_off, st_on, NL);
                prow += rstride;
                const int c = len_b - i + Wl + t, wk = c >> 5, Lk = wk / S, sk = wk - Lk * S;
                uint32_t vpw = 0u, vnw = 0u;
#pragma unroll
                for (int s = 0; s < S; ++s)
                    if (s == sk) { vpw = Vp[s]; vnw = Vn[s]; }
                vpw = __shfl_sync(FULL, vpw, Lk);
                vnw = __shfl_sync(FULL, vnw, Lk);
                colc += (int)((vpw >> (c & 31)) & 1u) - (int)((vnw >> (c & 31)) & 1u);
                if (colc < colbest) { colbest = colc; col_i = i; }
            }
        }
    }
    if (fail_row) { // certain: both sides of the strip are >= D/2 >= floor(i*R)/2
        res.fail_row = fail_row;
        res.cells = cells_upto(fail_row, D, len_b);
        band_cells += (long long)fail_row * 32 * NBw;
        return;
    }
    band_cells += (long long)rows_max * 32 * NBw;
    res.cells = cells_upto(len_a, D, len_b);

    // ---- goal_cell, seq_aligner.h:191-213
    __syncwarp();
#pragma unroll
    for (int s = 0; s < S; ++s) {
        planes[lane * S + s] = Hp[s];
        planes[T + lane * S + s] = Hn[s];
    }
    __syncwarp();
    int matlen_a, matlen_b, cost;
    if (goal_left) {
        matlen_a = col_i; matlen_b = len_b; cost = colbest;
    } else {
        // last row: cost(len_a, j), j in (len_a, len_b], from the final horizontal deltas; earliest strict minimum.
        // Frame bit of column j in the last row: c = j - len_a + Wl + t_last.  Each lane walks one frame word, an
        // exclusive warp scan of the word totals gives its starting cost; columns right of the frame cost more than Wgoal.
        const int t_last = (len_a - 1) & 31;
        const int c_first = 1 + Wl + t_last, c_last = min(len_b - len_a + Wl + t_last, 32 * NBw - 1);
        matlen_a = len_a;
        int best = cii, bestj = len_a, base = cii;
        for (int w0 = c_first >> 5; 32 * w0 <= c_last; w0 += 32) {
            const int w = w0 + lane;
            const int lo = max(c_first, 32 * w), hi = min(c_last, 32 * w + 31);
            int tot = 0, lmin = INT_MAX, lpos = 0;
            if (lo <= hi) {
                const uint32_t hp = planes[w], hn = planes[T + w];
                for (int c = lo; c <= hi; ++c) {
                    tot += (int)((hp >> (c & 31)) & 1u) - (int)((hn >> (c & 31)) & 1u);
                    if (tot < lmin) { lmin = tot; lpos = c; }
                }
            }
            int pre = tot; // inclusive scan over lanes
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(FULL, pre, d);
                if (lane >= d) pre += v;
            }
            const int total = __shfl_sync(FULL, pre, 31);
            int cand = lo <= hi ? base + (pre - tot) + lmin : INT_MAX;
            int candj = lpos - Wl - t_last + len_a;
            // lexicographic minimum of (cost, j) over the lanes
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const int oc = __shfl_xor_sync(FULL, cand, d), oj = __shfl_xor_sync(FULL, candj, d);
                if (oc < cand || (oc == cand && oj < candj)) { cand = oc; candj = oj; }
            }
            if (cand < best) { best = cand; bestj = candj; }
            base += total;
        }
        cost = best; matlen_b = bestj;
    }
    if (cost > Wgoal) { redo = 1; return; } // not certified: the full band decides
    {
        const int og = matlen_b - matlen_a; // the goal's offset; the path stays within [-(m - og)/2, (m + og)/2]
        if ((cost - og) / 2 > lo_eff || (cost + og) / 2 > hi_eff) { redo = 1; return; } // it may leave the stored strip
    }
    res.matlen_a = matlen_a; res.matlen_b = matlen_b; res.cost = cost;
    res.diag_cost = (a_len <= len_a && a_len <= len_b) ? cii : 0; // locator.cpp:86 (Q-L2)
    if ((double)matlen_b < len_b * (1 - R)) return; // seq_aligner.h:114
#ifdef PB_EXP_NO_TB // timing experiment only: no traceback (results are wrong)
    res.ret = matlen_b;
    return;
#endif

    // ---- find_path, seq_aligner.h:214-233.  Cell (i,j) sits at frame bit c = j - i + Wl + ((i-1)&31) of row i.  Lane r of a
    // window holds row i0w - r: the unit under the bit its row would have if the path kept its diagonal, plus the neighbouring
    // unit when that bit is within 8 of the unit's edge.  Windows are prefetched PB_NB_RING deep into shared memory (cp.async.cg).
    __syncwarp();
    uint32_t *ring = planes + 2 * T;
    int *meta = reinterpret_cast<int *>(ring + PB_NB_RING * 256);
    auto unit_of = [&](int w, int &b, int &n) {
        const int s = w % S;
        if (s < (S & ~1)) { b = w - (s & 1); n = 2; } else { b = w; n = 1; }
    };
    auto par_addr = [&](int row, int w) -> const uint2 * { // w inside the stored lanes (lane_units sees to that)
        const int L = w / S, s = w - L * S, Ls = L - L_lo;
        const uint32_t *rb = par + (size_t)(row - 1) * rstride;
        if (s < (S & ~1)) return reinterpret_cast<const uint2 *>(rb + ((s >> 1) * NL + Ls) * 4 + (s & 1) * 2);
        return reinterpret_cast<const uint2 *>(rb + (S / 2) * NL * 4 + Ls * 2);
    };
    const int w_lo = L_lo * S, w_hi = (L_hi + 1) * S - 1; // band words whose parents exist
    // this lane's units for a window whose first row is i0w, predicted diagonal kd = j - i + Wl
    auto lane_units = [&](int i0w, int kd, int &pb, int &pn, int &sb, int &sn) {
        const int row = i0w - lane;
        pb = pn = sb = sn = 0;
        if (row < 1) return;
        const int c = kd + ((row - 1) & 31);
        if (c < 32 * w_lo || c >= 32 * (w_hi + 1)) return;
        unit_of(c >> 5, pb, pn);
        const int pos = c - 32 * pb;
        if (pos < 8 && pb > w_lo) unit_of(pb - 1, sb, sn);
        else if (pos >= 32 * pn - 8 && pb + pn <= w_hi) unit_of(pb + pn, sb, sn);
    };
    auto fetch = [&](int slot, int i0w, int kd) {
        int pb, pn, sb, sn;
        lane_units(i0w, kd, pb, pn, sb, sn);
        const int row = i0w - lane;
        uint32_t *dst = ring + slot * 256 + 4 * lane;
        PB_CHECK_RANGE("ring slot", dst, 16, ring, ring + PB_NB_RING * 256);
        if (pn) PB_CHECK_RANGE("traceback prefetch", par_addr(row, pb), pn == 2 ? 16 : 8, par, par + par_words);
        if (sn) PB_CHECK_RANGE("traceback prefetch (2nd unit)", par_addr(row, sb), sn == 2 ? 16 : 8, par, par + par_words);
        // rows above the matrix / bits outside the frame: plain zero stores, never zero-fill copies (see pb_align.cu)
        if (!pn) *reinterpret_cast<uint4 *>(dst) = make_uint4(0u, 0u, 0u, 0u);
        else if (pn == 2) cp_async16(dst, par_addr(row, pb), 16);
        else cp_async8(dst, par_addr(row, pb), 8);
        if (sn == 2) cp_async16(dst + 128, par_addr(row, sb), 16);
        else if (sn == 1) cp_async8(dst + 128, par_addr(row, sb), 8);
        if (lane == 0) { meta[2 * slot] = i0w; meta[2 * slot + 1] = kd; }
        cp_async_commit();
    };
    int n = 0;
    {
        int i = matlen_a, j = matlen_b;
        const int guard = len_a + len_b + 1;
        int cur_slot = 0, cur_i0 = 0, wend = 0, cb0 = 0, cn0 = 0, cb1 = 0, cn1 = 0, tr = 0;
        uint4 U0 = make_uint4(0u, 0u, 0u, 0u), U1 = U0;
        bool have = false, cold = false;
        while (i > 0 && j > 0 && n < guard) {
            const int k = j - i + Wl;
            // is the current cell's word among the units lane r0 holds?
            bool need = !have || i <= wend;
            if (!need) {
                const int r0 = cur_i0 - i;
                const int w = (k + ((i - 1) & 31)) >> 5;
                const int b0 = __shfl_sync(FULL, cb0, r0), n0 = __shfl_sync(FULL, cn0, r0);
                const int b1 = __shfl_sync(FULL, cb1, r0), n1 = __shfl_sync(FULL, cn1, r0);
                need = !((unsigned)(w - b0) < (unsigned)n0 || (unsigned)(w - b1) < (unsigned)n1);
            }
            if (need) {
                const int nslot = cur_slot + 1 == PB_NB_RING ? 0 : cur_slot + 1;
                bool usual = have && i <= wend;
                if (usual) {
                    const int *m = meta + 2 * nslot;
                    int pb, pn, sb, sn;
                    usual = m[0] == i;
                    if (usual) { // does the prefetched window hold the current cell (its lane 0)?
                        const int kd = m[1], c = kd + ((i - 1) & 31), w = (k + ((i - 1) & 31)) >> 5;
                        pb = pn = sb = sn = 0;
                        if (c >= 32 * w_lo && c < 32 * (w_hi + 1)) {
                            unit_of(c >> 5, pb, pn);
                            const int pos = c - 32 * pb;
                            if (pos < 8 && pb > w_lo) unit_of(pb - 1, sb, sn);
                            else if (pos >= 32 * pn - 8 && pb + pn <= w_hi) unit_of(pb + pn, sb, sn);
                        }
                        usual = (unsigned)(w - pb) < (unsigned)pn || (unsigned)(w - sb) < (unsigned)sn;
                    }
                }
                __syncwarp();
                if (usual) { // refill the slot just walked with the window PB_NB_RING-1 ahead, on the current diagonal
                    fetch(cur_slot, i - 32 * (PB_NB_RING - 1), k);
                    cur_slot = nslot;
                } else { // cold start, or the path left the predicted units
                    cp_async_wait<0>();
                    __syncwarp();
#pragma unroll
                    for (int t = 0; t < PB_NB_RING; ++t) fetch(t, i - 32 * t, k);
                    cur_slot = 0;
                    cold = true;
                }
                cp_async_wait<PB_NB_RING - 1>();
                __syncwarp();
                cur_i0 = meta[2 * cur_slot];
                lane_units(cur_i0, meta[2 * cur_slot + 1], cb0, cn0, cb1, cn1);
                U0 = *reinterpret_cast<const uint4 *>(ring + cur_slot * 256 + 4 * lane);
                U1 = cn1 ? *reinterpret_cast<const uint4 *>(ring + cur_slot * 256 + 128 + 4 * lane) : make_uint4(0u, 0u, 0u, 0u);
                tr = (cur_i0 - lane - 1) & 31; // this lane's row position inside its block
                wend = cur_i0 - 32;
                have = true;
            }
            const int r0 = cur_i0 - i;
            // every lane looks at the bit its own row has on the current diagonal
            const int c = k + tr, w = c >> 5, kb = c & 31;
            const int d0 = w - cb0, d1 = w - cb1;
            const bool prim = (unsigned)d0 < (unsigned)cn0, sec = (unsigned)d1 < (unsigned)cn1;
            const int d = prim ? d0 : d1;
            const uint4 u = prim ? U0 : U1;
            const uint32_t mbit = (prim | sec) ? (((d ? u.z : u.x) >> kb) & 1u) : 0u;
            const uint32_t ibit = ((d ? u.w : u.y) >> kb) & 1u;
            const uint32_t Bm = __ballot_sync(FULL, mbit) >> r0;
            const uint32_t Av = __ballot_sync(FULL, prim | sec) >> r0; // lanes that hold their bit at all
            int run = (~Bm) ? __ffs(~Bm) - 1 : 32;
            const int lim = min(32 - r0, min(i, j));
            const bool stop_known = run < lim && ((Av >> run) & 1u); // the run ends on a cell this window holds
            run = min(run, lim);
            if (lane < run) opsrev[n + lane] = (uint8_t)PB_MATCH;
            n += run; i -= run; j -= run;
            // a fresh window that does not even hold its own first cell: the path left the frame (never for a certified goal)
            if (cold && run == 0 && !stop_known) { redo = 1; break; }
            cold = false;
            if (stop_known) {
                const uint32_t hb = __shfl_sync(FULL, ibit, r0 + run);
                if (lane == 0) opsrev[n] = (uint8_t)(hb ? PB_INSERT : PB_DELETE);
                ++n;
                if (hb) --j; else --i;
            }
        }
        cp_async_wait<0>();
        __syncwarp();
        if (redo) return;
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
