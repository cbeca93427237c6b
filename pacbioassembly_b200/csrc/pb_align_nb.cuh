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
// Per band word and row: 7 LOP3 + 2 SHF + 2 IADD3 (the full-band kernel: 11 + 6 + 2), and the strip is ~1.25 D wide instead of 2 D.
//
// No parents are stored by the forward pass.  Per 32-row block every lane writes a CHECKPOINT: its horizontal deltas at the top
// of the block (2 S words) and three 32-bit columns collected over the block's rows -- the vertical delta entering the lane's
// first word (+ bit, - bit) and the carry entering its multi-word add.  With those ONE THREAD reproduces the lane's
// 32 rows x S words of parents by itself (no shuffle, no ballot: everything that crossed the lane boundary is in the
// checkpoint), so the traceback recomputes just the (block, lane) tiles its path runs through: 32 tiles per round, one per
// thread, written to a small per-warp buffer that stays in L2 and is walked with the window ring.  Against stored parents
// (64 S words per lane and block) the forward pass writes 2 S + 3 -- DRAM writes drop ~20x, the aligner's scratch from tens of
// GB to a few -- and the traceback no longer reads scattered 16-byte units out of rows written long ago.  tools/narrow_model.c
// asserts that every tile a path touches is reproduced bit for bit.
#pragma once

#ifndef PB_NB_UNROLL
#define PB_NB_UNROLL 4 // rows of a full block per trip of the row loop (A/B on config 2: 2 -> 45.3 ms of K3, 4 -> 45.0, 8 -> 45.5, 16 -> 50.0, 32 -> 49.0: the instruction cache)
#endif
constexpr int NB_UNROLL = PB_NB_UNROLL;
#ifndef PB_NB_RING
#define PB_NB_RING 3 // traceback windows in flight in the strip pass (A/B on config 2: 3 -> 57.0 ms, 4 -> 57.9, 6 -> 61.0 of K3)
#endif
// per-warp shared memory behind the Eq planes and the staging buffer: the final deltas for the goal scan (2 x 32 S words), later
// the traceback's window ring (32 lanes x two {MATCH, INSERT} pairs + one info word per window), two meta words per window and
// the round's lane table
__host__ __device__ constexpr int nb_tb_words(int S)
{
    return ((2 * 32 * S > PB_NB_RING * 160 + 2 * PB_NB_RING + 16 ? 2 * 32 * S : PB_NB_RING * 160 + 2 * PB_NB_RING + 16) + 3) & ~3;
}
// scratch words of a warp slot: the tile buffer of a traceback round (32 tiles x 32 rows x S pairs) + one checkpoint per block
__host__ __device__ constexpr size_t nb_tile_words(int S) { return (size_t)32 * 32 * S * 2; }
__host__ __device__ constexpr size_t nb_ck_words(int S, int rows) { return (size_t)((rows + 31) >> 5) * (2 * S + 3) * 32; }

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
// aligned with the frame).  vinp / vinn / cina collect what enters the lane in this row (see the checkpoint above): the vertical
// deltas and the carry shift in from the right (row t of an n-row block ends at bit n-1-t).
// Returns the D0 word of slot 0.
template <int S>
__device__ __forceinline__ uint32_t row_step_nb(uint32_t (&Hp)[S], uint32_t (&Hn)[S], uint32_t (&Vp)[S], uint32_t (&Vn)[S],
                                                const uint32_t *__restrict__ pl, int lane, bool lane0, uint32_t &vinp, uint32_t &vinn,
                                                uint32_t &cina)
{
    uint32_t Eq[S], x[S], sum[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { Eq[s] = pl[s]; x[s] = Eq[s] & Hp[s]; }
    const uint32_t carry = CarryChain<S>::add(sum, x, Hp);
    uint32_t ones = sum[0];
#pragma unroll
    for (int s = 1; s < S; ++s) ones &= sum[s];
    const uint32_t G = __ballot_sync(FULL, carry);
    const uint32_t P = __ballot_sync(FULL, ones == 0xffffffffu);
    const uint32_t cin = ((((G | P) + G) ^ P) >> lane) & 1u; // the carry into this lane's block of S words
    cina = mad_lo(cina, 2u, cin); // an IMAD: the FMA pipe is idle, the integer pipe is the bound
    CarryChain<S>::inc(sum, cin);

    uint32_t d0w = 0u;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const uint32_t Xv = (sum[s] ^ Hp[s]) | Eq[s];
        Vp[s] = Hn[s] | ~(Xv | Hp[s]);
        Vn[s] = Hp[s] & Xv;
        if (s == 0) d0w = Xv | Hn[s];
    }
    // the vertical deltas of the column left of this lane's block enter at bit 0 of its first word; +1 at the frame's left edge
    uint32_t pprev = __shfl_up_sync(FULL, Vp[S - 1], 1), nprev = __shfl_up_sync(FULL, Vn[S - 1], 1);
    if (lane0) { pprev = 0x80000000u; nprev = 0u; }
    vinp = __funnelshift_l(pprev, vinp, 1);
    vinn = __funnelshift_l(nprev, vinn, 1);
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const uint32_t vps = __funnelshift_l(pprev, Vp[s], 1), vns = __funnelshift_l(nprev, Vn[s], 1);
        pprev = Vp[s];
        nprev = Vn[s];
        const uint32_t Xh = Eq[s] | Hn[s];
        Hp[s] = vns | ~(Xh | vps);
        Hn[s] = vps & Xh;
    }
    return d0w;
}

// The same row for ONE lane on its own, from what its checkpoint recorded: cb = the carry entering the lane's add, the top bits
// of pprev / nprev = the vertical delta entering its first word.  Writes the row's parents, {MATCH word, INSERT word} per band
// word (MATCH iff Eq | ~D0, INSERT iff the new h = +1), to out[s].
template <int S>
__device__ __forceinline__ void row_step_tile(uint32_t (&Hp)[S], uint32_t (&Hn)[S], const uint32_t *__restrict__ pl, uint32_t cb,
                                              uint32_t pprev, uint32_t nprev, uint2 *__restrict__ out)
{
    uint32_t Eq[S], x[S], sum[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { Eq[s] = pl[s]; x[s] = Eq[s] & Hp[s]; }
    CarryChain<S>::add_nc(sum, x, Hp);
    CarryChain<S>::inc(sum, cb);
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const uint32_t Xv = (sum[s] ^ Hp[s]) | Eq[s];
        const uint32_t vp = Hn[s] | ~(Xv | Hp[s]), vn = Hp[s] & Xv;
        const uint32_t Mw = Eq[s] | ~(Xv | Hn[s]);
        const uint32_t vps = __funnelshift_l(pprev, vp, 1), vns = __funnelshift_l(nprev, vn, 1);
        pprev = vp;
        nprev = vn;
        const uint32_t Xh = Eq[s] | Hn[s];
        Hp[s] = vns | ~(Xh | vps);
        Hn[s] = vps & Xh;
        out[s] = make_uint2(Mw, Hp[s]);
    }
}

// seq_aligner::align over the strip.  res.redo = 1: nothing certified, the full-band kernel must run this candidate.
// scr: this warp's scratch slot -- the tile buffer of a traceback round, then the checkpoints; opsrev: ent_cap words for the
// path's indel entries; tb: nb_tb_words(S) words of
// shared memory behind the planes and the staging buffer.
template <int S>
__device__ __forceinline__ void align_one_nb(const SeqView &A, int64_t a_bit, int a_len, const SeqView &B, int64_t b_bit, int b_len,
                                             double R, int maxn, int maxm, int g256, uint32_t *__restrict__ planes, int PW,
                                             uint32_t *__restrict__ scr, size_t scr_words, uint8_t *__restrict__ opsrev, int ent_cap,
                                             uint8_t *__restrict__ ops_out, uint32_t *__restrict__ raw, int RW, uint32_t *__restrict__ tb,
                                             uint64_t *bar, uint32_t &phase, AlnRes &res, int &redo, long long &band_cells, int (&tbc)[2])
{
    constexpr int T = 32 * S;
    constexpr int F = 2 * S + 3; // checkpoint fields per lane and block
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
    if (nb_tile_words(S) + nb_ck_words(S, rows_max) > scr_words) { redo = 1; return; } // the slot was sized for the item's own length
    uint32_t *const mb = scr;                    // tile buffer: [row of the block][band word of the lane][tile] {MATCH, INSERT}
    uint32_t *const ck = scr + nb_tile_words(S); // checkpoints: [block][field][lane]

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
        // checkpoint, part 1: the deltas the block's first row starts from
        uint32_t *ckb = ck + (size_t)q * (F * 32) + lane;
        PB_CHECK_RANGE("checkpoint of a block", ckb + (F - 1) * 32, 4, scr, scr + scr_words);
#pragma unroll
        for (int s = 0; s < S; ++s) { ckb[s * 32] = Hp[s]; ckb[(S + s) * 32] = Hn[s]; }
        uint32_t vinp = 0u, vinn = 0u, cina = 0u;
        // this block's 32 elements of seg_a: lane t keeps the plane offset of row t
        const uint32_t awh = load_window(A.hi, A.nwords, a_bit + i0 - 1), awl = load_window(A.lo, A.nwords, a_bit + i0 - 1);
        const int my_off = (int)(((awh >> lane) & 1u) * 2u + ((awl >> lane) & 1u)) * PW;
        const uint32_t *plq = planes + q + lane * S;
        const int tfast = max(0, min(32, nfast - i0 + 1)); // rows of this block with an early-failure test
        const int tall = min(32, rows_max - i0 + 1);
        if (tfast > 0) {
            uint32_t hist = 0u;
            if (tfast == 32) { // a full block: unrolled, the diagonal bit's mask is an immediate
#pragma unroll NB_UNROLL
                for (int t = 0; t < 32; ++t) {
                    const int off = __shfl_sync(FULL, my_off, t);
                    const uint32_t d0w = row_step_nb<S>(Hp, Hn, Vp, Vn, plq + off, lane, lane0, vinp, vinn, cina);
                    hist |= d0w & (1u << t); // row t's diagonal D0 bit is bit t of slot 0 in the diagonal's owner lane
                }
            } else {
                uint32_t tb1 = 1u;
                for (int t = 0; t < tfast; ++t) {
                    const int off = __shfl_sync(FULL, my_off, t);
                    const uint32_t d0w = row_step_nb<S>(Hp, Hn, Vp, Vn, plq + off, lane, lane0, vinp, vinn, cina);
                    hist |= d0w & tb1;
                    tb1 <<= 1;
                }
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
                row_step_nb<S>(Hp, Hn, Vp, Vn, plq + off, lane, lane0, vinp, vinn, cina);
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
        // checkpoint, part 2: what entered this lane in the block's rows
        ckb[(2 * S) * 32] = vinp; ckb[(2 * S + 1) * 32] = vinn; ckb[(2 * S + 2) * 32] = cina;
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
        tb[lane * S + s] = Hp[s];
        tb[T + lane * S + s] = Hn[s];
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
                const uint32_t hp = tb[w], hn = tb[T + w];
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
    res.matlen_a = matlen_a; res.matlen_b = matlen_b; res.cost = cost;
    res.diag_cost = (a_len <= len_a && a_len <= len_b) ? cii : 0; // locator.cpp:86 (Q-L2)
    if ((double)matlen_b < len_b * (1 - R)) return; // seq_aligner.h:114
#ifdef PB_EXP_NO_TB // timing experiment only: no traceback (results are wrong)
    res.ret = matlen_b;
    return;
#endif

    // ---- find_path, seq_aligner.h:214-233.  Cell (i,j) sits at frame bit c = j - i + Wl + ((i-1)&31) of row i, band word c >> 5,
    // lane (c >> 5) / S.  A ROUND recomputes the parents of NBK blocks going up from the current row, NLK adjacent lanes per
    // block around the path's predicted diagonal (its diagonal now, plus the drift per row seen so far): thread x takes
    // block r_qh - x / NLK, lane rtab[x / NLK] + x % NLK.  The walk goes through 32-row windows: lane r of a window holds row
    // i0w - r, the {MATCH, INSERT} pairs of the two band words around the bit its row has on the predicted diagonal (16 bits of
    // slack on either side); one ballot finds the run of MATCH steps, one shuffle the indel that ends it.  Windows are prefetched
    // PB_NB_RING deep (cp.async) from the tile buffer.  A cell outside the round (above its blocks, or in a lane it did not
    // recompute) starts a new round there.  The transcript is written at the end: MATCH everywhere, then the indels, which the
    // walk recorded as (position, op) entries.
    __syncwarp();
    constexpr int NLK = S == 1 ? 3 : 2, NBK = 32 / NLK; // A/B on config 2 (S >= 2): 2 lanes 47.1 ms, 3 lanes 48.0, 4 lanes 49.9 of K3
    uint32_t *ring = tb;                                           // PB_NB_RING windows x 32 lanes x {M, I} of two band words
    int *winfo = reinterpret_cast<int *>(ring + PB_NB_RING * 128); // per window and lane: first of its two words, which of them it has
    int *meta = winfo + PB_NB_RING * 32;                           // per window: first row, predicted diagonal
    int *rtab = meta + 2 * PB_NB_RING;                             // per block of the round: first recomputed lane
    uint2 *const mb2 = reinterpret_cast<uint2 *>(mb);              // tile buffer: [row of the block][tile][band word of the lane]
    const int L_last = (NBw - 1) / S; // last lane that holds valid frame words
    int r_qh = 0, r_nb = 0, r_sl = 0; // the current round: top block, blocks; drift of the diagonal in 1/256 bit per row
    int r_ai = matlen_a, r_ak = matlen_b - matlen_a + Wl; // where the drift is measured from
    bool r_have_sl = false;
    auto tile_of = [&](int row, int w) -> int { // the round's tile that holds band word w of `row`; -1: none
        if (row < 1 || (unsigned)w >= (unsigned)NBw) return -1;
        const int bi = r_qh - ((row - 1) >> 5);
        if (bi >= r_nb) return -1;
        const int dl = w / S - rtab[bi];
        return (unsigned)dl < (unsigned)NLK ? bi * NLK + dl : -1;
    };
    auto new_round = [&](int i, int k) {
        ++tbc[0];
        r_qh = (i - 1) >> 5;
        r_nb = min(r_qh + 1, NBK);
        // drift of the path's diagonal per row, from the stretch walked since the round before (not from the goal: the reference's
        // goal cell lies at j >= len_a, so a read that spans fewer reference bases than its length begins its way back with one
        // long run of INSERTs that says nothing about the rest)
        if (r_ai - i >= 96) {
            const int sl = max(-128, min(128, ((k - r_ak) * 256) / (r_ai - i)));
            r_sl = r_have_sl ? (r_sl + sl) / 2 : sl;
            r_have_sl = true;
            r_ai = i; r_ak = k;
        } else if (!r_have_sl) { r_ai = i; r_ak = k; } // still inside the run at the goal: anchor where it ends
        const int bi = lane / NLK, dl = lane - bi * NLK;
        const bool act = bi < r_nb;
        const int qq = r_qh - bi;
        // NLK lanes centred on the frame bit predicted for the middle of the block, rounded to a lane
        const int cmid = k + ((r_sl * (i - (32 * qq + 16))) >> 8) + 16;
        const int num = cmid - (NLK - 1) * (T / 2);
        const int La = max(0, min(num <= 0 ? 0 : num / T, L_last - NLK + 1));
        if (act && dl == 0) rtab[bi] = La;
        const int L = La + dl;
        uint32_t hp[S], hn[S], vp = 0u, vn = 0u, cc = 0u, awh = 0u, awl = 0u;
        int nr = 0;
#pragma unroll
        for (int s = 0; s < S; ++s) hp[s] = hn[s] = 0u;
        if (act) {
            const uint32_t *ckb = ck + (size_t)qq * (F * 32) + L;
            PB_CHECK_RANGE("checkpoint read", ckb + (F - 1) * 32, 4, scr, scr + scr_words);
#pragma unroll
            for (int s = 0; s < S; ++s) { hp[s] = __ldcg(ckb + s * 32); hn[s] = __ldcg(ckb + (S + s) * 32); }
            nr = min(32, rows_max - 32 * qq);
            // row t's entering deltas on top (<< t brings them to bit 31), its carry at bit t
            vp = __ldcg(ckb + (2 * S) * 32) << (32 - nr);
            vn = __ldcg(ckb + (2 * S + 1) * 32) << (32 - nr);
            cc = __brev(__ldcg(ckb + (2 * S + 2) * 32) << (32 - nr));
            awh = load_window(A.hi, A.nwords, a_bit + 32 * qq);
            awl = load_window(A.lo, A.nwords, a_bit + 32 * qq);
        }
        const uint32_t *plb = planes + qq + L * S;
        uint2 *out = mb2 + lane * S;
        for (int t = 0; t < 32; ++t) {
            if (t < nr) {
                const uint32_t *pl = plb + (int)(((awh >> t) & 1u) * 2u + ((awl >> t) & 1u)) * PW;
                PB_CHECK_RANGE("tile store", out + t * 32 * S + S - 1, 8, scr, scr + nb_tile_words(S));
                row_step_tile<S>(hp, hn, pl, (cc >> t) & 1u, vp << t, vn << t, out + t * 32 * S);
            }
        }
        __syncwarp();
    };
    // window `slot`: rows i0w, i0w-1, ..., one per lane, around the predicted diagonal kd = j - i + Wl
    auto fetch = [&](int slot, int i0w, int kd) {
        const int row = i0w - lane, t = (row - 1) & 31;
        const int w0 = (kd + t - 16) >> 5; // the row's bit on that diagonal lies in word w0 or w0 + 1, >= 16 bits from their ends
        const int bi = r_qh - ((row - 1) >> 5);
        const bool rowok = row >= 1 && bi < r_nb;
        const int La = rowok ? rtab[bi] : 0;
        // word w0 is slot s0 of lane L0; w0 + 1 the next slot, or slot 0 of the next lane (w0 = -1: lane -1, never in the round)
        const int L0 = (w0 + S) / S - 1, s0 = w0 - L0 * S;
        const int L1 = s0 + 1 < S ? L0 : L0 + 1, s1 = s0 + 1 < S ? s0 + 1 : 0;
        const bool ok0 = rowok && (unsigned)w0 < (unsigned)NBw && (unsigned)(L0 - La) < (unsigned)NLK;
        const bool ok1 = rowok && (unsigned)(w0 + 1) < (unsigned)NBw && (unsigned)(L1 - La) < (unsigned)NLK;
        const int tbase = t * 32 + bi * NLK - La; // tile index of lane L in this row: tbase + L
        uint32_t *dst = ring + slot * 128 + 4 * lane;
        PB_CHECK_RANGE("ring slot", dst, 16, ring, ring + PB_NB_RING * 128);
        // rows above the matrix / words outside the round: plain zero stores, never zero-fill copies (see pb_align.cu)
        if (ok0) {
            const uint2 *src = mb2 + (tbase + L0) * S + s0;
            PB_CHECK_RANGE("traceback prefetch", src, 8, scr, scr + nb_tile_words(S));
            cp_async8(dst, src, 8);
        } else *reinterpret_cast<uint2 *>(dst) = make_uint2(0u, 0u);
        if (ok1) {
            const uint2 *src = mb2 + (tbase + L1) * S + s1;
            PB_CHECK_RANGE("traceback prefetch (2nd word)", src, 8, scr, scr + nb_tile_words(S));
            cp_async8(dst + 2, src, 8);
        } else *reinterpret_cast<uint2 *>(dst + 2) = make_uint2(0u, 0u);
        winfo[slot * 32 + lane] = ((w0 + 1) << 2) | (ok1 ? 2 : 0) | (ok0 ? 1 : 0);
        if (lane == 0) { meta[2 * slot] = i0w; meta[2 * slot + 1] = kd; }
        cp_async_commit();
    };
    // What the walk records, per row of seg_a it leaves: how many INSERT steps it took along the row and whether it left it by
    // DELETE or by MATCH -- 16 bits per row, written 32 rows at a time; the transcript is laid out from that at the end.
    uint16_t *rowops = reinterpret_cast<uint16_t *>(opsrev);
    int i = matlen_a, j = matlen_b, nins = 0;
    {
        const int guard = 2 * (len_a + len_b) + 64; // windows: far more than any path needs
        int cur_slot = 0, cur_i0 = 0, kd = 0, obase = 0, carry = 0, nwin = 0;
        bool a0 = false, a1 = false;
        uint32_t avboth = 0u;
        uint4 U = make_uint4(0u, 0u, 0u, 0u);
        bool have = false, round_ok = false;
        while (i > 0 && j > 0) {
            const int k = j - i + Wl;
            bool fresh = false;
            if (++nwin > guard) { redo = 1; break; }
            if (have) { // the window is walked: the next one in the ring starts at row i, unless a cold start intervened
                const int nslot = cur_slot + 1 == PB_NB_RING ? 0 : cur_slot + 1;
                if (meta[2 * nslot] == i) {
                    __syncwarp();
                    fetch(cur_slot, i - 32 * (PB_NB_RING - 1), k + ((r_sl * 32 * (PB_NB_RING - 1)) >> 8)); // refill the slot just walked
                    cur_slot = nslot;
                } else have = false;
            }
            if (!have) { // cold start: (re)fill the ring from the current cell on its own diagonal
                cp_async_wait<0>();
                __syncwarp();
                const int w = (k + ((i - 1) & 31)) >> 5;
                if (!round_ok || tile_of(i, w) < 0) { // the cell is not in the current round: recompute from here
                    new_round(i, k);
                    round_ok = true;
                    if (tile_of(i, w) < 0) { redo = 1; break; } // outside the frame: never for a certified goal
                }
#pragma unroll
                for (int t = 0; t < PB_NB_RING; ++t) fetch(t, i - 32 * t, k + ((r_sl * 32 * t) >> 8));
                cur_slot = 0;
                ++tbc[1];
                fresh = true;
            }
            cp_async_wait<PB_NB_RING - 1>();
            __syncwarp();
            cur_i0 = meta[2 * cur_slot];
            kd = meta[2 * cur_slot + 1];
            {
                const int v = winfo[cur_slot * 32 + lane];
                U = *reinterpret_cast<const uint4 *>(ring + cur_slot * 128 + 4 * lane);
                a0 = (v & 1) != 0; a1 = (v & 2) != 0;
                // frame bit c of this lane's row sits at bit c - 32 w0 of its two words: obase + dk on diagonal kd + dk, and
                // obase is in [16, 48), so the diagonals dk in [-16, 16) lie inside the two words
                obase = kd + ((cur_i0 - lane - 1) & 31) - 32 * ((v >> 2) - 1);
                avboth = __ballot_sync(FULL, a0 && a1);
            }
            // ---- walk the window: r = row index (lane) of the current cell, dk = its diagonal relative to kd
            int r = cur_i0 - i, dk = k - kd;
            const int r_in = r, dk_in = dk;
            int ins = lane == r ? carry : 0; // INSERT steps along this lane's row
            uint32_t delmask = 0u;           // rows left by DELETE
            // careful: some row lacks a word, or the window reaches row 0 / column 0 -- availability and the matrix's edges are
            // then checked at every step; everywhere else the window is walked with two ballots per step and nothing else
            const bool careful = avboth != FULL || cur_i0 < 32 || cur_i0 - 31 + kd - 16 - Wl < 1;
            bool done = false;
            if (!careful) {
                while (r < 32 && (unsigned)(dk + 16) < 32u) { // until the window is walked or the path leaves the diagonals it was cut for
                    const int o = obase + dk;
                    const bool hi = (o & 32) != 0;
                    const uint32_t Bm = __ballot_sync(FULL, __funnelshift_r(hi ? U.z : U.x, 0u, o) & 1u);
                    const uint32_t Bi = __ballot_sync(FULL, __funnelshift_r(hi ? U.w : U.y, 0u, o) & 1u);
                    r += __clz(__brev(~(Bm >> r))); // MATCH steps: the zeros shifted in end the run at the window's end at the latest
                    if (r >= 32) break;
                    if ((Bi >> r) & 1u) { // INSERT: one column left along the row
                        --dk; ++nins;
                        ins += lane == r;
                    } else { // DELETE: up, on the next diagonal
                        ++dk;
                        delmask |= 1u << r;
                        ++r;
                    }
                }
            } else
                for (;;) {
                    if (r >= 32 || (unsigned)(dk + 16) >= 32u) break;
                    const int ci = cur_i0 - r, cj = ci + kd + dk - Wl;
                    if (ci <= 0 || cj <= 0) { done = true; break; }
                    const int lim = min(32 - r, min(ci, cj));
                    const int o = obase + dk;
                    const bool hi = (o & 32) != 0, in = hi ? a1 : a0;
                    const uint32_t Av = __ballot_sync(FULL, in);
                    if (!((Av >> r) & 1u)) break; // the window does not hold the current cell
                    const uint32_t Bm = __ballot_sync(FULL, in && (__funnelshift_r(hi ? U.z : U.x, 0u, o) & 1u));
                    const uint32_t Bi = __ballot_sync(FULL, __funnelshift_r(hi ? U.w : U.y, 0u, o) & 1u);
                    const int run = min(__clz(__brev(~(Bm >> r))), lim); // MATCH steps
                    r += run;
                    if (run == lim) continue; // the window's end, or row 0 / column 0: decided at the top
                    if (!((Av >> r) & 1u)) break;
                    if ((Bi >> r) & 1u) {
                        --dk; ++nins;
                        ins += lane == r;
                    } else {
                        ++dk;
                        delmask |= 1u << r;
                        ++r;
                    }
                }
            // rows the walk has left: their record; the row it stands on takes its INSERT count into the next window
            if (lane < r) {
                PB_CHECK_RANGE("row record", rowops + (cur_i0 - lane), 2, opsrev, opsrev + (size_t)ent_cap * 4);
                rowops[cur_i0 - lane] = (uint16_t)((ins << 1) | ((delmask >> lane) & 1u));
            }
            carry = r < 32 ? __shfl_sync(FULL, ins, r) : 0;
            if (fresh && !done && r == r_in && dk == dk_in) { redo = 1; break; } // a fresh window holds its own first cell (never reached for a certified goal)
            i = cur_i0 - r;
            j = i + kd + dk - Wl;
            have = r >= 32;
        }
        cp_async_wait<0>();
        __syncwarp();
        if (redo) return;
    }
    // i == 0: INSERT all the way along row 0; j == 0: DELETE all the way down column 0 (init_cell)
    const int i_end = i, tail_cnt = i == 0 ? max(j, 0) : i;
    const int tail_op = i == 0 ? PB_INSERT : PB_DELETE;
    const int n = tail_cnt + (matlen_a - i_end) + nins;
    if (ops_out) { // forward order: the tail, then row by row the step that enters the row and the INSERT steps along it
        for (int q = lane; q < tail_cnt; q += 32) ops_out[q] = (uint8_t)tail_op;
        int base = tail_cnt;
        for (int row0 = i_end + 1; row0 <= matlen_a; row0 += 32) {
            const int row = row0 + lane;
            const bool valid = row <= matlen_a;
            const uint32_t v = valid ? (uint32_t)__ldcg(rowops + row) : 0u;
            const int cnt = valid ? 1 + (int)(v >> 1) : 0;
            int pre = cnt; // inclusive scan over lanes
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int u = __shfl_up_sync(FULL, pre, d);
                if (lane >= d) pre += u;
            }
            uint8_t *dst = ops_out + base + pre - cnt;
            if (valid) {
                dst[0] = (uint8_t)((v & 1u) ? PB_DELETE : PB_MATCH);
                for (int e = 1; e < cnt; ++e) dst[e] = (uint8_t)PB_INSERT;
            }
            base += __shfl_sync(FULL, pre, 31);
        }
    }
    res.nedit = n;
    res.ret = matlen_b;
}
