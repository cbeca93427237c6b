/*
 * narrow_model.c -- CPU model of the NARROW, BLOCK-STATIONARY variant of K3 (development + CPU test aid).
 *
 * pb_align.cu's first-pass aligner does not compute the reference's whole band |i-j| <= D (seq_aligner.h:158-159).
 * It computes a narrower strip of offsets [-Wl, +Wr] around the main diagonal and PROVES, per alignment, that the
 * result it reports is the one the full band gives; what it cannot prove is handed to the full-band kernel (REDO).
 * This file is that arithmetic word by word as the warp executes it, checked against the oracle: every result the
 * model certifies must be identical (return value, fail row, goal cell, cost, diagonal cost, transcript).
 *
 * Why a strip is enough (o = j - i is a cell's offset; every off-diagonal step costs 1):
 *   (1) computed >= true everywhere: the strip only removes paths; cells outside it are treated as reachable at
 *       "left/up neighbour + 1", an upper bound of their true cost.
 *   (2) a path from (0,0) to a cell of offset o with cost c never climbs above offset (c+o)/2 nor below -(c-o)/2.
 *       So a cell whose true cost c satisfies (c+o)/2 <= Wr and (c-o)/2 <= Wl has all its optimal paths inside the
 *       strip and is computed exactly -- value and, by the same argument on its three neighbours, parent.
 *   (3) early failure (seq_aligner.h:185): computed(i,i) <= floor(i*R) passes for certain by (1); computed(i,i) >
 *       floor(i*R) fails for certain when floor(i*R)/2 <= min(Wl,Wr), by (2).  floor(i*R) <= D-1, so Wl,Wr >= D/2
 *       makes every decision exact.
 *   (4) goal + traceback, goal on the last row (len_a <= len_b, offsets o' >= 0): the computed minimum m <= D-1.  A
 *       goal-row cell that beats it has cost <= m and offset o' <= m, so its paths stay within [-m/2, +m]: exact if
 *       m <= Wr (m/2 <= D/2 <= Wl holds already).  Cells right of the strip cost more than Wr >= m.  Mirrored when
 *       the goal is on the last column (len_a > len_b): m <= Wl.
 *   (5) parents are only STORED where the path can be: a path of cost m ending at offset og never leaves
 *       [-(m - og)/2, (m + og)/2] (the walk below asserts it); the kernel checks those two numbers against the lanes it wrote.
 *   Not certified -> REDO with the full band.
 *
 * Parents are not stored at all by the kernel's forward pass.  Per 32-row block and lane (S words of the frame) it keeps a
 * CHECKPOINT: the lane's horizontal deltas at the top of the block (2 S words) and three 32-bit columns collected over the
 * block's rows -- the vertical delta entering the lane's first word (+ and - bit) and the carry entering the lane's
 * multi-word add.  With those a single thread reproduces the lane's 32 rows x S words of parents on its own (no neighbour
 * lanes needed), and the traceback recomputes just the (block, lane) tiles the path runs through.  The model keeps the
 * forward pass's parents only to assert that every tile the path touches is reproduced bit for bit.
 *
 * Block-stationary frame: for the 32 rows i0..i0+31 of a block the strip is held in COLUMN coordinates -- frame bit c
 * is column j = i0 - Wl + c for the whole block -- so nothing slides per row: Eq words are word-aligned plane words,
 * there is no per-row state shift and no moving edge mask.  Between blocks the state moves down by one whole word;
 * the word that enters on the right starts at h = +1 (upper bound, see (1)).  Row i0+t therefore covers offsets
 * [-Wl-t, NB-1-Wl-t]; the strip that is guaranteed for every row is [-Wl, NB-32-Wl].  The frame must stay inside the
 * reference's band: Wl + 32 <= D and NB - Wl <= D, so none of the reference's edge rules (no INSERT at o = -D, no
 * DELETE at o = +D) is ever in play.  Wl is a multiple of 32*S: the main diagonal then lives in slot 0 of one lane
 * and row t's diagonal bit is bit t.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../oracle/pb_oracle.h"

typedef struct {
    int32_t ret, len_a, len_b, max_dst, matlen_a, matlen_b, cost, diag_cost, nedit, fail_row;
    int32_t redo; /* 1: nothing here is certified, run the full band */
} model_out;

static inline int code_of(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

/* The strip for band half-width D in band class S (32 lanes x S words): Wl and the number of valid frame words.
 * goal_left: the goal cell is searched on the last column (len_a > len_b).  target: wanted width of the goal side.
 * Returns 0 when this class cannot hold a certified strip for D.  (Same function as nb_policy in pb_align.cu.) */
static int nb_policy(int D, int S, int goal_left, int target, int *Wl_out, int *NBw_out, int *Wgoal_out)
{ /* returns 0: no certified strip in this class; 1: valid, goal side limited by the class capacity; 2: target met (or the
     goal side is as wide as the reference band allows) */
    const int unit = 32 * S, cap = 1024 * S;
    const int Wh = D / 2; /* floor((D-1)/2) <= D/2: the non-goal side, and the least the goal side may have */
    if (target < Wh) target = Wh;
    int Wl, NB, full = 0;
    if (!goal_left) {
        Wl = (Wh + unit - 1) / unit * unit;
        if (Wl + 32 > D) return 0;
        NB = (Wl + 32 + target + 31) & ~31;  /* guaranteed right side = NB - 32 - Wl */
        if (NB >= ((D + Wl) & ~31)) { NB = (D + Wl) & ~31; full = 1; } /* frame inside the band: NB - Wl <= D */
        if (NB > cap) { NB = cap; full = 0; }
        if (NB - 32 - Wl < Wh) return 0;
        *Wgoal_out = NB - 32 - Wl;
    } else {
        int wl_max = target + unit - 1; /* no wider than asked (rounded up to the unit) */
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

static int g_force_S = 0;

/* a, b: ACGT only, forward views */
int model_align(const char *a, int a_len, const char *b, int b_len, double R, int maxn, int maxm, double g, model_out *out,
                uint8_t *ops, char *vals)
{
    int len_a, len_b, D;
    memset(out, 0, sizeof *out);
    out->ret = -1;
    if (b_len >= a_len) {
        len_a = a_len; D = 1 + (int)(len_a * R); len_b = b_len < len_a + D ? b_len : len_a + D;
    } else {
        len_b = b_len; D = 1 + (int)(len_b * R); len_a = a_len < len_b + D ? a_len : len_b + D;
    }
    out->len_a = len_a; out->len_b = len_b; out->max_dst = D;
    if (len_a >= maxn || D >= maxm) return -1;

    const int goal_left = len_a > len_b;
    const int target = (int)(g * D) + 1;
    int S = 0, Wl = 0, NBw = 0, Wgoal = 0;
    for (S = 1; S <= 16; ++S) {
        if (g_force_S && S != g_force_S) continue;
        const int pr = nb_policy(D, S, goal_left, target, &Wl, &NBw, &Wgoal);
        if (pr == 2 || (pr == 1 && g_force_S)) break;
    }
    if (S > 16) { out->redo = 1; return -1; }
    const int T = 32 * S;
    /* Eq planes: bit t <-> b index t - Wl (column j = t - Wl + 1) */
    const int PW = (len_a + 31) / 32 + T + 2;
    uint32_t *plane = (uint32_t *)calloc((size_t)4 * PW, 4);
    for (int x = 0; x < len_b; ++x) {
        int c = code_of(b[x]);
        if (c < 0) { free(plane); return -2; }
        int t = x + Wl;
        if ((t >> 5) < PW) plane[c * PW + (t >> 5)] |= 1u << (t & 31); /* columns no frame ever reaches are not needed */
    }
    uint32_t *par = (uint32_t *)malloc((size_t)(len_a + 1) * 2 * T * 4); /* [row][0=M,1=I][word] */
    /* checkpoints: [block][lane]: Hp[S], Hn[S] at the top of the block; vin+ / vin- / carry-in columns (bit t = row t) */
    const int nblk = (len_a + 31) / 32 + 1;
    uint32_t *ck_h = (uint32_t *)calloc((size_t)nblk * 32 * 2 * S, 4);
    uint32_t *ck_c = (uint32_t *)calloc((size_t)nblk * 32 * 3, 4);
    uint32_t Hp[512], Hn[512], Vp[512], Vn[512];
    for (int w = 0; w < T; ++w) {
        uint32_t hn = 0;
        for (int bit = 0; bit < 32; ++bit)
            if (w * 32 + bit < Wl) hn |= 1u << bit; /* columns j <= 0: fake cells cost(i,j) = i + |j| */
        Hn[w] = hn; Hp[w] = ~hn;
    }
    int cii = 0, colc = 0, colbest = 0, col_i = 0;
    int rows_max = len_a;
    if (goal_left && rows_max > len_b + Wl) rows_max = len_b + Wl; /* below that the last column has left the frame (cost > Wl) */
    for (int i = 1; i <= rows_max; ++i) {
        const int q = (i - 1) >> 5, t = (i - 1) & 31;
        if (t == 0 && q > 0) { /* next block: the frame moves right by one word */
            for (int w = 0; w + 1 < T; ++w) { Hp[w] = Hp[w + 1]; Hn[w] = Hn[w + 1]; }
            Hp[NBw - 1] = 0xFFFFFFFFu; Hn[NBw - 1] = 0u;
        }
        if (t == 0) /* checkpoint: the state the block's first row starts from */
            for (int L = 0; L < 32; ++L)
                for (int s2 = 0; s2 < S; ++s2) {
                    ck_h[((size_t)q * 32 + L) * 2 * S + s2] = Hp[L * S + s2];
                    ck_h[((size_t)q * 32 + L) * 2 * S + S + s2] = Hn[L * S + s2];
                }
        int ca = code_of(a[i - 1]);
        if (ca < 0) { free(plane); free(par); free(ck_h); free(ck_c); return -2; }
        const uint32_t *pl = plane + ca * PW + q;
        uint32_t carry = 0, pin = 1u, nin = 0u; /* vin = +1 at the frame's left edge */
        uint32_t *prow = par + (size_t)i * 2 * T;
        uint32_t d0diag = 0;
        for (int w = 0; w < T; ++w) {
            const uint32_t Eq = pl[w];
            if (w % S == 0) { /* what enters lane w / S in this row */
                uint32_t *cc = ck_c + ((size_t)q * 32 + w / S) * 3;
                cc[0] |= pin << t; cc[1] |= nin << t; cc[2] |= carry << t;
            }
            uint64_t s64 = (uint64_t)(Eq & Hp[w]) + Hp[w] + carry;
            carry = (uint32_t)(s64 >> 32);
            const uint32_t sum = (uint32_t)s64;
            const uint32_t Xv = (sum ^ Hp[w]) | Eq;
            Vp[w] = Hn[w] | ~(Xv | Hp[w]);
            Vn[w] = Hp[w] & Xv;
            const uint32_t D0 = Xv | Hn[w];
            const uint32_t Mm = Eq | ~D0;
            const uint32_t Xh = Eq | Hn[w];
            const uint32_t vps = (Vp[w] << 1) | pin, vns = (Vn[w] << 1) | nin;
            pin = Vp[w] >> 31; nin = Vn[w] >> 31;
            if (w == Wl / 32) d0diag = (D0 >> t) & 1u;
            Hp[w] = vns | ~(Xh | vps);
            Hn[w] = vps & Xh;
            prow[w] = Mm;
            prow[T + w] = Hp[w];
        }
        if (i <= len_b) {
            cii += 1 - (int)d0diag;
            if (i > 10 && (double)cii > i * R) { /* certain: Wl, Wr >= D/2 >= floor(i*R)/2 */
                out->fail_row = i;
                free(plane); free(par); free(ck_h); free(ck_c);
                return -1;
            }
            if (i == len_b) { colc = colbest = cii; col_i = i; }
        } else {
            const int c = len_b - i + Wl + t; /* column len_b in this row's frame; >= 0 by rows_max */
            colc += (int)((Vp[c >> 5] >> (c & 31)) & 1) - (int)((Vn[c >> 5] >> (c & 31)) & 1);
            if (colc < colbest) { colbest = colc; col_i = i; }
        }
    }
    int matlen_a, matlen_b, cost;
    const int t_last = (rows_max - 1) & 31;
    if (goal_left) {
        matlen_a = col_i; matlen_b = len_b; cost = colbest;
    } else {
        matlen_a = len_a; matlen_b = len_a; cost = cii;
        int c = cii;
        for (int j = len_a + 1; j <= len_b; ++j) {
            const int cc = j - len_a + Wl + t_last;
            if (cc >= 32 * NBw) break; /* right of the frame: cost > Wgoal */
            c += (int)((Hp[cc >> 5] >> (cc & 31)) & 1) - (int)((Hn[cc >> 5] >> (cc & 31)) & 1);
            if (c < cost) { cost = c; matlen_b = j; }
        }
    }
    if (cost > Wgoal) { out->redo = 1; free(plane); free(par); free(ck_h); free(ck_c); return -1; } /* (4): not certified */
    out->matlen_a = matlen_a; out->matlen_b = matlen_b; out->cost = cost;
    out->diag_cost = (a_len <= len_a && a_len <= len_b) ? cii : 0;
    if ((double)matlen_b < len_b * (1 - R)) { free(plane); free(par); free(ck_h); free(ck_c); return -1; }
    int n = 0, i = matlen_a, j = matlen_b;
    /* the kernel stores parents only where a path of this cost and goal offset can be: [-(m - og)/2, (m + og)/2] */
    const int og = matlen_b - matlen_a, o_lo = -((cost - og) / 2), o_hi = (cost + og) / 2;
    uint8_t *rev = (uint8_t *)malloc((size_t)len_a + len_b + 8);
    char *rv = (char *)malloc((size_t)len_a + len_b + 8);
    int last_q = -1, last_L = -1;
    while (i || j) {
        int op;
        if (i == 0) op = PBO_INSERT;
        else if (j == 0) op = PBO_DELETE;
        else {
            const int c = j - i + Wl + ((i - 1) & 31);
            if (c < 0 || c >= 32 * NBw) { fprintf(stderr, "certified path left the frame\n"); abort(); }
            if (j - i < o_lo || j - i > o_hi) { fprintf(stderr, "path offset %d outside [%d, %d] (cost %d, goal offset %d)\n", j - i, o_lo, o_hi, cost, og); abort(); }
            const uint32_t *prow = par + (size_t)i * 2 * T;
            { /* the (block, lane) tile this cell lies in, recomputed from its checkpoint alone: must equal the forward pass */
                const int q = (i - 1) >> 5, L = (c >> 5) / S;
                if (q != last_q || L != last_L) {
                    last_q = q; last_L = L;
                    uint32_t hp[16], hn[16];
                    for (int s2 = 0; s2 < S; ++s2) { hp[s2] = ck_h[((size_t)q * 32 + L) * 2 * S + s2]; hn[s2] = ck_h[((size_t)q * 32 + L) * 2 * S + S + s2]; }
                    const uint32_t *cc = ck_c + ((size_t)q * 32 + L) * 3;
                    for (int t2 = 0; t2 < 32 && 32 * q + t2 + 1 <= rows_max; ++t2) {
                        const int row = 32 * q + t2 + 1;
                        const uint32_t *pl2 = plane + code_of(a[row - 1]) * PW + q + L * S;
                        uint32_t cy = (cc[2] >> t2) & 1u, pin2 = (cc[0] >> t2) & 1u, nin2 = (cc[1] >> t2) & 1u;
                        const uint32_t *want = par + (size_t)row * 2 * T;
                        for (int s2 = 0; s2 < S; ++s2) {
                            const uint32_t Eq = pl2[s2];
                            uint64_t s64 = (uint64_t)(Eq & hp[s2]) + hp[s2] + cy;
                            cy = (uint32_t)(s64 >> 32);
                            const uint32_t sum = (uint32_t)s64;
                            const uint32_t Xv = (sum ^ hp[s2]) | Eq;
                            const uint32_t vp = hn[s2] | ~(Xv | hp[s2]), vn = hp[s2] & Xv;
                            const uint32_t Mm = Eq | ~(Xv | hn[s2]);
                            const uint32_t Xh = Eq | hn[s2];
                            const uint32_t vps = (vp << 1) | pin2, vns = (vn << 1) | nin2;
                            pin2 = vp >> 31; nin2 = vn >> 31;
                            hp[s2] = vns | ~(Xh | vps);
                            hn[s2] = vps & Xh;
                            if (Mm != want[L * S + s2] || hp[s2] != want[T + L * S + s2]) {
                                fprintf(stderr, "tile (block %d, lane %d) row %d word %d: recomputed parents differ\n", q, L, row, s2);
                                abort();
                            }
                        }
                    }
                }
            }
            if ((prow[c >> 5] >> (c & 31)) & 1) op = PBO_MATCH;
            else if ((prow[T + (c >> 5)] >> (c & 31)) & 1) op = PBO_INSERT;
            else op = PBO_DELETE;
        }
        rev[n] = (uint8_t)op;
        rv[n] = op == PBO_DELETE ? 0 : b[j - 1];
        ++n;
        if (op == PBO_MATCH) { --i; --j; } else if (op == PBO_INSERT) --j; else --i;
    }
    for (int k = 0; k < n; ++k) { ops[k] = rev[n - 1 - k]; vals[k] = rv[n - 1 - k]; }
    out->nedit = n;
    out->ret = matlen_b;
    free(rev); free(rv); free(plane); free(par); free(ck_h); free(ck_c);
    return matlen_b;
}

static uint64_t rs = 88172645463325252ull;
static uint32_t rnd(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); }

int main(int argc, char **argv)
{
    int ncases = argc > 1 ? atoi(argv[1]) : 3000;
    int maxlen = argc > 2 ? atoi(argv[2]) : 1500;
    double g = argc > 3 ? atof(argv[3]) : 0.8;
    g_force_S = argc > 4 ? atoi(argv[4]) : 0;
    static const char ACGT[4] = {'A', 'C', 'G', 'T'};
    char *a = malloc(maxlen * 3 + 64), *b = malloc(maxlen * 3 + 64);
    uint8_t *o1 = malloc(maxlen * 6 + 64), *o2 = malloc(maxlen * 6 + 64);
    char *v1 = malloc(maxlen * 6 + 64), *v2 = malloc(maxlen * 6 + 64);
    int nsucc = 0, nfail_early = 0, nredo = 0, nredo_succ = 0;
    for (int cs = 0; cs < ncases; ++cs) {
        int n = 200 + rnd() % maxlen;
        double rates[6] = {0.0, 0.02, 0.08, 0.15, 0.22, 0.3};
        double rate = rates[rnd() % 6];
        for (int k = 0; k < n; ++k) a[k] = ACGT[rnd() & 3];
        int m = 0;
        if (cs % 11 == 10) {
            m = 200 + rnd() % maxlen;
            for (int k = 0; k < m; ++k) b[k] = ACGT[rnd() & 3];
        } else {
            const int style = rnd() % 3; /* 0: balanced indels, 1: insertion-heavy, 2: deletion-heavy */
            for (int k = 0; k < n; ++k) {
                double u = (rnd() & 0xFFFFF) / 1048576.0;
                double pi = style == 1 ? 0.7 : style == 2 ? 0.15 : 0.4, pd = style == 2 ? 0.7 : style == 1 ? 0.15 : 0.4;
                if (u < rate * pi) { b[m++] = ACGT[rnd() & 3]; b[m++] = a[k]; }
                else if (u < rate * (pi + pd)) continue;
                else if (u < rate) b[m++] = ACGT[rnd() & 3];
                else b[m++] = a[k];
            }
            if (cs % 13 == 0) { /* a block indel: the path jumps far off the diagonal and stays there */
                int cut = rnd() % (n / 4 + 1), at = rnd() % (m / 2 + 1);
                if (at + cut < m) { memmove(b + at, b + at + cut, (size_t)(m - at - cut)); m -= cut; }
            }
            int extra = rnd() % (maxlen / 2 + 1);
            if (cs % 3 == 0) extra = 0;
            for (int k = 0; k < extra; ++k) b[m++] = ACGT[rnd() & 3];
            if (cs % 5 == 0 && m > 3) m -= rnd() % (m / 3 + 1);
            if (m == 0) b[m++] = 'A';
        }
        const char *pa = a, *pb = b; int la = n, lb = m;
        if (cs & 1) { pa = b; pb = a; la = m; lb = n; }
        double Rs[5] = {0.1, 0.15, 0.3, 0.3, 0.45};
        double R = Rs[rnd() % 5];
        pbo_align_out po; model_out mo;
        int r1 = pbo_align(pa, la, 1, pb, lb, 1, R, 26000, 6000, &po, o1, v1, (size_t)maxlen * 6);
        int r2 = model_align(pa, la, pb, lb, R, 26000, 6000, g, &mo, o2, v2);
        if (mo.redo) { ++nredo; nredo_succ += r1 >= 0; continue; }
        int bad = r1 != r2 || po.fail_row != mo.fail_row;
        if (!bad && r1 >= 0)
            bad = po.matlen_a != mo.matlen_a || po.matlen_b != mo.matlen_b || po.cost != mo.cost ||
                  po.diag_cost != mo.diag_cost || po.nedit != mo.nedit || memcmp(o1, o2, (size_t)po.nedit) ||
                  memcmp(v1, v2, (size_t)po.nedit);
        if (bad) {
            printf("MISMATCH case %d la=%d lb=%d R=%g: oracle ret=%d cost=%d ma=%d mb=%d ne=%d fr=%d dc=%d | model ret=%d cost=%d ma=%d mb=%d ne=%d fr=%d dc=%d\n",
                   cs, la, lb, R, r1, po.cost, po.matlen_a, po.matlen_b, po.nedit, po.fail_row, po.diag_cost, r2, mo.cost, mo.matlen_a,
                   mo.matlen_b, mo.nedit, mo.fail_row, mo.diag_cost);
            return 1;
        }
        nsucc += r1 >= 0;
        nfail_early += po.fail_row > 0;
    }
    printf("OK %d cases, %d aligned, %d early failures, %d redo (%d of them align)\n", ncases, nsucc, nfail_early, nredo, nredo_succ);
    return 0;
}
