/*
 * bitpar_model.c -- CPU model of the K3 kernel's arithmetic (development + CPU test aid).
 *
 * The CUDA aligner (pacbioassembly_b200/csrc/pb_align.cu) does not run the reference's cell-by-cell
 * recurrence (seq_aligner.h:151-190).  It runs a banded, bit-parallel formulation (Myers 1999 /
 * Hyyro 2003, transposed so that the state is the row's HORIZONTAL deltas and the band slides one
 * bit per row) that yields the same costs AND the same parents, including the reference's
 * tie-breaking (diag, then left if strictly smaller, then up if strictly smaller).
 * This file is that formulation written lane-by-lane the way the warp executes it (32 lanes x S
 * 32-bit words, ballot-style carry resolution), checked exhaustively against the oracle.  It is NOT
 * on the product path; tests/test_bitpar_model.py builds and runs it (links oracle/pb_oracle.c).
 *
 * Derivation (band half-width D = max_dst, band bit k of row i <-> column j = i - D + k):
 *   state  Hp/Hn : h_{i-1}(j) = cost(i-1,j) - cost(i-1,j-1) = +1 / -1, re-aligned to row i by a 1-bit
 *                  right shift.  Bits k > 2D are computed like cells but with Eq = 0 (`force` below is that
 *                  keep-mask); by induction their cost is cost(i,i+D) + (k-2D), so the bit sliding into
 *                  k = 2D is always +1: a fake out-of-band cell that can never win (U+1 >= Dg+2 > Dg+m).
 *                  vin = +1 at k = 0 likewise.
 *   Eq           : b[j-1] == a[i-1]
 *   Xv = (((Eq & Hp) + Hp) ^ Hp) | Eq ;  D0 = Xv | Hn            (cost(i,j) == cost(i-1,j-1))
 *   Vp = Hn | ~(Xv | Hp) ; Vn = Hp & Xv                            (vertical deltas of row i)
 *   Hp' = (Vn<<1) | ~(Eq | Hn | (Vp<<1|1)) ; Hn' = (Vp<<1|1) & (Eq | Hn)
 *   parent: MATCH iff Eq | ~D0 ; else INSERT iff Hp' ; else DELETE  (seq_aligner.h:164-173)
 *   columns j <= 0 of rows i < D are fake cells cost(i,j) = i + |j| (Hn = 1, Eq = 0), which satisfy
 *   the recurrence and reproduce init_cell (seq_aligner.h:139-150) for the real cells.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../oracle/pb_oracle.h"

#define LANES 32

typedef struct {
    int32_t ret, len_a, len_b, max_dst, matlen_a, matlen_b, cost, diag_cost, nedit, fail_row;
} model_out;

static inline int code_of(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

static uint32_t funnel_r(uint32_t lo, uint32_t hi, unsigned sh) /* (hi:lo) >> sh, low 32 bits */
{
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}

/* a, b: ACGT only, forward views.  ops/vals as in pbo_align. */
int model_align(const char *a, int a_len, const char *b, int b_len, double R, int maxn, int maxm, model_out *out,
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

    const int W = 2 * D + 1;
    const int NW = (W + 31) / 32;
    const int S = (NW + LANES - 1) / LANES;
    const int T = S * LANES; /* words computed per row */
    /* Eq planes: bit t <-> b index t - D */
    const int PW = (len_a + 31) / 32 + T + 2;
    uint32_t *plane = (uint32_t *)calloc((size_t)4 * PW, 4);
    for (int x = 0; x < len_b; ++x) {
        int c = code_of(b[x]);
        if (c < 0) { free(plane); return -2; }
        int t = x + D;
        plane[c * PW + (t >> 5)] |= 1u << (t & 31);
    }
    /* parents: rows 1..len_a, [row][plane 0=M,1=Hp'][s][lane] */
    uint32_t *par = (uint32_t *)malloc((size_t)(len_a + 1) * 2 * T * 4);
    uint32_t Hp[LANES][16], Hn[LANES][16], force[LANES][16], diagm[LANES][16];
    for (int L = 0; L < LANES; ++L)
        for (int s = 0; s < S; ++s) {
            int w = L * S + s;
            uint32_t hp = 0, hn = 0, f = 0, dm = 0;
            for (int bit = 0; bit < 32; ++bit) {
                int k = w * 32 + bit;
                if (k > D) hp |= 1u << bit; else hn |= 1u << bit;
                if (k <= 2 * D) f |= 1u << bit; /* keep mask: Eq is forced to 0 above the band's upper edge */
                if (k == D) dm |= 1u << bit;
            }
            Hp[L][s] = hp; Hn[L][s] = hn; force[L][s] = f; diagm[L][s] = dm;
        }

    int cii = 0;          /* cost(i,i) */
    int colc = 0, colbest = 0, col_i = 0; /* last-column tracking when len_a > len_b */
    for (int i = 1; i <= len_a; ++i) {
        int ca = code_of(a[i - 1]);
        if (ca < 0) { free(plane); free(par); return -2; }
        const uint32_t *pl = plane + ca * PW;
        const int q = (i - 1) >> 5;
        const unsigned sh = (unsigned)(i - 1) & 31;
        /* phase 1: state shift (needs bit 0 of the next lane's word 0) */
        uint32_t nx_hp[LANES], nx_hn[LANES];
        for (int L = 0; L < LANES; ++L) {
            nx_hp[L] = L + 1 < LANES ? Hp[L + 1][0] : 0xFFFFFFFFu;
            nx_hn[L] = L + 1 < LANES ? Hn[L + 1][0] : 0;
        }
        uint32_t Eq[LANES][16], sum[LANES][16];
        uint32_t G = 0, P = 0;
        for (int L = 0; L < LANES; ++L) {
            for (int s = 0; s < S; ++s) {
                uint32_t hp_hi = s + 1 < S ? Hp[L][s + 1] : nx_hp[L];
                uint32_t hn_hi = s + 1 < S ? Hn[L][s + 1] : nx_hn[L];
                Hp[L][s] = funnel_r(Hp[L][s], hp_hi, 1);
                Hn[L][s] = funnel_r(Hn[L][s], hn_hi, 1);
            }
            /* Eq words and the block add with carry-in 0 */
            uint32_t carry = 0, allones = 1;
            for (int s = 0; s < S; ++s) {
                int w = L * S + s;
                Eq[L][s] = funnel_r(pl[q + w], pl[q + w + 1], sh) & force[L][s];
                uint64_t t = (uint64_t)(Eq[L][s] & Hp[L][s]) + Hp[L][s] + carry;
                sum[L][s] = (uint32_t)t;
                carry = (uint32_t)(t >> 32);
                allones &= sum[L][s] == 0xFFFFFFFFu;
            }
            if (carry) G |= 1u << L;
            if (allones) P |= 1u << L;
        }
        /* ballot-style carry resolution: carry into lane L = bit L of ((G|P) + G) ^ P */
        uint32_t cin = ((G | P) + G) ^ P;
        /* phase 2 */
        uint32_t Vp[LANES][16], Vn[LANES][16], Xh[LANES][16], Mm[LANES][16];
        uint32_t d0diag = 0;
        for (int L = 0; L < LANES; ++L) {
            uint32_t c = (cin >> L) & 1;
            for (int s = 0; s < S; ++s) {
                uint32_t v = sum[L][s] + c;
                c = c & (v == 0);
                sum[L][s] = v;
                uint32_t Xv = (sum[L][s] ^ Hp[L][s]) | Eq[L][s];
                Vp[L][s] = Hn[L][s] | ~(Xv | Hp[L][s]);
                Vn[L][s] = Hp[L][s] & Xv;
                uint32_t D0 = Xv | Hn[L][s];
                Mm[L][s] = Eq[L][s] | ~D0;
                Xh[L][s] = Eq[L][s] | Hn[L][s];
                if (D0 & diagm[L][s]) d0diag = 1;
            }
        }
        /* phase 3: shift V left by one (top bits of the previous lane), new H */
        uint32_t *prow = par + (size_t)i * 2 * T;
        for (int L = 0; L < LANES; ++L) {
            uint32_t pin = L ? Vp[L - 1][S - 1] >> 31 : 1u; /* vin = +1 */
            uint32_t nin = L ? Vn[L - 1][S - 1] >> 31 : 0u;
            for (int s = 0; s < S; ++s) {
                uint32_t vps = (Vp[L][s] << 1) | pin, vns = (Vn[L][s] << 1) | nin;
                pin = Vp[L][s] >> 31; nin = Vn[L][s] >> 31;
                Hp[L][s] = vns | ~(Xh[L][s] | vps);
                Hn[L][s] = vps & Xh[L][s];
                prow[0 * T + s * LANES + L] = Mm[L][s];
                prow[1 * T + s * LANES + L] = Hp[L][s];
            }
        }
        cii += 1 - (int)d0diag;
        if (i > 10 && i <= len_b && (double)cii > i * R) { /* seq_aligner.h:185, fresh semantics (Q-D2) */
            out->fail_row = i;
            free(plane); free(par);
            return -1;
        }
        if (i == len_b) { colc = colbest = cii; col_i = i; }
        if (i > len_b) { /* only when len_a > len_b: vertical delta at column len_b */
            int k = len_b - i + D, w = k >> 5, L = w / S, s = w % S;
            colc += (int)((Vp[L][s] >> (k & 31)) & 1) - (int)((Vn[L][s] >> (k & 31)) & 1);
            if (colc < colbest) { colbest = colc; col_i = i; }
        }
    }
    /* goal_cell, seq_aligner.h:191-213 */
    int matlen_a, matlen_b, cost;
    if (len_a > len_b) {
        matlen_a = col_i; matlen_b = len_b; cost = colbest;
    } else {
        matlen_a = len_a; matlen_b = len_a; cost = cii;
        int c = cii;
        for (int j = len_a + 1; j <= len_b; ++j) {
            int k = j - len_a + D, w = k >> 5, L = w / S, s = w % S;
            c += (int)((Hp[L][s] >> (k & 31)) & 1) - (int)((Hn[L][s] >> (k & 31)) & 1);
            if (c < cost) { cost = c; matlen_b = j; }
        }
    }
    out->matlen_a = matlen_a; out->matlen_b = matlen_b; out->cost = cost;
    out->diag_cost = (a_len <= len_a && a_len <= len_b) ? cii : 0;
    if ((double)matlen_b < len_b * (1 - R)) { free(plane); free(par); return -1; }
    /* find_path, seq_aligner.h:214-233 */
    int n = 0, i = matlen_a, j = matlen_b;
    uint8_t *rev = (uint8_t *)malloc((size_t)len_a + len_b + 8);
    char *rv = (char *)malloc((size_t)len_a + len_b + 8);
    while (i || j) {
        int op;
        if (i == 0) op = PBO_INSERT;
        else if (j == 0) op = PBO_DELETE;
        else {
            int k = j - i + D, w = k >> 5, L = w / S, s = w % S;
            const uint32_t *prow = par + (size_t)i * 2 * T;
            if ((prow[s * LANES + L] >> (k & 31)) & 1) op = PBO_MATCH;
            else if ((prow[T + s * LANES + L] >> (k & 31)) & 1) op = PBO_INSERT;
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
    free(rev); free(rv); free(plane); free(par);
    return matlen_b;
}

static uint64_t rs = 88172645463325252ull;
static uint32_t rnd(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); }

int main(int argc, char **argv)
{
    int ncases = argc > 1 ? atoi(argv[1]) : 3000;
    int maxlen = argc > 2 ? atoi(argv[2]) : 700;
    static const char ACGT[4] = {'A', 'C', 'G', 'T'};
    char *a = malloc(maxlen * 3 + 64), *b = malloc(maxlen * 3 + 64);
    uint8_t *o1 = malloc(maxlen * 6 + 64), *o2 = malloc(maxlen * 6 + 64);
    char *v1 = malloc(maxlen * 6 + 64), *v2 = malloc(maxlen * 6 + 64);
    int nsucc = 0, nfail_early = 0;
    for (int cs = 0; cs < ncases; ++cs) {
        int n = 1 + rnd() % maxlen;
        double rates[5] = {0.0, 0.02, 0.08, 0.15, 0.3};
        double rate = rates[rnd() % 5];
        for (int k = 0; k < n; ++k) a[k] = ACGT[rnd() & 3];
        int m = 0;
        if (cs % 7 == 6) {
            m = 1 + rnd() % maxlen;
            for (int k = 0; k < m; ++k) b[k] = ACGT[rnd() & 3];
        } else {
            for (int k = 0; k < n; ++k) {
                double u = (rnd() & 0xFFFFF) / 1048576.0;
                if (u < rate * 0.5) { b[m++] = ACGT[rnd() & 3]; b[m++] = a[k]; }
                else if (u < rate * 0.8) continue;
                else if (u < rate) b[m++] = ACGT[rnd() & 3];
                else b[m++] = a[k];
            }
            int extra = rnd() % (maxlen / 2 + 1);
            if (cs % 3 == 0) extra = 0;
            for (int k = 0; k < extra; ++k) b[m++] = ACGT[rnd() & 3];
            if (cs % 5 == 0 && m > 3) m -= rnd() % (m / 3 + 1); /* truncated b: exercises len_a > len_b */
            if (m == 0) b[m++] = 'A';
        }
        const char *pa = a, *pb = b; int la = n, lb = m;
        if (cs & 1) { pa = b; pb = a; la = m; lb = n; }
        double Rs[5] = {0.05, 0.15, 0.3, 0.3, 0.45};
        double R = Rs[rnd() % 5];
        pbo_align_out po; model_out mo;
        int r1 = pbo_align(pa, la, 1, pb, lb, 1, R, 26000, 6000, &po, o1, v1, (size_t)maxlen * 6);
        int r2 = model_align(pa, la, pb, lb, R, 26000, 6000, &mo, o2, v2);
        int bad = r1 != r2 || po.fail_row != mo.fail_row;
        if (!bad && r1 >= 0)
            bad = po.matlen_a != mo.matlen_a || po.matlen_b != mo.matlen_b || po.cost != mo.cost ||
                  po.diag_cost != mo.diag_cost || po.nedit != mo.nedit || memcmp(o1, o2, (size_t)po.nedit) ||
                  memcmp(v1, v2, (size_t)po.nedit);
        if (bad) {
            printf("MISMATCH case %d la=%d lb=%d R=%g: oracle ret=%d cost=%d ma=%d mb=%d ne=%d fr=%d | model ret=%d cost=%d ma=%d mb=%d ne=%d fr=%d\n",
                   cs, la, lb, R, r1, po.cost, po.matlen_a, po.matlen_b, po.nedit, po.fail_row, r2, mo.cost, mo.matlen_a,
                   mo.matlen_b, mo.nedit, mo.fail_row);
            return 1;
        }
        nsucc += r1 >= 0;
        nfail_early += po.fail_row > 0;
    }
    printf("OK %d cases, %d aligned, %d early failures\n", ncases, nsucc, nfail_early);
    return 0;
}
