/*
 * pb_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see pb_oracle.h).
 *
 * Plain-C restatement of the reference's read-to-reference path.  Parity is
 * PINNED by tests/test_oracle.py against oracle/_ref (the unmodified reference
 * compiled from /root/reference) and the reference's own test vectors.
 */
#include "pb_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* L0                                                                        */
/* ------------------------------------------------------------------------- */

int pbo_c2i(int ch)
{ /* dna_seq.h:21 */
    return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3;
}

static const char k_codes[4] = {'A', 'C', 'G', 'T'}; /* dna_seq.h:30 */

uint32_t pbo_encode(const char *text, size_t avail)
{ /* dna_seq.h:86-96 + t2b :147-159.  Byte k of the word (k=0 is the LSB on the
     little-endian reference platform) holds bases 4k..4k+3, first base in bits 7:6. */
    uint32_t w = 0;
    for (int k = 0; k < 16; ++k) {
        int ch = (size_t)k < avail ? (unsigned char)text[k] : 0;
        uint32_t code = (uint32_t)pbo_c2i(ch);
        int byte = k >> 2, slot = k & 3;
        w |= code << (8 * byte + 6 - 2 * slot);
    }
    return w;
}

void pbo_decode(uint32_t code, char *out16)
{ /* dna_seq.h:101-107 + b2t :160-176 */
    for (int k = 0; k < 16; ++k) {
        int byte = k >> 2, slot = k & 3;
        out16[k] = k_codes[(code >> (8 * byte + 6 - 2 * slot)) & 3];
    }
}

size_t pbo_text2bin(const char *text, size_t tlen, uint8_t *out, size_t cap)
{ /* dna_seq.h:113-127 */
    size_t blen = 4 + (tlen + 3) / 4;
    if (cap < blen) return 0;
    uint32_t l32 = (uint32_t)tlen;
    memcpy(out, &l32, 4); /* native little-endian u32 length */
    uint8_t *pb = out + 4;
    for (size_t i = 0; i < tlen; i += 4) {
        uint8_t b = 0;
        for (size_t k = 0; k < 4 && i + k < tlen; ++k)
            b |= (uint8_t)(pbo_c2i((unsigned char)text[i + k]) << (6 - 2 * k));
        *pb++ = b;
    }
    return blen;
}

size_t pbo_bin2text(const uint8_t *rec, char *out, size_t cap)
{ /* dna_seq.h:133-145 */
    uint32_t tlen;
    memcpy(&tlen, rec, 4);
    if (cap <= tlen) return 0;
    const uint8_t *pb = rec + 4;
    for (size_t i = 0; i < tlen; ++i)
        out[i] = k_codes[(pb[i >> 2] >> (6 - 2 * (i & 3))) & 3];
    out[tlen] = '\0';
    return tlen;
}

static uint8_t rec_byte(const uint8_t *rec, size_t rec_bytes, size_t off)
{
    return off < rec_bytes ? rec[off] : 0;
}

uint32_t pbo_seed_at(const uint8_t *rec, size_t rec_bytes, int pos, int quirk)
{ /* dna_seq.h:62-76 */
    if (quirk && (pos & 3) == 0) {
        /* Q-S1: *((unsigned*)(pbin + 4 + pos)) -- byte offset pos, not pos>>2 */
        uint32_t w = 0;
        for (int k = 0; k < 4; ++k)
            w |= (uint32_t)rec_byte(rec, rec_bytes, 4 + (size_t)pos + k) << (8 * k);
        return w;
    }
    size_t base = 4 + ((size_t)pos >> 2);
    unsigned ls = ((unsigned)pos & 3u) << 1, rs = 8 - ls;
    uint32_t w = 0;
    for (int k = 0; k < 4; ++k) {
        uint8_t hi = rec_byte(rec, rec_bytes, base + k);
        uint8_t lo = rec_byte(rec, rec_bytes, base + k + 1);
        uint8_t v = (uint8_t)((hi << ls) | (ls ? (lo >> rs) : 0));
        w |= (uint32_t)v << (8 * k);
    }
    return w;
}

uint32_t pbo_parse_pattern(const char *pat)
{ /* spaced_seed.cpp:166-180 */
    char dnapat[17] = "AAAAAAAAAAAAAAAA";
    size_t len = strlen(pat);
    if (len > 16) len = 16;
    for (size_t i = 0; i < len; ++i) dnapat[i] = pat[i] == '1' ? 'T' : 'A';
    return pbo_encode(dnapat, 16);
}

/* ------------------------------------------------------------------------- */
/* L1: seed index.  Observable behaviour of hash_map<unsigned, list<int>>:    */
/* key equality + per-key insertion order (common.h:54).                      */
/* ------------------------------------------------------------------------- */

struct pbo_index {
    size_t n;         /* entries */
    size_t nkeys;     /* distinct keys */
    uint32_t *keys;   /* [n] sorted (stable) */
    int32_t *pos;     /* [n] */
    uint32_t *ukeys;  /* [nkeys] */
    size_t *ustart;   /* [nkeys+1] */
};

static void stable_radix(uint32_t *keys, int32_t *pos, size_t n)
{
    uint32_t *k2 = (uint32_t *)malloc(n * sizeof *k2);
    int32_t *p2 = (int32_t *)malloc(n * sizeof *p2);
    for (int pass = 0; pass < 4; ++pass) {
        size_t cnt[257] = {0};
        int sh = pass * 8;
        for (size_t i = 0; i < n; ++i) cnt[((keys[i] >> sh) & 255) + 1]++;
        for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
        for (size_t i = 0; i < n; ++i) {
            size_t o = cnt[(keys[i] >> sh) & 255]++;
            k2[o] = keys[i];
            p2[o] = pos[i];
        }
        memcpy(keys, k2, n * sizeof *k2);
        memcpy(pos, p2, n * sizeof *p2);
    }
    free(k2);
    free(p2);
}

pbo_index *pbo_index_build(const char *ref, size_t len, uint32_t mask, int policy)
{
    size_t cap = policy == PBO_POLICY_LOCATOR ? len : 40000;
    uint32_t *keys = (uint32_t *)malloc((cap + 1) * sizeof *keys);
    int32_t *pos = (int32_t *)malloc((cap + 1) * sizeof *pos);
    size_t n = 0;
    if (policy == PBO_POLICY_LOCATOR) { /* locator.cpp:62-66 */
        for (size_t i = 0; i < len; ++i) {
            uint32_t k = pbo_encode(ref + i, len - i) & mask;
            if (k) { keys[n] = k; pos[n] = (int32_t)i; ++n; }
        }
    } else { /* ref_seq.h:291-311 */
        long L = (long)len, nmax = L - 16;
        long nhead = nmax < 20000 ? nmax : 20000;
        for (long i = 0; i < nhead; ++i) {
            uint32_t k = pbo_encode(ref + i, len - (size_t)i) & mask;
            if (k) { keys[n] = k; pos[n] = (int32_t)i; ++n; }
        }
        long ntail = L - 20000 - 16;
        if (ntail > 20000) ntail = 20000;
        for (long i = 0; i < ntail; ++i) {
            long p = L - 16 - i;
            uint32_t k = pbo_encode(ref + p, len - (size_t)p) & mask;
            if (k) { keys[n] = k; pos[n] = (int32_t)(L - i - 16); ++n; }
        }
    }
    stable_radix(keys, pos, n);
    pbo_index *ix = (pbo_index *)calloc(1, sizeof *ix);
    ix->n = n;
    ix->keys = keys;
    ix->pos = pos;
    size_t nk = 0;
    for (size_t i = 0; i < n; ++i)
        if (i == 0 || keys[i] != keys[i - 1]) ++nk;
    ix->nkeys = nk;
    ix->ukeys = (uint32_t *)malloc((nk + 1) * sizeof *ix->ukeys);
    ix->ustart = (size_t *)malloc((nk + 1) * sizeof *ix->ustart);
    nk = 0;
    for (size_t i = 0; i < n; ++i)
        if (i == 0 || keys[i] != keys[i - 1]) { ix->ukeys[nk] = keys[i]; ix->ustart[nk] = i; ++nk; }
    ix->ustart[nk] = n;
    return ix;
}

void pbo_index_free(pbo_index *ix)
{
    if (!ix) return;
    free(ix->keys); free(ix->pos); free(ix->ukeys); free(ix->ustart); free(ix);
}

size_t pbo_index_nkeys(const pbo_index *ix) { return ix->nkeys; }
size_t pbo_index_nentries(const pbo_index *ix) { return ix->n; }

size_t pbo_index_find(const pbo_index *ix, uint32_t key, const int32_t **pos)
{
    size_t lo = 0, hi = ix->nkeys;
    while (lo < hi) {
        size_t mid = (lo + hi) >> 1;
        if (ix->ukeys[mid] < key) lo = mid + 1; else hi = mid;
    }
    if (lo == ix->nkeys || ix->ukeys[lo] != key) { if (pos) *pos = NULL; return 0; }
    if (pos) *pos = ix->pos + ix->ustart[lo];
    return ix->ustart[lo + 1] - ix->ustart[lo];
}

/* ------------------------------------------------------------------------- */
/* L2: banded DP                                                             */
/* ------------------------------------------------------------------------- */

typedef struct {
    int32_t *cost;
    uint8_t *par;
    size_t cap;
} dp_ws;

static int dp_ws_reserve(dp_ws *ws, size_t cells)
{
    if (cells <= ws->cap) return 0;
    free(ws->cost); free(ws->par);
    ws->cost = (int32_t *)malloc(cells * sizeof(int32_t));
    ws->par = (uint8_t *)malloc(cells);
    ws->cap = ws->cost && ws->par ? cells : 0;
    return ws->cap ? 0 : -1;
}

static void dp_ws_free(dp_ws *ws) { free(ws->cost); free(ws->par); ws->cost = NULL; ws->par = NULL; ws->cap = 0; }

/* wa / wb == NULL: the reference's unit costs (seq_aligner.h:136-137).  Otherwise the labelled EXTENSION of BASELINE config 3
 * ("quality-weighted scoring"): the same recurrence with the two scoring hooks replaced by per-element table lookups --
 * match(a_i, b_j) = (a_i != b_j) ? wa[i] : 0, indel = wa[i] for a DELETE (a_i skipped), wb[j] for an INSERT (b_j skipped) --
 * and the early-failure line scaled by fail_scale (1 = the reference's cost(i,i) > i*R).  The reference pins no result for it. */
static int align_ws_w(dp_ws *ws, const char *a, int a_len, int a_stride,
                      const char *b, int b_len, int b_stride,
                      const uint8_t *wa, const uint8_t *wb, double fail_scale,
                      double R, int maxn, int maxm,
                      pbo_align_out *out, uint8_t *ops, char *vals, size_t cap);

static int align_ws(dp_ws *ws, const char *a, int a_len, int a_stride,
                    const char *b, int b_len, int b_stride,
                    double R, int maxn, int maxm,
                    pbo_align_out *out, uint8_t *ops, char *vals, size_t cap)
{
    return align_ws_w(ws, a, a_len, a_stride, b, b_len, b_stride, NULL, NULL, 1.0, R, maxn, maxm, out, ops, vals, cap);
}

static int align_ws_w(dp_ws *ws, const char *a, int a_len, int a_stride,
                      const char *b, int b_len, int b_stride,
                      const uint8_t *wa, const uint8_t *wb, double fail_scale,
                      double R, int maxn, int maxm,
                      pbo_align_out *out, uint8_t *ops, char *vals, size_t cap)
{
    int len_a, len_b, max_dst;
    memset(out, 0, sizeof *out);
    out->ret = -1;
    /* seq_aligner.h:94-102 */
    if (b_len >= a_len) {
        len_a = a_len;
        max_dst = 1 + (int)(len_a * R);
        len_b = b_len < len_a + max_dst ? b_len : len_a + max_dst;
    } else {
        len_b = b_len;
        max_dst = 1 + (int)(len_b * R);
        len_a = a_len < len_b + max_dst ? a_len : len_b + max_dst;
    }
    out->len_a = len_a; out->len_b = len_b; out->max_dst = max_dst;
    /* seq_aligner.h:104-107, domain per SURVEY Q-D3 */
    if (len_a >= maxn || max_dst >= maxm) return -1;

    const size_t W = 2 * (size_t)max_dst + 1;
    if (dp_ws_reserve(ws, ((size_t)len_a + 1) * W)) return -1;
    int32_t *C = ws->cost;
    uint8_t *P = ws->par;
#define IDX(i, j) ((size_t)(i) * W + (size_t)((j) - (i) + max_dst))
    /* init_cell, seq_aligner.h:139-150 */
    C[IDX(0, 0)] = 0; P[IDX(0, 0)] = 0;
    for (int i = 1; i <= max_dst && i <= len_a; ++i) { C[IDX(i, 0)] = C[IDX(i - 1, 0)] + (wa ? wa[i - 1] : 1); P[IDX(i, 0)] = PBO_DELETE; }
    for (int j = 1; j <= max_dst; ++j) { C[IDX(0, j)] = C[IDX(0, j - 1)] + (wb && j <= b_len ? wb[j - 1] : 1); P[IDX(0, j)] = PBO_INSERT; }

    /* search, seq_aligner.h:151-190 */
    int64_t cells = 0;
    for (int i = 1; i <= len_a; ++i) {
        char c = a[(ptrdiff_t)(i - 1) * a_stride];
        const int w_i = wa ? wa[i - 1] : 1; /* weight of a_i: what a mismatch at it or its deletion costs */
        int beg = i - max_dst > 1 ? i - max_dst : 1;
        int end = i + max_dst < len_b ? i + max_dst : len_b;
        for (int j = beg; j <= end; ++j) {
            char d = b[(ptrdiff_t)(j - 1) * b_stride];
            int t, cost = C[IDX(i - 1, j - 1)] + (c != d ? w_i : 0);
            int src = PBO_MATCH;
            if (i - j < max_dst && (t = C[IDX(i, j - 1)] + (wb ? wb[j - 1] : 1)) < cost) { cost = t; src = PBO_INSERT; }
            if (j - i < max_dst && (t = C[IDX(i - 1, j)] + w_i) < cost) { cost = t; src = PBO_DELETE; }
            C[IDX(i, j)] = cost;
            P[IDX(i, j)] = (uint8_t)src;
        }
        if (end >= beg) cells += end - beg + 1;
        /* early failure :185 ; cell (i,i) unwritten for i>len_b reads 0 (fresh, Q-D2) */
        if (i > 10 && i <= len_b && (double)C[IDX(i, i)] > i * R * fail_scale) {
            out->fail_row = i;
            out->cells = cells;
            return -1;
        }
    }
    out->cells = cells;

    /* goal_cell, seq_aligner.h:191-213 */
    int matlen_a, matlen_b;
    if (len_a > len_b) {
        matlen_a = len_b; matlen_b = len_b;
        int best = C[IDX(len_b, len_b)];
        for (int i = len_b + 1; i <= len_a; ++i)
            if (C[IDX(i, len_b)] < best) { best = C[IDX(i, len_b)]; matlen_a = i; }
    } else {
        matlen_a = len_a; matlen_b = len_a;
        int best = C[IDX(len_a, len_a)];
        for (int j = len_a + 1; j <= len_b; ++j)
            if (C[IDX(len_a, j)] < best) { best = C[IDX(len_a, j)]; matlen_b = j; }
    }
    out->matlen_a = matlen_a; out->matlen_b = matlen_b;
    out->cost = C[IDX(matlen_a, matlen_b)];
    out->diag_cost = (a_len <= len_a && a_len <= len_b) ? C[IDX(a_len, a_len)] : 0;
    /* seq_aligner.h:114 */
    if ((double)matlen_b < len_b * (1 - R)) return -1;

    /* find_path, seq_aligner.h:214-233 (iterative, then reversed) */
    size_t ne = 0;
    {
        int i = matlen_a, j = matlen_b;
        for (;;) {
            int p = P[IDX(i, j)];
            if (p == 0) break;
            ++ne;
            if (p == PBO_MATCH) { --i; --j; }
            else if (p == PBO_INSERT) { --j; }
            else { --i; }
        }
        out->nedit = (int32_t)ne;
        if (ops && ne <= cap) {
            size_t k = ne;
            i = matlen_a; j = matlen_b;
            for (;;) {
                int p = P[IDX(i, j)];
                if (p == 0) break;
                --k;
                ops[k] = (uint8_t)p;
                if (vals) vals[k] = p == PBO_DELETE ? 0 : b[(ptrdiff_t)(j - 1) * b_stride];
                if (p == PBO_MATCH) { --i; --j; }
                else if (p == PBO_INSERT) { --j; }
                else { --i; }
            }
        }
    }
#undef IDX
    out->ret = matlen_b;
    return matlen_b;
}

int pbo_align(const char *a, int a_len, int a_stride,
              const char *b, int b_len, int b_stride,
              double R, int maxn, int maxm,
              pbo_align_out *out, uint8_t *ops, char *vals, size_t cap)
{
    dp_ws ws = {0, 0, 0};
    int r = align_ws(&ws, a, a_len, a_stride, b, b_len, b_stride, R, maxn, maxm, out, ops, vals, cap);
    dp_ws_free(&ws);
    return r;
}

int pbo_align_weighted(const char *a, int a_len, const uint8_t *wa, const char *b, int b_len, const uint8_t *wb,
                       double R, double fail_scale, int maxn, int maxm, pbo_align_out *out, uint8_t *ops, size_t cap)
{
    dp_ws ws = {0, 0, 0};
    int r = align_ws_w(&ws, a, a_len, 1, b, b_len, 1, wa, wb, fail_scale, R, maxn, maxm, out, ops, NULL, cap);
    dp_ws_free(&ws);
    return r;
}

/* ------------------------------------------------------------------------- */
/* locate loop                                                               */
/* ------------------------------------------------------------------------- */

typedef struct {
    const pbo_index *ix;
    const char *ref; size_t ref_len;
    const char *reads; const int64_t *offs; const int32_t *lens;
    const int64_t *kept; int64_t *next, nk; /* threads take the next kept read from a shared counter */
    uint32_t mask; double R; int ntrial, maxn, maxm;
    pbo_locate_rec *recs;
    uint8_t *ops_out; const int64_t *ops_off;
} locate_job;

static void *locate_thread(void *arg)
{
    locate_job *jb = (locate_job *)arg;
    dp_ws ws = {0, 0, 0};
    uint8_t *ops = NULL; size_t ops_cap = 0;
    for (;;) {
        const int64_t k = __atomic_fetch_add(jb->next, 1, __ATOMIC_RELAXED);
        if (k >= jb->nk) break;
        int64_t r = jb->kept[k];
        const char *seq = jb->reads + jb->offs[r];
        int len = jb->lens[r];
        pbo_locate_rec *rec = &jb->recs[k];
        memset(rec, 0, sizeof *rec);
        rec->nseq = (int32_t)k;
        if (jb->ops_out) {
            size_t need = (size_t)len * 2 + (size_t)jb->maxm + 16;
            if (need > ops_cap) { free(ops); ops = (uint8_t *)malloc(need); ops_cap = need; }
        }
        int found = 0;
        for (int j = 0; j < jb->ntrial && !found; ++j) { /* locator.cpp:74 */
            uint32_t key = pbo_encode(seq + j, (size_t)(len - j) + 1 /* NUL terminator readable */) & jb->mask;
            const int32_t *plist;
            size_t cnt = pbo_index_find(jb->ix, key, &plist);
            for (size_t c = 0; c < cnt; ++c) { /* locator.cpp:79 */
                int pos = plist[c];
                pbo_align_out ao;
                rec->ncand++;
                int ret = align_ws(&ws, seq + j, len - j, 1, jb->ref + pos, (int)(jb->ref_len - (size_t)pos), 1,
                                   jb->R, jb->maxn, jb->maxm, &ao, jb->ops_out ? ops : NULL, NULL, ops_cap);
                rec->cells += ao.cells;
                if (ret > 0) { /* locator.cpp:82-88 */
                    found = 1;
                    rec->found = 1; rec->j = j; rec->pos = pos; rec->cost = ao.cost;
                    rec->seg_len = len - j; rec->diag_cost = ao.diag_cost;
                    rec->matlen_a = ao.matlen_a; rec->matlen_b = ao.matlen_b; rec->nedit = ao.nedit;
                    if (jb->ops_out) memcpy(jb->ops_out + jb->ops_off[k], ops, (size_t)ao.nedit);
                    break;
                }
            }
        }
    }
    free(ops);
    dp_ws_free(&ws);
    return NULL;
}

int64_t pbo_locate(const pbo_index *ix, const char *ref, size_t ref_len,
                   const char *reads, const int64_t *offs, const int32_t *lens,
                   int64_t nreads, uint32_t mask, double R, int ntrial, int minlen,
                   int maxn, int maxm, int nthreads, pbo_locate_rec *recs,
                   uint8_t *ops_out, const int64_t *ops_off)
{
    int64_t *kept = (int64_t *)malloc((size_t)(nreads + 1) * sizeof *kept);
    int64_t nk = 0;
    for (int64_t r = 0; r < nreads; ++r)
        if (lens[r] >= minlen) kept[nk++] = r; /* locator.cpp:72, Q-L1 */
    if (nthreads < 1) nthreads = 1;
    if ((int64_t)nthreads > nk) nthreads = nk > 0 ? (int)nk : 1;
    locate_job *jobs = (locate_job *)calloc((size_t)nthreads, sizeof *jobs);
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof *th);
    int64_t next = 0; /* dynamic schedule: read lengths vary 40-fold, a static split leaves threads idle */
    for (int t = 0; t < nthreads; ++t) {
        locate_job *jb = &jobs[t];
        jb->ix = ix; jb->ref = ref; jb->ref_len = ref_len;
        jb->reads = reads; jb->offs = offs; jb->lens = lens; jb->kept = kept;
        jb->next = &next; jb->nk = nk;
        jb->mask = mask; jb->R = R; jb->ntrial = ntrial; jb->maxn = maxn; jb->maxm = maxm;
        jb->recs = recs; jb->ops_out = ops_out; jb->ops_off = ops_off;
        if (nthreads == 1) locate_thread(jb);
        else pthread_create(&th[t], NULL, locate_thread, jb);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(jobs); free(th); free(kept);
    return nk;
}

/* ------------------------------------------------------------------------- */
/* assembler-side trial loop (spaced_seed.cpp:261-299, 424-436)              */
/* ------------------------------------------------------------------------- */

typedef struct {
    const pbo_index *ix; const char *ref; size_t ref_len;
    const uint8_t *bin; size_t bin_bytes; const size_t *rec_off; int64_t k0, k1;
    uint32_t mask; double R; int max_trial, min_overlap, maxn, maxm, quirk;
    pbo_overlap_rec *recs;
} overlap_job;

static uint32_t image_seed_at(const uint8_t *bin, size_t bin_bytes, size_t rec, int pos, int quirk)
{ /* dna_seq::seed_at over the whole image: bytes past the image read 0 */
    return pbo_seed_at(bin + rec, bin_bytes - rec, pos, quirk);
}

static void *overlap_thread(void *arg)
{
    overlap_job *jb = (overlap_job *)arg;
    dp_ws ws = {0, 0, 0};
    char *txt = NULL; size_t txt_cap = 0;
    for (int64_t k = jb->k0; k < jb->k1; ++k) {
        const size_t rec = jb->rec_off[k];
        uint32_t slen;
        memcpy(&slen, jb->bin + rec, 4);
        pbo_overlap_rec *out = &jb->recs[k];
        memset(out, 0, sizeof *out);
        out->id = (int32_t)k;
        if (slen + 2 > txt_cap) { free(txt); txt_cap = slen + 64; txt = (char *)malloc(txt_cap); }
        pbo_bin2text(jb->bin + rec, txt, txt_cap); /* set_active_seg, spaced_seed.cpp:109-118 */
        int found = 0;
        for (int j = 0; j < jb->max_trial && !found; ++j) {
            for (int side = 0; side < 2 && !found; ++side) { /* spaced_seed.cpp:426 */
                const int forward = side == 0;
                const long pos = forward ? j : (long)slen - j - 16;
                const uint32_t key = image_seed_at(jb->bin, jb->bin_bytes, rec, (int)pos, jb->quirk) & jb->mask;
                const int32_t *plist;
                const size_t cnt = pbo_index_find(jb->ix, key, &plist);
                if (!cnt) continue;
                const long s_offset = forward ? pos : pos + 15;
                const long s_len = forward ? (long)slen - s_offset : s_offset + 1;
                if (s_len < jb->min_overlap) continue; /* spaced_seed.cpp:280 */
                for (size_t c = 0; c < cnt && !found; ++c) {
                    const long r_offset = forward ? plist[c] : plist[c] + 15;
                    const long r_len = forward ? (long)jb->ref_len - r_offset : r_offset + 1; /* ref_seq.h:282-286 */
                    pbo_align_out ao;
                    out->ncand++;
                    /* note the argument order: a = reference view, b = read view (ref_seq.h:264) */
                    const int ret = align_ws(&ws, jb->ref + r_offset, (int)r_len, forward ? 1 : -1, txt + s_offset, (int)s_len,
                                             forward ? 1 : -1, jb->R, jb->maxn, jb->maxm, &ao, NULL, NULL, 0);
                    out->cells += ao.cells;
                    if (ret < 0) continue;
                    if (ao.matlen_a < jb->min_overlap) continue; /* ref_seq.h:265 */
                    found = 1;
                    out->found = 1; out->j = j; out->ref_pos = plist[c]; out->cost = ao.cost; out->read_pos = (int32_t)pos;
                    out->dir = forward ? 1 : -1; out->matlen_a = ao.matlen_a; out->matlen_b = ao.matlen_b; out->nedit = ao.nedit;
                }
            }
        }
    }
    free(txt);
    dp_ws_free(&ws);
    return NULL;
}

int64_t pbo_overlap(const pbo_index *ix, const char *ref, size_t ref_len, const uint8_t *bin, size_t bin_bytes,
                    int min_excl, int max_excl, uint32_t mask, double R, int max_trial, int min_overlap, int maxn,
                    int maxm, int quirk, int nthreads, pbo_overlap_rec *recs)
{
    size_t cap = 1024, nk = 0;
    size_t *rec_off = (size_t *)malloc(cap * sizeof *rec_off);
    for (size_t p = 0; p + 4 <= bin_bytes;) { /* open_binary, spaced_seed.cpp:330-342 */
        uint32_t l;
        memcpy(&l, bin + p, 4);
        if ((long)l > min_excl && (long)l < max_excl) {
            if (nk == cap) { cap *= 2; rec_off = (size_t *)realloc(rec_off, cap * sizeof *rec_off); }
            rec_off[nk++] = p;
        }
        p += 4 + ((size_t)l + 3) / 4;
    }
    if (recs) {
        if (nthreads < 1) nthreads = 1;
        if ((size_t)nthreads > nk) nthreads = nk ? (int)nk : 1;
        overlap_job *jobs = (overlap_job *)calloc((size_t)nthreads, sizeof *jobs);
        pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof *th);
        for (int t = 0; t < nthreads; ++t) {
            overlap_job *jb = &jobs[t];
            jb->ix = ix; jb->ref = ref; jb->ref_len = ref_len; jb->bin = bin; jb->bin_bytes = bin_bytes; jb->rec_off = rec_off;
            jb->k0 = (int64_t)(nk * t / nthreads); jb->k1 = (int64_t)(nk * (t + 1) / nthreads);
            jb->mask = mask; jb->R = R; jb->max_trial = max_trial; jb->min_overlap = min_overlap;
            jb->maxn = maxn; jb->maxm = maxm; jb->quirk = quirk; jb->recs = recs;
            if (nthreads == 1) overlap_thread(jb);
            else pthread_create(&th[t], NULL, overlap_thread, jb);
        }
        if (nthreads > 1)
            for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
        free(jobs); free(th);
    }
    free(rec_off);
    return (int64_t)nk;
}

/* ============================================================================================================
 * Consensus voting and the UNLOCKED assembler rounds (ref_seq.h:25-41 apply_edits, :47-183 base_vote / vote_box,
 * :207-256 ref_seq ctor / append / prepend, :259-276 try_align, :317-362 evolve / elect; spaced_seed.cpp:408-453).
 * ============================================================================================================ */

typedef struct { uint16_t sel[4], sup[4]; int32_t total; } vbox; /* vote_box: selection, suppliment (unsigned short), total */

struct pbo_cons {
    char *txt;          /* txt_buf: [pre, post) is readable; beg/end delimit the current iteration's reference */
    size_t cap;         /* txt has 3*cap bytes; beg starts at cap (MAX_SEQ_LEN in the reference) */
    long beg, end, pre, post;
    vbox *box;          /* consensus list, box[0] <-> txt[pre] */
    size_t nbox, boxcap;
};

static int vmax(const uint16_t *v) { int m = v[0]; for (int k = 1; k < 4; ++k) if (v[k] > m) m = v[k]; return m; }
static char vwinner(const uint16_t *v) { int m = vmax(v); return m == v[0] ? 'A' : (m == v[1] ? 'C' : (m == v[2] ? 'G' : 'T')); } /* :94-98 */

static void cons_reserve(pbo_cons *c, size_t n)
{
    if (n <= c->boxcap) return;
    size_t nc = c->boxcap ? c->boxcap : 1024;
    while (nc < n) nc *= 2;
    c->box = (vbox *)realloc(c->box, nc * sizeof(vbox));
    c->boxcap = nc;
}

pbo_cons *pbo_cons_create(const char *text, size_t len, int weight, size_t cap)
{ /* ref_seq(const char*, int, bool, int w), ref_seq.h:218-225: vote_box(c, w) has selection[C2I(c)] = w and total = 1 */
    pbo_cons *c = (pbo_cons *)calloc(1, sizeof *c);
    c->cap = cap < len + 16 ? len + 16 : cap;
    c->txt = (char *)calloc(3 * c->cap, 1);
    c->beg = c->pre = (long)c->cap;
    c->end = c->post = c->beg + (long)len;
    memcpy(c->txt + c->beg, text, len);
    cons_reserve(c, len + 1);
    for (size_t i = 0; i < len; ++i) {
        vbox b; memset(&b, 0, sizeof b);
        b.sel[pbo_c2i(text[i])] = (uint16_t)weight;
        b.total = 1;
        c->box[i] = b;
    }
    c->nbox = len;
    return c;
}
void pbo_cons_free(pbo_cons *c) { if (c) { free(c->txt); free(c->box); free(c); } }
size_t pbo_cons_length(const pbo_cons *c) { return (size_t)(c->end - c->beg); }     /* ref_seq::length() */
size_t pbo_cons_extent(const pbo_cons *c, long *before) { if (before) *before = c->beg - c->pre; return (size_t)(c->post - c->pre); }
const char *pbo_cons_text(const pbo_cons *c) { return c->txt + c->beg; }

void pbo_cons_append(pbo_cons *c, const char *seg, int len)
{ /* ref_seq.h:227-233 */
    if (len <= 0) return;
    memmove(c->txt + c->post, seg, (size_t)len);
    c->post += len;
    cons_reserve(c, c->nbox + (size_t)len);
    for (int i = 0; i < len; ++i) {
        vbox b; memset(&b, 0, sizeof b);
        b.sel[pbo_c2i(seg[i])] = 1; b.total = 1;
        c->box[c->nbox++] = b;
    }
}
void pbo_cons_prepend(pbo_cons *c, const char *seg, int len)
{ /* ref_seq.h:235-243: the text is copied in order, the boxes are pushed to the front last char first */
    if (len <= 0) return;
    c->pre -= len;
    memmove(c->txt + c->pre, seg, (size_t)len);
    cons_reserve(c, c->nbox + (size_t)len);
    memmove(c->box + len, c->box, c->nbox * sizeof(vbox));
    for (int i = 0; i < len; ++i) {
        vbox b; memset(&b, 0, sizeof b);
        b.sel[pbo_c2i(seg[i])] = 1; b.total = 1;
        c->box[i] = b;
    }
    c->nbox += (size_t)len;
}

void pbo_cons_elect(pbo_cons *c, int pos, const uint8_t *ops, const char *vals, int nedit, int forward)
{ /* elect (ref_seq.h:351-361) + apply_edits (:25-41): both iterators start on the box of text position pos */
    long idx = pos + c->beg - c->pre;
    const long step = forward ? 1 : -1;
    for (int k = 0; k < nedit; ++k) {
        if (ops[k] == PBO_DELETE) { c->box[idx].total++; idx += step; }
        else if (ops[k] == PBO_MATCH) { c->box[idx].sel[pbo_c2i(vals[k])]++; c->box[idx].total++; idx += step; }
        else { /* INSERT: forward "--it; supply; ++it" = the box before; a reverse_iterator supplies the box it stands on */
            const long t = forward ? idx - 1 : idx;
            if (t >= 0 && t < (long)c->nbox) c->box[t].sup[pbo_c2i(vals[k])]++; /* t < 0 is undefined behaviour in the reference */
        }
    }
}

void pbo_cons_evolve(pbo_cons *c)
{ /* ref_seq.h:317-348 */
    vbox *out = (vbox *)malloc((2 * c->nbox + 2) * sizeof(vbox));
    size_t no = 0; /* boxes kept so far = the list in front of `cur` */
    c->end = c->pre = c->beg = (long)c->cap;
    char *p = c->txt + c->beg;
    for (size_t i = 0; i < c->nbox; ++i) {
        vbox cur = c->box[i], ins;
        int has_ins = 0;
        if (2 * vmax(cur.sup) > cur.total) { /* has_supply(0.5): split the suppliment off as a box of its own, right after */
            memset(&ins, 0, sizeof ins);
            memcpy(ins.sel, cur.sup, sizeof cur.sup);
            ins.total = cur.total;
            memset(cur.sup, 0, sizeof cur.sup);
            has_ins = 1;
        }
        if (2 * vmax(cur.sel) > cur.total) { /* is_valid(0.5) */
            *p++ = vwinner(cur.sel); ++c->end;
            out[no++] = cur;
        } else if (no > 0) { /* erased: its selection is absorbed by the previous box's suppliment */
            for (int k = 0; k < 4; ++k) out[no - 1].sup[k] = (uint16_t)(out[no - 1].sup[k] + cur.sel[k]);
        }
        if (has_ins) { /* visited next: no suppliment, and valid by the same inequality that created it */
            *p++ = vwinner(ins.sel); ++c->end;
            out[no++] = ins;
        }
    }
    c->post = c->end;
    free(c->box);
    c->box = out; c->nbox = no; c->boxcap = 2 * c->nbox + 2;
    if (c->boxcap < no) c->boxcap = no;
}

int64_t pbo_cons_votes(const pbo_cons *c, int32_t *out9)
{ /* dump of the list for tests: per box sel[4], sup[4], total */
    if (out9)
        for (size_t i = 0; i < c->nbox; ++i) {
            for (int k = 0; k < 4; ++k) { out9[9 * i + k] = c->box[i].sel[k]; out9[9 * i + 4 + k] = c->box[i].sup[k]; }
            out9[9 * i + 8] = c->box[i].total;
        }
    return (int64_t)c->nbox;
}

/* The assembler's rounds with an UNLOCKED reference (spaced_seed.cpp:408-453): round r uses round_masks[r] (the reference
 * draws them with rand()); every remaining read runs the trial loop against the CURRENT text -- a success votes
 * (elect) and, when the whole reference view was consumed, grows the text (append / prepend), so later reads of the same
 * round see the longer text (ref_seq.h:266-276); found reads leave the pool; evolve() closes the round.
 * cons_out receives each round's consensus (cons_stride bytes apart, cons_len[r] chars); found_round[k] = round (1-based) in
 * which kept read k was found (0: never), recs[k] = the record of that success.  Returns the number of kept reads. */
int64_t pbo_assemble(const char *ref_text, size_t ref_len, int weight, const uint8_t *bin, size_t bin_bytes, int min_excl,
                     int max_excl, const uint32_t *round_masks, int nrounds, double R, int max_trial, int min_overlap,
                     int maxn, int maxm, int quirk, char *cons_out, size_t cons_stride, int32_t *cons_len,
                     int32_t *found_round, pbo_overlap_rec *recs)
{
    size_t cap = 1024, nk = 0;
    size_t *rec_off = (size_t *)malloc(cap * sizeof *rec_off);
    for (size_t p = 0; p + 4 <= bin_bytes;) {
        uint32_t l;
        memcpy(&l, bin + p, 4);
        if ((long)l > min_excl && (long)l < max_excl) {
            if (nk == cap) { cap *= 2; rec_off = (size_t *)realloc(rec_off, cap * sizeof *rec_off); }
            rec_off[nk++] = p;
        }
        p += 4 + ((size_t)l + 3) / 4;
    }
    if (!cons_out) { free(rec_off); return (int64_t)nk; }
    for (size_t k = 0; k < nk; ++k) { found_round[k] = 0; memset(&recs[k], 0, sizeof recs[k]); recs[k].id = (int32_t)k; }
    pbo_cons *c = pbo_cons_create(ref_text, ref_len, weight, 800000);
    dp_ws ws = {0, 0, 0};
    char *txt = (char *)malloc((size_t)max_excl + 64);
    size_t ecap = (size_t)max_excl * 2 + (size_t)maxm + 64;
    uint8_t *ops = (uint8_t *)malloc(ecap);
    char *vals = (char *)malloc(ecap);
    for (int r = 0; r < nrounds; ++r) {
        const uint32_t mask = round_masks[r];
        pbo_index *ix = pbo_index_build(c->txt + c->beg, (size_t)(c->end - c->beg), mask, 1); /* get_seedmap */
        for (size_t k = 0; k < nk; ++k) {
            if (found_round[k]) continue; /* erased from indices */
            const size_t rec = rec_off[k];
            uint32_t slen;
            memcpy(&slen, bin + rec, 4);
            pbo_bin2text(bin + rec, txt, (size_t)max_excl + 64);
            int found = 0;
            recs[k].ncand = 0; recs[k].cells = 0; /* per round: the record describes the round in which the read was found */
            for (int j = 0; j < max_trial && !found; ++j)
                for (int side = 0; side < 2 && !found; ++side) {
                    const int forward = side == 0;
                    const long pos = forward ? j : (long)slen - j - 16;
                    const uint32_t key = image_seed_at(bin, bin_bytes, rec, (int)pos, quirk) & mask;
                    const int32_t *plist;
                    const size_t cnt = pbo_index_find(ix, key, &plist);
                    if (!cnt) continue;
                    const long s_offset = forward ? pos : pos + 15;
                    const long s_len = forward ? (long)slen - s_offset : s_offset + 1;
                    if (s_len < min_overlap) continue;
                    for (size_t q = 0; q < cnt && !found; ++q) {
                        const long r_offset = forward ? plist[q] : plist[q] + 15;
                        /* get_accessor (ref_seq.h:282-286) on the text as it is NOW */
                        const long r_len = forward ? c->post - c->beg - r_offset : r_offset + c->beg - c->pre + 1;
                        pbo_align_out ao;
                        recs[k].ncand++;
                        const int ret = align_ws(&ws, c->txt + c->beg + r_offset, (int)r_len, forward ? 1 : -1, txt + s_offset, (int)s_len,
                                                 forward ? 1 : -1, R, maxn, maxm, &ao, ops, vals, ecap);
                        recs[k].cells += ao.cells;
                        if (ret < 0 || ao.matlen_a < min_overlap) continue;
                        found = 1;
                        pbo_cons_elect(c, (int)r_offset, ops, vals, ao.nedit, forward); /* ref_seq.h:266 */
                        if (ao.matlen_a == r_len) { /* the read runs past the reference: grow (ref_seq.h:267-275) */
                            const int add_len = (int)s_len - ao.matlen_b;
                            if (forward) pbo_cons_append(c, txt + s_offset + ao.matlen_b, add_len);
                            else pbo_cons_prepend(c, txt, add_len); /* pt(length-1) of a backward accessor = the read's first base */
                        }
                        found_round[k] = r + 1;
                        recs[k].found = 1; recs[k].j = j; recs[k].ref_pos = plist[q]; recs[k].cost = ao.cost; recs[k].read_pos = (int32_t)pos;
                        recs[k].dir = forward ? 1 : -1; recs[k].matlen_a = ao.matlen_a; recs[k].matlen_b = ao.matlen_b; recs[k].nedit = ao.nedit;
                    }
                }
        }
        pbo_index_free(ix);
        pbo_cons_evolve(c);
        const size_t n = (size_t)(c->end - c->beg);
        cons_len[r] = (int32_t)n;
        memcpy(cons_out + (size_t)r * cons_stride, c->txt + c->beg, n < cons_stride ? n : cons_stride);
    }
    free(txt); free(ops); free(vals); free(rec_off);
    dp_ws_free(&ws);
    pbo_cons_free(c);
    return (int64_t)nk;
}
