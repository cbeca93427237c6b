/*
 * ref_shim.cpp -- C entry points into the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).
 *
 * This file contains no reference code.  It #includes the reference's own sources from
 * where they lie (-I/root/reference/src, see oracle/Makefile) and exposes them through a
 * plain C ABI so that tests/ can pin oracle/pb_oracle.c (the restatement) and the CUDA path
 * against the real thing, and so that bench.py --impl reference can time the real thing.
 * Output: oracle/_ref/libpbref.so (git-ignored, travels to the GPU box prebuilt).
 *
 * Only the *drivers* are restated here, because the shipped ones cannot run the configured
 * sizes (locator.cpp:26-27 uses fixed char[800000] buffers; MAX_SEQ_LEN is an unconditional
 * #define).  All arithmetic -- encode/seed_at/text2bin/parse_pattern/get_seedmap/align -- is
 * executed by the reference's code.
 *
 * Fresh-state shim (SURVEY Q-D2): seq_aligner.h:185 reads cost(i,i) for i > len_b, a cell the
 * current call never writes.  Before every align() we zero exactly those cells, which is what a
 * freshly constructed aligner would hold, so results do not depend on call history.
 */
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <stdint.h>
#include <pthread.h>
#include <new>
#include <vector>

/* spaced_seed.cpp is a program; rename its main so parse_pattern() (spaced_seed.cpp:166-180)
 * can be called as compiled from the reference source. */
#define main pb_ref_spaced_seed_main
#include "spaced_seed.cpp"
#undef main

typedef seq_aligner<40000, 6000> loc_aligner; /* locator.cpp:23-24,68 */

namespace {

template <class A> A *fresh_aligner(double R)
{
    /* calloc: zero pages are mapped lazily, so the 1.25 / 1.92 GB object costs only what is touched */
    void *mem = calloc(1, sizeof(A));
    if (!mem) return NULL;
    A *al = reinterpret_cast<A *>(mem);
    al->R = R;
    return al;
}

struct align_out { /* mirrors pbo_align_out up to nedit (oracle/pb_oracle.h) */
    int32_t ret, len_a, len_b, max_dst, matlen_a, matlen_b, cost, diag_cost, nedit, fail_row;
    int64_t cells;
};

template <class A, int MAXN, int MAXM>
int do_align(A *al, char *a, int a_len, int a_fwd, char *b, int b_len, int b_fwd, double R, align_out *out,
             uint8_t *ops, char *vals, size_t cap)
{
    memset(out, 0, sizeof *out);
    out->ret = -1;
    al->R = R;
    /* same derivation as seq_aligner.h:94-102, only to know which cells to pre-zero */
    int len_a, len_b, max_dst;
    if (b_len >= a_len) {
        len_a = a_len;
        max_dst = 1 + (int)(len_a * R);
        len_b = std::min(b_len, len_a + max_dst);
    } else {
        len_b = b_len;
        max_dst = 1 + (int)(len_b * R);
        len_a = std::min(a_len, len_b + max_dst);
    }
    out->len_a = len_a; out->len_b = len_b; out->max_dst = max_dst;
    if (len_a >= MAXN || max_dst >= MAXM) return -1; /* domain, SURVEY Q-D3 */
    al->max_dst = max_dst;
    for (int i = len_b + 1; i <= len_a; ++i) al->set_cost(i, i, 0); /* Q-D2 */

    seq_accessor ac_a(a, a_fwd != 0, a_len), ac_b(b, b_fwd != 0, b_len);
    int ret = al->align(&ac_a, &ac_b);
    out->ret = ret;
    if (ret < 0) return ret;
    out->len_a = al->len_a; out->len_b = al->len_b; out->max_dst = al->max_dst;
    out->matlen_a = al->matlen_a; out->matlen_b = al->matlen_b;
    out->cost = al->final_cost();
    out->diag_cost = (a_len <= al->len_a && a_len <= al->len_b) ? al->get_cost(a_len, a_len) : 0; /* Q-L2 */
    out->nedit = al->nedit;
    if (ops && (size_t)al->nedit <= cap)
        for (int k = 0; k < al->nedit; ++k) {
            ops[k] = (uint8_t)al->edits[k].op;
            if (vals) vals[k] = al->edits[k].op == DELETE ? 0 : al->edits[k].val;
        }
    return ret;
}

t_aligner *g_tal = NULL;
loc_aligner *g_lal = NULL;
hash_table *g_map = NULL;

} // namespace

extern "C" {

unsigned pbref_encode(const char *text16) { return dna_seq::encode(text16); }
void pbref_decode(unsigned code, char *out16) { dna_seq::decode(code, out16); }
unsigned pbref_text2bin(const char *text, unsigned char *out, unsigned cap) { return dna_seq::text2bin(text, out, cap); }
unsigned pbref_bin2text(const unsigned char *rec, char *out, unsigned cap) { return dna_seq::bin2text(rec, out, cap); }
unsigned pbref_seed_at(unsigned char *rec, int pos) { return dna_seq::seed_at(rec, pos); }
unsigned pbref_parse_pattern(const char *pat) { return parse_pattern(pat); }
int pbref_c2i(int ch) { char x = (char)ch; return C2I(x); }

/* which: 0 = t_aligner (seq_aligner<26000,6000>), 1 = locator's seq_aligner<40000,6000> */
int pbref_align(int which, char *a, int a_len, int a_fwd, char *b, int b_len, int b_fwd, double R, align_out *out,
                uint8_t *ops, char *vals, size_t cap)
{
    if (which == 0) {
        if (!g_tal) g_tal = fresh_aligner<t_aligner>(R);
        return do_align<t_aligner, MAX_READ_LEN + MAX_DIFF_LEN, MAX_DIFF_LEN>(g_tal, a, a_len, a_fwd, b, b_len, b_fwd, R,
                                                                            out, ops, vals, cap);
    }
    if (!g_lal) g_lal = fresh_aligner<loc_aligner>(R);
    return do_align<loc_aligner, 40000, 6000>(g_lal, a, a_len, a_fwd, b, b_len, b_fwd, R, out, ops, vals, cap);
}

/* ---- seed index ---------------------------------------------------------- */

/* policy 0: locator.cpp:62-66 over a heap copy of `ref` (16 NUL bytes appended, which is what the
 * shipped global zero-initialised buffer holds past the string).  policy 1: ref_seq::get_seedmap
 * (ref_seq.h:291-311; len must be < MAX_SEQ_LEN).  Returns number of keys (hash_table::size()). */
long pbref_index_build(const char *ref, long len, unsigned mask, int policy)
{
    if (!g_map) g_map = new hash_table(1 << 20);
    g_map->clear();
    if (policy == 0) {
        char *contig = (char *)calloc((size_t)len + 32, 1);
        memcpy(contig, ref, (size_t)len);
        for (long i = 0; i < len; ++i) {
            int sd = dna_seq::encode(contig + i);
            if (sd & mask) (*g_map)[sd & mask].push_back((int)i);
        }
        free(contig);
    } else {
        if (len >= MAX_SEQ_LEN) return -1;
        ref_seq *pr = new ref_seq(ref, (int)len, true);
        pr->get_seedmap(*g_map, mask);
        delete pr;
    }
    return (long)g_map->size();
}

/* hash_table::find; copies up to cap positions in list order, returns the list length (0 = end()) */
long pbref_index_find(unsigned key, int *out, long cap)
{
    if (!g_map) return 0;
    sm_it it = g_map->find(key);
    if (it == g_map->end()) return 0;
    long n = 0;
    for (std::list<int>::iterator p = it->second.begin(); p != it->second.end(); ++p, ++n)
        if (n < cap) out[n] = *p;
    return n;
}

/* ---- locate loop (locator.cpp:57-92 restated over heap buffers) ----------- */

struct locate_rec { /* mirrors pbo_locate_rec */
    int32_t nseq, found, j, pos, cost, seg_len, diag_cost, matlen_a, matlen_b, nedit, ncand;
    int64_t cells;
};

struct locate_job {
    const hash_table *map; char *contig; long contig_len;
    const char *reads; const int64_t *offs; const int32_t *lens; const int64_t *kept;
    int64_t *next, nk; unsigned mask; double R; int ntrial; /* threads take the next kept read from a shared counter */
    locate_rec *recs; uint8_t *ops_out; const int64_t *ops_off;
};

static void *locate_thread(void *arg)
{
    locate_job *jb = (locate_job *)arg;
    loc_aligner *al = fresh_aligner<loc_aligner>(jb->R);
    char *sequence = (char *)malloc(40000 + 64);
    for (;;) {
        const int64_t k = __atomic_fetch_add(jb->next, 1, __ATOMIC_RELAXED);
        if (k >= jb->nk) break;
        int64_t r = jb->kept[k];
        int len = jb->lens[r];
        locate_rec *rec = &jb->recs[k];
        memset(rec, 0, sizeof *rec);
        rec->nseq = (int32_t)k;
        if (len >= 40000) continue; /* outside the aligner's domain: every align() returns -1 */
        memcpy(sequence, jb->reads + jb->offs[r], (size_t)len);
        memset(sequence + len, 0, 32);
        bool found = false;
        for (int j = 0; j < jb->ntrial && !found; ++j) { /* locator.cpp:74 */
            int seed = dna_seq::encode(sequence + j) & jb->mask;
            hash_table::const_iterator sit = jb->map->find(seed);
            if (sit == jb->map->end()) continue;
            for (std::list<int>::const_iterator it = sit->second.begin(); it != sit->second.end(); ++it) {
                align_out ao;
                rec->ncand++;
                int ret = do_align<loc_aligner, 40000, 6000>(al, sequence + j, len - j, 1, jb->contig + *it,
                                                             (int)(jb->contig_len - *it), 1, jb->R, &ao, NULL, NULL, 0);
                if (ret > 0) { /* locator.cpp:82-88 */
                    found = true;
                    rec->found = 1; rec->j = j; rec->pos = *it; rec->cost = ao.cost; rec->seg_len = len - j;
                    rec->diag_cost = ao.diag_cost; rec->matlen_a = ao.matlen_a; rec->matlen_b = ao.matlen_b;
                    rec->nedit = ao.nedit;
                    if (jb->ops_out)
                        for (int e = 0; e < al->nedit; ++e) jb->ops_out[jb->ops_off[k] + e] = (uint8_t)al->edits[e].op;
                    break;
                }
            }
        }
    }
    free(sequence);
    free(al);
    return NULL;
}

/* The locator's setup (locator.cpp:57-66): contig copy + seed map.  Built once, reused by pbref_locator_run, the way
 * the reference program builds its map once and then streams reads. */
struct locator_state { char *contig; long len; hash_table *map; unsigned mask; };

void *pbref_locator_open(const char *ref, long ref_len, unsigned mask)
{
    locator_state *st = new locator_state();
    st->contig = (char *)calloc((size_t)ref_len + 32, 1);
    memcpy(st->contig, ref, (size_t)ref_len);
    st->len = ref_len;
    st->mask = mask;
    st->map = new hash_table(1 << 23); /* locator.cpp:28 */
    for (long i = 0; i < ref_len; ++i) { /* locator.cpp:62-66 */
        int sd = dna_seq::encode(st->contig + i);
        if (sd & mask) (*st->map)[sd & mask].push_back((int)i);
    }
    return st;
}

void pbref_locator_close(void *h)
{
    locator_state *st = (locator_state *)h;
    if (!st) return;
    free(st->contig);
    delete st->map;
    delete st;
}

/* Same contract as pbo_locate (oracle/pb_oracle.h) with the locator's seq_aligner<40000,6000>. */
int64_t pbref_locator_run(void *h, const char *reads, const int64_t *offs, const int32_t *lens, int64_t nreads, double R,
                          int ntrial, int minlen, int nthreads, locate_rec *recs, uint8_t *ops_out, const int64_t *ops_off)
{
    locator_state *st = (locator_state *)h;
    int64_t *kept = (int64_t *)malloc((size_t)(nreads + 1) * sizeof *kept);
    int64_t nk = 0;
    for (int64_t r = 0; r < nreads; ++r)
        if (lens[r] >= minlen) kept[nk++] = r; /* locator.cpp:72 */
    if (nthreads < 1) nthreads = 1;
    if ((int64_t)nthreads > nk) nthreads = nk > 0 ? (int)nk : 1;
    locate_job *jobs = (locate_job *)calloc((size_t)nthreads, sizeof *jobs);
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof *th);
    int64_t next = 0; /* dynamic schedule: read lengths vary 40-fold, a static split leaves threads idle */
    for (int t = 0; t < nthreads; ++t) {
        locate_job *jb = &jobs[t];
        jb->map = st->map; jb->contig = st->contig; jb->contig_len = st->len;
        jb->reads = reads; jb->offs = offs; jb->lens = lens; jb->kept = kept;
        jb->next = &next; jb->nk = nk;
        jb->mask = st->mask; jb->R = R; jb->ntrial = ntrial;
        jb->recs = recs; jb->ops_out = ops_out; jb->ops_off = ops_off;
        if (nthreads == 1) locate_thread(jb);
        else pthread_create(&th[t], NULL, locate_thread, jb);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(jobs); free(th); free(kept);
    return nk;
}

int64_t pbref_locate(const char *ref, long ref_len, const char *reads, const int64_t *offs, const int32_t *lens,
                     int64_t nreads, unsigned mask, double R, int ntrial, int minlen, int nthreads,
                     locate_rec *recs, uint8_t *ops_out, const int64_t *ops_off)
{
    void *h = pbref_locator_open(ref, ref_len, mask);
    int64_t nk = pbref_locator_run(h, reads, offs, lens, nreads, R, ntrial, minlen, nthreads, recs, ops_out, ops_off);
    pbref_locator_close(h);
    return nk;
}

/* ---- assembler-side trial loop driven through the reference's OWN code ----------------------------------------
 * ref_seq (locked: no voting / growth), ref_seq::get_seedmap, dna_seq::seed_at, dna_seq::bin2text, ref_seq::try_align
 * and seq_aligner::align are the reference's; only the two loops of spaced_seed.cpp:424-436 and :261-299 are
 * restated, so that the fresh-state shim (Q-D2) can be applied before every align. */
struct overlap_rec {
    int32_t id, found, j, ref_pos, cost, read_pos, dir, matlen_a, matlen_b, nedit, ncand, _pad;
    int64_t cells;
};

int64_t pbref_overlap(const char *ref_text, long ref_len, unsigned char *bin, long bin_bytes, int min_excl, int max_excl,
                      unsigned mask, double R, int max_trial, overlap_rec *recs)
{
    if (ref_len >= MAX_SEQ_LEN) return -1;
    ref_seq *pr = new ref_seq(ref_text, (int)ref_len, true); /* locked */
    hash_table *map = new hash_table(1 << 20);
    pr->get_seedmap(*map, mask); /* ref_seq.h:291-311 */
    t_aligner *al = fresh_aligner<t_aligner>(R);
    char *txt = (char *)malloc(MAX_READ_LEN + 64);
    int64_t k = 0;
    for (long off = 0; off + 4 <= bin_bytes;) { /* open_binary, spaced_seed.cpp:330-342 */
        unsigned slen = *((unsigned *)(bin + off));
        if (slen > (unsigned)min_excl && slen < (unsigned)max_excl) {
            overlap_rec *out = &recs[k];
            memset(out, 0, sizeof *out);
            out->id = (int32_t)k;
            dna_seq::bin2text(bin + off, txt, slen + 1); /* set_active_seg */
            bool found = false;
            for (int j = 0; j < max_trial && !found; ++j)
                for (int side = 0; side < 2 && !found; ++side) {
                    bool forward = side == 0;
                    long pos = forward ? j : (long)slen - j - 16;
                    sm_it sit = map->find(dna_seq::seed_at(bin + off, (int)pos) & mask); /* spaced_seed.cpp:265 */
                    if (sit == map->end()) continue;
                    int s_offset = forward ? pos : pos + 16 - 1;
                    int s_len = forward ? (int)slen - s_offset : s_offset + 1;
                    seq_accessor ac_seg(txt + s_offset, forward, s_len);
                    if (s_len < OVERLAP_MIN) continue;
                    for (std::list<int>::iterator it = sit->second.begin(); it != sit->second.end() && !found; ++it) {
                        int r_offset = forward ? (*it) : (*it) + 16 - 1;
                        /* fresh-state shim: zero the cells align() will read without writing (Q-D2) */
                        seq_accessor ac_ref = pr->get_accessor(r_offset, forward);
                        int la, lb, md;
                        if (ac_seg.length() >= ac_ref.length()) { la = ac_ref.length(); md = 1 + (int)(la * R); lb = std::min(ac_seg.length(), la + md); }
                        else { lb = ac_seg.length(); md = 1 + (int)(lb * R); la = std::min(ac_ref.length(), lb + md); }
                        out->ncand++;
                        if (la < MAX_READ_LEN + MAX_DIFF_LEN && md < MAX_DIFF_LEN) {
                            al->max_dst = md;
                            for (int i = lb + 1; i <= la; ++i) al->set_cost(i, i, 0);
                        }
                        al->R = R;
                        if (pr->try_align(al, r_offset, &ac_seg)) { /* ref_seq.h:259-266 */
                            found = true;
                            out->found = 1; out->j = j; out->ref_pos = *it; out->cost = al->final_cost(); out->read_pos = (int32_t)pos;
                            out->dir = forward ? 1 : -1; out->matlen_a = al->matlen_a; out->matlen_b = al->matlen_b; out->nedit = al->nedit;
                        }
                    }
                }
            ++k;
        }
        off += 4 + ((long)slen + 3) / 4;
    }
    free(txt); free(al);
    delete map; delete pr;
    return k;
}

/* ---- unlocked assembler rounds through the reference's OWN ref_seq (voting, growth, evolve) --------------------------
 * ref_seq::try_align (elect / append / prepend, ref_seq.h:259-276), ref_seq::evolve (:317-348), get_seedmap, seed_at, bin2text
 * and align are the reference's; the loops of spaced_seed.cpp:408-453 and :261-299 are restated (fresh-state shim before every
 * align, seeds given per round instead of rand()).  Same outputs as pbo_assemble. */
int64_t pbref_assemble(const char *ref_text, long ref_len, int weight, unsigned char *bin, long bin_bytes, int min_excl, int max_excl,
                       const unsigned *round_masks, int nrounds, double R, int max_trial, char *cons_out, long cons_stride,
                       int32_t *cons_len, int32_t *found_round, overlap_rec *recs)
{
    if (ref_len >= MAX_SEQ_LEN) return -1;
    std::vector<long> rec_off;
    for (long off = 0; off + 4 <= bin_bytes;) {
        unsigned slen = *((unsigned *)(bin + off));
        if (slen > (unsigned)min_excl && slen < (unsigned)max_excl) rec_off.push_back(off);
        off += 4 + ((long)slen + 3) / 4;
    }
    const int64_t nk = (int64_t)rec_off.size();
    if (!cons_out) return nk;
    for (int64_t k = 0; k < nk; ++k) { found_round[k] = 0; memset(&recs[k], 0, sizeof recs[k]); recs[k].id = (int32_t)k; }
    ref_seq *pr = new ref_seq(ref_text, (int)ref_len, false, weight); /* unlocked */
    hash_table *map = new hash_table(1 << 20);
    t_aligner *al = fresh_aligner<t_aligner>(R);
    char *txt = (char *)malloc(MAX_READ_LEN + 64);
    for (int r = 0; r < nrounds; ++r) {
        const unsigned mask = round_masks[r];
        pr->get_seedmap(*map, mask);
        for (int64_t k = 0; k < nk; ++k) {
            if (found_round[k]) continue;
            const long off = rec_off[k];
            unsigned slen = *((unsigned *)(bin + off));
            dna_seq::bin2text(bin + off, txt, slen + 1);
            overlap_rec *out = &recs[k];
            bool found = false;
            out->ncand = 0; /* per round */
            for (int j = 0; j < max_trial && !found; ++j)
                for (int side = 0; side < 2 && !found; ++side) {
                    bool forward = side == 0;
                    long pos = forward ? j : (long)slen - j - 16;
                    sm_it sit = map->find(dna_seq::seed_at(bin + off, (int)pos) & mask);
                    if (sit == map->end()) continue;
                    int s_offset = forward ? pos : pos + 16 - 1;
                    int s_len = forward ? (int)slen - s_offset : s_offset + 1;
                    seq_accessor ac_seg(txt + s_offset, forward, s_len);
                    if (s_len < OVERLAP_MIN) continue;
                    for (std::list<int>::iterator it = sit->second.begin(); it != sit->second.end() && !found; ++it) {
                        int r_offset = forward ? (*it) : (*it) + 16 - 1;
                        seq_accessor ac_ref = pr->get_accessor(r_offset, forward);
                        int la, lb, md;
                        if (ac_seg.length() >= ac_ref.length()) { la = ac_ref.length(); md = 1 + (int)(la * R); lb = std::min(ac_seg.length(), la + md); }
                        else { lb = ac_seg.length(); md = 1 + (int)(lb * R); la = std::min(ac_ref.length(), lb + md); }
                        out->ncand++;
                        if (la < MAX_READ_LEN + MAX_DIFF_LEN && md < MAX_DIFF_LEN) {
                            al->max_dst = md;
                            for (int i = lb + 1; i <= la; ++i) al->set_cost(i, i, 0);
                        }
                        al->R = R;
                        if (pr->try_align(al, r_offset, &ac_seg)) { /* votes and grows */
                            found = true;
                            found_round[k] = r + 1;
                            out->found = 1; out->j = j; out->ref_pos = *it; out->cost = al->final_cost(); out->read_pos = (int32_t)pos;
                            out->dir = forward ? 1 : -1; out->matlen_a = al->matlen_a; out->matlen_b = al->matlen_b; out->nedit = al->nedit;
                        }
                    }
                }
        }
        pr->evolve();
        const int n = (int)pr->length();
        cons_len[r] = n;
        seq_accessor ac = pr->get_accessor(0, true);
        for (int i = 0; i < n && i < cons_stride; ++i) cons_out[(long)r * cons_stride + i] = ac.next();
    }
    free(txt); free(al);
    delete map; delete pr;
    return nk;
}

} /* extern "C" */
