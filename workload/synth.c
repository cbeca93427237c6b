/*
 * synth.c -- deterministic synthetic PacBio-like workload generator (bench/test input only).
 *
 * Not part of the product path and not part of the oracle: it only manufactures inputs
 * (iid-uniform references; CLR-like reads with an indel-heavy error
 * model, forward strand only because the reference never reverse-complements).
 * Everything is driven by splitmix64 so the same (seed, index) gives the same bytes on any
 * platform; no rand()/libstdc++ distributions are used.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t mix(uint64_t seed, uint64_t idx)
{
    uint64_t s = seed * 0xD1342543DE82EF95ull + idx * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    return splitmix64(&s) ^ s;
}
static inline double u01(uint64_t *s) { return (double)(splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }

static const char ACGT[4] = {'A', 'C', 'G', 'T'};

/* iid-uniform ACGT; base i depends only on (seed, i/32) so generation can be chunked */
void pbs_reference(uint64_t seed, int64_t len, char *out)
{
    for (int64_t w = 0; w * 32 < len; ++w) {
        uint64_t bits = mix(seed, (uint64_t)w);
        int64_t n = len - w * 32 < 32 ? len - w * 32 : 32;
        for (int64_t k = 0; k < n; ++k) out[w * 32 + k] = ACGT[(bits >> (2 * k)) & 3];
    }
}

/* read lengths: log-normal with the given arithmetic mean and sigma_log, clipped to [lo,hi];
 * if sigma_log <= 0, uniform in [lo,hi]. */
void pbs_read_lengths(uint64_t seed, int64_t nreads, double mean, double sigma_log, int lo, int hi, int32_t *lens)
{
    double mu = log(mean) - 0.5 * sigma_log * sigma_log;
    for (int64_t r = 0; r < nreads; ++r) {
        uint64_t s = mix(seed ^ 0xA5A5A5A5ull, (uint64_t)r);
        double L;
        if (sigma_log > 0) {
            double u1 = u01(&s), u2 = u01(&s);
            if (u1 < 1e-300) u1 = 1e-300;
            double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
            L = exp(mu + sigma_log * z);
        } else {
            L = lo + u01(&s) * (double)(hi - lo + 1);
        }
        int32_t v = (int32_t)L;
        if (v < lo) v = lo;
        if (v > hi) v = hi;
        lens[r] = v;
    }
}

typedef struct {
    uint64_t seed; const char *ref; int64_t ref_len; int64_t r0, r1;
    const int32_t *lens; const int64_t *offs; char *out; int64_t *starts;
    double p_ins, p_del, p_sub;
} job_t;

static void gen_read(const job_t *jb, int64_t r)
{
    uint64_t s = mix(jb->seed, (uint64_t)r);
    int32_t len = jb->lens[r];
    char *o = jb->out + jb->offs[r];
    /* leave room so that the read (almost always) fits inside the reference */
    int64_t span = (int64_t)((double)len * 1.25) + 64;
    int64_t maxstart = jb->ref_len - span;
    if (maxstart < 1) maxstart = 1;
    int64_t p = (int64_t)(splitmix64(&s) % (uint64_t)maxstart);
    if (jb->starts) jb->starts[r] = p;
    int32_t n = 0;
    while (n < len) {
        if (p >= jb->ref_len) { o[n++] = ACGT[splitmix64(&s) & 3]; continue; }
        double u = u01(&s);
        if (u < jb->p_ins) { o[n++] = ACGT[splitmix64(&s) & 3]; continue; } /* inserted base, source not consumed */
        u -= jb->p_ins;
        char c = jb->ref[p++];
        if (u < jb->p_del) continue; /* source base dropped */
        u -= jb->p_del;
        if (u < jb->p_sub) {
            int k = (c == 'A') ? 0 : (c == 'C') ? 1 : (c == 'G') ? 2 : 3;
            c = ACGT[(k + 1 + (int)(splitmix64(&s) % 3)) & 3];
        }
        o[n++] = c;
    }
}

static void *gen_thread(void *arg)
{
    const job_t *jb = (const job_t *)arg;
    for (int64_t r = jb->r0; r < jb->r1; ++r) gen_read(jb, r);
    return NULL;
}

/* reads r in [0,nreads): text written at out+offs[r], lens[r] bases, no terminator.
 * starts (optional) receives the true sampling position. */
void pbs_reads(uint64_t seed, const char *ref, int64_t ref_len, int64_t nreads, const int32_t *lens,
               const int64_t *offs, double p_ins, double p_del, double p_sub, int nthreads, char *out,
               int64_t *starts)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    job_t jobs[256];
    pthread_t th[256];
    for (int t = 0; t < nthreads; ++t) {
        job_t *jb = &jobs[t];
        jb->seed = seed; jb->ref = ref; jb->ref_len = ref_len;
        jb->r0 = nreads * t / nthreads; jb->r1 = nreads * (t + 1) / nthreads;
        jb->lens = lens; jb->offs = offs; jb->out = out; jb->starts = starts;
        jb->p_ins = p_ins; jb->p_del = p_del; jb->p_sub = p_sub;
        if (nthreads == 1) gen_thread(jb);
        else pthread_create(&th[t], NULL, gen_thread, jb);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
}

/* DP-sweep pair: b = source of length blen; a = mutated copy of b[0:alen') with
 * edit k placed at source position >= ceil(2k/R)+12 (so that cost(i,i) <= i*R holds on every row and
 * the early-failure test of seq_aligner.h:185 never fires); edits alternate sub / ins / del. */
void pbs_sweep_pair(uint64_t seed, int64_t idx, int alen, int blen, double R, int nedits, char *a, char *b, int *alen_out)
{
    uint64_t s = mix(seed, (uint64_t)idx);
    for (int k = 0; k < blen; ++k) b[k] = ACGT[splitmix64(&s) & 3];
    int spacing = (int)ceil(2.0 / R) + 1;
    int n = 0, p = 0, k = 0, next_edit = 12 + spacing;
    while (n < alen && p < blen) {
        if (k < nedits && p >= next_edit) {
            int kind = k % 3;
            ++k;
            next_edit = p + spacing + (int)(splitmix64(&s) % (uint64_t)(spacing + 1));
            if (kind == 0) { /* substitution */
                char c = b[p++];
                int q = (c == 'A') ? 0 : (c == 'C') ? 1 : (c == 'G') ? 2 : 3;
                a[n++] = ACGT[(q + 1 + (int)(splitmix64(&s) % 3)) & 3];
            } else if (kind == 1) { /* insertion into a */
                a[n++] = ACGT[splitmix64(&s) & 3];
            } else { /* deletion from a */
                ++p;
            }
            continue;
        }
        a[n++] = b[p++];
    }
    *alen_out = n;
}

/* A read batch as the .bin image the assembler reads (binary_test.cpp:55-63: per record a native u32 length, then
 * ceil(len/4) bytes of 4 bases each, first base in bits 7:6, A=0 C=1 G=2 other=3; records back to back).  Returns the
 * image size; out == NULL only sizes it.  Bench / test INPUT tooling: the product's own packer is pb_text2bin. */
int64_t pbs_pack_bin(const char *text, const int64_t *offs, const int32_t *lens, int64_t n, uint8_t *out)
{
    int64_t p = 0;
    for (int64_t r = 0; r < n; ++r) {
        const int32_t l = lens[r];
        const int64_t body = ((int64_t)l + 3) / 4;
        if (out) {
            uint32_t l32 = (uint32_t)l;
            memcpy(out + p, &l32, 4);
            const char *t = text + offs[r];
            uint8_t *o = out + p + 4;
            for (int64_t b = 0; b < body; ++b) {
                uint8_t v = 0;
                for (int k = 0; k < 4; ++k) {
                    const int64_t i = 4 * b + k;
                    int c = 0;
                    if (i < l) { const char ch = t[i]; c = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3; }
                    v = (uint8_t)((v << 2) | c);
                }
                o[b] = v;
            }
        }
        p += 4 + body;
    }
    return p;
}
