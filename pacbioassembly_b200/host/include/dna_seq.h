// dna_seq.h -- `dna_seq` statics and `seq_accessor` with the reference's interface (src/dna_seq.h), every
// codec call executed by the GPU library (batch of one through the C ABI; the batched entry points are what
// production callers use).
#pragma once

#include <assert.h>
#include <string.h>

#include "common.h"
#include "pb_runtime.hpp"

#define C2I(x) ((x == 'A') ? 0 : ((x == 'C') ? 1 : (x == 'G' ? 2 : 3))) // dna_seq.h:21
#define I2C(x) ((x == 0) ? 'A' : ((x == 1) ? 'C' : (x == 2 ? 'G' : 'T'))) // dna_seq.h:23
#define N_SEQ_WORD 16
#define N_SEQ_BYTE 4

class dna_seq {
public:
    // dna_seq.h:62-76.  Canonical value = encode(text+pos); define PB_SEED_AT_QUIRK to get the reference's
    // pos%4==0 branch, which reads the word at byte offset pos (SURVEY Q-S1).
    static t_seed seed_at(unsigned char *pbin, int pos)
    {
        unsigned len;
        memcpy(&len, pbin, 4);
        uint32_t out = 0;
        int32_t p = pos;
#ifdef PB_SEED_AT_QUIRK
        const int quirk = 1;
#else
        const int quirk = 0;
#endif
        // bytes past the record read as 0 in both modes (the reference reads whatever follows in memory)
        size_t rec_bytes = 4 + ((size_t)len + 3) / 4;
        pb::check(pb_seed_at_batch(pb::ctx(), pbin, rec_bytes, &p, 1, quirk, &out), "pb_seed_at_batch");
        return out;
    }
    static char value_at(unsigned char bv, int idx)
    { // dna_seq.h:78-80 (a table lookup, no arithmetic worth a kernel)
        static const char codes[4] = {'A', 'C', 'G', 'T'};
        return codes[(bv >> ((~idx & 0x3) << 1)) & 0x3];
    }
    // dna_seq.h:86-96: ptext must have 16 readable chars, like the reference
    static unsigned encode(const char *ptext)
    {
        uint32_t out = 0;
        int64_t off = 0;
        pb::check(pb_encode_batch(pb::ctx(), ptext, 16, &off, 1, &out), "pb_encode_batch");
        return out;
    }
    static void decode(unsigned code, char *ptext)
    { // dna_seq.h:101-107
        uint32_t c = code;
        pb::check(pb_decode_batch(pb::ctx(), &c, 1, ptext), "pb_decode_batch");
    }
    static unsigned text2bin(const char *ptext, unsigned char *pbin, unsigned buflen)
    { // dna_seq.h:113-127
        size_t tlen = strlen(ptext), written = 0;
        assert(buflen >= 4 + (tlen + 3) / 4);
        pb::check(pb_text2bin(pb::ctx(), ptext, tlen, pbin, buflen, &written), "pb_text2bin");
        return (unsigned)written;
    }
    static unsigned bin2text(const unsigned char *pbin, char *ptext, unsigned buflen)
    { // dna_seq.h:133-145
        size_t tlen = 0;
        unsigned l;
        memcpy(&l, pbin, 4);
        assert(buflen > l);
        pb::check(pb_bin2text(pb::ctx(), pbin, ptext, buflen, &tlen), "pb_bin2text");
        return (unsigned)tlen;
    }
};

// Directional, non-owning view over a text sequence (dna_seq.h:185-233).  Pure host pointer arithmetic: there is
// nothing to compute, the view is what gets handed to the aligner as (pointer, stride, length).
class seq_accessor {
public:
    seq_accessor(char *p, bool f, int l) : pdna(p), pcur(p), len(l), cnt(0), forward(f) {}
    int length() { return len; }
    bool is_forward() { return forward; }
    bool has_more() { return cnt < len; }
    char next() { ++cnt; return forward ? *pcur++ : *pcur--; }
    void reset(int pos) { cnt = pos; pcur = forward ? pdna + pos : pdna - pos; }
    char at(int i) { return forward ? *(pdna + i) : *(pdna - i); }
    char *pt(int i) { return forward ? (pdna + i) : (pdna - i); }
private:
    char *pdna;
    char *pcur;
    int len;
    int cnt;
    bool forward;
};
