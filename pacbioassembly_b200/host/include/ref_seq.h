// ref_seq.h -- `ref_seq`, `vote_box`, `base_vote` with the reference's public interface (src/ref_seq.h), so that code
// written against the reference's header -- test/ref_test.cpp, the assembler's round loop -- compiles against this directory.
//
// The voting state lives on the GPU (pb_consensus_*, csrc/pb_cons.cu): one vote box per base as a struct of arrays, votes cast
// by a kernel that walks the transcript, evolve() as count -> scan -> write -> absorb.  This class keeps what callers hold
// pointers into -- the text buffer that get_accessor() hands out views of (ref_seq.h:282-286) -- and forwards everything
// that computes: align (K3), elect (vote kernel), evolve, get_seedmap (K1 + index build).  `vote_box` / `base_vote` are the
// small host value classes test/ref_test.cpp exercises directly; the device boxes follow the same rules (pb_cons.cu).
#pragma once

#include <assert.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.h"
#include "dna_seq.h"
#include "seq_aligner.h"

// Votes of the reads overlapping one place, per base (ref_seq.h:47-99)
class base_vote {
public:
    base_vote() { reset(); }
    base_vote(char c, int n = 1) { reset(); add_char(c, n); }
    void add_char(char c) { ++n_[C2I(c)]; }
    void add_char(int c, int n) { n_[C2I(c)] = (unsigned short)(n_[C2I(c)] + n); }
    void add_code(int c) { ++n_[c]; }
    void reset() { n_[0] = n_[1] = n_[2] = n_[3] = 0; }
    void absorb(base_vote &other)
    { // take the other's votes; it is left empty
        for (int k = 0; k < 4; ++k) n_[k] = (unsigned short)(n_[k] + other.n_[k]);
        other.reset();
    }
    int max_vote() { return *std::max_element(n_, n_ + 4); }
    char winner()
    { // the first of A, C, G, T holding the maximum
        const int m = max_vote();
        for (int k = 0; k < 3; ++k)
            if (n_[k] == m) return I2C(k);
        return 'T';
    }
private:
    unsigned short n_[4];
};

// One place of the reference: the votes on its base and on a base to insert right behind it (ref_seq.h:107-183)
class vote_box {
public:
    vote_box() : total(0) {}
    vote_box(char c, int n = 1) : selection(c, n), total(1) {}
    base_vote selection;  // votes on the base at this place
    base_vote suppliment; // votes on a base to insert behind it
    int total;            // reads that covered this place
    void select(char c) { selection.add_char(c); ++total; } // MATCH (or mismatch): the read's base
    void ignore() { ++total; }                               // DELETE: the read skips this base
    void supply(char c) { suppliment.add_char(c); }          // INSERT: the read has a base behind this one
    void split(vote_box *other)
    { // the suppliment becomes a box of its own
        other->selection = suppliment;
        other->total = total;
        suppliment.reset();
    }
    bool is_valid(double ratio) { return selection.max_vote() > ratio * total; }
    bool has_supply(double ratio) { return suppliment.max_vote() > ratio * total; }
    char get_vote() { return selection.winner(); }
    char get_supply() { return suppliment.winner(); }
};

class ref_seq {
public:
    // from a packed record (ref_seq.h:207-213): every base starts with one vote
    ref_seq(const t_bseq *pseq, bool lk = false) : locked(lk), txt_((size_t)3 * MAX_SEQ_LEN + 16)
    {
        beg = pre = MAX_SEQ_LEN;
        end = post = beg + (int)dna_seq::bin2text(pseq, &txt_[beg], MAX_SEQ_LEN);
        create(1);
    }
    // from text with an initial weight (ref_seq.h:218-225)
    ref_seq(const char *ptxt, int len, bool l, int w = 1) : locked(l), txt_((size_t)3 * MAX_SEQ_LEN + 16)
    {
        beg = pre = MAX_SEQ_LEN;
        end = post = beg + len;
        memcpy(&txt_[beg], ptxt, (size_t)len);
        create(w);
    }
    ~ref_seq() { if (cons_) pb_consensus_free(cons_); }
    ref_seq(const ref_seq &) = delete;
    ref_seq &operator=(const ref_seq &) = delete;

    void append(char *pseg, int len)
    { // ref_seq.h:227-233
        memmove(&txt_[post], pseg, (size_t)len);
        pb::check(pb_consensus_append(pb::ctx(), cons_, &txt_[post], len), "pb_consensus_append");
        post += len;
    }
    void prepend(char *pseg, int len)
    { // ref_seq.h:235-243
        pre -= len;
        memmove(&txt_[pre], pseg, (size_t)len);
        pb::check(pb_consensus_prepend(pb::ctx(), cons_, &txt_[pre], len), "pb_consensus_prepend");
    }
    bool contained(int pos) { return pos + beg >= pre && pos + beg < post; }
    unsigned length() { return (unsigned)(end - beg); }

    // ref_seq.h:259-277: align(reference view, segment) -- mind the argument order -- then the OVERLAP_MIN gate, the votes and,
    // when the whole reference view was consumed, the growth by what is left of the segment
    bool try_align(t_aligner *paligner, int pos, seq_accessor *pac_seg)
    {
        const bool forward = pac_seg->is_forward();
        seq_accessor ac_ref = get_accessor(pos, forward);
        if (paligner->align(&ac_ref, pac_seg) < 0) return false;
        if (paligner->matlen_a < OVERLAP_MIN) return false;
        if (locked) return true;
        elect(pos, &paligner->edits[0], paligner->nedit, forward);
        if (paligner->matlen_a == ac_ref.length()) {
            const int rest = pac_seg->length() - paligner->matlen_b;
            if (forward) append(pac_seg->pt(paligner->matlen_b), rest);
            else prepend(pac_seg->pt(pac_seg->length() - 1), rest);
        }
        return true;
    }
    seq_accessor get_accessor(int pos, bool forward)
    { // ref_seq.h:282-286: forward views run to post, backward views down to pre
        assert(contained(pos));
        return seq_accessor(&txt_[beg + pos], forward, forward ? post - beg - pos : pos + beg - pre + 1);
    }
    // ref_seq.h:291-311: head / tail windows of [beg, end) under the mask; one GPU pass (K1 + index build)
    unsigned get_seedmap(hash_table &seedmap, t_seed sd_pat) { return seedmap.build_refseq(&txt_[beg], (size_t)(end - beg), sd_pat); }

    void evolve()
    { // ref_seq.h:317-348 on the device; the host text is refreshed from the winners
        if (locked) return;
        pb::check(pb_consensus_evolve(pb::ctx(), cons_), "pb_consensus_evolve");
        const int n = (int)pb_consensus_length(cons_);
        end = pre = beg = MAX_SEQ_LEN;
        pb::check(pb_consensus_text(pb::ctx(), cons_, 0, &txt_[beg], (size_t)n + 1), "pb_consensus_text");
        end = post = beg + n;
    }
    // ref_seq.h:351-361 + apply_edits (:25-41): one transcript's votes.  edit.val carries the segment's elements, so the
    // segment is rebuilt from the transcript and voted by the same kernel the batched path uses.
    void elect(int pos, edit *pedit, int nedit, bool forward)
    {
        std::vector<char> seg;
        std::vector<uint8_t> ops((size_t)nedit + 1);
        int na = 0;
        for (int k = 0; k < nedit; ++k) {
            ops[k] = (uint8_t)pedit[k].op;
            if (pedit[k].op != DELETE) seg.push_back(pedit[k].val);
            if (pedit[k].op != INSERT) ++na;
        }
        const int nb = (int)seg.size();
        if (!forward) std::reverse(seg.begin(), seg.end()); // memory order of a backward view
        if (seg.empty()) seg.push_back('A');
        const int64_t off = 0, ops_off = 0;
        const int32_t len = (int32_t)seg.size();
        pb_seqset *s = nullptr;
        pb::check(pb_seqset_from_text(pb::ctx(), seg.data(), &off, &len, nullptr, 1, &s), "pb_seqset_from_text");
        pb_overlap_rec rec;
        memset(&rec, 0, sizeof rec);
        rec.id = 0; rec.found = 1; rec.dir = forward ? 1 : -1;
        // the vote kernel places a backward view 15 elements behind the seed position it was found at (spaced_seed.cpp:275-276)
        rec.ref_pos = forward ? pos : pos - 15;
        rec.read_pos = forward ? 0 : nb - 16;
        rec.matlen_a = na; rec.matlen_b = nb; rec.nedit = nedit;
        pb::check(pb_consensus_elect_batch(pb::ctx(), cons_, s, &rec, 1, ops.data(), &ops_off), "pb_consensus_elect_batch");
        pb_seqset_free(s);
    }

private:
    void create(int weight)
    {
        pb::check(pb_consensus_create(pb::ctx(), &txt_[beg], end - beg, weight, &cons_), "pb_consensus_create");
    }
    int beg, end, pre, post; // origin / end of the current iteration, extension before / after (ref_seq.h:363-366)
    bool locked;
    std::vector<char> txt_;  // the reference's txt_buf[3*MAX_SEQ_LEN]: what accessors point into
    pb_consensus *cons_ = nullptr;
};
