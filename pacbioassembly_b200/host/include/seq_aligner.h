// seq_aligner.h -- `seq_aligner<MAXN,MAXM>` with the reference's public interface (src/seq_aligner.h:58-81),
// executed by the CUDA banded aligner (K3) through pb_align_batch (a batch of one; use the batched C ABI for
// throughput).  Fresh-state semantics: nothing carries over between calls (SURVEY Q-D2); after a failed align()
// the result fields are unspecified, as in the reference (Q-D4).
#pragma once

#include <string.h>

#include <vector>

#include "common.h"
#include "dna_seq.h"
#include "pb_runtime.hpp"

enum OP { MATCH = 1, INSERT, DELETE }; // seq_aligner.h:32-36

typedef struct {
    enum OP op; // operation
    char val;   // value inserted / matched: seg_b's element (unset for DELETE, as in the reference)
} edit;

template <int MAXN, int MAXM> class seq_aligner {
public:
    seq_aligner() : R(MAXR), edits(MAXN + MAXM), ops_(MAXN + MAXM + 64) {}
    seq_aligner(double r) : R(r), edits(MAXN + MAXM), ops_(MAXN + MAXM + 64) {}
    double R;     // ratio of difference allowed (public and mutable, as in the reference)
    int len_a = 0, len_b = 0, max_dst = 0, matlen_a = 0, matlen_b = 0;
    std::vector<edit> edits; // edits[0..nedit): transform seg_a into seg_b[0:matlen_b]
    int nedit = 0;

    // seq_aligner.h:92-125: returns matlen_b, or -1
    int align(seq_accessor *seg_a, seq_accessor *seg_b)
    {
        const int32_t la = seg_a->length(), lb = seg_b->length();
        const int32_t sa = seg_a->is_forward() ? 1 : -1, sb = seg_b->is_forward() ? 1 : -1;
        const int64_t off = 0, ops_off = 0;
        if ((size_t)la + lb + 64 > ops_.size()) ops_.resize((size_t)la + lb + 64);
        pb_align_out out;
        pb::check(pb_align_batch(pb::ctx(), seg_a->pt(0), &off, &la, &sa, seg_b->pt(0), &off, &lb, &sb, 1, R, MAXN, MAXM, &out,
                                 ops_.data(), &ops_off), "pb_align_batch");
        len_a = out.len_a; len_b = out.len_b; max_dst = out.max_dst;
        if (out.ret < 0) {
            if (out.fail_row == 0 && out.cells == 0 && (out.len_a >= MAXN || out.max_dst >= MAXM))
                LOG("segment too long: %d\n", out.len_a); // seq_aligner.h:105
            return -1;
        }
        matlen_a = out.matlen_a; matlen_b = out.matlen_b; nedit = out.nedit;
        cost_ = out.cost; diag_cost_ = out.diag_cost; a_len_ = la;
        int j = 0;
        for (int k = 0; k < nedit; ++k) { // edit.val = seg_b->at(j-1) under MATCH / INSERT (seq_aligner.h:219,225)
            edits[k].op = (OP)ops_[k];
            if (ops_[k] != DELETE) edits[k].val = seg_b->at(j++);
        }
        return out.ret;
    }
    int final_cost() { return cost_; } // seq_aligner.h:130
    // seq_aligner.h:131.  The reference exposes its whole DP matrix; its callers read two cells of it: the goal cell
    // (final_cost) and the main-diagonal cell at seg_a's end (locator.cpp:86).  Those two are served; any other
    // cell aborts rather than return something made up.
    int get_cost(int i, int j)
    {
        if (i == matlen_a && j == matlen_b) return cost_;
        if (i == a_len_ && j == a_len_) return diag_cost_;
        LOG("seq_aligner::get_cost(%d,%d): only the goal cell and (|a|,|a|) are kept on the host\n", i, j);
        abort();
    }

private:
    std::vector<uint8_t> ops_;
    int cost_ = 0, diag_cost_ = 0, a_len_ = 0;
};

typedef seq_aligner<MAX_READ_LEN + MAX_DIFF_LEN, MAX_DIFF_LEN> t_aligner; // seq_aligner.h:260
