"""Host side of the assembler's rounds with an UNLOCKED reference (spaced_seed.cpp:408-453) over the C ABI.

In the reference every successful try_align votes into the consensus at once and, when the read runs past an end of the
reference, grows the reference text (ref_seq::try_align, ref_seq.h:259-276), so a later read of the same round is aligned
against the longer text.  Votes are counter increments and commute; growth does not.  A round therefore runs as passes:
one pb_overlap_batch over the reads not yet visited, against the text as it is now; the results are exact up to and
including the first read that grows the reference (everything before it saw the same text in the reference's own order);
those matches are voted in one pb_consensus_elect_batch, the growth is applied, and the next pass starts behind that read.
A round without growth is one pass.  pb_consensus_evolve closes the round (ref_seq::evolve, ref_seq.h:317-348).
"""
from __future__ import annotations

import numpy as np

from .api import OVERLAP_DTYPE, POLICY_REFSEQ, Context


def kept_records(image: bytes, min_excl: int = 500, max_excl: int = 20000) -> list[bytes]:
    """open_binary (spaced_seed.cpp:309-345): records back to back, keep min_excl < len < max_excl"""
    recs, p = [], 0
    while p + 4 <= len(image):
        l = int.from_bytes(image[p:p + 4], "little")
        n = 4 + (l + 3) // 4
        if min_excl < l < max_excl:
            recs.append(image[p:p + n])
        p += n
    return recs


def assemble_rounds(ctx: Context, ref_text, image: bytes, round_masks, weight: int = 1, R: float = 0.3, max_trial: int = 32,
                    seed_at_quirk: int = 1, min_excl: int = 500, max_excl: int = 20000, log=None):
    """Returns (consensus text per round, found_round int32[nkept] (0 = never), OVERLAP_DTYPE records of the successes,
    number of pb_overlap_batch passes per round)."""
    records = kept_records(image, min_excl, max_excl)
    nk = len(records)
    cons = ctx.consensus(ref_text, weight)
    reads = ctx.seqset_from_bin(image, min_excl, max_excl)  # once: passes name the reads they visit by id (pb_overlap_subset)
    pool = list(range(nk))
    found_round = np.zeros(nk, dtype=np.int32)
    out = np.zeros(nk, dtype=OVERLAP_DTYPE)
    out["id"] = np.arange(nk)
    texts, passes = [], []
    for rnd, mask in enumerate(round_masks):
        cur = cons.seqset(full=False)
        ix = ctx.index(cur, int(mask), policy=POLICY_REFSEQ)  # ref_seq::get_seedmap over [beg, end)
        pending = list(pool)
        npass = 0
        window = 256  # reads per pass: what lies behind a growing read is recomputed, so look ahead only as far as growth is rare
        while pending:
            before, total = cons.extent()
            full = cons.seqset(full=True)
            chunk = pending[:window]
            recs, ops, ops_off = ctx.overlap(ix, reads, want_ops="raw", ref=full, ids=chunk, ref_shift=before, R=R, max_trial=max_trial,
                                             seed_at_quirk=seed_at_quirk)
            npass += 1
            # the first match that consumes its whole reference view grows the text (ref_seq.h:267): results behind it are void
            fwd = recs["dir"] == 1
            r_off = np.where(fwd, recs["ref_pos"], recs["ref_pos"] + 15).astype(np.int64)
            a_len = np.where(fwd, (total - before) - r_off, r_off + before + 1)
            grows = np.nonzero((recs["found"] == 1) & (recs["matlen_a"] == a_len))[0]
            stop = int(grows[0]) + 1 if len(grows) else len(chunk)
            window = max(64, window // 2) if len(grows) else window * 4
            batch = recs[:stop].copy()
            cons.elect(reads, batch, ops, ops_off[:stop])
            if len(grows):
                g = batch[stop - 1]
                text = ctx.bin2text(records[pending[stop - 1]])
                if g["dir"] == 1:
                    s_off = int(g["read_pos"])
                    add = (len(text) - s_off) - int(g["matlen_b"])
                    cons.append(text[s_off + int(g["matlen_b"]): s_off + int(g["matlen_b"]) + add])
                else:  # pt(length-1) of a backward accessor is the read's first base (ref_seq.h:273)
                    add = (int(g["read_pos"]) + 16) - int(g["matlen_b"])
                    cons.prepend(text[:add])
            for i in np.nonzero(batch["found"] == 1)[0]:
                k = pending[i]
                found_round[k] = rnd + 1
                out[k] = batch[i]
                if log:
                    log(f"found {k} at cost {int(batch[i]['cost'])}:\tref_ml={int(batch[i]['matlen_a'])},\tseg_ml={int(batch[i]['matlen_b'])}")
            pool = [k for k in pool if not found_round[k]]
            pending = pending[stop:]
            full.free()
        ix.free()
        cur.free()
        cons.evolve()
        texts.append(cons.text())
        passes.append(npass)
    reads.free()
    cons.free()
    return texts, found_round, out, passes
