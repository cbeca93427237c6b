// pb_pin_ring.h -- bookkeeping of the pinned staging ring behind pb_h2d (pb_ctx.cu): which bytes of the ring a staged copy gets
// and which earlier chunks must have been copied out before those bytes are written again.  Plain C++ (no CUDA) so that the
// discipline is testable on the CPU (tests/test_abi.py builds tools/pin_ring_test.cpp).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

struct PbRingBook {
    struct Chunk { size_t off, bytes; uint64_t id; };
    size_t ring = 0, head = 0;
    uint64_t next_id = 1;
    std::vector<Chunk> live; // in the order they were placed

    // Places `bytes` (rounded up to 256) in the ring; returns the offset and the id of the new chunk, and moves every live chunk
    // that overlaps the place out of `live` into `retire` -- the caller waits for those before writing.  bytes <= ring.
    size_t place(size_t bytes, uint64_t *id, std::vector<Chunk> *retire)
    {
        const size_t need = (bytes + 255) & ~(size_t)255;
        if (head + need > ring) head = 0; // a lap that wraps early leaves chunks near the end of the ring older than those at its start
        const size_t lo = head, hi = head + need;
        size_t w = 0;
        for (size_t r = 0; r < live.size(); ++r) { // all of them are looked at, not only the oldest
            if (live[r].off + live[r].bytes <= lo || live[r].off >= hi) live[w++] = live[r];
            else retire->push_back(live[r]);
        }
        live.resize(w);
        *id = next_id++;
        live.push_back({lo, need, *id});
        head = hi;
        return lo;
    }
};
