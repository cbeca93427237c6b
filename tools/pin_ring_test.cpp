// CPU test of the staging ring's bookkeeping (csrc/pb_pin_ring.h): random copy sizes, copies that complete late; a place must
// never be handed out while a chunk that has not been retired overlaps it, and every place lies inside the ring.
#include <stdio.h>
#include <stdlib.h>

#include <map>

#include "../pacbioassembly_b200/csrc/pb_pin_ring.h"

int main(int argc, char **argv)
{
    const int rounds = argc > 1 ? atoi(argv[1]) : 200000;
    unsigned long long s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (size_t ring : {(size_t)1 << 20, (size_t)3 << 20, (size_t)32 << 20}) {
        PbRingBook rb;
        rb.ring = ring;
        std::map<uint64_t, PbRingBook::Chunk> pending; // chunks nobody has waited for yet
        long retired = 0, laps = 0;
        size_t last = 0;
        for (int k = 0; k < rounds; ++k) {
            size_t bytes = 1 + rnd() % (ring / 2);
            if (rnd() % 4) bytes = 1 + bytes % (ring / 16 + 1); // mostly small, sometimes up to half the ring
            uint64_t id;
            std::vector<PbRingBook::Chunk> retire;
            const size_t off = rb.place(bytes, &id, &retire);
            const size_t need = (bytes + 255) & ~(size_t)255;
            if (off % 256 || off + need > ring) { printf("FAIL place outside ring\n"); return 1; }
            if (off < last) ++laps;
            last = off;
            for (auto &c : retire) { pending.erase(c.id); ++retired; }
            for (auto &kv : pending) // nothing that is still pending may overlap the new place
                if (!(kv.second.off + kv.second.bytes <= off || kv.second.off >= off + need)) { printf("FAIL overlap with a live chunk\n"); return 1; }
            pending[id] = {off, need, id};
            if (pending.size() != rb.live.size()) { printf("FAIL bookkeeping out of step\n"); return 1; }
        }
        printf("ring %zu: %d places, %ld laps, %ld retired, %zu live at the end\n", ring, rounds, laps, retired, rb.live.size());
    }
    printf("OK\n");
    return 0;
}
