// visual_align -- gapped display of pairwise alignments, with the reference tool's input and output (src/visual_align.cpp):
//
//     visual_align < pairs        pairs = whitespace-separated strings, taken two at a time: reference, then segment
//     stdout per pair: final cost, the reference with '-' where the segment has an extra base, the segment with '-' where
//     the reference has one.
//
// Written against the drop-in seq_aligner<> (GPU banded aligner through the C ABI).  The direction of the call is the
// reference's: align(&seg, &ref), so INSERT consumes a reference base and DELETE a segment base (visual_align.cpp:41,54-65).
// One deliberate difference: when align() fails the reference still prints whatever the previous alignment left in its aligner
// (SURVEY Q-D4); here a failed pair prints the "cannot align" lines on stderr and nothing on stdout.
#include <iostream>
#include <string>

#include "common.h"
#include "dna_seq.h"
#include "seq_aligner.h"

int main()
{
    std::string ref_str, seg_str;
    t_aligner *aligner = new t_aligner();
    while (std::cin >> ref_str >> seg_str) {
        seq_accessor ref((char *)ref_str.c_str(), true, (int)ref_str.length());
        seq_accessor seg((char *)seg_str.c_str(), true, (int)seg_str.length());
        if (aligner->align(&seg, &ref) <= 0) {
            std::cerr << "cannot align" << std::endl << ref_str << std::endl << seg_str << std::endl;
            continue;
        }
        std::string gapped_ref, gapped_seg;
        size_t iref = 0, iseg = 0;
        for (int i = 0; i < aligner->nedit; ++i) {
            switch (aligner->edits[i].op) {
                case MATCH: gapped_ref += ref_str[iref++]; gapped_seg += seg_str[iseg++]; break;
                case INSERT: gapped_ref += ref_str[iref++]; gapped_seg += '-'; break;
                default: gapped_ref += '-'; gapped_seg += seg_str[iseg++]; break;
            }
        }
        std::cout << aligner->final_cost() << std::endl << gapped_ref << std::endl << gapped_seg << std::endl;
    }
    delete aligner;
    pb::shutdown();
    return EXIT_SUCCESS;
}
