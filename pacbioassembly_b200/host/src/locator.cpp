// locator -- read-to-contig locator with the reference driver's command line and output (src/locator.cpp):
//
//     locator contig_file pattern [R] < reads        ->  stdout: nseq \t pos \t cost \t len-j \t cost(len-j,len-j)
//
// Written against the GPU library: the contig is indexed once on the device (K1 + index build), reads are
// streamed from stdin in batches and mapped by pb_locate_batch (K1 -> K2 -> K3a -> K3).  Semantics follow
// locator.cpp:41-96 line by line where observable: first whitespace token of the contig file, '1' = care in the
// pattern, reads shorter than 500 skipped without consuming a sequence number (Q-L1), first successful candidate in
// (offset j, list order) wins, R = 0.15 unless given.  Progress lines go to stderr like the reference's LOG().
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.h"
#include "pb_runtime.hpp"

int main(int argc, char *argv[])
{
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); // one hardware queue per band-class stream of the aligner (this process is ours)
    if (argc <= 2) {
        fprintf(stderr, "usage: locator contig_file seed [ratio] < seq_file\n");
        return EXIT_FAILURE;
    }
    FILE *fp = fopen(argv[1], "r");
    if (!fp) { perror(argv[1]); return EXIT_FAILURE; }
    std::string contig;
    { // fscanf("%s"): first whitespace-delimited token (locator.cpp:48-49), without the fixed 800000-byte buffer
        int c;
        while ((c = fgetc(fp)) != EOF && (c == ' ' || c == '\n' || c == '\t' || c == '\r')) {}
        for (; c != EOF && c != ' ' && c != '\n' && c != '\t' && c != '\r'; c = fgetc(fp)) contig.push_back((char)c);
        fclose(fp);
    }
    // locator.cpp:56-60, "convert N to A": the loop never advances its pointer, so what it does is turn contig[0] into 'A' when
    // it is 'N' -- before the seed map is built and before any alignment reads the contig.  Mirrored, not fixed.
    if (!contig.empty() && contig[0] == 'N') contig[0] = 'A';
    const unsigned mask = pb_parse_pattern(argv[2]); // locator.cpp:51-54
    pb_locate_params prm;
    pb_locate_default_params(&prm);
    if (argc > 3) prm.R = atof(argv[3]);

    hash_table seedmap(1 << 23);
    seedmap.build_locator(contig.data(), contig.size(), mask); // locator.cpp:62-66

    const size_t BATCH_BASES = (size_t)256 << 20;
    std::vector<char> text;
    std::vector<int64_t> off;
    std::vector<int32_t> len;
    std::vector<pb_locate_rec> recs;
    long nseq = 0;
    bool eof = false;
    std::string tok;
    while (!eof) {
        text.clear(); off.clear(); len.clear();
        while (text.size() < BATCH_BASES) { // scanf("%s") tokens (locator.cpp:70)
            int c;
            while ((c = getchar()) != EOF && (c == ' ' || c == '\n' || c == '\t' || c == '\r')) {}
            if (c == EOF) { eof = true; break; }
            tok.clear();
            for (; c != EOF && c != ' ' && c != '\n' && c != '\t' && c != '\r'; c = getchar()) tok.push_back((char)c);
            off.push_back((int64_t)text.size());
            len.push_back((int32_t)tok.size());
            text.insert(text.end(), tok.begin(), tok.end());
            if (c == EOF) { eof = true; break; }
        }
        if (off.empty()) break;
        recs.resize(off.size());
        int64_t nkept = 0;
        pb::check(pb_locate_batch(pb::ctx(), seedmap.index(), seedmap.reference(), 0, text.data(), off.data(), len.data(),
                                  (int64_t)off.size(), &prm, recs.data(), &nkept, nullptr, nullptr), "pb_locate_batch");
        for (int64_t k = 0; k < nkept; ++k) {
            const pb_locate_rec &r = recs[k];
            if (r.found) printf("%ld\t%d\t%d\t%d\t%d\n", nseq + r.nseq, r.pos, r.cost, r.seg_len, r.diag_cost); // locator.cpp:84-86
            if (!((nseq + r.nseq + 1) & 0xFFF)) LOG("%ld sequences processed\n", nseq + r.nseq + 1);
        }
        nseq += nkept;
    }
    LOG("totally %ld sequences processed\n", nseq);
    seedmap.clear();
    pb::shutdown();
    return EXIT_SUCCESS;
}
