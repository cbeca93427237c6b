// spaced_seed -- the assembler's probe / verify rounds against a LOCKED reference, with the reference driver's command
// line, stderr log, stdout and dump file (src/spaced_seed.cpp):
//
//     spaced_seed -l [-f ref_file] [-r ratio] [-d dumpfile] [-m nround] [-t ntrials] bin seedfile
//
// Per round (spaced_seed.cpp:408-453): pick a seed (random while rounds succeed, then seeds[nfailure-1]), rebuild the
// reference's seed map (ref_seq::get_seedmap), and for every remaining read run the trial loop
//     for j < max_trial: try_align(read, j, +1) || try_align(read, slen-j-16, -1)
// whose first success removes the read from the pool, logs "found <id> at cost ..." and, with -d, dumps the matched part
// of the reference and of the read (in accessor order: backward matches come out reversed) for visual_align.  On the GPU a
// round is one pb_index_build + one pb_overlap_batch over the remaining reads (K1' -> K2 -> K3a -> K3).
//
// Scope: the locked mode only.  Without -l the reference votes every match into its consensus and grows its text DURING the
// round (ref_seq::try_align -> elect / append / prepend, ref_seq.h:266-276), which makes every alignment depend on the ones
// before it; that loop is outside this library's path (out of scope) and is refused loudly.
// Known differences, both documented in DESIGN.md: (1) the reference keeps fgets' trailing newline as the last element of a
// -f reference (spaced_seed.cpp:197-201); it is stripped here.  (2) "#trials" (a DBG counter of non-empty probes) is not printed.
// Random choices use rand() seeded with time(0) like the reference, or with $PB_SRAND for repeatable runs.
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "common.h"
#include "pb_runtime.hpp"

#define SEQ_THRESHOLD 500 // spaced_seed.cpp:36
#define handle_error(msg) do { perror(msg); exit(EXIT_FAILURE); } while (0)

static const char *usage_str = "usage: %s [options] bin seedfile\n"
    "options: [-f:r:d:m:t:lh]\n"
    "   -h          Get help and usage.\n"
    "   -f file     Use the string from file as starting reference (first line: the\n"
    "               sequence, second line: an integer weight, unused while locked).\n"
    "               Without this option a random segment is the reference.\n"
    "   -r ratio    Ratio of difference (0.3 by default) allowed.\n"
    "   -d dumpfile Dump matched segments.\n"
    "   -m nround   Maximum number of round of iteration.\n"
    "   -t ntrials  Number of seeding trial for each segment.\n"
    "   -l          Lock reference during iteration (required by this build).\n";

struct seq_index { int id; size_t offset; unsigned len; }; // spaced_seed.cpp:60-66

static std::string record_text(const std::vector<unsigned char> &buf, const seq_index &s)
{
    std::string t(s.len + 1, '\0');
    size_t tl = 0;
    pb::check(pb_bin2text(pb::ctx(), buf.data() + s.offset, &t[0], t.size(), &tl), "pb_bin2text");
    t.resize(tl);
    return t;
}

// dump_seq (spaced_seed.cpp:127-134) of an accessor view: `length` elements from offset o, walking up or down
static void dump_view(FILE *fp, const std::string &text, long o, bool forward, int length)
{
    for (int i = 0; i < length; ++i) fputc(text[(size_t)(forward ? o + i : o - i)], fp);
    fputc('\n', fp);
}

int main(int argc, char *argv[])
{
    double ratio = MAXR;
    bool locked = false;
    int max_round = INT_MAX, max_trial = 32, opt;
    FILE *fpref = NULL, *fpdump = NULL;
    if (argc < 3) { fprintf(stderr, usage_str, argv[0]); return EXIT_FAILURE; }
    while ((opt = getopt(argc, argv, "f:r:d:m:t:lh")) != -1) {
        switch (opt) {
            case 'h': fprintf(stdout, usage_str, argv[0]); return EXIT_SUCCESS;
            case 'f': if (!(fpref = fopen(optarg, "r"))) handle_error("failed to read ref_file"); break;
            case 'd': if (!(fpdump = fopen(optarg, "w"))) handle_error("failed to create dump file"); break;
            case 'r': ratio = atof(optarg); break;
            case 'l': locked = true; break;
            case 'm': max_round = atoi(optarg); break;
            case 't': max_trial = atoi(optarg); break;
            default: fprintf(stderr, usage_str, argv[0]); exit(EXIT_FAILURE);
        }
    }
    if (optind + 2 > argc) { fprintf(stderr, usage_str, argv[0]); return EXIT_FAILURE; }
    if (!locked) {
        fprintf(stderr, "%s: only the locked-reference mode (-l) runs on this path; consensus voting and growth during a round "
                        "(ref_seq::elect / append / prepend) are not part of it\n", argv[0]);
        return EXIT_FAILURE;
    }

    // open_binary (spaced_seed.cpp:309-345): records back to back, keep 500 < len < 20000, ids = rank among kept
    std::vector<unsigned char> buf;
    {
        FILE *fp = fopen(argv[optind], "rb");
        if (!fp) handle_error("open");
        unsigned char chunk[1 << 16];
        size_t n;
        while ((n = fread(chunk, 1, sizeof chunk, fp)) > 0) buf.insert(buf.end(), chunk, chunk + n);
        fclose(fp);
    }
    std::vector<seq_index> indices;
    for (size_t offset = 0; offset + 4 <= buf.size();) {
        unsigned seq_len;
        memcpy(&seq_len, &buf[offset], 4);
        if (seq_len > SEQ_THRESHOLD && seq_len < MAX_READ_LEN) indices.push_back({(int)indices.size(), offset, seq_len});
        offset += 4 + ((size_t)seq_len + 3) / 4;
    }
    LOG("indices: size %d\n", (int)indices.size());
    LOG("number of seeding trial: %d\n", max_trial);

    // init (spaced_seed.cpp:186-233)
    const char *sr = getenv("PB_SRAND");
    srand(sr ? (unsigned)atol(sr) : (unsigned)time(0));
    std::string ref;
    if (fpref) {
        int c;
        while ((c = fgetc(fpref)) != EOF && c != '\n') ref.push_back((char)c);
        int weight = 1;
        if (fscanf(fpref, "%d", &weight) != 1) weight = 1;
        LOG("reference weight: %d\n", weight);
        fclose(fpref);
    } else {
        if (indices.empty()) { fprintf(stderr, "no usable read in %s\n", argv[optind]); return EXIT_FAILURE; }
        const seq_index &s = indices[(size_t)rand() % indices.size()];
        ref = record_text(buf, s);
        LOG("%d selected as the initial reference.\n", s.id);
    }
    LOG("ref_len: %d\n", (int)ref.size());
    std::vector<unsigned> seeds;
    std::vector<std::string> seed_names;
    {
        FILE *fp = fopen(argv[optind + 1], "r");
        if (!fp) handle_error("failed to open seedfile");
        char line[1024];
        while (fgets(line, sizeof line, fp)) {
            line[strlen(line) - 1] = '\0'; // the reference drops the last character of every line (spaced_seed.cpp:225)
            seeds.push_back(pb_parse_pattern(line));
            LOG("seed %s: %08x\n", line, seeds.back());
        }
        fclose(fp);
    }
    if (seeds.empty()) { fprintf(stderr, "no seed pattern in %s\n", argv[optind + 1]); return EXIT_FAILURE; }

    pb_ctx *ctx = pb::ctx();
    pb_seqset *refset = nullptr;
    {
        const int64_t off = 0;
        const int32_t len = (int32_t)ref.size();
        pb::check(pb_seqset_from_text(ctx, ref.data(), &off, &len, nullptr, 1, &refset), "pb_seqset_from_text");
    }
    pb_overlap_params prm;
    pb_overlap_default_params(&prm);
    prm.R = ratio;
    prm.max_trial = max_trial;
    prm.seed_at_quirk = 1; // dna_seq::seed_at as shipped (SURVEY Q-S1)

    size_t nfailure = 0;
    std::vector<unsigned char> image;
    std::vector<pb_overlap_rec> recs;
    for (int nround = 1; nround <= max_round; ++nround) {
        const unsigned seed = nfailure == 0 ? seeds[(size_t)rand() % seeds.size()] : seeds[nfailure - 1];
        LOG("--------------- round %d ---------\n", nround);
        LOG("seed: %08x\n", seed);
        pb_index *ix = nullptr;
        pb::check(pb_index_build(ctx, refset, 0, seed, PB_POLICY_REFSEQ, &ix), "pb_index_build");
        LOG("seedmap size: %d\n", (int)pb_index_nscanned(ix));
        LOG("reference length: %d\n", (int)ref.size());
        int nmatches = 0;
        if (!indices.empty()) {
            image.clear();
            for (const seq_index &s : indices) image.insert(image.end(), buf.begin() + s.offset, buf.begin() + s.offset + 4 + (s.len + 3) / 4);
            pb_seqset *reads = nullptr;
            pb::check(pb_seqset_from_bin(ctx, image.data(), image.size(), SEQ_THRESHOLD, MAX_READ_LEN, &reads), "pb_seqset_from_bin");
            recs.resize(indices.size());
            pb::check(pb_overlap_batch(ctx, ix, refset, 0, reads, &prm, recs.data(), nullptr, nullptr), "pb_overlap_batch");
            pb_seqset_free(reads);
            std::vector<seq_index> left;
            int count = 0;
            for (size_t k = 0; k < indices.size(); ++k) {
                const pb_overlap_rec &r = recs[k];
                if (r.found) {
                    LOG("found %d at cost %d:\tref_ml=%d,\tseg_ml=%d\n", indices[k].id, r.cost, r.matlen_a, r.matlen_b);
                    ++nmatches;
                    if (fpdump) { // spaced_seed.cpp:286-292
                        const bool forward = r.dir == 1;
                        const std::string seg = record_text(buf, indices[k]);
                        dump_view(fpdump, ref, forward ? r.ref_pos : r.ref_pos + 15, forward, r.matlen_a);
                        dump_view(fpdump, seg, forward ? r.read_pos : r.read_pos + 15, forward, r.matlen_b);
                        fflush(fpdump);
                    }
                } else {
                    left.push_back(indices[k]);
                }
                if (!(++count & 0xFFFF)) LOG("%d sequences processed\n", count);
            }
            indices.swap(left);
        }
        pb_index_free(ix);
        LOG("#matches: %d\n", nmatches);
        if (nmatches != 0) nfailure = 0; // reset only if we have found some match
        else if (++nfailure == seeds.size()) break; // stop once every seed has been tried
        // evolve() is a no-op on a locked reference (ref_seq.h:317): the "consensus" printed per round is the reference
        fwrite(ref.data(), 1, ref.size(), stdout);
        fputc('\n', stdout);
    }
    if (fpdump) fclose(fpdump);
    pb_seqset_free(refset);
    pb::shutdown();
    return EXIT_SUCCESS;
}
