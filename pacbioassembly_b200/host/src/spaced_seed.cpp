// spaced_seed -- the assembler's probe / verify / vote rounds, with the reference driver's command line, stderr log, stdout
// and dump file (src/spaced_seed.cpp):
//
//     spaced_seed [-l] [-f ref_file] [-r ratio] [-d dumpfile] [-m nround] [-t ntrials] bin seedfile
//
// Per round (spaced_seed.cpp:408-453): pick a seed (random while rounds succeed, then seeds[nfailure-1]), rebuild the
// reference's seed map (ref_seq::get_seedmap), and for every remaining read run the trial loop
//     for j < max_trial: try_align(read, j, +1) || try_align(read, slen-j-16, -1)
// whose first success removes the read from the pool, logs "found <id> at cost ..." and, with -d, dumps the matched part
// of the reference and of the read (in accessor order: backward matches come out reversed) for visual_align.
//
// Locked (-l): the reference never changes; a round is one pb_index_build + one pb_overlap_subset over the remaining reads
// (K1' -> K2 -> K3a -> K3).
// Unlocked: every match also votes its transcript into the consensus (ref_seq::elect) and a read that runs past an end
// grows the reference text, which later reads of the same round then see (ref_seq::try_align, ref_seq.h:259-276); evolve()
// rewrites the text from the votes when the round ends and that consensus goes to stdout.  Votes commute, growth does not:
// a round runs as passes of pb_overlap_subset over the reads not yet visited against the text as it is now (ref_shift = what
// has grown in front); results are exact up to and including the first growing read, those matches are voted in one
// pb_consensus_elect_batch, the growth is applied (pb_consensus_append / prepend) and the next pass starts behind that read.
// A round without growth is one pass.  pb_consensus_evolve closes the round.
//
// Known differences, both documented in DESIGN.md: (1) the reference keeps fgets' trailing newline as the last element of a
// -f reference (spaced_seed.cpp:197-201); it is stripped here.  (2) "#trials" (a DBG counter of non-empty probes) is not printed.
// Random choices use rand() seeded with time(0) like the reference, or with $PB_SRAND for repeatable runs.
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
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
    "               sequence, second line: an integer weight: its initial votes).\n"
    "               Without this option a random segment is the reference.\n"
    "   -r ratio    Ratio of difference (0.3 by default) allowed.\n"
    "   -d dumpfile Dump matched segments.\n"
    "   -m nround   Maximum number of round of iteration.\n"
    "   -t ntrials  Number of seeding trial for each segment.\n"
    "   -l          Lock reference during iteration.\n";

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
    int weight = 1;
    if (fpref) {
        int c;
        while ((c = fgetc(fpref)) != EOF && c != '\n') ref.push_back((char)c);
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
    pb_seqset *refset = nullptr;   // locked: the reference, uploaded once
    pb_consensus *cons = nullptr;  // unlocked: text + vote boxes on the device
    if (locked) {
        const int64_t off = 0;
        const int32_t len = (int32_t)ref.size();
        pb::check(pb_seqset_from_text(ctx, ref.data(), &off, &len, nullptr, 1, &refset), "pb_seqset_from_text");
    } else {
        pb::check(pb_consensus_create(ctx, ref.data(), (int64_t)ref.size(), weight, &cons), "pb_consensus_create");
    }
    pb_overlap_params prm;
    pb_overlap_default_params(&prm);
    prm.R = ratio;
    prm.max_trial = max_trial;
    prm.seed_at_quirk = 1; // dna_seq::seed_at as shipped (SURVEY Q-S1)
    prm.want_ops = locked ? 0 : 1;

    // the reads go to the device once, as the whole image: a pass names the reads it visits by id, and the shipped seed_at
    // (raw image bytes at pos%4==0, SURVEY Q-S1) reads past a record's end into its neighbours exactly as the reference does
    pb_seqset *reads = nullptr;
    pb::check(pb_seqset_from_bin(ctx, buf.data(), buf.size(), SEQ_THRESHOLD, MAX_READ_LEN, &reads), "pb_seqset_from_bin");
    size_t nfailure = 0;
    std::vector<int32_t> ids;
    std::vector<pb_overlap_rec> recs;
    std::vector<uint8_t> ops;
    std::vector<int64_t> ops_off;
    std::string full; // the readable text [pre, post) of this pass
    for (int nround = 1; nround <= max_round; ++nround) {
        const unsigned seed = nfailure == 0 ? seeds[(size_t)rand() % seeds.size()] : seeds[nfailure - 1];
        LOG("--------------- round %d ---------\n", nround);
        LOG("seed: %08x\n", seed);
        pb_seqset *cur = refset; // the text [beg, end) the seed map is built over
        if (!locked) pb::check(pb_consensus_seqset(ctx, cons, 0, &cur), "pb_consensus_seqset");
        pb_index *ix = nullptr;
        pb::check(pb_index_build(ctx, cur, 0, seed, PB_POLICY_REFSEQ, &ix), "pb_index_build");
        LOG("seedmap size: %d\n", (int)pb_index_nscanned(ix));
        LOG("reference length: %d\n", locked ? (int)ref.size() : (int)pb_consensus_length(cons));
        int nmatches = 0, count = 0;
        std::vector<seq_index> todo(indices), left;
        size_t window = locked ? todo.size() : 256; // unlocked: reads per pass -- what lies behind a growing read is recomputed
        while (!todo.empty()) {
            std::vector<seq_index> pending(todo.begin(), todo.begin() + (long)std::min(window, todo.size()));
            int64_t before = 0, total = (int64_t)ref.size();
            pb_seqset *view = refset;
            if (!locked) {
                pb::check(pb_consensus_extent(cons, &before, &total), "pb_consensus_extent");
                pb::check(pb_consensus_seqset(ctx, cons, 1, &view), "pb_consensus_seqset");
                if (fpdump) {
                    full.assign((size_t)total + 1, '\0');
                    pb::check(pb_consensus_text(ctx, cons, 1, &full[0], full.size()), "pb_consensus_text");
                    full.resize((size_t)total);
                }
            }
            ids.clear();
            for (const seq_index &s : pending) ids.push_back(s.id);
            recs.resize(pending.size());
            if (!locked) {
                ops_off.resize(pending.size());
                int64_t ext = 0;
                for (size_t k = 0; k < pending.size(); ++k) {
                    ops_off[k] = ext;
                    ext += (3 * (int64_t)pending[k].len + 2 * prm.maxm + 16 + 15) & ~(int64_t)15;
                }
                if (ops.size() < (size_t)ext + 16) ops.resize((size_t)ext + 16);
            }
            prm.ref_shift = (int32_t)before;
            pb::check(pb_overlap_subset(ctx, ix, view, 0, reads, ids.data(), (int64_t)ids.size(), &prm, recs.data(), locked ? nullptr : ops.data(),
                                        locked ? nullptr : ops_off.data()), "pb_overlap_subset");
            // unlocked: the first match that consumes its whole reference view grows the text (ref_seq.h:267); what was
            // computed for the reads behind it is void, they go into the next pass
            size_t stop = pending.size();
            bool grows = false;
            if (!locked)
                for (size_t k = 0; k < pending.size(); ++k) {
                    const pb_overlap_rec &r = recs[k];
                    if (!r.found) continue;
                    const int64_t r_off = r.dir == 1 ? r.ref_pos : r.ref_pos + 15;
                    const int64_t a_len = r.dir == 1 ? (total - before) - r_off : r_off + before + 1; // get_accessor, ref_seq.h:282-286
                    if (r.matlen_a == a_len) { stop = k + 1; grows = true; break; }
                }
            if (!locked) pb::check(pb_consensus_elect_batch(ctx, cons, reads, recs.data(), (int64_t)stop, ops.data(), ops_off.data()), "pb_consensus_elect_batch");
            for (size_t k = 0; k < stop; ++k) {
                const pb_overlap_rec &r = recs[k];
                if (r.found) {
                    LOG("found %d at cost %d:\tref_ml=%d,\tseg_ml=%d\n", pending[k].id, r.cost, r.matlen_a, r.matlen_b);
                    ++nmatches;
                    const bool forward = r.dir == 1;
                    std::string seg;
                    if (fpdump || (grows && k + 1 == stop)) seg = record_text(buf, pending[k]);
                    if (fpdump) { // spaced_seed.cpp:286-292
                        dump_view(fpdump, locked ? ref : full, before + (forward ? r.ref_pos : r.ref_pos + 15), forward, r.matlen_a);
                        dump_view(fpdump, seg, forward ? r.read_pos : r.read_pos + 15, forward, r.matlen_b);
                        fflush(fpdump);
                    }
                    if (grows && k + 1 == stop) { // ref_seq.h:267-275
                        if (forward) {
                            const int add = ((int)seg.size() - r.read_pos) - r.matlen_b;
                            pb::check(pb_consensus_append(ctx, cons, seg.data() + r.read_pos + r.matlen_b, add), "pb_consensus_append");
                        } else { // pt(length-1) of a backward accessor is the read's first base
                            const int add = (r.read_pos + 16) - r.matlen_b;
                            pb::check(pb_consensus_prepend(ctx, cons, seg.data(), add), "pb_consensus_prepend");
                        }
                    }
                } else {
                    left.push_back(pending[k]);
                }
                if (!(++count & 0xFFFF)) LOG("%d sequences processed\n", count);
            }
            if (!locked) pb_seqset_free(view);
            todo.erase(todo.begin(), todo.begin() + (long)stop);
            window = grows ? std::max<size_t>(64, window / 2) : window * 4;
        }
        indices.swap(left);
        pb_index_free(ix);
        if (!locked) pb_seqset_free(cur);
        LOG("#matches: %d\n", nmatches);
        if (nmatches != 0) nfailure = 0; // reset only if we have found some match
        else if (++nfailure == seeds.size()) break; // stop once every seed has been tried
        if (!locked) { // pref->evolve(), then the consensus goes to stdout (spaced_seed.cpp:449-452)
            pb::check(pb_consensus_evolve(ctx, cons), "pb_consensus_evolve");
            ref.assign((size_t)pb_consensus_length(cons) + 1, '\0');
            pb::check(pb_consensus_text(ctx, cons, 0, &ref[0], ref.size()), "pb_consensus_text");
            ref.resize(ref.size() - 1);
        }
        fwrite(ref.data(), 1, ref.size(), stdout);
        fputc('\n', stdout);
    }
    if (fpdump) fclose(fpdump);
    pb_seqset_free(reads);
    if (refset) pb_seqset_free(refset);
    if (cons) pb_consensus_free(cons);
    pb::shutdown();
    return EXIT_SUCCESS;
}
