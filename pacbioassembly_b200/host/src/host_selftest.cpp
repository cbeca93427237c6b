// host_selftest -- exercises the header-level drop-in classes (dna_seq, seq_accessor, seq_aligner<>, hash_table)
// on the GPU with the same known answers the reference's unit tests assert (test/dna_test.cpp:20-60,
// test/aligner_test.cpp:44-98, test/ref_test.cpp:119-128).  Own code, own harness (include/gtest shim).
#include <gtest/gtest.h>

#include <string>

#include <dna_seq.h>
#include <seq_aligner.h>

static char dna_str[] = "ACGTGTCATCGGATCAACCGGTT";

TEST(host_dna_seq, codec_known_answers)
{
    unsigned char bin_buf[14];
    char txt_buf[41];
    EXPECT_EQ(10u, dna_seq::text2bin(dna_str, bin_buf, 14));
    EXPECT_EQ(23u, dna_seq::bin2text(bin_buf, txt_buf, 41));
    EXPECT_STREQ(dna_str, txt_buf);
    EXPECT_EQ(0x34DAB41Bu, dna_seq::seed_at(bin_buf, 0));
    EXPECT_EQ(0xD068D36Eu, dna_seq::seed_at(bin_buf, 1));
    EXPECT_EQ(0x41A34DBBu, dna_seq::seed_at(bin_buf, 2));
    EXPECT_EQ(0xAF058D36u, dna_seq::seed_at(bin_buf, 7));
    EXPECT_EQ(0x0000001bu, dna_seq::encode("ACGTAAAAAAAAAAAA"));
    EXPECT_EQ(0x1b000000u, dna_seq::encode("AAAAAAAAAAAAACGT"));
    char dec[17] = {0};
    dna_seq::decode(0x34DAB41Bu, dec);
    EXPECT_STREQ("ACGTGTCATCGGATCA", dec);
    EXPECT_EQ('G', dna_seq::value_at(0x1B, 2));
}

TEST(host_seq_accessor, views)
{
    seq_accessor f(dna_str, true, 4);
    EXPECT_EQ('A', f.next()); EXPECT_EQ('C', f.next()); EXPECT_EQ('G', f.at(2));
    f.reset(2);
    EXPECT_EQ('G', f.next()); EXPECT_EQ('T', f.next()); EXPECT_EQ(false, f.has_more());
    seq_accessor b(dna_str + 4, false, 3);
    EXPECT_EQ('G', b.next()); EXPECT_EQ('T', b.next()); EXPECT_EQ('G', b.next()); EXPECT_EQ('T', b.at(1));
}

static void replay(seq_accessor *ref, t_aligner *al)
{ // every MATCH / INSERT carries seg_b's next element
    int j = 0;
    for (int i = 0; i < al->nedit; ++i)
        if (al->edits[i].op != DELETE) { EXPECT_EQ(ref->at(j), al->edits[i].val); ++j; }
}

TEST(host_seq_aligner, small_known_answers)
{
    char dna_ref[] = "ACGTAACCGGTT", seg1[] = "CGTAAGC", seg2[] = "GTAACGGGTTAA", seg3[] = "TCGTAAC";
    t_aligner al;
    { seq_accessor r(dna_ref, true, 8), s(seg1, true, 7); EXPECT_EQ(7, al.align(&s, &r)); EXPECT_EQ(2, al.final_cost()); replay(&r, &al); }
    { seq_accessor r(dna_ref, true, 8), s(seg3, true, 7); EXPECT_EQ(7, al.align(&s, &r)); EXPECT_EQ(1, al.final_cost()); replay(&r, &al); }
    { seq_accessor r(dna_ref + 7, false, 7), s(seg1 + 6, false, 7); EXPECT_EQ(7, al.align(&s, &r)); EXPECT_EQ(1, al.final_cost()); replay(&r, &al); }
    { seq_accessor r(dna_ref + 2, true, 10), s(seg2, true, 12); EXPECT_EQ(10, al.align(&s, &r)); EXPECT_EQ(1, al.final_cost()); replay(&r, &al); }
    {
        seq_accessor r(dna_ref, true, 10), s(dna_ref + 1, true, 9);
        EXPECT_EQ(10, al.align(&s, &r)); EXPECT_EQ(10, al.nedit); EXPECT_EQ(INSERT, al.edits[0].op); EXPECT_EQ(1, al.final_cost());
        EXPECT_EQ(9, al.align(&r, &s)); EXPECT_EQ(10, al.nedit); EXPECT_EQ(DELETE, al.edits[0].op); EXPECT_EQ(1, al.final_cost());
        EXPECT_EQ(al.final_cost(), al.get_cost(al.matlen_a, al.matlen_b));
    }
}

TEST(host_hash_table, seedmap_known_answers)
{
    char dna_txt[] = "ACGTAACCGGTTAAACCCGGGTTTTGCAAAAAAAAAAAAAAAA"; // 43 bases
    const int sz = (int)strlen(dna_txt);
    hash_table seedmap;
    EXPECT_EQ((unsigned)(sz - 16), seedmap.build_refseq(dna_txt, sz, 0xFFFFFFFFu));
    EXPECT_EQ((size_t)(sz - 15 - 1), seedmap.size());
    for (int i = 0; i < sz - 16; ++i) EXPECT_TRUE(seedmap.find(dna_seq::encode(dna_txt + i)) != seedmap.end());
    EXPECT_TRUE(seedmap.find(dna_seq::encode(dna_txt + sz - 15)) == seedmap.end());
    sm_it it = seedmap.find(dna_seq::encode(dna_txt + 3));
    EXPECT_EQ(1u, it->second.size());
    EXPECT_EQ(3, it->second.front());
    // locator policy: every position, list order ascending
    char rep[] = "ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT";
    seedmap.build_locator(rep, strlen(rep), 0xFFFFFFFFu);
    it = seedmap.find(dna_seq::encode(rep));
    EXPECT_TRUE(it != seedmap.end());
    int want = 0;
    for (std::list<int>::iterator p = it->second.begin(); p != it->second.end(); ++p, want += 4) EXPECT_EQ(want, *p);
    EXPECT_EQ(28, want); // positions 0,4,...,24 hold the full 16-mer
}
