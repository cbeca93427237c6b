// gtest/gtest.h -- a few-line stand-in for GoogleTest (absent from this image, no network), enough to compile and
// run the reference's own test/*.cpp unmodified against this directory's headers.  Written from scratch.
#pragma once

#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace testing {
class Test {
public:
    virtual ~Test() {}
    virtual void SetUp() {}
    virtual void TearDown() {}
    virtual void TestBody() = 0;
};
struct Registry {
    struct Item { std::string name; std::function<Test *()> make; };
    static std::vector<Item> &items() { static std::vector<Item> v; return v; }
    static int &failures() { static int f = 0; return f; }
};
struct Registrar {
    Registrar(const char *suite, const char *name, std::function<Test *()> mk)
    { Registry::items().push_back({std::string(suite) + "." + name, mk}); }
};
template <class A, class B> bool cmp_eq(const A &a, const B &b) { return a == b; }
inline void report(const char *file, int line, const char *expr)
{
    ++Registry::failures();
    std::fprintf(stderr, "%s:%d: Failure: %s\n", file, line, expr);
}
inline void InitGoogleTest(int *, char **) {}
} // namespace testing

#define PB_GT_CLASS(s, n) s##_##n##_Test
#define PB_GT_DEFINE(s, n, base)                                                                       \
    class PB_GT_CLASS(s, n) : public base { public: void TestBody() override; };                        \
    static ::testing::Registrar pb_gt_reg_##s##_##n(#s, #n, [] { return (::testing::Test *)new PB_GT_CLASS(s, n)(); }); \
    void PB_GT_CLASS(s, n)::TestBody()
#define TEST(s, n) PB_GT_DEFINE(s, n, ::testing::Test)
#define TEST_F(f, n) PB_GT_DEFINE(f, n, f)

#define PB_GT_CHECK(cond, text) do { if (!(cond)) ::testing::report(__FILE__, __LINE__, text); } while (0)
#define EXPECT_EQ(a, b) PB_GT_CHECK(::testing::cmp_eq((a), (b)), "EXPECT_EQ(" #a ", " #b ")")
#define EXPECT_NE(a, b) PB_GT_CHECK(!::testing::cmp_eq((a), (b)), "EXPECT_NE(" #a ", " #b ")")
#define EXPECT_LE(a, b) PB_GT_CHECK((a) <= (b), "EXPECT_LE(" #a ", " #b ")")
#define EXPECT_LT(a, b) PB_GT_CHECK((a) < (b), "EXPECT_LT(" #a ", " #b ")")
#define EXPECT_GE(a, b) PB_GT_CHECK((a) >= (b), "EXPECT_GE(" #a ", " #b ")")
#define EXPECT_GT(a, b) PB_GT_CHECK((a) > (b), "EXPECT_GT(" #a ", " #b ")")
#define EXPECT_TRUE(a) PB_GT_CHECK((a), "EXPECT_TRUE(" #a ")")
#define EXPECT_FALSE(a) PB_GT_CHECK(!(a), "EXPECT_FALSE(" #a ")")
#define EXPECT_STREQ(a, b) PB_GT_CHECK(std::strcmp((a), (b)) == 0, "EXPECT_STREQ(" #a ", " #b ")")
#define ASSERT_EQ EXPECT_EQ
#define ASSERT_TRUE EXPECT_TRUE

inline int RUN_ALL_TESTS()
{
    int ran = 0;
    for (auto &it : ::testing::Registry::items()) {
        int before = ::testing::Registry::failures();
        ::testing::Test *t = it.make();
        t->SetUp();
        t->TestBody();
        t->TearDown();
        delete t;
        ++ran;
        std::printf("[%s] %s\n", ::testing::Registry::failures() == before ? "  OK  " : "FAILED", it.name.c_str());
        std::fflush(stdout); // a later test may assert() its way out of the process
    }
    std::printf("%d tests, %d failures\n", ran, ::testing::Registry::failures());
    return ::testing::Registry::failures() ? 1 : 0;
}
