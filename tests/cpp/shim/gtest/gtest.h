// Minimal stand-in for <gtest/gtest.h> (GoogleTest is not in this image): just enough for the reference's
// test/*.cpp sources -- TEST, EXPECT_EQ / EXPECT_TRUE / EXPECT_NE / ASSERT_*, RUN_ALL_TESTS -- to compile
// UNMODIFIED and report pass / fail through the exit code.  Test infrastructure only.
#pragma once
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

namespace testing {
struct Registry {
    struct Item {
        std::string name;
        std::function<void()> fn;
    };
    static std::vector<Item>& items()
    {
        static std::vector<Item> v;
        return v;
    }
    static int& failures()
    {
        static int f = 0;
        return f;
    }
};
struct Registrar {
    Registrar(const char* suite, const char* name, std::function<void()> fn)
    {
        Registry::items().push_back({std::string(suite) + "." + name, fn});
    }
};
inline void InitGoogleTest(int*, char**) {}
} // namespace testing

#define TEST(suite, name)                                                                          \
    static void suite##_##name##_body();                                                           \
    static ::testing::Registrar suite##_##name##_reg(#suite, #name, suite##_##name##_body);        \
    static void suite##_##name##_body()

#define HEON_GT_CHECK(cond, text)                                                                  \
    do                                                                                             \
    {                                                                                              \
        if (!(cond))                                                                               \
        {                                                                                          \
            std::printf("%s:%d: Failure: %s\n", __FILE__, __LINE__, text);                         \
            ++::testing::Registry::failures();                                                     \
        }                                                                                          \
    } while (0)
#define EXPECT_EQ(a, b) HEON_GT_CHECK((a) == (b), #a " == " #b)
#define EXPECT_NE(a, b) HEON_GT_CHECK((a) != (b), #a " != " #b)
#define EXPECT_TRUE(a) HEON_GT_CHECK((a), #a)
#define EXPECT_FALSE(a) HEON_GT_CHECK(!(a), "!" #a)
#define EXPECT_LT(a, b) HEON_GT_CHECK((a) < (b), #a " < " #b)
#define EXPECT_LE(a, b) HEON_GT_CHECK((a) <= (b), #a " <= " #b)
#define EXPECT_GT(a, b) HEON_GT_CHECK((a) > (b), #a " > " #b)
#define ASSERT_EQ(a, b) EXPECT_EQ(a, b)
#define ASSERT_TRUE(a) EXPECT_TRUE(a)

inline int RUN_ALL_TESTS()
{
    int bad = 0;
    for (auto& t : ::testing::Registry::items())
    {
        const int before = ::testing::Registry::failures();
        std::printf("[ RUN      ] %s\n", t.name.c_str());
        try
        {
            t.fn();
        }
        catch (const std::exception& e)
        {
            std::printf("  exception: %s\n", e.what());
            ++::testing::Registry::failures();
        }
        const bool ok = ::testing::Registry::failures() == before;
        std::printf("[ %s ] %s\n", ok ? "      OK" : " FAILED ", t.name.c_str());
        bad += ok ? 0 : 1;
    }
    std::printf("%d test(s), %d failed\n", (int) ::testing::Registry::items().size(), bad);
    return bad ? 1 : 0;
}
