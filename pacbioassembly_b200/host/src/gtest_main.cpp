// main() for the test drivers (the reference links GoogleTest's gtest_main; CMakeLists.txt:54-56)
#include <gtest/gtest.h>

#include "pb_runtime.hpp"

int main(int argc, char **argv)
{
    ::testing::InitGoogleTest(&argc, argv);
    int rc = RUN_ALL_TESTS();
    pb::shutdown();
    return rc;
}
