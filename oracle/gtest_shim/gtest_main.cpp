// TEST INFRASTRUCTURE: main() for the gtest shim (the reference links GTest::gtest_main).
#include "gtest/gtest.h"
int main(int argc, char** argv) {
    ::testing::InitGoogleTest(&argc, argv);
    return RUN_ALL_TESTS();
}
