// test/main.cpp — the reference's entry point (test/main.cpp:1-7): one 4096 x 4096 self-check.
#include "tester.hpp"

int main()
{
    SparseSgemvTester harness(4096, 4096);
    harness.RunTest();
    return 0;
}
