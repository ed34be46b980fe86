// TEST INFRASTRUCTURE — stand-in for Catch2 v3 (not installed here), just large enough for the
// reference's own test sources (/root/reference/tests/*.cpp, *.hpp) to compile and run unmodified:
// TEST_CASE registers a function, REQUIRE throws on failure, run_all_test_cases() is the runner.
// Used (a) by the oracle build, which includes tests/TransFuncs.hpp, and (b) by
// stencilstream_b200/tools/build_reference_tests.py, which builds the reference's cuda unit tests
// against this backend.
#pragma once
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace catch_standin {
struct TestCase {
    const char *name;
    void (*fn)();
};
inline std::vector<TestCase> &registry() {
    static std::vector<TestCase> cases;
    return cases;
}
struct Registrar {
    Registrar(const char *name, void (*fn)()) { registry().push_back({name, fn}); }
};
/// Runs every registered test case; returns the number of failed ones.
inline int run_all_test_cases() {
    int failed = 0;
    for (auto const &tc : registry()) {
        try {
            tc.fn();
            std::printf("[ OK ] %s\n", tc.name);
        } catch (std::exception const &e) {
            failed++;
            std::printf("[FAIL] %s: %s\n", tc.name, e.what());
        }
    }
    std::printf("%zu test case(s), %d failed\n", registry().size(), failed);
    return failed;
}
} // namespace catch_standin

#define CATCH_STANDIN_CAT2(a, b) a##b
#define CATCH_STANDIN_CAT(a, b) CATCH_STANDIN_CAT2(a, b)
#define CATCH_STANDIN_TEST_CASE(fn, ...)                                                           \
    static void fn();                                                                              \
    static ::catch_standin::Registrar CATCH_STANDIN_CAT(fn, _registrar)(                           \
        CATCH_STANDIN_FIRST(__VA_ARGS__), &fn);                                                    \
    static void fn()
#define CATCH_STANDIN_FIRST(first, ...) first
#define TEST_CASE(...) CATCH_STANDIN_TEST_CASE(CATCH_STANDIN_CAT(catch_standin_test_, __COUNTER__), __VA_ARGS__)

#define REQUIRE(cond)                                                                              \
    do {                                                                                           \
        if (!(cond))                                                                               \
            throw std::logic_error(std::string("REQUIRE failed: " #cond " (") + __FILE__ + ":" +   \
                                   std::to_string(__LINE__) + ")");                                \
    } while (0)
