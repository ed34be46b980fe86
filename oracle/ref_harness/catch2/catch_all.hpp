// TEST INFRASTRUCTURE — stand-in so that /root/reference/tests/TransFuncs.hpp (which includes Catch2,
// not installed here) can be compiled for the oracle. Only the host-only HostTransFunc uses REQUIRE.
#pragma once
#include <stdexcept>
#define REQUIRE(cond)                                                                              \
    do {                                                                                           \
        if (!(cond))                                                                               \
            throw std::logic_error("REQUIRE failed: " #cond);                                      \
    } while (0)
