/*
 * StencilStream-B200 — backend-independent helpers.
 *
 * Counterpart of the reference's StencilStream/internal/Helpers.hpp, of which only two things are
 * reachable from the cuda hot path: `int_ceil_div` (reference Helpers.hpp:46-48, used at
 * cuda/StencilUpdate.hpp:210) and the STENCILSTREAM_NAMED_* kernel-naming macros
 * (reference Helpers.hpp:24-30). The FPGA pipe-word plumbing of that file is out of scope.
 */
#pragma once
#include <cstddef>
#include <cstdint>
#include <sycl/sycl.hpp>

// Kernel-naming macros: kept for source compatibility with user code that targets several backends.
#if defined(STENCILSTREAM_NAMED_KERNELS)
    #define STENCILSTREAM_NAMED_SINGLE_TASK(Name, argument) single_task<class Name>(argument)
    #define STENCILSTREAM_NAMED_PARALLEL_FOR(Name, range, kernel)                                  \
        parallel_for<class Name>(range, kernel)
#else
    #define STENCILSTREAM_NAMED_SINGLE_TASK(Name, argument) single_task(argument)
    #define STENCILSTREAM_NAMED_PARALLEL_FOR(Name, range, kernel) parallel_for(range, kernel)
#endif

namespace stencil {
namespace internal {

/// Integer division that rounds towards +inf (for non-negative operands).
template <typename T> STST_HD inline constexpr T int_ceil_div(T a, T b) {
    return (a + b - T(1)) / b;
}

/// Round `a` up to the next multiple of `b`.
template <typename T> STST_HD inline constexpr T round_up(T a, T b) {
    return int_ceil_div<T>(a, b) * b;
}

} // namespace internal
} // namespace stencil
