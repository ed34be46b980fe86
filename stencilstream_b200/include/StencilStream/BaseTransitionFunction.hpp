/*
 * StencilStream-B200 — defaults for transition functions that use neither time-dependent values
 * nor sub-iterations nor a radius other than one.
 *
 * API-compatible with the reference's `stencil::BaseTransitionFunction`
 * (reference StencilStream/BaseTransitionFunction.hpp:40-81). `get_time_dependent_value` is
 * __host__ __device__; the B200 backend only ever calls it on the host, once per iteration
 * (reference semantics: cuda/StencilUpdate.hpp:224).
 */
#pragma once
#include "Concepts.hpp"
#include <variant>

namespace stencil {

class BaseTransitionFunction {
  public:
    /// `std::monostate` has exactly one value: the TDV feature is switched off.
    using TimeDependentValue = std::monostate;

    /// A radius of one yields the 3x3 Moore neighbourhood.
    static constexpr std::size_t stencil_radius = 1;

    /// One sweep over the grid per iteration.
    static constexpr std::size_t n_subiterations = 1;

    STST_HD constexpr std::monostate get_time_dependent_value(std::size_t) const { return {}; }
};

} // namespace stencil
