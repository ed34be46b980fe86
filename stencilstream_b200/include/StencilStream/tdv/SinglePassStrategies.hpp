/*
 * StencilStream-B200 — time-dependent-value (TDV) strategies for single-pass updaters.
 *
 * API counterpart of the reference's StencilStream/tdv/SinglePassStrategies.hpp (concepts :47-101,
 * InlineStrategy :114-144, PrecomputeOnDeviceStrategy :155-192, PrecomputeOnHostStrategy :203-264).
 * In the reference these strategies parameterise the FPGA backends; its cpu/cuda backends ignore them
 * and call `get_time_dependent_value` on the host once per sweep (cuda/StencilUpdate.hpp:224). User
 * code nevertheless names them (examples/fdtd/src/fdtd.cpp:39-45), so the three strategy types and the
 * four concepts exist here with the same shapes.
 *
 * On this backend one scheme serves all three names, the one the fused kernel needs anyway: the host
 * evaluates the values of the requested iterations once (`GlobalState`), every launch receives the
 * window of its fused iterations BY VALUE as a kernel parameter (`KernelArgument`, which is its own
 * `LocalState`; cf. internal::TdvArray in cuda/internal/TileKernel.hpp). `InlineStrategy` alone keeps
 * the reference's meaning — evaluate on use — and is therefore only device-callable if the user's
 * `get_time_dependent_value` is `STST_HD`.
 */
#pragma once
#include "../Concepts.hpp"

#include <sycl/sycl.hpp>

#include <algorithm>
#include <concepts>
#include <cstddef>
#include <vector>

namespace stencil {
namespace tdv {
namespace single_pass {

/// Pass-local state: answers `get_time_dependent_value(i)` for the pass-local iteration i.
template <typename T, typename TransFunc>
concept LocalState = stencil::concepts::TransitionFunction<TransFunc> &&
                     requires(T const &state, std::size_t i) {
                         {
                             state.get_time_dependent_value(i)
                         } -> std::same_as<typename TransFunc::TimeDependentValue>;
                     };

/// What the host hands to a launch; the launch builds its LocalState from it.
template <typename T, typename TransFunc>
concept KernelArgument = stencil::concepts::TransitionFunction<TransFunc> && std::copyable<T> &&
                         LocalState<typename T::LocalState, TransFunc> &&
                         std::constructible_from<typename T::LocalState, T const &>;

/// Host-resident state for one `operator()` call: built from (functor, iteration offset, iterations).
template <typename T, typename TransFunc>
concept GlobalState = stencil::concepts::TransitionFunction<TransFunc> &&
                      std::constructible_from<T, TransFunc, std::size_t, std::size_t> &&
                      KernelArgument<typename T::KernelArgument, TransFunc> &&
                      std::constructible_from<typename T::KernelArgument, T &, sycl::handler &,
                                              std::size_t, std::size_t>;

/// A strategy is a GlobalState template over (functor, iterations per pass).
template <typename T, typename TransFunc, std::size_t max_n_iterations>
concept Strategy =
    stencil::concepts::TransitionFunction<TransFunc> &&
    GlobalState<typename T::template GlobalState<TransFunc, max_n_iterations>, TransFunc>;

namespace detail {

/// Host precompute -> per-pass window shipped by value.
template <stencil::concepts::TransitionFunction TransFunc, std::size_t max_n_iterations>
class WindowedGlobalState {
  public:
    using TDV = typename TransFunc::TimeDependentValue;

    WindowedGlobalState(TransFunc trans_func, std::size_t iteration_offset, std::size_t n_iterations)
        : first_iteration(iteration_offset), values(n_iterations) {
        for (std::size_t i = 0; i < n_iterations; i++)
            values[i] = trans_func.get_time_dependent_value(iteration_offset + i);
    }

    class KernelArgument {
      public:
        using LocalState = KernelArgument;

        /// Window of the pass that starts at global iteration `i_iteration`.
        KernelArgument(WindowedGlobalState &global_state, sycl::handler &, std::size_t i_iteration,
                       std::size_t n_iterations)
            : window{} {
            const std::size_t begin = i_iteration >= global_state.first_iteration
                                          ? i_iteration - global_state.first_iteration
                                          : 0;
            const std::size_t available =
                begin < global_state.values.size() ? global_state.values.size() - begin : 0;
            const std::size_t count = std::min({n_iterations, max_n_iterations, available});
            for (std::size_t i = 0; i < count; i++)
                window[i] = global_state.values[begin + i];
        }

        STST_HD TDV get_time_dependent_value(std::size_t i) const { return window[i]; }

      private:
        TDV window[max_n_iterations];
    };

  private:
    std::size_t first_iteration;
    std::vector<TDV> values;
};

} // namespace detail

/// Evaluate the TDV function where it is used (no precomputation).
struct InlineStrategy {
    template <stencil::concepts::TransitionFunction TransFunc, std::size_t max_n_iterations>
    class GlobalState {
      public:
        using TDV = typename TransFunc::TimeDependentValue;

        GlobalState(TransFunc trans_func, std::size_t, std::size_t) : trans_func(trans_func) {}

        class KernelArgument {
          public:
            using LocalState = KernelArgument;

            KernelArgument(GlobalState &global_state, sycl::handler &, std::size_t i_iteration,
                           std::size_t)
                : trans_func(global_state.trans_func), first_iteration(i_iteration) {}

            KernelArgument(GlobalState &global_state, std::size_t i_iteration)
                : trans_func(global_state.trans_func), first_iteration(i_iteration) {}

            STST_HD TDV get_time_dependent_value(std::size_t i) const {
                return trans_func.get_time_dependent_value(first_iteration + i);
            }

          private:
            TransFunc trans_func;
            std::size_t first_iteration;
        };

      private:
        TransFunc trans_func;
    };
};

/// Precompute per pass. On this backend: host precompute, window by value (see file comment).
struct PrecomputeOnDeviceStrategy {
    template <stencil::concepts::TransitionFunction TransFunc, std::size_t max_n_iterations>
    using GlobalState = detail::WindowedGlobalState<TransFunc, max_n_iterations>;
};

/// Precompute all requested iterations on the host, ship a window per pass.
struct PrecomputeOnHostStrategy {
    template <stencil::concepts::TransitionFunction TransFunc, std::size_t max_n_iterations>
    using GlobalState = detail::WindowedGlobalState<TransFunc, max_n_iterations>;
};

} // namespace single_pass
} // namespace tdv
} // namespace stencil
