/*
 * StencilStream-B200 — the concepts that tie transition functions, grids and updaters together.
 *
 * Same names and requirements as the reference (StencilStream/Concepts.hpp:61-77 TransitionFunction,
 * :85-89 GridAccessor, :114-129 Grid, :157-172 StencilUpdate) so that user code and the reference's
 * own `static_assert(concepts::Grid<...>)` checks (tests/cuda/Grid.cpp:28) hold for this backend.
 */
#pragma once
#include "Stencil.hpp"
#include <concepts>
#include <sycl.hpp>
#include <type_traits>

namespace stencil {
namespace concepts {

/**
 * A transition function provides the types `Cell` (semiregular) and `TimeDependentValue`
 * (copyable), the `std::size_t` constants `stencil_radius >= 1` and `n_subiterations >= 1`, a pure
 * call operator mapping a stencil to the central cell's next value, and a pure
 * `get_time_dependent_value(i_iteration)`.
 */
template <typename T>
concept TransitionFunction =
    std::semiregular<typename T::Cell> && std::copyable<typename T::TimeDependentValue> &&
    std::same_as<decltype(T::stencil_radius), const std::size_t> && (T::stencil_radius >= 1) &&
    std::same_as<decltype(T::n_subiterations), const std::size_t> && (T::n_subiterations >= 1) &&
    requires(T const &tf, std::size_t i_iteration,
             Stencil<typename T::Cell, T::stencil_radius, typename T::TimeDependentValue> const &st) {
        { tf(st) } -> std::same_as<typename T::Cell>;
        {
            tf.get_time_dependent_value(i_iteration)
        } -> std::same_as<typename T::TimeDependentValue>;
    };

/// Host-side element access, either `ac[sycl::id<2>]` or `ac[r][c]`.
template <typename Accessor, typename Cell>
concept GridAccessor = requires(Accessor ac, std::size_t r, std::size_t c) {
    { ac[sycl::id<2>(r, c)] } -> std::same_as<Cell &>;
    { ac[r][c] } -> std::same_as<Cell &>;
};

/**
 * A regular two-dimensional grid of cells: constructible from (rows, columns), a `sycl::range<2>`
 * or a `sycl::buffer<Cell, 2>`; copyable from/to such buffers; reports its extent; can create an
 * equally-sized sibling; and exposes a `GridAccessor<mode>` class template for host access.
 */
template <typename G, typename Cell>
concept Grid =
    requires(G &grid, sycl::buffer<Cell, 2> buffer, std::size_t r, std::size_t c, Cell cell) {
        { G(r, c) } -> std::same_as<G>;
        { G(sycl::range<2>(r, c)) } -> std::same_as<G>;
        { G(buffer) } -> std::same_as<G>;
        { grid.copy_from_buffer(buffer) } -> std::same_as<void>;
        { grid.copy_to_buffer(buffer) } -> std::same_as<void>;
        { grid.get_grid_height() } -> std::convertible_to<std::size_t>;
        { grid.get_grid_width() } -> std::convertible_to<std::size_t>;
        { grid.get_grid_range() } -> std::convertible_to<sycl::range<2>>;
        { grid.make_similar() } -> std::same_as<G>;
        {
            typename G::template GridAccessor<sycl::access::mode::read_write>(grid)
        } -> GridAccessor<Cell>;
    };

/**
 * A grid updater: constructed from its nested `Params` aggregate (which must carry
 * `transition_function`, `halo_value`, `iteration_offset`, `n_iterations`), hands out a live
 * reference to those parameters, and maps an input grid to a new, updated grid.
 */
template <typename SU, typename TF, typename G>
concept StencilUpdate =
    TransitionFunction<TF> && Grid<G, typename TF::Cell> &&
    std::is_class_v<typename SU::Params> &&
    requires(SU stencil_update, G &grid, typename SU::Params params) {
        { SU(params) } -> std::same_as<SU>;
        { stencil_update.get_params() } -> std::same_as<typename SU::Params &>;
        { stencil_update(grid) } -> std::same_as<G>;
        { params.transition_function } -> std::same_as<TF &>;
        { params.halo_value } -> std::same_as<typename TF::Cell &>;
        { params.iteration_offset } -> std::same_as<std::size_t &>;
        { params.n_iterations } -> std::same_as<std::size_t &>;
    };

} // namespace concepts
} // namespace stencil
