/*
 * StencilStream-B200 — the neighbourhood buffer handed to a transition function.
 *
 * API-compatible restatement of the reference's `stencil::Stencil<Cell, radius, TDV>`
 * (reference StencilStream/Stencil.hpp:45-181): same template parameters, constructors, the two
 * indexing schemes (signed, centre-origin `stencil[r][c]`; unsigned, north-west-origin
 * `stencil[sycl::id<2>]`) and the same public, immutable metadata members. Every member is
 * __host__ __device__: instances live entirely in registers inside the sm_100a sweep kernels
 * (cuda/internal/TileKernel.hpp) and are also usable on the host (unit tests, oracle).
 */
#pragma once
#include "internal/Helpers.hpp"
#include <concepts>
#include <limits>
#include <sycl/id.hpp>
#include <sycl/range.hpp>
#include <variant>

namespace stencil {

template <typename Cell, std::size_t stencil_radius, typename TimeDependentValue = std::monostate>
    requires std::semiregular<Cell> && (stencil_radius >= 1)
class Stencil {
  public:
    /// Edge length of the (square) neighbourhood: `2 * stencil_radius + 1`.
    static constexpr std::size_t diameter = 2 * stencil_radius + 1;
    static_assert(diameter <= std::size_t(std::numeric_limits<int>::max()));

    /**
     * Metadata-only constructor; the cells are value-initialised and meant to be filled through
     * `operator[](sycl::id<2>)`.
     *
     * \param id Global (row, column) position of the central cell.
     * \param grid_range Global (height, width) of the grid the cell belongs to.
     * \param iteration Iteration index of the cells held by the stencil.
     * \param subiteration Sub-iteration index of the cells held by the stencil.
     * \param tdv The time-dependent value of `iteration`.
     */
    STST_HD Stencil(sycl::id<2> id, sycl::range<2> grid_range, std::size_t iteration,
                    std::size_t subiteration, TimeDependentValue tdv)
        : id(id), iteration(iteration), subiteration(subiteration), grid_range(grid_range),
          time_dependent_value(tdv), cells() {}

    /// Same as above, additionally copying the neighbourhood out of `raw`.
    STST_HD Stencil(sycl::id<2> id, sycl::range<2> grid_range, std::size_t iteration,
                    std::size_t subiteration, TimeDependentValue tdv,
                    Cell raw[diameter][diameter])
        : id(id), iteration(iteration), subiteration(subiteration), grid_range(grid_range),
          time_dependent_value(tdv), cells() {
#pragma unroll
        for (std::size_t i = 0; i < diameter * diameter; i++) {
            cells[i / diameter][i % diameter] = raw[i / diameter][i % diameter];
        }
    }

    /**
     * Proxy produced by the signed row subscript; its own subscript selects the column. Offsets
     * are relative to the central cell and lie in `[-stencil_radius, +stencil_radius]`.
     */
    template <std::signed_integral index_t>
        requires(stencil_radius <= std::size_t(std::numeric_limits<index_t>::max()))
    class StencilSubscript {
      public:
        STST_HD StencilSubscript(Stencil const &stencil, index_t r) : stencil(stencil), r(r) {}

        STST_HD Cell const &operator[](index_t c) const {
            return stencil.cells[r + index_t(stencil_radius)][c + index_t(stencil_radius)];
        }

      private:
        Stencil const &stencil;
        index_t r;
    };

    /// Signed, centre-origin access: `stencil[dr][dc]`.
    template <std::signed_integral index_t>
    STST_HD StencilSubscript<index_t> operator[](index_t r) const
        requires(stencil_radius <= std::size_t(std::numeric_limits<index_t>::max()))
    {
        return StencilSubscript<index_t>(*this, r);
    }

    /// Unsigned access; `(0, 0)` is the north-western corner of the neighbourhood.
    STST_HD Cell const &operator[](sycl::id<2> uid) const { return cells[uid[0]][uid[1]]; }

    /// Unsigned, mutable access; `(0, 0)` is the north-western corner of the neighbourhood.
    STST_HD Cell &operator[](sycl::id<2> uid) { return cells[uid[0]][uid[1]]; }

    /// Global position of the central cell.
    const sycl::id<2> id;
    /// Iteration index of the cells in the stencil.
    const std::size_t iteration;
    /// Sub-iteration index of the cells in the stencil.
    const std::size_t subiteration;
    /// Global range of the grid.
    const sycl::range<2> grid_range;
    /// Time-dependent value of the current iteration.
    const TimeDependentValue time_dependent_value;

  private:
    Cell cells[diameter][diameter];
};

} // namespace stencil
