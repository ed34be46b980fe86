/*
 * TEST INFRASTRUCTURE — parity oracle, never part of the product path.
 *
 * Shared driver for oracle/_ref/liboracle_ref.so: runs the REFERENCE's own
 * `stencil::cpu::StencilUpdate<F>` (/root/reference/StencilStream/cpu/StencilUpdate.hpp:40-229),
 * included unmodified from where it lies, on host arrays handed over through a C interface.
 * The only stand-in is the SYCL shim in stencilstream_b200/compat (no SYCL compiler exists here);
 * its `handler::parallel_for` is an OpenMP loop over rows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may load the
 * resulting library.
 */
#pragma once
#include <StencilStream/cpu/StencilUpdate.hpp>

#include "window.hpp"

#include <cstddef>
#include <cstring>

namespace oracle_ref {

template <typename F>
int run_cpu_backend(F transition_function, typename F::Cell halo_value, const void *cells_in,
                    void *cells_out, std::size_t rows, std::size_t cols,
                    std::size_t iteration_offset, std::size_t n_iterations) {
    using Cell = typename F::Cell;
    using Grid = stencil::cpu::Grid<Cell>;
    using Update = stencil::cpu::StencilUpdate<F>;

    Grid input(rows, cols);
    {
        typename Grid::template GridAccessor<sycl::access::mode::read_write> ac(input);
        std::memcpy(static_cast<void *>(ac.get_pointer()), cells_in, rows * cols * sizeof(Cell));
    }
    Update update({
        .transition_function = transition_function,
        .halo_value = halo_value,
        .iteration_offset = iteration_offset,
        .n_iterations = n_iterations,
        .blocking = true,
    });
    Grid output = update(input);
    {
        typename Grid::template GridAccessor<sycl::access::mode::read> ac(output);
        std::memcpy(cells_out, static_cast<const void *>(ac.get_pointer()),
                    rows * cols * sizeof(Cell));
    }
    return 0;
}

/**
 * Presents the reference backend's crop-local stencil to `Inner` with GLOBAL coordinates: same
 * cells, `id` shifted by the crop origin, `grid_range` = the whole grid. The reference cpu backend
 * itself stays unmodified; it substitutes `halo_value` beyond the crop, which is what the whole
 * grid would present where the crop ends at the grid's border and WRONG where the grid continues:
 * after n iterations only cells further than n * n_subiterations * radius from such a crop edge are
 * those of the whole-grid run (domain of dependence). The callers (tests/window_oracle.py) crop
 * with that margin and compare the remaining core only.
 */
template <typename Inner> struct Shifted {
    using Cell = typename Inner::Cell;
    using TimeDependentValue = typename Inner::TimeDependentValue;
    static constexpr std::size_t stencil_radius = Inner::stencil_radius;
    static constexpr std::size_t n_subiterations = Inner::n_subiterations;
    using StencilImpl = stencil::Stencil<Cell, stencil_radius, TimeDependentValue>;

    Inner inner;
    Window window;

    Cell operator()(StencilImpl const &local) const {
        StencilImpl global(sycl::id<2>(local.id[0] + window.row0, local.id[1] + window.col0),
                           sycl::range<2>(window.global_rows, window.global_cols), local.iteration,
                           local.subiteration, local.time_dependent_value);
        for (std::size_t r = 0; r < StencilImpl::diameter; r++)
            for (std::size_t c = 0; c < StencilImpl::diameter; c++)
                global[sycl::id<2>(r, c)] = local[sycl::id<2>(r, c)];
        return inner(global);
    }

    TimeDependentValue get_time_dependent_value(std::size_t i_iteration) const {
        return inner.get_time_dependent_value(i_iteration);
    }
};

/// run_cpu_backend on a crop (see Shifted); without an active window, the plain run.
template <typename F>
int run_cpu_backend(F transition_function, typename F::Cell halo_value, const void *cells_in,
                    void *cells_out, std::size_t rows, std::size_t cols,
                    std::size_t iteration_offset, std::size_t n_iterations, const Window *window) {
    if (!window || !window->active())
        return run_cpu_backend(transition_function, halo_value, cells_in, cells_out, rows, cols,
                               iteration_offset, n_iterations);
    return run_cpu_backend(Shifted<F>{transition_function, *window}, halo_value, cells_in, cells_out,
                           rows, cols, iteration_offset, n_iterations);
}

template <typename Cell> Cell cell_or_default(const void *halo) {
    Cell c = Cell();
    if (halo)
        std::memcpy(static_cast<void *>(&c), halo, sizeof(Cell));
    return c;
}

} // namespace oracle_ref

#define ORACLE_EXPORT extern "C" __attribute__((visibility("default")))

#define ORACLE_REF_SIGNATURE                                                                       \
    const void *params, const void *halo, const void *cells_in, void *cells_out, std::size_t rows, \
        std::size_t cols, std::size_t iteration_offset, std::size_t n_iterations,                  \
        const oracle_ref::Window *window
