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
        std::size_t cols, std::size_t iteration_offset, std::size_t n_iterations
