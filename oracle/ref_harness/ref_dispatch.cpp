// TEST INFRASTRUCTURE — oracle: name-based entry point of oracle/_ref/liboracle_ref.so.
#include <cstddef>
#include <cstring>
#include <exception>
#include <string>

#include "window.hpp"

#define ORACLE_EXPORT extern "C" __attribute__((visibility("default")))

#define ORACLE_REF_SIGNATURE                                                                       \
    const void *params, const void *halo, const void *cells_in, void *cells_out, std::size_t rows, \
        std::size_t cols, std::size_t iteration_offset, std::size_t n_iterations,                  \
        const oracle_ref::Window *window

extern "C" {
int oracle_ref_conway(ORACLE_REF_SIGNATURE);
int oracle_ref_jacobi5(ORACLE_REF_SIGNATURE);
int oracle_ref_jacobi9(ORACLE_REF_SIGNATURE);
int oracle_ref_jacobi_r2(ORACLE_REF_SIGNATURE);
int oracle_ref_jacobi_r3(ORACLE_REF_SIGNATURE);
int oracle_ref_hotspot(ORACLE_REF_SIGNATURE);
int oracle_ref_fdtd(ORACLE_REF_SIGNATURE);
int oracle_ref_convection_pt(ORACLE_REF_SIGNATURE);
int oracle_ref_convection_thermal(ORACLE_REF_SIGNATURE);
int oracle_ref_kat(ORACLE_REF_SIGNATURE);
int oracle_ref_kat_r2(ORACLE_REF_SIGNATURE);
}

namespace {
thread_local std::string g_error;
struct Entry {
    const char *name;
    int (*fn)(ORACLE_REF_SIGNATURE);
};
const Entry entries[] = {
    {"conway", oracle_ref_conway},
    {"jacobi5", oracle_ref_jacobi5},
    {"jacobi9", oracle_ref_jacobi9},
    {"jacobi_r2", oracle_ref_jacobi_r2},
    {"jacobi_r3", oracle_ref_jacobi_r3},
    {"hotspot", oracle_ref_hotspot},
    {"fdtd", oracle_ref_fdtd},
    {"convection_pt", oracle_ref_convection_pt},
    {"convection_thermal", oracle_ref_convection_thermal},
    {"kat", oracle_ref_kat},
    {"kat_r2", oracle_ref_kat_r2},
};
} // namespace

ORACLE_EXPORT const char *oracle_last_error(void) { return g_error.c_str(); }

ORACLE_EXPORT const char *oracle_kind(void) { return "reference"; }

namespace {
int dispatch(const char *workload, ORACLE_REF_SIGNATURE) {
    for (auto const &e : entries) {
        if (std::strcmp(e.name, workload) == 0) {
            try {
                return e.fn(params, halo, cells_in, cells_out, rows, cols, iteration_offset,
                            n_iterations, window);
            } catch (std::exception const &ex) {
                g_error = ex.what();
                return -4;
            }
        }
    }
    g_error = std::string("unknown workload: ") + workload;
    return -1;
}
} // namespace

/// Run `n_iterations` of the named workload on the reference cpu backend. Returns 0 on success.
ORACLE_EXPORT int oracle_run(const char *workload, const void *params, const void *halo,
                             const void *cells_in, void *cells_out, std::size_t rows,
                             std::size_t cols, std::size_t iteration_offset,
                             std::size_t n_iterations) {
    return dispatch(workload, params, halo, cells_in, cells_out, rows, cols, iteration_offset,
                    n_iterations, nullptr);
}

/// The same on a crop of a larger grid: `cells_in` holds rows [row0, row0 + rows) x columns
/// [col0, col0 + cols) of a global_rows x global_cols grid; transition functions see global
/// coordinates (oracle_ref::Shifted in harness.hpp states which output cells are exact).
ORACLE_EXPORT int oracle_run_window2d(const char *workload, const void *params, const void *halo,
                                      const void *cells_in, void *cells_out, std::size_t rows,
                                      std::size_t cols, std::size_t row0, std::size_t col0,
                                      std::size_t global_rows, std::size_t global_cols,
                                      std::size_t iteration_offset, std::size_t n_iterations) {
    if (row0 + rows > global_rows || col0 + cols > global_cols) {
        g_error = "window exceeds the grid";
        return -2;
    }
    const oracle_ref::Window window{row0, col0, global_rows, global_cols};
    return dispatch(workload, params, halo, cells_in, cells_out, rows, cols, iteration_offset,
                    n_iterations, &window);
}
