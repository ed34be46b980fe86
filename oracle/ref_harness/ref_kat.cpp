// TEST INFRASTRUCTURE — oracle: the reference's self-checking test functor FPGATransFunc<radius>
// (tests/TransFuncs.hpp:55-104) on the reference cpu backend, as tests/StencilUpdateTest.hpp:30-63
// drives it. Catch2 is replaced by the one-macro stub in oracle/ref_harness/catch2.
#include "tests/TransFuncs.hpp"
#include "harness.hpp"
#include <stst_workloads.h>

static_assert(sizeof(Cell) == sizeof(stst_kat_cell));

ORACLE_EXPORT int oracle_ref_kat(ORACLE_REF_SIGNATURE) {
    (void)params;
    return oracle_ref::run_cpu_backend(FPGATransFunc<1>(), oracle_ref::cell_or_default<Cell>(halo),
                                       cells_in, cells_out, rows, cols, iteration_offset,
                                       n_iterations, window);
}

ORACLE_EXPORT int oracle_ref_kat_r2(ORACLE_REF_SIGNATURE) {
    (void)params;
    return oracle_ref::run_cpu_backend(FPGATransFunc<2>(), oracle_ref::cell_or_default<Cell>(halo),
                                       cells_in, cells_out, rows, cols, iteration_offset,
                                       n_iterations, window);
}
