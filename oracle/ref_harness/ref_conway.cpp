// TEST INFRASTRUCTURE — oracle: reference Conway functor (examples/conway/conway.cpp:35-56) on the
// reference cpu backend. The example source is compiled in place; its main() is renamed.
#define main reference_conway_example_main
#include "examples/conway/conway.cpp"
#undef main
#include "harness.hpp"

ORACLE_EXPORT int oracle_ref_conway(ORACLE_REF_SIGNATURE) {
    (void)params;
    return oracle_ref::run_cpu_backend(ConwayKernel(), oracle_ref::cell_or_default<bool>(halo),
                                       cells_in, cells_out, rows, cols, iteration_offset,
                                       n_iterations, window);
}
