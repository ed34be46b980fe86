// TEST INFRASTRUCTURE — oracle: reference HotSpot functor (examples/hotspot/hotspot.cpp:57-97) on the
// reference cpu backend. The example source is compiled in place; its main() is renamed.
#define main reference_hotspot_example_main
#include "examples/hotspot/hotspot.cpp"
#undef main
#include "harness.hpp"
#include <stst_workloads.h>

static_assert(sizeof(HotspotCell) == sizeof(stst_hotspot_cell));

ORACLE_EXPORT int oracle_ref_hotspot(ORACLE_REF_SIGNATURE) {
    const auto *p = static_cast<const stst_hotspot_params *>(params);
    HotspotKernel kernel{.Rx_1 = p->Rx_1, .Ry_1 = p->Ry_1, .Rz_1 = p->Rz_1, .Cap_1 = p->Cap_1};
    return oracle_ref::run_cpu_backend(kernel, oracle_ref::cell_or_default<HotspotCell>(halo),
                                       cells_in, cells_out, rows, cols, iteration_offset,
                                       n_iterations, window);
}
