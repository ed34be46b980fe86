// TEST INFRASTRUCTURE — oracle: reference FDTD functor Kernel<CoefResolver>
// (examples/fdtd/src/Kernel.hpp:52-141, material/CoefResolver.hpp) on the reference cpu backend.
// The example source is compiled in place (MATERIAL=0 -> coef resolver); its main() is renamed.
//
// The functor keeps its derived constants in private members that only its constructor (fed by the
// JSON-driven Parameters class) can set. The oracle needs to set them from the C parameter block, so
// `private` is made public for the reference sources — after every system/third-party header they
// use has already been included, so that only the reference's own classes are affected.
#include <nlohmann/json.hpp>
#include <sycl/ext/intel/ac_types/ac_int.hpp>
#include <sycl/sycl.hpp>

#include <bit>
#include <cmath>
#include <deque>
#include <fstream>
#include <optional>
#include <unistd.h>

#define MATERIAL 0
#define TDVS_TYPE 0
#define private public
#define main reference_fdtd_example_main
#include "examples/fdtd/src/fdtd.cpp"
#undef main
#undef private
#include "harness.hpp"
#include <stst_workloads.h>

#include <new>

static_assert(sizeof(CellImpl) == sizeof(stst_fdtd_cell));

namespace {
KernelImpl make_kernel(const stst_fdtd_params &p) {
    // KernelImpl is trivially copyable; build it member by member from the parameter block.
    alignas(KernelImpl) unsigned char raw[sizeof(KernelImpl)] = {};
    KernelImpl *k = reinterpret_cast<KernelImpl *>(raw);
    k->dt = p.dt;
    k->t_0 = p.t_0;
    k->tau = p.tau;
    k->omega = p.omega;
    k->cutoff_iteration = p.cutoff_iteration;
    k->detect_iteration = p.detect_iteration;
    k->source_radius_squared = p.source_radius_squared;
    k->source_r = p.source_r;
    k->source_c = p.source_c;
    k->source_distance_bound = p.source_distance_bound;
    k->double_center_rc = p.double_center_rc;
    return *k;
}
} // namespace

ORACLE_EXPORT int oracle_ref_fdtd(ORACLE_REF_SIGNATURE) {
    KernelImpl kernel = make_kernel(*static_cast<const stst_fdtd_params *>(params));
    return oracle_ref::run_cpu_backend(kernel, oracle_ref::cell_or_default<CellImpl>(halo),
                                       cells_in, cells_out, rows, cols, iteration_offset,
                                       n_iterations, window);
}

// Derive the functor constants, grid size and step counts from an experiment JSON file exactly as
// the reference does (Parameters.hpp + Kernel.hpp:64-78), for checking the Python-side derivation.
ORACLE_EXPORT int oracle_ref_fdtd_derive(const char *json_path, stst_fdtd_params *out,
                                      std::size_t *grid_wh, std::size_t *n_timesteps,
                                      std::size_t *n_snap_timesteps) {
    std::string path(json_path);
    char arg0[] = "oracle", arg1[] = "-c";
    char *argv[] = {arg0, arg1, path.data(), nullptr};
    optind = 1;
    Parameters parameters(3, argv);
    KernelImpl k(parameters, MaterialResolver(parameters));
    out->dt = k.dt;
    out->t_0 = k.t_0;
    out->tau = k.tau;
    out->omega = k.omega;
    out->cutoff_iteration = k.cutoff_iteration;
    out->detect_iteration = k.detect_iteration;
    out->source_radius_squared = k.source_radius_squared;
    out->source_r = k.source_r;
    out->source_c = k.source_c;
    out->source_distance_bound = k.source_distance_bound;
    out->double_center_rc = k.double_center_rc;
    *grid_wh = parameters.grid_range()[0];
    *n_timesteps = parameters.n_timesteps();
    *n_snap_timesteps = parameters.n_snap_timesteps().value_or(0);
    return 0;
}

// Initial grid exactly as examples/fdtd/src/fdtd.cpp:193-216 builds it (that loop lives inside the
// example's main(), so it is restated here against the reference's own Parameters/CellImpl types).
ORACLE_EXPORT int oracle_ref_fdtd_initial_grid(const char *json_path, void *cells_out,
                                            std::size_t n_cells) {
    std::string path(json_path);
    char arg0[] = "oracle", arg1[] = "-c";
    char *argv[] = {arg0, arg1, path.data(), nullptr};
    optind = 1;
    Parameters parameters(3, argv);
    const std::size_t wh = parameters.grid_range()[0];
    if (n_cells != wh * wh)
        return -1;
    CellImpl *cells = static_cast<CellImpl *>(cells_out);
    for (size_t r = 0; r < parameters.grid_range()[1]; r++) {
        for (size_t c = 0; c < parameters.grid_range()[0]; c++) {
            float a = float(r) - float(parameters.grid_range()[0]) / 2.0;
            float b = float(c) - float(parameters.grid_range()[1]) / 2.0;
            float distance = parameters.dx * std::sqrt(a * a + b * b);
            float radius = 0.0;
            for (size_t i = 0; i <= parameters.rings.size(); i++) {
                if (i < parameters.rings.size()) {
                    radius += parameters.rings[i].radius;
                    if (distance < radius) {
                        cells[r * wh + c] = CellImpl::from_parameters(parameters, i);
                        break;
                    }
                } else {
                    cells[r * wh + c] = CellImpl::from_parameters(parameters, i);
                }
            }
        }
    }
    return 0;
}
