// TEST INFRASTRUCTURE — oracle: reference mantle-convection functors
// (examples/convection/convection.cpp:76-242) on the reference cpu backend. The example source is
// compiled in place; its main() is renamed.
#define main reference_convection_example_main
#include "examples/convection/convection.cpp"
#undef main
#include "harness.hpp"
#include <stst_workloads.h>

static_assert(sizeof(ThermalConvectionCell) == sizeof(stst_convection_cell));

ORACLE_EXPORT int oracle_ref_convection_pt(ORACLE_REF_SIGNATURE) {
    const auto *p = static_cast<const stst_convection_pt_params *>(params);
    PseudoTransientKernel kernel{
        .nx = p->nx,
        .ny = p->ny,
        .roh0_g_alpha = p->roh0_g_alpha,
        .delta_eta_delta_T = p->delta_eta_delta_T,
        .eta0 = p->eta0,
        .deltaT = p->deltaT,
        .dx = p->dx,
        .dy = p->dy,
        .delta_tau_iter = p->delta_tau_iter,
        .beta = p->beta,
        .rho = p->rho,
        .dampX = p->dampX,
        .dampY = p->dampY,
        .DcT = p->DcT,
    };
    return oracle_ref::run_cpu_backend(kernel,
                                       oracle_ref::cell_or_default<ThermalConvectionCell>(halo),
                                       cells_in, cells_out, rows, cols, iteration_offset,
                                       n_iterations, window);
}

ORACLE_EXPORT int oracle_ref_convection_thermal(ORACLE_REF_SIGNATURE) {
    const auto *p = static_cast<const stst_convection_thermal_params *>(params);
    ThermalSolverKernel kernel{
        .nx = p->nx, .ny = p->ny, .dx = p->dx, .dy = p->dy, .dt = p->dt, .DcT = p->DcT};
    return oracle_ref::run_cpu_backend(kernel,
                                       oracle_ref::cell_or_default<ThermalConvectionCell>(halo),
                                       cells_in, cells_out, rows, cols, iteration_offset,
                                       n_iterations, window);
}
