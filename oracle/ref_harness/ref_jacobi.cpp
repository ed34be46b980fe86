// TEST INFRASTRUCTURE — oracle: reference Jacobi functors (examples/jacobi/kernels.hpp) on the
// reference cpu backend. examples/jacobi/jacobi.cpp has no CPU branch (:22-35), hence this driver.
// The radius-2/3 star variants have no reference functor; the B200 repo's own functor definition
// (stencilstream_b200/csrc/workloads/functors.hpp, host-compiled here against the REFERENCE's
// Stencil/BaseTransitionFunction headers) is run through the reference's radius-generic cpu backend.
#include <sycl/sycl.hpp>
#define JACOBI_KERNEL Jacobi5General
#include "examples/jacobi/kernels.hpp"
#include "harness.hpp"
#include <stst_workloads.h>
#include "../../stencilstream_b200/csrc/workloads/functors.hpp"

#include <cstdio>
#include <string>
#include <vector>

void print_usage(int, char **) { throw std::invalid_argument("jacobi oracle: bad argument count"); }

namespace {
// The reference functors take their coefficients as argv strings parsed with atof()
// (kernels.hpp:257-263). "%.9g" round-trips every float exactly through that path.
template <typename Kernel> Kernel make_from_coefficients(const float *coef, int n) {
    std::vector<std::string> storage = {"oracle", "0", "0", "0", "out"};
    for (int i = 0; i < n; i++) {
        char text[64];
        std::snprintf(text, sizeof(text), "%.9g", double(coef[i]));
        storage.emplace_back(text);
    }
    std::vector<char *> argv;
    for (auto &s : storage)
        argv.push_back(s.data());
    return Kernel(int(argv.size()), argv.data());
}
} // namespace

ORACLE_EXPORT int oracle_ref_jacobi5(ORACLE_REF_SIGNATURE) {
    const auto *p = static_cast<const stst_jacobi5_params *>(params);
    Jacobi5General kernel = make_from_coefficients<Jacobi5General>(p->coef, 5);
    return oracle_ref::run_cpu_backend(kernel, oracle_ref::cell_or_default<float>(halo), cells_in,
                                       cells_out, rows, cols, iteration_offset, n_iterations, window);
}

ORACLE_EXPORT int oracle_ref_jacobi9(ORACLE_REF_SIGNATURE) {
    const auto *p = static_cast<const stst_jacobi9_params *>(params);
    Jacobi9General kernel = make_from_coefficients<Jacobi9General>(&p->coef[0][0], 9);
    return oracle_ref::run_cpu_backend(kernel, oracle_ref::cell_or_default<float>(halo), cells_in,
                                       cells_out, rows, cols, iteration_offset, n_iterations, window);
}

ORACLE_EXPORT int oracle_ref_jacobi_r2(ORACLE_REF_SIGNATURE) {
    stst_workloads::JacobiStarRule<2> rule;
    rule.p = *static_cast<const stst_jacobi_star_params *>(params);
    return oracle_ref::run_cpu_backend(rule, oracle_ref::cell_or_default<float>(halo), cells_in,
                                       cells_out, rows, cols, iteration_offset, n_iterations, window);
}

ORACLE_EXPORT int oracle_ref_jacobi_r3(ORACLE_REF_SIGNATURE) {
    stst_workloads::JacobiStarRule<3> rule;
    rule.p = *static_cast<const stst_jacobi_star_params *>(params);
    return oracle_ref::run_cpu_backend(rule, oracle_ref::cell_or_default<float>(halo), cells_in,
                                       cells_out, rows, cols, iteration_offset, n_iterations, window);
}
