/*
 * libstst_workloads — instantiates the header-only B200 backend (StencilStream/cuda/*.hpp) for the
 * "light" transition functions of workloads/functors.hpp. One translation unit per group so that
 * nvcc compiles the kernel templates of the groups in parallel; see workloads/model.hpp.
 */
#include "workloads/model.hpp"

namespace stst_model {
void register_light(std::vector<WorkloadEntry> &entries) {
    entries.push_back(make_entry<ConwayRule, stst_conway_params>("conway"));
    entries.push_back(make_entry<Jacobi5Rule, stst_jacobi5_params>("jacobi5"));
    entries.push_back(make_entry<Jacobi9Rule, stst_jacobi9_params>("jacobi9"));
    entries.push_back(make_entry<JacobiStarRule<2>, stst_jacobi_star_params>("jacobi_r2"));
    entries.push_back(make_entry<JacobiStarRule<3>, stst_jacobi_star_params>("jacobi_r3"));
}
} // namespace stst_model
