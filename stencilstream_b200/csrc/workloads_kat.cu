/*
 * libstst_workloads — instantiates the header-only B200 backend (StencilStream/cuda/*.hpp) for the
 * "kat" transition functions of workloads/functors.hpp. One translation unit per group so that
 * nvcc compiles the kernel templates of the groups in parallel; see workloads/model.hpp.
 */
#include "workloads/model.hpp"

namespace stst_model {
void register_kat(std::vector<WorkloadEntry> &entries) {
    entries.push_back(make_entry<KatRule<1>, stst_kat_params>("kat"));
    entries.push_back(make_entry<KatRule<2>, stst_kat_params>("kat_r2"));
}
} // namespace stst_model
