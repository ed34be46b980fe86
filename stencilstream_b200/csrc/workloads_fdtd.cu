/*
 * libstst_workloads — instantiates the header-only B200 backend (StencilStream/cuda/*.hpp) for the
 * "fdtd" transition functions of workloads/functors.hpp. One translation unit per group so that
 * nvcc compiles the kernel templates of the groups in parallel; see workloads/model.hpp.
 */
#include "workloads/model.hpp"

namespace stst_model {
void register_fdtd(std::vector<WorkloadEntry> &entries) {
    entries.push_back(make_entry<FdtdCoefRule, stst_fdtd_params>("fdtd"));
}
} // namespace stst_model
