/*
 * libstst_workloads — instantiates the header-only B200 backend (StencilStream/cuda/*.hpp) for the
 * "convection" transition functions of workloads/functors.hpp. One translation unit per group so that
 * nvcc compiles the kernel templates of the groups in parallel; see workloads/model.hpp.
 */
#include "workloads/model.hpp"

namespace stst_model {
void register_convection(std::vector<WorkloadEntry> &entries) {
    entries.push_back(make_entry<ConvectionPseudoTransientRule, stst_convection_pt_params>("convection_pt"));
    entries.push_back(make_entry<ConvectionThermalRule, stst_convection_thermal_params>("convection_thermal"));
}
} // namespace stst_model
