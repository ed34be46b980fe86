// Host-only probe of the launch planner's tile geometry (cuda/internal/Planner.hpp: shape_for and the
// shared-memory accounting of TileKernel.hpp). Compiled with g++ — no device code, no CUDA runtime
// call — and driven by tests/test_planner_cpu.py, which checks the printed shapes.
//   usage: planner_shapes <cell: f32|hotspot|fdtd|convection> <k> <block_x> <smem_budget> <single_mask>
#include <StencilStream/cuda/internal/Planner.hpp>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <tuple>

using namespace stencil::cuda::internal;

struct HotspotLike {
    float temp, power;
    static constexpr auto fields = std::make_tuple(&HotspotLike::temp, &HotspotLike::power);
};
struct FdtdLike {
    float ex, ey, hz, hz_sum, ca, cb, da, db;
    static constexpr auto fields =
        std::make_tuple(&FdtdLike::ex, &FdtdLike::ey, &FdtdLike::hz, &FdtdLike::hz_sum, &FdtdLike::ca,
                        &FdtdLike::cb, &FdtdLike::da, &FdtdLike::db);
};
struct ConvectionLike {
    double f[11];
};

template <typename Cell>
int report(unsigned n_sub, unsigned k, unsigned block_x, std::size_t budget, unsigned single) {
    constexpr unsigned cw = unsigned(column_group_width<Cell>());
    const unsigned align = column_alignment<Cell>(tma_capable<Cell>());
    const TileShape s = shape_for<Cell>(k, n_sub, 1, cw, align, block_x, 0, budget, 1u << 20, single);
    std::printf("{\"feasible\": %s, \"halo\": %u, \"hpad\": %u, \"tile_h\": %u, \"tile_w\": %u, "
                "\"rows\": %u, \"cols\": %u, \"smem_bytes\": %zu, \"efficiency\": %.6f, \"cw\": %u, "
                "\"n_planes\": %zu, \"buffer_bytes\": %zu, \"second_buffer_bytes\": %zu}\n",
                s.feasible ? "true" : "false", s.halo, s.hpad, s.tile_h, s.tile_w, s.rows, s.cols,
                s.smem_bytes, s.efficiency, cw, CellLayout<Cell>::n_planes,
                tile_buffer_bytes<Cell>(s.rows, s.cols), tile_buffer_bytes<Cell>(s.rows, s.cols, single));
    return 0;
}

int main(int argc, char **argv) {
    if (argc != 6)
        return 2;
    const unsigned k = unsigned(std::atoi(argv[2])), block_x = unsigned(std::atoi(argv[3]));
    const std::size_t budget = std::size_t(std::atoll(argv[4]));
    const unsigned single = unsigned(std::strtoul(argv[5], nullptr, 0));
    if (!std::strcmp(argv[1], "f32"))
        return report<float>(1, k, block_x, budget, single);
    if (!std::strcmp(argv[1], "hotspot"))
        return report<HotspotLike>(1, k, block_x, budget, single);
    if (!std::strcmp(argv[1], "fdtd"))
        return report<FdtdLike>(2, k, block_x, budget, single);
    if (!std::strcmp(argv[1], "convection"))
        return report<ConvectionLike>(3, k, block_x, budget, single);
    return 2;
}
