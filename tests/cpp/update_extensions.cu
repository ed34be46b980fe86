// C++-level tests of stencil::cuda::StencilUpdate beyond what the reference's own unit tests cover:
//   * stencil radii larger than the per-thread column group (the reference's backends accept any
//     radius, reference StencilStream/Concepts.hpp:61-77; double-precision radius 3: 2 columns per
//     thread; an 88-byte cell with radius 2: 1 column per thread),
//   * Cell::constant_fields (declared pass-through; STST_VERIFY_CONSTANT_FIELDS catches a functor
//     that breaks the declaration),
//   * Params::cuda_devices: one call spread over several row slabs from C++ (a device may be listed
//     twice, so this runs on a one-GPU box).
// Every case is checked against a plain host loop that restates the reference sweep
// (reference StencilStream/cpu/StencilUpdate.hpp:185-223) for the functor at hand.
// Runner: the Catch2 stand-in of oracle/ref_harness/catch2 (test infrastructure); built by
// stencilstream_b200/tools/build_reference_tests.py (`build_own`).
#include <StencilStream/BaseTransitionFunction.hpp>
#include <StencilStream/cuda/StencilUpdate.hpp>
#include <catch2/catch_all.hpp>

#include <cstdlib>
#include <cstring>
#include <vector>

using namespace stencil;

// ---- host restatement of the generation loop for any functor ------------------------------------------
template <typename F>
std::vector<typename F::Cell> host_update(F const &f, std::vector<typename F::Cell> cells, std::size_t rows,
                                          std::size_t cols, typename F::Cell halo, std::size_t offset,
                                          std::size_t n) {
    using Cell = typename F::Cell;
    constexpr long R = long(F::stencil_radius);
    std::vector<Cell> next(cells.size());
    for (std::size_t it = offset; it < offset + n; it++) {
        for (std::size_t sub = 0; sub < F::n_subiterations; sub++) {
            for (std::size_t r = 0; r < rows; r++) {
                for (std::size_t c = 0; c < cols; c++) {
                    Stencil<Cell, F::stencil_radius, typename F::TimeDependentValue> st(
                        sycl::id<2>(r, c), sycl::range<2>(rows, cols), it, sub,
                        f.get_time_dependent_value(it));
                    for (long dr = -R; dr <= R; dr++)
                        for (long dc = -R; dc <= R; dc++) {
                            const long rr = long(r) + dr, cc = long(c) + dc;
                            const bool in = rr >= 0 && cc >= 0 && rr < long(rows) && cc < long(cols);
                            st[sycl::id<2>(std::size_t(dr + R), std::size_t(dc + R))] =
                                in ? cells[std::size_t(rr) * cols + std::size_t(cc)] : halo;
                        }
                    next[r * cols + c] = f(st);
                }
            }
            cells.swap(next);
        }
    }
    return cells;
}

template <typename F>
std::vector<typename F::Cell> device_update(typename cuda::StencilUpdate<F>::Params params,
                                            std::vector<typename F::Cell> const &cells, std::size_t rows,
                                            std::size_t cols, cuda::StencilUpdate<F> **keep = nullptr) {
    using Cell = typename F::Cell;
    cuda::Grid<Cell> grid(rows, cols);
    grid.copy_from_host(cells.data());
    static std::vector<std::unique_ptr<cuda::StencilUpdate<F>>> updates;
    updates.push_back(std::make_unique<cuda::StencilUpdate<F>>(params));
    if (keep)
        *keep = updates.back().get();
    cuda::Grid<Cell> result = (*updates.back())(grid);
    std::vector<Cell> out(rows * cols);
    result.copy_to_host(out.data());
    return out;
}

template <typename Cell> bool same_bytes(std::vector<Cell> const &a, std::vector<Cell> const &b) {
    return a.size() == b.size() && std::memcmp(a.data(), b.data(), a.size() * sizeof(Cell)) == 0;
}

// ---- radius 3 on double cells: column group width 2 < radius ----------------------------------------
struct WideDouble : public BaseTransitionFunction {
    using Cell = double;
    static constexpr std::size_t stencil_radius = 3;
    STST_HD Cell operator()(Stencil<Cell, 3> const &st) const {
        double sum = 0.0;
        for (int d = -3; d <= 3; d++)
            sum = sum + st[d][0] * (1.0 / 16.0) + st[0][d] * (1.0 / 32.0);
        return sum + st[-3][3] * 0.001 - st[3][-3] * 0.002; // corners of the window too
    }
};

TEST_CASE("radius 3 on 8-byte cells (radius > column group)", "[cuda::StencilUpdate]") {
    const std::size_t rows = 83, cols = 301;
    std::vector<double> cells(rows * cols);
    for (std::size_t i = 0; i < cells.size(); i++)
        cells[i] = double((i * 2654435761u) % 1000) / 1000.0;
    for (unsigned fused : {1u, 2u, 0u}) {
        auto got = device_update<WideDouble>({.transition_function = WideDouble{}, .halo_value = 0.25,
                                              .n_iterations = 5, .blocking = true,
                                              .fused_iterations = fused},
                                             cells, rows, cols);
        auto want = host_update(WideDouble{}, cells, rows, cols, 0.25, 0, 5);
        double worst = 0.0;
        for (std::size_t i = 0; i < want.size(); i++)
            worst = std::max(worst, std::abs(got[i] - want[i]));
        REQUIRE(worst <= 1e-12); // FMA contraction only
    }
}

// ---- radius 2 on a fat cell: one column per thread --------------------------------------------------
struct FatCell { // 11 doubles, like the mantle-convection cell of the reference
    double a, b, f2, f3, f4, f5, f6, f7, f8, f9, stamp;
    static constexpr auto fields =
        std::make_tuple(&FatCell::a, &FatCell::b, &FatCell::f2, &FatCell::f3, &FatCell::f4, &FatCell::f5,
                        &FatCell::f6, &FatCell::f7, &FatCell::f8, &FatCell::f9, &FatCell::stamp);
};
struct FatWide : public BaseTransitionFunction {
    using Cell = FatCell;
    static constexpr std::size_t stencil_radius = 2;
    static constexpr std::size_t n_subiterations = 2;
    STST_HD Cell operator()(Stencil<Cell, 2> const &st) const {
        Cell next = st[0][0];
        if (st.subiteration == 0)
            next.a = 0.2 * (st[-2][0].b + st[2][0].b + st[0][-2].b + st[0][2].b + st[0][0].a);
        else
            next.b = 0.25 * (st[-1][-2].a + st[1][2].a + st[2][1].a + st[-2][-1].a);
        next.f5 = st[0][0].f5 + st[0][1].f9 * 0.5;
        next.stamp = double(st.id[0] * 1000 + st.id[1]) + double(st.iteration);
        return next;
    }
};

TEST_CASE("radius 2 on an 88-byte cell (one column per thread)", "[cuda::StencilUpdate]") {
    const std::size_t rows = 61, cols = 97;
    std::vector<FatCell> cells(rows * cols);
    for (std::size_t i = 0; i < cells.size(); i++) {
        double *v = reinterpret_cast<double *>(&cells[i]);
        for (int j = 0; j < 11; j++)
            v[j] = double((i * 40503u + j * 977u) % 512) / 64.0;
    }
    FatCell halo{};
    halo.a = 1.0;
    halo.b = -1.0;
    auto got = device_update<FatWide>({.transition_function = FatWide{}, .halo_value = halo,
                                       .iteration_offset = 3, .n_iterations = 4, .blocking = true},
                                      cells, rows, cols);
    auto want = host_update(FatWide{}, cells, rows, cols, halo, 3, 4);
    double worst = 0.0;
    for (std::size_t i = 0; i < want.size(); i++)
        for (int j = 0; j < 11; j++)
            worst = std::max(worst, std::abs(reinterpret_cast<const double *>(&got[i])[j] -
                                             reinterpret_cast<const double *>(&want[i])[j]));
    REQUIRE(worst <= 1e-9);
}

// ---- Cell::constant_fields ------------------------------------------------------------------------------
struct HeatCell {
    float temp, source, mask;
    int tag;
    static constexpr auto fields =
        std::make_tuple(&HeatCell::temp, &HeatCell::source, &HeatCell::mask, &HeatCell::tag);
    static constexpr auto constant_fields = std::make_tuple(&HeatCell::source, &HeatCell::mask);
};
template <bool kHonest> struct Heat : public BaseTransitionFunction {
    using Cell = HeatCell;
    STST_HD Cell operator()(Stencil<Cell, 1> const &st) const {
        Cell next = st[0][0];
        next.temp = st[0][0].mask * 0.25f * (st[-1][0].temp + st[1][0].temp + st[0][-1].temp + st[0][1].temp) +
                    st[0][0].source;
        next.tag = st[0][0].tag + 1;
        if (!kHonest && st.iteration == 5 && st.id[0] == 7 && st.id[1] == 9)
            next.source = 99.0f; // breaks the declaration
        return next;
    }
};

static std::vector<HeatCell> heat_cells(std::size_t rows, std::size_t cols) {
    std::vector<HeatCell> cells(rows * cols);
    for (std::size_t i = 0; i < cells.size(); i++)
        cells[i] = HeatCell{float(i % 17), float(i % 5) * 0.125f, (i % 7) ? 1.0f : 0.5f, int(i % 3)};
    return cells;
}

TEST_CASE("Cell::constant_fields: declared pass-through", "[cuda::StencilUpdate]") {
    static_assert(cuda::internal::constant_fields_mask<HeatCell>() == 0b0110);
    const std::size_t rows = 150, cols = 530;
    auto cells = heat_cells(rows, cols);
    cuda::StencilUpdate<Heat<true>> *update = nullptr;
    auto got = device_update<Heat<true>>({.transition_function = Heat<true>{}, .halo_value = HeatCell{},
                                          .n_iterations = 11, .blocking = true, .fused_iterations = 3},
                                         cells, rows, cols, &update);
    auto want = host_update(Heat<true>{}, cells, rows, cols, HeatCell{}, 0, 11);
    for (std::size_t i = 0; i < want.size(); i++) {
        REQUIRE(std::abs(got[i].temp - want[i].temp) <= 1e-4f * (1.0f + std::abs(want[i].temp)));
        REQUIRE(got[i].source == want[i].source);
        REQUIRE(got[i].mask == want[i].mask);
        REQUIRE(got[i].tag == want[i].tag);
    }
    REQUIRE(update->get_passthrough_planes() == 0b0110); // decided at compile time ...
    REQUIRE(update->get_n_speculation_redos() == 0);      // ... nothing to repeat
    REQUIRE(update->get_n_launches() == 4);               // ... and no observing launch: ceil(11 / 3)
}

TEST_CASE("Cell::constant_fields: a broken declaration is reported on request", "[cuda::StencilUpdate]") {
    setenv("STST_VERIFY_CONSTANT_FIELDS", "1", 1);
    const std::size_t rows = 40, cols = 64;
    auto cells = heat_cells(rows, cols);
    bool thrown = false;
    try {
        device_update<Heat<false>>({.transition_function = Heat<false>{}, .halo_value = HeatCell{},
                                    .n_iterations = 8, .blocking = true},
                                   cells, rows, cols);
    } catch (std::logic_error const &) {
        thrown = true;
    }
    REQUIRE(thrown);
}

// ---- Params::cuda_devices from C++ ------------------------------------------------------------------------
TEST_CASE("Params::cuda_devices: one call on several row slabs", "[cuda::StencilUpdate]") {
    const std::size_t rows = 257, cols = 300;
    auto cells = heat_cells(rows, cols);
    int n_devices = 0;
    STST_RT_CHECK(stst_device_count(&n_devices));
    std::vector<int> devices;
    for (int i = 0; i < 3; i++)
        devices.push_back(i % n_devices);
    cuda::StencilUpdate<Heat<true>> *update = nullptr;
    auto got = device_update<Heat<true>>({.transition_function = Heat<true>{}, .halo_value = HeatCell{},
                                          .n_iterations = 13, .blocking = true, .fused_iterations = 2,
                                          .cuda_devices = devices},
                                         cells, rows, cols, &update);
    auto single = device_update<Heat<true>>({.transition_function = Heat<true>{}, .halo_value = HeatCell{},
                                             .n_iterations = 13, .blocking = true, .fused_iterations = 2},
                                            cells, rows, cols);
    REQUIRE(update->get_n_slabs() == 3);
    REQUIRE(same_bytes(got, single)); // same kernels, same arithmetic: bit-identical to one device
}

int main() { return catch_standin::run_all_test_cases() == 0 ? 0 : 1; }
