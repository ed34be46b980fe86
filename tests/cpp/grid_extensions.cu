// C++-level tests of the B200 extensions of stencil::cuda::Grid (max_abs, single-field copies,
// direct copies from/to ordinary host memory) — what a StencilStream user would call from an
// application instead of the host loops of reference examples/convection/convection.cpp:412-438 and
// examples/fdtd/src/fdtd.cpp:114-166. Runner: the Catch2 stand-in of oracle/ref_harness/catch2
// (test infrastructure); built by stencilstream_b200/tools/build_reference_tests.py (`build_own`).
#include <StencilStream/BaseTransitionFunction.hpp>
#include <StencilStream/cuda/StencilUpdate.hpp>
#include <catch2/catch_all.hpp>

#include <cmath>
#include <limits>
#include <vector>

using namespace stencil;

struct FlowCell {
    double T, Vx;
    float tag;
    int count;
    static constexpr auto fields =
        std::make_tuple(&FlowCell::T, &FlowCell::Vx, &FlowCell::tag, &FlowCell::count);
};

// Decays T towards the mean of its four neighbours, advects nothing; enough to make the device
// copy newer than any host image.
struct Relax : public BaseTransitionFunction {
    using Cell = FlowCell;
    STST_HD Cell operator()(Stencil<Cell, 1> const &st) const {
        Cell next = st[0][0];
        next.T = 0.25 * (st[-1][0].T + st[1][0].T + st[0][-1].T + st[0][1].T);
        next.count = st[0][0].count + 1;
        return next;
    }
};

using Grid = cuda::Grid<FlowCell>;

static FlowCell cell_at(std::size_t r, std::size_t c) {
    return FlowCell{std::sin(0.37 * r) * std::cos(0.11 * c) - 0.2, double(r) - 1.5 * double(c),
                    float(int(r * 31 + c * 17) % 97) - 48.0f, int(r) - int(c)};
}

static Grid make_grid(std::size_t rows, std::size_t cols) {
    Grid grid(rows, cols);
    Grid::GridAccessor<sycl::access::mode::read_write> ac(grid);
    for (std::size_t r = 0; r < rows; r++)
        for (std::size_t c = 0; c < cols; c++)
            ac[r][c] = cell_at(r, c);
    return grid;
}

template <typename Get> static double host_max_abs(Grid &grid, std::size_t rows, std::size_t cols, Get get) {
    Grid::GridAccessor<sycl::access::mode::read> ac(grid);
    double m = -std::numeric_limits<double>::infinity();
    for (std::size_t r = 0; r < rows; r++)
        for (std::size_t c = 0; c < cols; c++)
            if (std::abs(double(get(ac[r][c]))) > m)
                m = std::abs(double(get(ac[r][c])));
    return m;
}

TEST_CASE("cuda::Grid::max_abs (B200 extension)", "[cuda::Grid]") {
    const std::size_t rows = 157, cols = 301;
    Grid grid = make_grid(rows, cols);
    REQUIRE(Grid::plane_of<&FlowCell::T>() == 0);
    REQUIRE(Grid::plane_of<&FlowCell::count>() == 3);
    REQUIRE(grid.max_abs<&FlowCell::T>(rows, cols) ==
            host_max_abs(grid, rows, cols, [](FlowCell const &c) { return c.T; }));
    REQUIRE(grid.max_abs<&FlowCell::Vx>(rows - 1, cols - 2) ==
            host_max_abs(grid, rows - 1, cols - 2, [](FlowCell const &c) { return c.Vx; }));
    REQUIRE(grid.max_abs<&FlowCell::tag>(rows, 7) ==
            host_max_abs(grid, rows, 7, [](FlowCell const &c) { return c.tag; }));
    REQUIRE(grid.max_abs<&FlowCell::count>(3, cols) ==
            host_max_abs(grid, 3, cols, [](FlowCell const &c) { return c.count; }));
    REQUIRE(grid.max_abs<&FlowCell::T>(0, cols) == -std::numeric_limits<double>::infinity());
    // several norms in one pass, as the convection loop needs them
    std::vector<double> norms = grid.max_abs({{0, rows, cols}, {1, rows, cols - 1}, {3, rows - 1, cols}});
    REQUIRE(norms.size() == 3);
    REQUIRE(norms[1] == host_max_abs(grid, rows, cols - 1, [](FlowCell const &c) { return c.Vx; }));
    bool thrown = false;
    try {
        grid.max_abs({{4, rows, cols}});
    } catch (std::invalid_argument const &) {
        thrown = true;
    }
    REQUIRE(thrown);
}

TEST_CASE("cuda::Grid::max_abs sees the result of an update", "[cuda::Grid]") {
    const std::size_t rows = 96, cols = 200;
    Grid grid = make_grid(rows, cols);
    cuda::StencilUpdate<Relax> update({.transition_function = Relax{}, .n_iterations = 5});
    Grid result = update(grid);
    const double on_device = result.max_abs<&FlowCell::T>(rows, cols); // before any download
    REQUIRE(on_device == host_max_abs(result, rows, cols, [](FlowCell const &c) { return c.T; }));
    REQUIRE(result.max_abs<&FlowCell::count>(rows, cols) ==
            host_max_abs(result, rows, cols, [](FlowCell const &c) { return c.count; }));
}

TEST_CASE("cuda::Grid single-field copies (B200 extension)", "[cuda::Grid]") {
    const std::size_t rows = 64, cols = 130;
    Grid grid = make_grid(rows, cols);
    sycl::buffer<double, 2> vx(sycl::range<2>(rows, cols));
    grid.copy_field_to_buffer<&FlowCell::Vx>(vx);
    {
        sycl::host_accessor ac(vx, sycl::read_only);
        for (std::size_t r = 0; r < rows; r++)
            for (std::size_t c = 0; c < cols; c++)
                REQUIRE(ac[r][c] == cell_at(r, c).Vx);
    }
    bool thrown = false;
    try {
        sycl::buffer<double, 2> wrong(sycl::range<2>(rows, cols + 1));
        grid.copy_field_to_buffer<&FlowCell::Vx>(wrong);
    } catch (std::range_error const &) {
        thrown = true;
    }
    REQUIRE(thrown);
    std::vector<int> counts(rows * cols, 7);
    grid.copy_plane_from_host(Grid::plane_of<&FlowCell::count>(), counts.data());
    Grid::GridAccessor<sycl::access::mode::read> ac(grid);
    REQUIRE(ac[5][9].count == 7);
    REQUIRE(ac[5][9].T == cell_at(5, 9).T); // other fields untouched
}

TEST_CASE("cuda::Grid copies from and to ordinary host memory", "[cuda::Grid]") {
    const std::size_t rows = 700, cols = 1100; // 18 MB of 24-byte cells: more than one staging slot
    std::vector<FlowCell> cells(rows * cols);
    for (std::size_t r = 0; r < rows; r++)
        for (std::size_t c = 0; c < cols; c++)
            cells[r * cols + c] = cell_at(r, c);
    Grid grid(rows, cols);
    grid.copy_from_host(cells.data());
    std::vector<FlowCell> back(rows * cols);
    grid.copy_to_host(back.data());
    REQUIRE(std::memcmp(back.data(), cells.data(), cells.size() * sizeof(FlowCell)) == 0);
    Grid::GridAccessor<sycl::access::mode::read> ac(grid);
    REQUIRE(ac[rows - 1][cols - 1].Vx == cell_at(rows - 1, cols - 1).Vx);
}

int main() { return catch_standin::run_all_test_cases() == 0 ? 0 : 1; }
