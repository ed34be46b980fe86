/*
 * StencilStream-B200 — `stencil::cuda::StencilUpdate<F, split_cell_structure>`: the generation loop.
 *
 * Drop-in for the reference's updater (reference StencilStream/cuda/StencilUpdate.hpp:41-445):
 * same template parameters, the same `Params` aggregate (field names and order are API because user
 * code uses designated initialisers; B200-only fields come last), the same methods
 * (`operator()(GridImpl&)`, `get_params`, `get_n_processed_cells`, `get_walltime`,
 * `get_kernel_runtime`) and the same observable semantics:
 *
 *   - the source grid is never modified; the result is a different allocation (for
 *     `n_iterations == 0` the result aliases the source, like the reference's non-split path, :275),
 *   - iteration i in [iteration_offset, iteration_offset + n_iterations) runs sub-iterations
 *     0 .. F::n_subiterations-1 as full-grid sweeps (:212-215),
 *   - the time-dependent value of iteration i is `transition_function.get_time_dependent_value(i)`,
 *     evaluated ON THE HOST (:224) — here once per iteration up front, shipped per launch as a
 *     kernel-parameter array (the scheme of the reference's
 *     tdv::single_pass::PrecomputeOnHostStrategy, StencilStream/tdv/SinglePassStrategies.hpp:203-264),
 *   - cells outside the grid read as `halo_value` in every sweep (:241-252),
 *   - `params` is re-read on every call (:223), calls are asynchronous unless `blocking` (:133-135),
 *   - `n_processed_cells += n_iterations * H * W`, walltime accumulates (:137-141).
 *
 * What is different: instead of one launch per sweep, ceil(n_iterations / k) launches of the fused,
 * shared-memory-tiled kernel in cuda/internal/TileKernel.hpp, planned by cuda/internal/Planner.hpp.
 * `split_cell_structure` is accepted for source compatibility; the device layout is decided by
 * `Grid<Cell>` from `Cell::fields` alone (planes whenever the cell lists its fields), so both
 * values select the same code path and produce the same results.
 */
#pragma once
#include "../Concepts.hpp"
#include "../Stencil.hpp"
#include "Grid.hpp"
#include "internal/Helpers.hpp"
#include "internal/Launch.hpp"
#include "internal/Planner.hpp"
#include "internal/Runtime.hpp"
#include "internal/TileKernel.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

namespace stencil {
namespace cuda {

template <concepts::TransitionFunction F, bool split_cell_structure = false> class StencilUpdate {
  private:
    using Cell = typename F::Cell;
    using TDV = typename F::TimeDependentValue;
    using Layout = internal::CellLayout<Cell>;

    static_assert(std::is_trivially_copyable_v<F>,
                  "transition functions are copied into kernel parameters and must be trivially "
                  "copyable");
    static_assert(std::is_trivially_copyable_v<TDV>,
                  "time-dependent values are shipped as kernel parameters and must be trivially "
                  "copyable");

  public:
    /// The grid type this updater consumes and produces.
    using GridImpl = Grid<Cell>;

    struct Params {
        /// The transition function instance; may carry runtime parameters.
        F transition_function;

        /// Value presented for every cell outside the grid, in every sweep.
        Cell halo_value = Cell();

        /// Iteration index of the input grid; added to the local iteration counter.
        std::size_t iteration_offset = 0;

        /// Number of iterations to compute.
        std::size_t n_iterations = 1;

        /// Kept for source compatibility; the CUDA device is chosen by `cuda_device`.
        sycl::device device = sycl::device();

        /// Wait for the result before returning from `operator()`.
        bool blocking = false;

        /// Record device-side timestamps around every launch (see get_kernel_runtime()).
        bool profiling = false;

        // ---- B200 extensions (keep last: designated initialisers of reference code stay valid) ----

        /// CUDA device ordinal; negative: the device the source grid lives on.
        int cuda_device = -1;

        /// Iterations fused per launch (temporal blocking depth k); 0 picks one automatically.
        unsigned fused_iterations = 0;

        /// Upper bound for the output tile height; 0 lets the planner fill shared memory.
        unsigned tile_rows = 0;
    };

    StencilUpdate(Params params)
        : params(params), n_processed_cells(0), walltime(0.0), n_launches(0), last_plan(),
          tensor_maps(), profile_events() {}

    /**
     * Compute `n_iterations` iterations of the source grid and return the result as a new grid.
     * The source grid is not modified.
     */
    GridImpl operator()(GridImpl &source_grid) {
        auto walltime_start = std::chrono::high_resolution_clock::now();

        GridImpl result_grid = run_simulation(source_grid);

        if (params.blocking) {
            STST_RT_CHECK(stst_stream_synchronize(source_grid.get_storage().stream));
        }

        auto walltime_end = std::chrono::high_resolution_clock::now();
        walltime += std::chrono::duration<double>(walltime_end - walltime_start).count();
        n_processed_cells +=
            params.n_iterations * source_grid.get_grid_height() * source_grid.get_grid_width();
        return result_grid;
    }

    /// Live reference to the parameters; changes apply to the next call of `operator()`.
    Params &get_params() { return params; }

    /// Accumulated `n_iterations * height * width` over all calls.
    std::size_t get_n_processed_cells() const { return n_processed_cells; }

    /// Accumulated host-side time spent in `operator()` (meaningful as compute time if `blocking`).
    double get_walltime() const { return walltime; }

    /**
     * Accumulated device-side runtime in seconds of all fused launches submitted while
     * `Params::profiling` was set. Waits for those launches to finish.
     */
    double get_kernel_runtime() const {
        double seconds = 0.0;
        for (auto const &pair : profile_events) {
            STST_RT_CHECK(stst_event_synchronize(pair.second->get()));
            float ms = 0.0f;
            STST_RT_CHECK(stst_event_elapsed_ms(pair.first->get(), pair.second->get(), &ms));
            seconds += double(ms) * 1e-3;
        }
        return seconds;
    }

    /// B200 extension: number of kernel launches submitted so far.
    std::size_t get_n_launches() const { return n_launches; }

    /// B200 extension: the plan used by the most recent call.
    internal::LaunchPlan const &get_last_plan() const { return last_plan; }

  private:
    GridImpl run_simulation(GridImpl &source_grid) {
        if (params.n_iterations == 0) {
            return source_grid;
        }

        auto &source = source_grid.get_storage();
        const unsigned grid_h = unsigned(source.height);
        const unsigned grid_w = unsigned(source.width);

        GridImpl swap_grid_a = source_grid.make_similar();
        if (grid_h == 0 || grid_w == 0)
            return swap_grid_a;

        const bool trace = std::getenv("STST_TRACE") != nullptr;
        auto t_begin = std::chrono::high_resolution_clock::now();
        auto lap = [&](const char *what) {
            if (!trace)
                return;
            auto now = std::chrono::high_resolution_clock::now();
            std::fprintf(stderr, "[stst] %-22s %9.3f ms\n", what,
                         std::chrono::duration<double, std::milli>(now - t_begin).count());
            t_begin = now;
        };

        source.require_device();
        lap("require_device");

        const internal::LaunchPlan plan = internal::make_plan<F>(
            source.device, grid_h, grid_w, params.n_iterations, params.fused_iterations,
            params.tile_rows);
        last_plan = plan;
        lap("make_plan");

        const std::size_t k = plan.fused_iterations;
        const std::size_t n_full = params.n_iterations / k;
        const std::size_t n_tail = params.n_iterations % k;
        const std::size_t launches = n_full + (n_tail ? 1 : 0);

        GridImpl swap_grid_b = (launches > 1) ? source_grid.make_similar() : swap_grid_a;
        swap_grid_a.get_storage().allocate_device();
        swap_grid_b.get_storage().allocate_device();
        lap("allocate scratch grids");

        GridImpl *pass_source = &source_grid;
        GridImpl *pass_target = &swap_grid_a;
        std::size_t iteration = params.iteration_offset;
        for (std::size_t l = 0; l < launches; l++) {
            const unsigned n_gens = unsigned((l < n_full) ? k : n_tail);
            launch(plan, pass_source->get_storage(), pass_target->get_storage(), iteration, n_gens);
            iteration += n_gens;
            if (l == 0) {
                pass_source = &swap_grid_a;
                pass_target = &swap_grid_b;
            } else {
                std::swap(pass_source, pass_target);
            }
        }
        pass_source->get_storage().device_written();
        lap("submit launches");
        return *pass_source;
    }

    void launch(internal::LaunchPlan const &plan, internal::GridStorage<Cell> &src,
                internal::GridStorage<Cell> &dst, std::size_t iteration0, unsigned n_gens) {
        using namespace internal;
        LaunchRegion region{};
        region.device = src.device;
        region.grid_h = unsigned(src.height);
        region.grid_w = unsigned(src.width);
        region.buf_row0 = 0;
        region.buf_rows = src.height;
        region.out_row_lo = 0;
        region.out_row_hi = int(src.height);
        region.tile_h = 0;

        std::shared_ptr<Event> start, stop;
        if (params.profiling) {
            start = std::make_shared<Event>(true);
            stop = std::make_shared<Event>(true);
            start->record(src.stream);
        }

        SweepLauncher<F>::launch(plan, params.transition_function, params.halo_value, src.planes,
                                 dst.planes, nullptr, region, iteration0, n_gens, tensor_maps,
                                 src.stream);
        n_launches++;

        if (params.profiling) {
            stop->record(src.stream);
            profile_events.emplace_back(std::move(start), std::move(stop));
        }
    }

    Params params;
    std::size_t n_processed_cells;
    double walltime;
    std::size_t n_launches;
    internal::LaunchPlan last_plan;
    internal::TensorMapCache<Cell> tensor_maps;
    std::vector<std::pair<std::shared_ptr<internal::Event>, std::shared_ptr<internal::Event>>>
        profile_events;
};

} // namespace cuda
} // namespace stencil
