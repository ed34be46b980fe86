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
 * Fields that a transition function only passes through (HotSpot's `power`, FDTD's material
 * coefficients) are detected at run time and no longer copied between the tile buffers — speculative
 * plane pass-through, verified by every launch and transparently repeated without the speculation if a
 * launch ever sees such a field change (see `run_speculative` below and run_tile in TileKernel.hpp).
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
#include "internal/SlabUpdate.hpp"
#include "internal/TileKernel.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

namespace stencil {
namespace cuda {

template <concepts::TransitionFunction F, bool split_cell_structure = false> class StencilUpdate {
  public:
    // (public: nvcc's host-side rewriting of designated initialisers such as
    // `{.halo_value = MyCell{}}` names the aggregate's member types through these aliases)
    using Cell = typename F::Cell;
    using TDV = typename F::TimeDependentValue;

  private:
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

        /// Kept for source compatibility (the shim's sycl::device carries no information); an update
        /// runs on the CUDA device its source grid lives on, see `cuda_device`.
        sycl::device device = sycl::device();

        /// Wait for the result before returning from `operator()`.
        bool blocking = false;

        /// Record device-side timestamps around every launch (see get_kernel_runtime()).
        bool profiling = false;

        // ---- B200 extensions (keep last: designated initialisers of reference code stay valid) ----

        /// CUDA device ordinal the update is expected to run on; negative (default): wherever the
        /// source grid lives. An update never migrates a grid: a non-negative ordinal that differs
        /// from the source grid's device makes `operator()` throw std::invalid_argument.
        int cuda_device = -1;

        /// Iterations fused per launch (temporal blocking depth k); 0 picks one automatically.
        unsigned fused_iterations = 0;

        /// Upper bound for the output tile height; 0 lets the planner fill shared memory.
        unsigned tile_rows = 0;

        /// CUDA devices to spread one update over: the grid is cut into row slabs, one per entry
        /// (an ordinal may appear more than once), each slab runs the generation loop on its device
        /// and pushes its boundary rows into its neighbours over NVLink (cuda/internal/SlabUpdate.hpp).
        /// Source and result grid stay ordinary single-device grids: the slabs are filled from the
        /// source and gathered into the result with device-to-device copies, inside `operator()`.
        /// Empty (default): the environment variable STST_DEVICES ("0,1,2,3" or "0-7"), else the
        /// device of the source grid alone. The reference's updater takes one device
        /// (reference cuda/StencilUpdate.hpp:83, :124-127); this is how a program written against
        /// it — the unmodified examples — uses a whole 8 x B200 box.
        std::vector<int> cuda_devices = {};
    };

    StencilUpdate(Params params)
        : params(params), n_processed_cells(0), walltime(0.0), n_launches(0), last_plan(),
          tensor_maps(), profile_events() {}

    StencilUpdate(StencilUpdate const &other)
        : params(other.params), n_processed_cells(other.n_processed_cells),
          walltime(other.walltime), n_launches(other.n_launches), last_plan(other.last_plan),
          tensor_maps(), profile_events(other.profile_events), spec_probed(other.spec_probed),
          n_spec_redos(other.n_spec_redos) {
        // what was observed about the transition function travels with the copy; the device words
        // (spec_flags) are per object and allocated on first use
        for (unsigned q = 0; q < internal::max_spec_subiterations; q++)
            spec_keep[q] = other.spec_keep[q];
    }
    StencilUpdate &operator=(StencilUpdate const &) = delete;

    ~StencilUpdate() {
        if (spec_flags)
            internal::device_free(spec_device, spec_flags, internal::default_stream(spec_device));
        if (spec_host_flags)
            internal::pinned_free(spec_host_flags);
    }

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

    /// B200 extension: number of row slabs (GPUs) the most recent call ran on (1: not sharded).
    std::size_t get_n_slabs() const { return shards ? shards->slabs.size() : 1; }

    /// B200 extension: the plan used by the most recent call.
    internal::LaunchPlan const &get_last_plan() const { return last_plan; }

    /// B200 extension: planes (bit i = plane i) that currently pass through every sub-iteration
    /// without being copied (speculative plane pass-through), and how many times an update had to
    /// be repeated because a launch saw such a plane change.
    unsigned get_passthrough_planes() const {
        if (shards)
            return shards->slabs.front()->passthrough_planes();
        if (!spec_probed)
            return 0;
        unsigned m = all_planes;
        for (unsigned q = 0; q < n_sub; q++)
            m &= spec_keep[q];
        return m;
    }
    std::size_t get_n_speculation_redos() const { return n_spec_redos; }

  private:
    static constexpr unsigned n_sub = unsigned(F::n_subiterations);
    static constexpr unsigned all_planes =
        Layout::n_planes >= 32 ? ~0u : ((1u << Layout::n_planes) - 1u);

    static bool speculation_enabled() {
        if constexpr (!internal::speculation_capable<F>())
            return false;
        static const bool enabled = internal::env_long("STST_SPECULATE", 1) != 0;
        return enabled;
    }

    /// Planes of fields the cell type declares constant (`Cell::constant_fields`, see
    /// cuda/internal/Helpers.hpp), if using them pays (same rule as for detected ones).
    static constexpr unsigned declared_constant = internal::constant_fields_mask<Cell>();
    static_assert(declared_constant != ~0u || Layout::n_planes >= 32,
                  "every entry of Cell::constant_fields must also be listed in Cell::fields");

    static bool declared_passthrough_pays() {
        if constexpr (declared_constant == 0 || !internal::speculation_capable<F>()) {
            return false;
        } else {
            std::size_t bytes = 0;
            for (std::size_t i = 0; i < Layout::n_planes; i++)
                if ((declared_constant >> i) & 1u)
                    bytes += Layout::plane_bytes(i);
            return speculation_enabled() && 4 * bytes >= sizeof(Cell) && sizeof(Cell) <= 64;
        }
    }

    /**
     * The generation loop with DECLARED constant fields: the pass-through kernels with a keep mask
     * known at compile time. No observing launch and no read-back — the call stays asynchronous.
     * With STST_VERIFY_CONSTANT_FIELDS=1 the kernels' own check of the kept planes is read after the
     * last launch (one stream synchronisation) and a violated declaration throws std::logic_error.
     */
    GridImpl run_declared(GridImpl &source_grid) {
        using namespace internal;
        auto &source = source_grid.get_storage();
        const unsigned grid_h = unsigned(source.height), grid_w = unsigned(source.width);
        source.require_device();
        if (spec_flags && spec_device != source.device) {
            device_free(spec_device, spec_flags, default_stream(spec_device));
            spec_flags = nullptr;
        }
        const std::size_t flag_bytes = sizeof(unsigned) * (max_spec_subiterations + 1);
        if (!spec_flags) {
            spec_device = source.device;
            spec_flags = static_cast<unsigned *>(device_alloc(spec_device, flag_bytes, source.stream));
            STST_RT_CHECK(stst_memset_async(spec_flags, 0, flag_bytes, source.stream));
        }
        Speculation spec{};
        for (unsigned q = 0; q < n_sub; q++)
            spec.keep[q] = spec_keep[q] = declared_constant & all_planes;
        spec.flags = spec_flags;
        spec_probed = true;
        const LaunchPlan plan =
            make_plan<F>(source.device, grid_h, grid_w, params.n_iterations, params.fused_iterations,
                         params.tile_rows, spec.single_planes(n_sub, all_planes));
        last_plan = plan;
        GridImpl swap_a = source_grid.make_similar();
        GridImpl swap_b = (params.n_iterations > plan.fused_iterations) ? source_grid.make_similar()
                                                                         : swap_a;
        swap_a.get_storage().allocate_device();
        swap_b.get_storage().allocate_device();
        GridImpl *pass_source = &source_grid, *pass_target = &swap_a;
        std::size_t iteration = params.iteration_offset, remaining = params.n_iterations;
        bool first = true;
        while (remaining > 0) {
            const unsigned n_gens = unsigned(std::min<std::size_t>(remaining, plan.fused_iterations));
            launch(plan, pass_source->get_storage(), pass_target->get_storage(), iteration, n_gens,
                   &spec);
            iteration += n_gens;
            remaining -= n_gens;
            if (first) {
                pass_source = &swap_a;
                pass_target = &swap_b;
                first = false;
            } else {
                std::swap(pass_source, pass_target);
            }
        }
        pass_source->get_storage().device_written();
        static const bool verify = env_long("STST_VERIFY_CONSTANT_FIELDS", 0) != 0;
        if (verify) {
            unsigned host_flags[max_spec_subiterations + 1] = {};
            STST_RT_CHECK(stst_memcpy_d2h_async(host_flags, spec_flags, flag_bytes, source.stream));
            STST_RT_CHECK(stst_stream_synchronize(source.stream));
            if (host_flags[max_spec_subiterations] != 0)
                throw std::logic_error(
                    "StencilStream-B200: the transition function changed a field that its cell type "
                    "lists in Cell::constant_fields (plane mask " +
                    std::to_string(host_flags[max_spec_subiterations]) + ")");
        }
        return *pass_source;
    }

    /**
     * The generation loop with speculative plane pass-through.
     *
     * The first launch this updater ever submits also *observes*: per sub-iteration, which planes
     * came out of the transition function different from the centre cell that went in. Planes that
     * never changed are from then on left in place by the tile sweeps instead of being copied from
     * tile buffer to tile buffer, and planes untouched in every sub-iteration need only one tile
     * buffer, so the tiles grow. Every launch still compares what the function returned for those
     * planes with what went in (for a genuinely passed-through field the compiler folds that away)
     * and reports any difference through a device flag, which is read when all launches of the call
     * have been submitted. If a launch reports a difference — the field does change after all, e.g.
     * FDTD's `hz_sum` once `iteration >= detect_iteration` — that plane is struck from the masks and
     * the WHOLE call is repeated from the (never modified) source grid. The result is therefore
     * always what the unspeculated loop computes; the price is that such calls end with a stream
     * synchronisation even when `blocking` is false.
     */
    GridImpl run_speculative(GridImpl &source_grid) {
        using namespace internal;
        auto &source = source_grid.get_storage();
        const unsigned grid_h = unsigned(source.height), grid_w = unsigned(source.width);
        source.require_device();
        if (spec_flags && spec_device != source.device) {
            device_free(spec_device, spec_flags, default_stream(spec_device));
            spec_flags = nullptr;
        }
        if (!spec_flags) {
            spec_device = source.device;
            spec_flags = static_cast<unsigned *>(
                device_alloc(spec_device, sizeof(unsigned) * (max_spec_subiterations + 1),
                             source.stream));
        }
        if (!spec_host_flags)
            spec_host_flags = static_cast<unsigned *>(
                pinned_alloc(sizeof(unsigned) * (max_spec_subiterations + 1)));
        unsigned *host_flags = spec_host_flags;
        // launches and profiling events of an attempt that is discarded must not be counted
        const std::size_t launches_before = n_launches;
        const std::size_t events_before = profile_events.size();
        auto read_flags = [&] {
            STST_RT_CHECK(stst_memcpy_d2h_async(host_flags, spec_flags,
                                                sizeof(unsigned) * (max_spec_subiterations + 1),
                                                source.stream));
            STST_RT_CHECK(stst_stream_synchronize(source.stream));
        };
        auto clear_flags = [&] {
            STST_RT_CHECK(stst_memset_async(spec_flags, 0,
                                            sizeof(unsigned) * (max_spec_subiterations + 1),
                                            source.stream));
        };

        {
            for (;;) {
                clear_flags();
                GridImpl swap_a = source_grid.make_similar();
                GridImpl swap_b = source_grid.make_similar();
                swap_a.get_storage().allocate_device();
                GridImpl *pass_source = &source_grid, *pass_target = &swap_a;
                std::size_t iteration = params.iteration_offset;
                std::size_t remaining = params.n_iterations;
                bool first = true;
                auto advance = [&](LaunchPlan const &plan, Speculation const &spec) {
                    const unsigned n_gens =
                        unsigned(std::min<std::size_t>(remaining, plan.fused_iterations));
                    pass_target->get_storage().allocate_device();
                    launch(plan, pass_source->get_storage(), pass_target->get_storage(), iteration,
                           n_gens, &spec);
                    iteration += n_gens;
                    remaining -= n_gens;
                    if (first) {
                        pass_source = &swap_a;
                        pass_target = &swap_b;
                        first = false;
                    } else {
                        std::swap(pass_source, pass_target);
                    }
                };

                if (!spec_probed) {
                    // observe with the unspeculated plan, then read what changed
                    const LaunchPlan plan0 = make_plan<F>(source.device, grid_h, grid_w,
                                                          params.n_iterations,
                                                          params.fused_iterations, params.tile_rows);
                    Speculation observe{};
                    observe.probe = true;
                    observe.flags = spec_flags;
                    advance(plan0, observe);
                    last_plan = plan0;
                    read_flags();
                    for (unsigned q = 0; q < n_sub; q++)
                        spec_keep[q] = ~host_flags[q] & all_planes;
                    spec_probed = true;
                    drop_unprofitable_speculation();
                    clear_flags();
                }

                Speculation spec{};
                for (unsigned q = 0; q < n_sub; q++)
                    spec.keep[q] = spec_keep[q];
                spec.flags = spec_flags;
                if (remaining > 0) {
                    const LaunchPlan plan = make_plan<F>(
                        source.device, grid_h, grid_w, params.n_iterations, params.fused_iterations,
                        params.tile_rows, spec.single_planes(n_sub, all_planes));
                    last_plan = plan;
                    while (remaining > 0)
                        advance(plan, spec);
                }
                read_flags();
                const unsigned violated = host_flags[max_spec_subiterations];
                if (violated == 0) {
                    pass_source->get_storage().device_written();
                    return *pass_source;
                }
                // A plane believed to pass through did change: never speculate on it again and
                // recompute this call from the untouched source grid.
                for (unsigned q = 0; q < n_sub; q++)
                    spec_keep[q] &= ~violated;
                drop_unprofitable_speculation();
                n_spec_redos++;
                n_launches = launches_before;
                profile_events.resize(events_before);
            }
        }
    }

    /// Pass-through pays through the planes that need only ONE tile buffer (taller tiles) and the
    /// stores it saves; the kernels that support it carry per-plane buffer bookkeeping, which costs
    /// registers. Measured (profiles/r01_s3_sweep_speculation.log): HotSpot +11 %, FDTD +19 % with
    /// half of the cell passing through, mantle convection -11 % with one field of eleven. Below a
    /// quarter of the cell's bytes the unspeculated kernels are used.
    void drop_unprofitable_speculation() {
        unsigned single = all_planes;
        for (unsigned q = 0; q < n_sub; q++)
            single &= spec_keep[q];
        std::size_t bytes = 0;
        for (std::size_t i = 0; i < Layout::n_planes; i++)
            if ((single >> i) & 1u)
                bytes += Layout::plane_bytes(i);
        // ... and cells beyond 64 bytes run at the register limit of their 512-thread CTAs: the
        // per-plane buffer bookkeeping spills there (convection: 9.2 -> 8.2 GCell-updates/s even
        // with four of eleven fields constant in the benchmark input)
        if (4 * bytes < sizeof(Cell) || sizeof(Cell) > 64) {
            for (unsigned q = 0; q < n_sub; q++)
                spec_keep[q] = 0;
        }
    }

    bool speculation_has_nothing_left() const {
        if (!spec_probed)
            return false;
        for (unsigned q = 0; q < n_sub; q++)
            if (spec_keep[q] != 0)
                return false;
        return true;
    }

    GridImpl run_simulation(GridImpl &source_grid) {
        if (params.n_iterations == 0) {
            return source_grid;
        }
        if (params.cuda_device >= 0 && params.cuda_device != source_grid.get_storage().device)
            throw std::invalid_argument(
                "StencilStream-B200: Params::cuda_device is " + std::to_string(params.cuda_device) +
                " but the source grid lives on device " +
                std::to_string(source_grid.get_storage().device) +
                "; updates run where their grid is (create the grid on that device)");
        {
            const std::vector<int> devices = resolve_devices();
            if (devices.size() > 1 && source_grid.get_grid_height() > 0 &&
                source_grid.get_grid_width() > 0) {
                if (ensure_shards(source_grid, devices))
                    return run_sharded(source_grid);
            } else {
                shards.reset();
            }
        }
        if constexpr (declared_constant != 0 && internal::speculation_capable<F>()) {
            // fields declared constant by the cell type: no run-time detection on top of that
            if (declared_passthrough_pays() && source_grid.get_grid_height() > 0 &&
                source_grid.get_grid_width() > 0)
                return run_declared(source_grid);
        } else if constexpr (internal::speculation_capable<F>()) {
            if (speculation_enabled() && !speculation_has_nothing_left() &&
                source_grid.get_grid_height() > 0 && source_grid.get_grid_width() > 0)
                return run_speculative(source_grid);
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

    // ---- one update spread over several GPUs of the box (Params::cuda_devices / STST_DEVICES) ----------

    using Slab = internal::SlabUpdate<F>;

    struct ShardSet {
        ~ShardSet() {
            // a slab may still be pushing rows into a neighbour: wait for all before freeing any
            for (auto &slab : slabs)
                if (slab)
                    slab->synchronize();
        }
        std::vector<int> devices;
        std::size_t grid_h = 0, grid_w = 0;
        unsigned fused_override = 0, tile_rows = 0;
        std::vector<std::unique_ptr<Slab>> slabs;
        std::vector<std::unique_ptr<internal::Event>> done; ///< per slab: result rows are in place
        bool speculating = false;
    };

    std::vector<int> resolve_devices() const {
        if (!params.cuda_devices.empty())
            return params.cuda_devices;
        std::vector<int> devices;
        const char *env = std::getenv("STST_DEVICES");
        if (!env || !*env)
            return devices;
        // "0,1,2,3", "0-7" or a mixture ("0-3,6,7")
        const std::string text(env);
        std::size_t pos = 0;
        while (pos < text.size()) {
            std::size_t end = text.find(',', pos);
            if (end == std::string::npos)
                end = text.size();
            const std::string item = text.substr(pos, end - pos);
            const std::size_t dash = item.find('-', 1);
            try {
                if (dash == std::string::npos) {
                    devices.push_back(std::stoi(item));
                } else {
                    const int lo = std::stoi(item.substr(0, dash)), hi = std::stoi(item.substr(dash + 1));
                    for (int d = lo; d <= hi; d++)
                        devices.push_back(d);
                }
            } catch (std::exception const &) {
                throw std::invalid_argument("StencilStream-B200: cannot parse STST_DEVICES=\"" + text +
                                            "\"");
            }
            pos = end + 1;
        }
        return devices;
    }

    /// Build (or reuse) the slabs for this grid and device list. Returns false if the grid cannot be
    /// cut that way (fewer rows per slab than the ghost depth needs): the caller then runs unsharded.
    bool ensure_shards(GridImpl &source_grid, std::vector<int> const &devices) {
        using namespace internal;
        auto &source = source_grid.get_storage();
        if (shards && shards->devices == devices && shards->grid_h == source.height &&
            shards->grid_w == source.width && shards->fused_override == params.fused_iterations &&
            shards->tile_rows == params.tile_rows)
            return true;
        shards.reset();
        int n_devices = 0;
        STST_RT_CHECK(stst_device_count(&n_devices));
        for (int d : devices)
            if (d < 0 || d >= n_devices)
                throw std::invalid_argument("StencilStream-B200: no CUDA device " + std::to_string(d) +
                                            " (cuda_devices / STST_DEVICES)");
        for (std::size_t count = std::min<std::size_t>(devices.size(), source.height); count > 1;
             count--) {
            auto set = std::make_unique<ShardSet>();
            set->devices = devices;
            set->grid_h = source.height;
            set->grid_w = source.width;
            set->fused_override = params.fused_iterations;
            set->tile_rows = params.tile_rows;
            try {
                unsigned fused = params.fused_iterations;
                for (int attempt = 0; attempt < 2; attempt++) {
                    set->slabs.clear();
                    unsigned k_min = ~0u, k_max = 0;
                    for (std::size_t i = 0; i < count; i++) {
                        typename Slab::Config cfg{};
                        cfg.grid_rows = source.height;
                        cfg.grid_cols = source.width;
                        partition_rows(source.height, count, i, cfg.row_lo, cfg.row_hi);
                        cfg.device = devices[i];
                        cfg.fused_iterations = fused;
                        cfg.tile_rows = params.tile_rows;
                        cfg.overlap = true;
                        set->slabs.push_back(std::make_unique<Slab>(cfg));
                        const unsigned k = set->slabs.back()->get_plan().fused_iterations;
                        k_min = std::min(k_min, k);
                        k_max = std::max(k_max, k);
                    }
                    if (k_min == k_max)
                        break;
                    fused = k_min; // all slabs must fuse the same depth: it fixes the ghost rows
                }
            } catch (std::invalid_argument const &) {
                continue; // a slab would own fewer rows than its neighbour's ghost depth: use fewer
            }
            // neighbours push into each other, the grid's device copies to and from every slab
            for (std::size_t i = 0; i < count; i++) {
                const int me = devices[i];
                auto allow = [&](int a, int b) {
                    if (a == b)
                        return;
                    int can = 0;
                    STST_RT_CHECK(stst_peer_can_access(a, b, &can));
                    if (!can)
                        throw std::runtime_error("StencilStream-B200: devices " + std::to_string(a) +
                                                 " and " + std::to_string(b) +
                                                 " cannot access each other's memory");
                    STST_RT_CHECK(stst_peer_enable(a, b));
                };
                if (i > 0) {
                    allow(me, devices[i - 1]);
                    auto const &up = set->slabs[i - 1]->get_config();
                    set->slabs[i]->attach(SlabSide::up, set->slabs[i - 1]->device_base(), up.row_lo,
                                          up.row_hi);
                }
                if (i + 1 < count) {
                    allow(me, devices[i + 1]);
                    auto const &down = set->slabs[i + 1]->get_config();
                    set->slabs[i]->attach(SlabSide::down, set->slabs[i + 1]->device_base(),
                                          down.row_lo, down.row_hi);
                }
                allow(me, source.device);
                allow(source.device, me);
                // (a CUDA event belongs to the device that is current when it is created, and can
                // only be recorded on that device's streams)
                STST_RT_CHECK(stst_set_device(me));
                set->done.push_back(std::make_unique<Event>());
            }
            set->speculating = speculation_enabled();
            if (set->speculating)
                for (auto &slab : set->slabs)
                    set->speculating = slab->enable_speculation(true) && set->speculating;
            if (!set->speculating)
                for (auto &slab : set->slabs)
                    slab->enable_speculation(false);
            shards = std::move(set);
            return true;
        }
        return false;
    }

    /**
     * The generation loop on row slabs, one per device: fill the slabs (owned and ghost rows) from
     * the source grid, advance all of them pass by pass — the passes of the slabs are enqueued
     * round-robin, so that no device's launch queue fills up with work that waits for a neighbour
     * whose work is not enqueued yet — and gather the owned rows into the result grid. Everything is
     * stream-ordered; the result grid's stream waits for the gather. With plane pass-through active
     * the call ends by collecting the slabs' verification flags, and is repeated from the (never
     * modified) source grid if any slab saw a kept plane change.
     */
    GridImpl run_sharded(GridImpl &source_grid) {
        using namespace internal;
        auto &source = source_grid.get_storage();
        source.require_device();
        GridImpl result = source_grid.make_similar();
        auto &target = result.get_storage();
        target.allocate_device();
        STST_RT_CHECK(stst_set_device(source.device)); // `ready` and the profiling events live there
        Event ready;
        const std::size_t k = shards->slabs.front()->get_plan().fused_iterations;
        std::size_t launches_before = 0;
        for (auto &slab : shards->slabs)
            launches_before += slab->get_n_launches();

        std::shared_ptr<Event> start, stop;
        if (params.profiling) {
            start = std::make_shared<Event>(true);
            stop = std::make_shared<Event>(true);
            start->record(source.stream);
        }
        for (;;) {
            ready.record(source.stream);
            for (auto &slab : shards->slabs)
                slab->load_from_grid(source.planes, source.device, ready.get());
            std::size_t iteration = params.iteration_offset, remaining = params.n_iterations;
            while (remaining > 0) {
                const std::size_t n_gens = std::min(remaining, k);
                for (auto &slab : shards->slabs)
                    slab->run(params.transition_function, params.halo_value, iteration, n_gens);
                iteration += n_gens;
                remaining -= n_gens;
            }
            unsigned violated = 0;
            if (shards->speculating) {
                bool useful = false;
                for (auto &slab : shards->slabs) {
                    violated |= slab->take_violations();
                    useful = useful || slab->passthrough_planes() != 0;
                }
                if (violated == 0 && !useful) {
                    shards->speculating = false; // nothing left to pass through: drop the protocol
                    for (auto &slab : shards->slabs)
                        slab->enable_speculation(false);
                }
            }
            if (violated == 0)
                break;
            for (auto &slab : shards->slabs)
                slab->drop_passthrough(violated);
            n_spec_redos++;
        }
        for (std::size_t i = 0; i < shards->slabs.size(); i++) {
            shards->slabs[i]->store_to_grid(target.planes, target.device, shards->done[i]->get());
            STST_RT_CHECK(stst_stream_wait_event(target.stream, shards->done[i]->get()));
        }
        if (params.profiling) {
            stop->record(target.stream);
            profile_events.emplace_back(std::move(start), std::move(stop));
        }
        target.device_written();
        last_plan = shards->slabs.front()->get_active_plan();
        std::size_t launches_after = 0;
        for (auto &slab : shards->slabs)
            launches_after += slab->get_n_launches();
        n_launches += launches_after - launches_before;
        // leave the thread on the grid's device, as a single-device update does
        STST_RT_CHECK(stst_set_device(source.device));
        return result;
    }

    void launch(internal::LaunchPlan const &plan, internal::GridStorage<Cell> &src,
                internal::GridStorage<Cell> &dst, std::size_t iteration0, unsigned n_gens,
                internal::Speculation const *spec = nullptr) {
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
            STST_RT_CHECK(stst_set_device(src.device)); // events belong to the current device
            start = std::make_shared<Event>(true);
            stop = std::make_shared<Event>(true);
            start->record(src.stream);
        }

        SweepLauncher<F>::launch(plan, params.transition_function, params.halo_value, src.planes,
                                 dst.planes, nullptr, region, iteration0, n_gens, tensor_maps,
                                 src.stream, spec);
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
    // speculative plane pass-through (run_speculative)
    bool spec_probed = false;
    unsigned spec_keep[internal::max_spec_subiterations] = {};
    unsigned *spec_flags = nullptr;
    unsigned *spec_host_flags = nullptr; ///< pinned landing zone of spec_flags, allocated once
    int spec_device = 0;
    std::size_t n_spec_redos = 0;
    // row slabs of the multi-GPU path (run_sharded); rebuilt when grid shape or device list change
    std::unique_ptr<ShardSet> shards;
};

} // namespace cuda
} // namespace stencil
