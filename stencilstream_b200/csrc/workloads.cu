/*
 * libstst_workloads — instantiates the header-only B200 backend (StencilStream/cuda/*.hpp) for the
 * transition functions in workloads/functors.hpp and exposes the resulting Grid / StencilUpdate
 * objects through the C ABI of include/stst_workloads.h.
 *
 * All compute goes through stencil::cuda::StencilUpdate -> fused_sweep_kernel; nothing in this file
 * computes cells on the host.
 */
#include <StencilStream/cuda/StencilUpdate.hpp>
#include <StencilStream/cuda/internal/SlabUpdate.hpp>
#include <stst_workloads.h>

#include "workloads/functors.hpp"

#include <algorithm>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

using namespace stst_workloads;
namespace sc = stencil::cuda;

thread_local std::string g_error;

int report(int code, std::string message) {
    g_error = std::move(message);
    return code;
}

// ---- type-erased object model -------------------------------------------------------------------

struct GridBase {
    virtual ~GridBase() = default;
    const char *workload = nullptr;
    virtual std::size_t rows() const = 0;
    virtual std::size_t cols() const = 0;
    virtual std::size_t cell_bytes() const = 0;
    virtual void copy_from_host(const void *cells) = 0;
    virtual void copy_to_host(void *cells) = 0;
    virtual void sync_to_device() = 0;
    virtual void *host_accessor(int mode) = 0;
    virtual bool host_pinned() = 0;
    virtual void max_abs(const stst_field_extent *extents, std::size_t n, double *out) = 0;
    virtual std::size_t field_bytes(std::size_t field) const = 0;
    virtual void copy_field(std::size_t field, void *host, bool to_device) = 0;
    virtual GridBase *share() = 0;
    virtual GridBase *make_similar() = 0;
};

template <typename Cell> struct GridHolder final : GridBase {
    sc::Grid<Cell> grid;
    GridHolder(const char *name, sc::Grid<Cell> g) : grid(std::move(g)) { workload = name; }
    std::size_t rows() const override { return grid.get_grid_height(); }
    std::size_t cols() const override { return grid.get_grid_width(); }
    std::size_t cell_bytes() const override { return sizeof(Cell); }
    void copy_from_host(const void *cells) override {
        // copy_from_buffer semantics without the intermediate sycl::buffer
        grid.copy_from_host(static_cast<const Cell *>(cells));
    }
    void copy_to_host(void *cells) override {
        grid.copy_to_host(static_cast<Cell *>(cells));
    }
    void sync_to_device() override {
        grid.get_storage().require_device();
        sc::internal::check(stst_stream_synchronize(grid.get_storage().stream), "stream sync");
    }
    void *host_accessor(int mode) override {
        if (mode == 0) {
            typename sc::Grid<Cell>::template GridAccessor<sycl::access::mode::read> ac(grid);
            return const_cast<void *>(static_cast<const void *>(ac.get_pointer()));
        }
        typename sc::Grid<Cell>::template GridAccessor<sycl::access::mode::read_write> ac(grid);
        return static_cast<void *>(ac.get_pointer());
    }
    bool host_pinned() override { return grid.get_storage().host_mirror_is_pinned(); }
    void max_abs(const stst_field_extent *extents, std::size_t n, double *out) override {
        std::vector<sc::FieldExtent> list(n);
        for (std::size_t q = 0; q < n; q++)
            list[q] = sc::FieldExtent{extents[q].field, extents[q].rows, extents[q].cols};
        const std::vector<double> result = grid.max_abs(list);
        std::copy(result.begin(), result.end(), out);
    }
    std::size_t field_bytes(std::size_t field) const override {
        return sc::Grid<Cell>::plane_element_bytes(field);
    }
    void copy_field(std::size_t field, void *host, bool to_device) override {
        if (to_device)
            grid.copy_plane_from_host(field, host);
        else
            grid.copy_plane_to_host(field, host);
    }
    GridBase *share() override { return new GridHolder(workload, grid); }
    GridBase *make_similar() override { return new GridHolder(workload, grid.make_similar()); }
};

struct UpdateBase {
    virtual ~UpdateBase() = default;
    const char *workload = nullptr;
    virtual void set_params(const stst_update_params &p) = 0;
    virtual GridBase *apply(GridBase &source) = 0;
    virtual void stats(stst_update_stats &s) = 0;
};

template <typename F, typename ParamBlock> struct UpdateHolder final : UpdateBase {
    using Update = sc::StencilUpdate<F>;
    using Cell = typename F::Cell;
    std::unique_ptr<Update> update;

    static typename Update::Params convert(const stst_update_params &p) {
        if (p.transition_function_bytes != sizeof(ParamBlock) || p.transition_function == nullptr)
            throw std::invalid_argument("transition_function_bytes does not match the workload's "
                                        "parameter struct");
        if (p.halo_value != nullptr && p.halo_value_bytes != sizeof(Cell))
            throw std::invalid_argument("halo_value_bytes does not match the workload's cell type");
        typename Update::Params out{};
        std::memcpy(&out.transition_function.p, p.transition_function, sizeof(ParamBlock));
        if (p.halo_value != nullptr)
            std::memcpy(static_cast<void *>(&out.halo_value), p.halo_value, sizeof(Cell));
        out.iteration_offset = p.iteration_offset;
        out.n_iterations = p.n_iterations;
        out.blocking = p.blocking != 0;
        out.profiling = p.profiling != 0;
        out.cuda_device = p.cuda_device;
        out.fused_iterations = p.fused_iterations;
        out.tile_rows = p.tile_rows;
        return out;
    }

    UpdateHolder(const char *name, const stst_update_params &p)
        : update(std::make_unique<Update>(convert(p))) {
        workload = name;
    }

    void set_params(const stst_update_params &p) override { update->get_params() = convert(p); }

    GridBase *apply(GridBase &source) override {
        auto *typed = dynamic_cast<GridHolder<Cell> *>(&source);
        if (!typed)
            throw std::invalid_argument("grid belongs to a workload with a different cell type");
        sc::Grid<Cell> result = (*update)(typed->grid);
        return new GridHolder<Cell>(source.workload, result);
    }

    void stats(stst_update_stats &s) override {
        std::memset(&s, 0, sizeof(s));
        s.n_processed_cells = update->get_n_processed_cells();
        s.walltime = update->get_walltime();
        s.kernel_runtime = update->get_kernel_runtime();
        s.n_launches = update->get_n_launches();
        auto const &plan = update->get_last_plan();
        s.fused_iterations = plan.fused_iterations;
        s.tile_h = plan.tile_h;
        s.tile_w = plan.tile_w;
        s.block_x = plan.block_x;
        s.block_y = plan.block_y;
        s.use_tma = plan.use_tma ? 1u : 0u;
        s.smem_bytes = plan.smem_bytes;
        s.passthrough_planes = update->get_passthrough_planes();
        s.speculation_redos = update->get_n_speculation_redos();
    }
};

struct SlabBase {
    virtual ~SlabBase() = default;
    const char *workload = nullptr;
    std::vector<void *> ipc_mappings;
    virtual std::size_t cell_bytes() const = 0;
    virtual void info(stst_slab_info &out) = 0;
    virtual void *device_base() = 0;
    virtual int device() const = 0;
    virtual void attach(int side, void *mapped, std::size_t lo, std::size_t hi) = 0;
    virtual void upload(const void *cells, std::size_t first_row, std::size_t n_rows) = 0;
    virtual void download(void *cells, std::size_t first_row, std::size_t n_rows) = 0;
    virtual void exchange() = 0;
    virtual void max_abs(const stst_field_extent *extents, std::size_t n, double *out) = 0;
    virtual std::size_t field_bytes(std::size_t field) const = 0;
    virtual void download_field(std::size_t field, void *host, std::size_t first_row,
                                std::size_t n_rows) = 0;
    virtual void update(const stst_update_params &p) = 0;
    virtual void synchronize() = 0;
    virtual void record(void *event) = 0;
    virtual void copy_from(SlabBase &other) = 0;
    virtual sc::internal::PlaneSet current_planes() = 0;
    virtual std::size_t ghost() const = 0;
    virtual bool enable_speculation(bool on) = 0;
    virtual void backup() = 0;
    virtual void restore() = 0;
    virtual unsigned take_violations() = 0;
    virtual void drop_passthrough(unsigned planes) = 0;
};

template <typename F, typename ParamBlock> struct SlabHolder final : SlabBase {
    using Slab = sc::internal::SlabUpdate<F>;
    using Cell = typename F::Cell;
    std::unique_ptr<Slab> slab;

    SlabHolder(const char *name, typename Slab::Config const &cfg)
        : slab(std::make_unique<Slab>(cfg)) {
        workload = name;
    }
    ~SlabHolder() override {
        slab.reset();
        for (void *m : ipc_mappings)
            (void)stst_ipc_close_mem_handle(m);
    }
    std::size_t cell_bytes() const override { return sizeof(Cell); }
    void info(stst_slab_info &out) override {
        std::memset(&out, 0, sizeof(out));
        auto const &cfg = slab->get_config();
        auto const &plan = slab->get_active_plan();
        out.grid_rows = cfg.grid_rows;
        out.grid_cols = cfg.grid_cols;
        out.row_lo = cfg.row_lo;
        out.row_hi = cfg.row_hi;
        out.ghost_rows = slab->ghost_rows();
        out.device_bytes = slab->device_bytes();
        out.n_launches = slab->get_n_launches();
        out.epoch = slab->get_epoch();
        out.device = cfg.device;
        out.fused_iterations = plan.fused_iterations;
        out.tile_h = plan.tile_h;
        out.tile_w = plan.tile_w;
        out.block_x = plan.block_x;
        out.block_y = plan.block_y;
        out.use_tma = plan.use_tma ? 1u : 0u;
        out.overlap = cfg.overlap ? 1u : 0u;
        out.smem_bytes = plan.smem_bytes;
        out.passthrough_planes = slab->passthrough_planes();
    }
    void *device_base() override { return slab->device_base(); }
    int device() const override { return slab->get_config().device; }
    void attach(int side, void *mapped, std::size_t lo, std::size_t hi) override {
        slab->attach(side == 0 ? sc::internal::SlabSide::up : sc::internal::SlabSide::down, mapped,
                     lo, hi);
    }
    void upload(const void *cells, std::size_t first_row, std::size_t n_rows) override {
        slab->upload_rows(static_cast<const Cell *>(cells), first_row, n_rows);
    }
    void download(void *cells, std::size_t first_row, std::size_t n_rows) override {
        slab->download_rows(static_cast<Cell *>(cells), first_row, n_rows);
    }
    void exchange() override { slab->exchange_halos(); }
    void max_abs(const stst_field_extent *extents, std::size_t n, double *out) override {
        std::vector<std::size_t> planes(n), rows(n), cols(n);
        for (std::size_t q = 0; q < n; q++) {
            planes[q] = extents[q].field;
            rows[q] = extents[q].rows;
            cols[q] = extents[q].cols;
        }
        slab->max_abs(n, planes.data(), rows.data(), cols.data(), out);
    }
    std::size_t field_bytes(std::size_t field) const override {
        return sc::Grid<Cell>::plane_element_bytes(field);
    }
    void download_field(std::size_t field, void *host, std::size_t first_row,
                        std::size_t n_rows) override {
        slab->download_plane_rows(field, host, first_row, n_rows);
    }
    void update(const stst_update_params &p) override {
        auto params = UpdateHolder<F, ParamBlock>::convert(p);
        slab->run(params.transition_function, params.halo_value, params.iteration_offset,
                  params.n_iterations);
        if (params.blocking)
            slab->synchronize();
    }
    void synchronize() override { slab->synchronize(); }
    void record(void *event) override { slab->record(event); }
    sc::internal::PlaneSet current_planes() override { return slab->current_planes(); }
    std::size_t ghost() const override { return slab->ghost_rows(); }
    void copy_from(SlabBase &other) override {
        stst_slab_info mine, theirs;
        info(mine);
        other.info(theirs);
        if (other.cell_bytes() != cell_bytes() || theirs.grid_rows != mine.grid_rows ||
            theirs.grid_cols != mine.grid_cols || theirs.row_lo != mine.row_lo ||
            theirs.row_hi != mine.row_hi)
            throw std::range_error("The source slab has not the same rows, columns or cell type");
        if (other.device() != device())
            throw std::invalid_argument("the two slabs live on different devices");
        other.synchronize(); // its current generation must be complete before it is read
        slab->copy_owned_rows_from(other.current_planes(), other.ghost());
    }
    bool enable_speculation(bool on) override { return slab->enable_speculation(on); }
    void backup() override { slab->backup(); }
    void restore() override { slab->restore(); }
    unsigned take_violations() override { return slab->take_violations(); }
    void drop_passthrough(unsigned planes) override { slab->drop_passthrough(planes); }
};

struct WorkloadEntry {
    const char *name;
    stst_workload_info info;
    GridBase *(*make_grid)(const char *, std::size_t, std::size_t, int);
    UpdateBase *(*make_update)(const char *, const stst_update_params &);
    SlabBase *(*make_slab)(const char *, std::size_t, std::size_t, std::size_t, std::size_t, int,
                           unsigned, unsigned, bool);
};

template <typename F, typename ParamBlock> WorkloadEntry make_entry(const char *name) {
    using Cell = typename F::Cell;
    WorkloadEntry e{};
    e.name = name;
    e.info.cell_bytes = sizeof(Cell);
    e.info.params_bytes = sizeof(ParamBlock);
    e.info.n_planes = sc::internal::CellLayout<Cell>::n_planes;
    e.info.stencil_radius = F::stencil_radius;
    e.info.n_subiterations = F::n_subiterations;
    e.info.bytes_per_cell_iteration = 2 * sizeof(Cell) * F::n_subiterations;
    e.make_grid = [](const char *n, std::size_t r, std::size_t c, int device) -> GridBase * {
        if (device < 0)
            device = sc::internal::default_device_ordinal();
        return new GridHolder<Cell>(n, sc::Grid<Cell>(r, c, device));
    };
    e.make_update = [](const char *n, const stst_update_params &p) -> UpdateBase * {
        return new UpdateHolder<F, ParamBlock>(n, p);
    };
    e.make_slab = [](const char *n, std::size_t grid_rows, std::size_t grid_cols, std::size_t row_lo,
                     std::size_t row_hi, int device, unsigned fused, unsigned tile_rows,
                     bool overlap) -> SlabBase * {
        if (device < 0)
            device = sc::internal::default_device_ordinal();
        typename sc::internal::SlabUpdate<F>::Config cfg{grid_rows, grid_cols, row_lo, row_hi,
                                                          device,    fused,     tile_rows, overlap};
        return new SlabHolder<F, ParamBlock>(n, cfg);
    };
    return e;
}

const std::vector<WorkloadEntry> &registry() {
    static const std::vector<WorkloadEntry> entries = {
        make_entry<ConwayRule, stst_conway_params>("conway"),
        make_entry<Jacobi5Rule, stst_jacobi5_params>("jacobi5"),
        make_entry<Jacobi9Rule, stst_jacobi9_params>("jacobi9"),
        make_entry<JacobiStarRule<2>, stst_jacobi_star_params>("jacobi_r2"),
        make_entry<JacobiStarRule<3>, stst_jacobi_star_params>("jacobi_r3"),
        make_entry<HotspotRule, stst_hotspot_params>("hotspot"),
        make_entry<FdtdCoefRule, stst_fdtd_params>("fdtd"),
        make_entry<ConvectionPseudoTransientRule, stst_convection_pt_params>("convection_pt"),
        make_entry<ConvectionThermalRule, stst_convection_thermal_params>("convection_thermal"),
        make_entry<KatRule<1>, stst_kat_params>("kat"),
        make_entry<KatRule<2>, stst_kat_params>("kat_r2"),
    };
    return entries;
}

const WorkloadEntry *find(const char *name) {
    if (!name)
        return nullptr;
    for (auto const &e : registry())
        if (std::strcmp(e.name, name) == 0)
            return &e;
    return nullptr;
}

template <typename Fn> int guarded(Fn &&fn) {
    try {
        return fn();
    } catch (std::range_error const &e) {
        return report(STST_ERR_RANGE, e.what());
    } catch (std::invalid_argument const &e) {
        return report(STST_ERR_INVALID_ARGUMENT, e.what());
    } catch (std::exception const &e) {
        return report(STST_ERR_RUNTIME, e.what());
    } catch (...) {
        return report(STST_ERR_RUNTIME, "unknown exception");
    }
}

} // namespace

struct stst_grid {
    std::unique_ptr<GridBase> impl;
};
struct stst_update {
    std::unique_ptr<UpdateBase> impl;
};
struct stst_slab {
    std::unique_ptr<SlabBase> impl;
};

#define STST_EXPORT extern "C" __attribute__((visibility("default")))

STST_EXPORT int stst_workloads_abi_version(void) { return STST_WORKLOADS_ABI_VERSION; }

STST_EXPORT const char *stst_workloads_last_error(void) { return g_error.c_str(); }

STST_EXPORT int stst_workload_count(void) { return int(registry().size()); }

STST_EXPORT const char *stst_workload_name(int index) {
    if (index < 0 || index >= int(registry().size()))
        return nullptr;
    return registry()[index].name;
}

STST_EXPORT int stst_workload_get_info(const char *workload, stst_workload_info *info) {
    const WorkloadEntry *e = find(workload);
    if (!e)
        return report(STST_ERR_UNKNOWN_WORKLOAD, std::string("unknown workload: ") +
                                                     (workload ? workload : "(null)"));
    if (!info)
        return report(STST_ERR_INVALID_ARGUMENT, "info is null");
    *info = e->info;
    return STST_OK;
}

STST_EXPORT int stst_grid_create(const char *workload, size_t rows, size_t cols, int device,
                                 stst_grid **grid) {
    const WorkloadEntry *e = find(workload);
    if (!e)
        return report(STST_ERR_UNKNOWN_WORKLOAD, std::string("unknown workload: ") +
                                                     (workload ? workload : "(null)"));
    if (!grid)
        return report(STST_ERR_INVALID_ARGUMENT, "grid is null");
    return guarded([&] {
        *grid = new stst_grid{std::unique_ptr<GridBase>(e->make_grid(e->name, rows, cols, device))};
        return STST_OK;
    });
}

STST_EXPORT int stst_grid_share(stst_grid *grid, stst_grid **other) {
    if (!grid || !other)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        *other = new stst_grid{std::unique_ptr<GridBase>(grid->impl->share())};
        return STST_OK;
    });
}

STST_EXPORT int stst_grid_make_similar(stst_grid *grid, stst_grid **other) {
    if (!grid || !other)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        *other = new stst_grid{std::unique_ptr<GridBase>(grid->impl->make_similar())};
        return STST_OK;
    });
}

STST_EXPORT int stst_grid_destroy(stst_grid *grid) {
    delete grid;
    return STST_OK;
}

STST_EXPORT int stst_grid_shape(const stst_grid *grid, size_t *rows, size_t *cols) {
    if (!grid || !rows || !cols)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    *rows = grid->impl->rows();
    *cols = grid->impl->cols();
    return STST_OK;
}

STST_EXPORT int stst_grid_copy_from_host(stst_grid *grid, const void *cells, size_t bytes) {
    if (!grid || (!cells && bytes != 0))
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    if (bytes != grid->impl->rows() * grid->impl->cols() * grid->impl->cell_bytes())
        return report(STST_ERR_RANGE, "The target buffer has not the same size as the grid");
    return guarded([&] {
        grid->impl->copy_from_host(cells);
        return STST_OK;
    });
}

STST_EXPORT int stst_grid_copy_to_host(stst_grid *grid, void *cells, size_t bytes) {
    if (!grid || (!cells && bytes != 0))
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    if (bytes != grid->impl->rows() * grid->impl->cols() * grid->impl->cell_bytes())
        return report(STST_ERR_RANGE, "The target buffer has not the same size as the grid");
    return guarded([&] {
        grid->impl->copy_to_host(cells);
        return STST_OK;
    });
}

STST_EXPORT int stst_grid_sync_to_device(stst_grid *grid) {
    if (!grid)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        grid->impl->sync_to_device();
        return STST_OK;
    });
}

STST_EXPORT int stst_grid_host_accessor(stst_grid *grid, int mode, void **cells) {
    if (!grid || !cells || mode < 0 || mode > 2)
        return report(STST_ERR_INVALID_ARGUMENT, "bad argument");
    return guarded([&] {
        *cells = grid->impl->host_accessor(mode);
        return STST_OK;
    });
}

STST_EXPORT int stst_grid_host_image_is_pinned(stst_grid *grid, int *pinned) {
    if (!grid || !pinned)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    *pinned = grid->impl->host_pinned() ? 1 : 0;
    return STST_OK;
}

STST_EXPORT int stst_grid_max_abs(stst_grid *grid, const stst_field_extent *extents, size_t n,
                                  double *out) {
    if (!grid || (n != 0 && (!extents || !out)))
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        grid->impl->max_abs(extents, n, out);
        return STST_OK;
    });
}

namespace {
int grid_copy_field(stst_grid *grid, size_t field, void *host, size_t bytes, bool to_device) {
    if (!grid || (!host && bytes != 0))
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        const size_t want = grid->impl->rows() * grid->impl->cols() * grid->impl->field_bytes(field);
        if (bytes != want)
            return report(STST_ERR_RANGE, "The target buffer has not the same size as the field");
        grid->impl->copy_field(field, host, to_device);
        return STST_OK;
    });
}
} // namespace

STST_EXPORT int stst_grid_copy_field_to_host(stst_grid *grid, size_t field, void *values,
                                             size_t bytes) {
    return grid_copy_field(grid, field, values, bytes, false);
}

STST_EXPORT int stst_grid_copy_field_from_host(stst_grid *grid, size_t field, const void *values,
                                               size_t bytes) {
    return grid_copy_field(grid, field, const_cast<void *>(values), bytes, true);
}

STST_EXPORT int stst_update_create(const char *workload, const stst_update_params *params,
                                   stst_update **update) {
    const WorkloadEntry *e = find(workload);
    if (!e)
        return report(STST_ERR_UNKNOWN_WORKLOAD, std::string("unknown workload: ") +
                                                     (workload ? workload : "(null)"));
    if (!params || !update)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        *update = new stst_update{std::unique_ptr<UpdateBase>(e->make_update(e->name, *params))};
        return STST_OK;
    });
}

STST_EXPORT int stst_update_set_params(stst_update *update, const stst_update_params *params) {
    if (!update || !params)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        update->impl->set_params(*params);
        return STST_OK;
    });
}

STST_EXPORT int stst_update_apply(stst_update *update, stst_grid *source, stst_grid **result) {
    if (!update || !source || !result)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        *result = new stst_grid{std::unique_ptr<GridBase>(update->impl->apply(*source->impl))};
        return STST_OK;
    });
}

STST_EXPORT int stst_update_get_stats(stst_update *update, stst_update_stats *stats) {
    if (!update || !stats)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        update->impl->stats(*stats);
        return STST_OK;
    });
}

STST_EXPORT int stst_update_destroy(stst_update *update) {
    delete update;
    return STST_OK;
}


// ---- row slabs ---------------------------------------------------------------------------------------

STST_EXPORT int stst_slab_create(const char *workload, size_t grid_rows, size_t grid_cols,
                                 size_t row_lo, size_t row_hi, int device,
                                 unsigned fused_iterations, unsigned tile_rows, int overlap,
                                 stst_slab **slab) {
    const WorkloadEntry *e = find(workload);
    if (!e)
        return report(STST_ERR_UNKNOWN_WORKLOAD, std::string("unknown workload: ") +
                                                     (workload ? workload : "(null)"));
    if (!slab)
        return report(STST_ERR_INVALID_ARGUMENT, "slab is null");
    if (row_hi <= row_lo || row_hi > grid_rows || grid_cols == 0)
        return report(STST_ERR_INVALID_ARGUMENT, "illegal slab row range");
    return guarded([&] {
        *slab = new stst_slab{std::unique_ptr<SlabBase>(
            e->make_slab(e->name, grid_rows, grid_cols, row_lo, row_hi, device, fused_iterations,
                         tile_rows, overlap != 0))};
        return STST_OK;
    });
}

STST_EXPORT int stst_slab_destroy(stst_slab *slab) {
    delete slab;
    return STST_OK;
}

STST_EXPORT int stst_slab_get_info(stst_slab *slab, stst_slab_info *info) {
    if (!slab || !info)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        slab->impl->info(*info);
        return STST_OK;
    });
}

STST_EXPORT int stst_slab_get_ipc_handle(stst_slab *slab, unsigned char handle[64]) {
    if (!slab || !handle)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        sc::internal::check(stst_ipc_get_mem_handle(slab->impl->device_base(), handle),
                            "stst_ipc_get_mem_handle");
        return STST_OK;
    });
}

STST_EXPORT int stst_slab_attach_ipc(stst_slab *slab, int side, const unsigned char handle[64],
                                     size_t peer_row_lo, size_t peer_row_hi) {
    if (!slab || !handle || side < 0 || side > 1)
        return report(STST_ERR_INVALID_ARGUMENT, "bad argument");
    return guarded([&] {
        sc::internal::check(stst_set_device(slab->impl->device()), "stst_set_device");
        void *mapped = nullptr;
        sc::internal::check(stst_ipc_open_mem_handle(handle, &mapped), "stst_ipc_open_mem_handle");
        slab->impl->ipc_mappings.push_back(mapped);
        slab->impl->attach(side, mapped, peer_row_lo, peer_row_hi);
        return STST_OK;
    });
}

STST_EXPORT int stst_slab_attach_local(stst_slab *slab, int side, stst_slab *peer) {
    if (!slab || !peer || side < 0 || side > 1)
        return report(STST_ERR_INVALID_ARGUMENT, "bad argument");
    return guarded([&] {
        stst_slab_info theirs;
        peer->impl->info(theirs);
        if (peer->impl->device() != slab->impl->device()) {
            int can = 0;
            sc::internal::check(stst_peer_can_access(slab->impl->device(), peer->impl->device(), &can),
                                "stst_peer_can_access");
            if (!can)
                throw std::runtime_error("the two slabs' devices cannot access each other");
            sc::internal::check(stst_peer_enable(slab->impl->device(), peer->impl->device()),
                                "stst_peer_enable");
        }
        slab->impl->attach(side, peer->impl->device_base(), theirs.row_lo, theirs.row_hi);
        return STST_OK;
    });
}

namespace {
int slab_copy_rows(stst_slab *slab, void *cells, size_t bytes, size_t first_row, size_t n_rows,
                   bool whole, bool to_device) {
    if (!slab || (!cells && bytes != 0))
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        stst_slab_info info;
        slab->impl->info(info);
        const size_t owned = info.row_hi - info.row_lo;
        if (whole) {
            first_row = 0;
            n_rows = owned;
        }
        if (first_row > owned || n_rows > owned - first_row ||
            bytes != n_rows * info.grid_cols * slab->impl->cell_bytes())
            return report(STST_ERR_RANGE, "The target buffer has not the same size as the slab rows");
        if (to_device)
            slab->impl->upload(cells, first_row, n_rows);
        else
            slab->impl->download(cells, first_row, n_rows);
        return STST_OK;
    });
}
} // namespace

STST_EXPORT int stst_slab_copy_from_host(stst_slab *slab, const void *cells, size_t bytes) {
    return slab_copy_rows(slab, const_cast<void *>(cells), bytes, 0, 0, true, true);
}

STST_EXPORT int stst_slab_copy_to_host(stst_slab *slab, void *cells, size_t bytes) {
    return slab_copy_rows(slab, cells, bytes, 0, 0, true, false);
}

STST_EXPORT int stst_slab_copy_rows_from_host(stst_slab *slab, size_t first_row, size_t n_rows,
                                              const void *cells, size_t bytes) {
    return slab_copy_rows(slab, const_cast<void *>(cells), bytes, first_row, n_rows, false, true);
}

STST_EXPORT int stst_slab_copy_rows_to_host(stst_slab *slab, size_t first_row, size_t n_rows,
                                            void *cells, size_t bytes) {
    return slab_copy_rows(slab, cells, bytes, first_row, n_rows, false, false);
}

STST_EXPORT int stst_slab_exchange_halos(stst_slab *slab) {
    if (!slab)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        slab->impl->exchange();
        return STST_OK;
    });
}

STST_EXPORT int stst_slab_max_abs(stst_slab *slab, const stst_field_extent *extents, size_t n,
                                  double *out) {
    if (!slab || (n != 0 && (!extents || !out)))
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        slab->impl->max_abs(extents, n, out);
        return STST_OK;
    });
}

STST_EXPORT int stst_slab_copy_field_rows_to_host(stst_slab *slab, size_t field, size_t first_row,
                                                  size_t n_rows, void *values, size_t bytes) {
    if (!slab || (!values && bytes != 0))
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        stst_slab_info info;
        slab->impl->info(info);
        const size_t owned = info.row_hi - info.row_lo;
        if (first_row > owned || n_rows > owned - first_row ||
            bytes != n_rows * info.grid_cols * slab->impl->field_bytes(field))
            return report(STST_ERR_RANGE, "The target buffer has not the same size as the field rows");
        slab->impl->download_field(field, values, first_row, n_rows);
        return STST_OK;
    });
}

STST_EXPORT int stst_slab_update(stst_slab *slab, const stst_update_params *params) {
    if (!slab || !params)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        slab->impl->update(*params);
        return STST_OK;
    });
}

STST_EXPORT int stst_slab_copy_from_slab(stst_slab *slab, stst_slab *source) {
    if (!slab || !source)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        slab->impl->copy_from(*source->impl);
        return STST_OK;
    });
}

STST_EXPORT int stst_slab_enable_speculation(stst_slab *slab, int enable, int *enabled) {
    if (!slab)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        const bool on = slab->impl->enable_speculation(enable != 0);
        if (enabled)
            *enabled = on ? 1 : 0;
        return STST_OK;
    });
}

STST_EXPORT int stst_slab_backup(stst_slab *slab) {
    if (!slab)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        slab->impl->backup();
        return STST_OK;
    });
}

STST_EXPORT int stst_slab_restore(stst_slab *slab) {
    if (!slab)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        slab->impl->restore();
        return STST_OK;
    });
}

STST_EXPORT int stst_slab_take_violations(stst_slab *slab, unsigned *planes) {
    if (!slab || !planes)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        *planes = slab->impl->take_violations();
        return STST_OK;
    });
}

STST_EXPORT int stst_slab_drop_passthrough(stst_slab *slab, unsigned planes) {
    if (!slab)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        slab->impl->drop_passthrough(planes);
        return STST_OK;
    });
}

STST_EXPORT int stst_slab_synchronize(stst_slab *slab) {
    if (!slab)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        slab->impl->synchronize();
        return STST_OK;
    });
}

STST_EXPORT int stst_slab_record_event(stst_slab *slab, void *event) {
    if (!slab || !event)
        return report(STST_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        slab->impl->record(event);
        return STST_OK;
    });
}
