/*
 * TEST INFRASTRUCTURE — parity oracle. NOT part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may build, load or call this.
 *
 * Plain-C restatement of the reference algorithm for the StencilStream generation loop:
 *
 *   - the sweep driver follows stencil::cpu::StencilUpdate::operator() and run_iter
 *     (/root/reference/StencilStream/cpu/StencilUpdate.hpp:109-142 ping-pong and loop order,
 *      :185-223 one sweep: bounds-checked (2r+1)^2 gather, `halo_value` outside the grid, the
 *      time-dependent value evaluated once per iteration at :197), which is also what the cuda
 *      backend computes (StencilStream/cuda/StencilUpdate.hpp:212-273);
 *   - each transition function follows the example/test functor cited at its definition.
 *
 * Pinning: oracle/_ref/liboracle_ref.so is the reference's own cpu backend and example sources
 * compiled in place; tests/test_oracle.py checks this restatement against it bit for bit (when the
 * reference tree is present) and against the golden vectors in tests/golden/ that were generated
 * from it (tests/golden/generate.py). The reference itself ships no numeric golden files for this
 * path; its self-checking test functor (tests/TransFuncs.hpp) is restated below as "kat".
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off -fopenmp -fPIC -shared (see oracle/recipes.py).
 * -ffp-contract=off keeps a*b+c as two rounded operations, like the reference oracle build.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "stst_workloads.h"

#define MAX_DIAMETER 7

/* What a transition function sees: the reference's Stencil object (StencilStream/Stencil.hpp:165-180). */
typedef struct stencil_view {
    const void *cell[MAX_DIAMETER][MAX_DIAMETER]; /* [row][col], north-west origin */
    size_t id[2];                                 /* global (row, column) of the centre */
    size_t grid_range[2];                         /* (rows, columns) */
    size_t iteration, subiteration;
    double tdv_f64;  /* time-dependent value, if it is a float (stored widened, exact) */
    size_t tdv_size; /* time-dependent value, if it is a size_t */
    int radius;
} stencil_view;

typedef void (*transition_fn)(const void *params, const stencil_view *st, void *next);
typedef void (*tdv_fn)(const void *params, size_t i_iteration, stencil_view *st);

typedef struct workload_def {
    const char *name;
    size_t cell_bytes;
    int radius;
    int n_subiterations;
    transition_fn fn;
    tdv_fn tdv;
} workload_def;

#define NB(type, st, dr, dc) ((const type *)(st)->cell[(dr) + (st)->radius][(dc) + (st)->radius])

/* ---- Conway: examples/conway/conway.cpp:35-56 ---------------------------------------------------- */
static void conway_fn(const void *params, const stencil_view *st, void *next) {
    (void)params;
    int alive = 0;
    for (int r = -1; r <= 1; r++)
        for (int c = -1; c <= 1; c++)
            if (*NB(uint8_t, st, r, c) && !(r == 0 && c == 0))
                alive += 1;
    uint8_t out;
    if (*NB(uint8_t, st, 0, 0))
        out = (alive == 2 || alive == 3);
    else
        out = (alive == 3);
    *(uint8_t *)next = out;
}

/* ---- Jacobi: examples/jacobi/kernels.hpp:236-272 (5-point), :274-319 (9-point) ---------------------- */
static void jacobi5_fn(const void *params, const stencil_view *st, void *next) {
    const stst_jacobi5_params *p = (const stst_jacobi5_params *)params;
    *(float *)next = p->coef[0] * *NB(float, st, -1, 0) + p->coef[1] * *NB(float, st, 0, -1) +
                     p->coef[2] * *NB(float, st, 1, 0) + p->coef[3] * *NB(float, st, 0, 1) +
                     p->coef[4] * *NB(float, st, 0, 0);
}

static void jacobi9_fn(const void *params, const stencil_view *st, void *next) {
    const stst_jacobi9_params *p = (const stst_jacobi9_params *)params;
    float sum = 0.0f;
    for (int r = -1; r <= 1; r++)
        for (int c = -1; c <= 1; c++)
            sum += p->coef[r + 1][c + 1] * *NB(float, st, r, c);
    *(float *)next = sum;
}

/* Radius-R star stencil: defined by this project (no reference functor), see
 * stencilstream_b200/csrc/workloads/functors.hpp JacobiStarRule. */
static void jacobi_star_fn(const void *params, const stencil_view *st, void *next) {
    const stst_jacobi_star_params *p = (const stst_jacobi_star_params *)params;
    float acc = p->centre * *NB(float, st, 0, 0);
    for (int d = 1; d <= st->radius; d++)
        acc += p->arm[d - 1] * (((*NB(float, st, -d, 0) + *NB(float, st, 0, -d)) +
                                 *NB(float, st, d, 0)) +
                                *NB(float, st, 0, d));
    *(float *)next = acc;
}

/* ---- HotSpot: examples/hotspot/hotspot.cpp:64-97 ------------------------------------------------------ */
static void hotspot_fn(const void *params, const stencil_view *st, void *next) {
    const stst_hotspot_params *p = (const stst_hotspot_params *)params;
    const float amb_temp = 80.0f;
    float power = NB(stst_hotspot_cell, st, 0, 0)->power;
    float old = NB(stst_hotspot_cell, st, 0, 0)->temp;
    float top = NB(stst_hotspot_cell, st, -1, 0)->temp;
    float bottom = NB(stst_hotspot_cell, st, 1, 0)->temp;
    float left = NB(stst_hotspot_cell, st, 0, -1)->temp;
    float right = NB(stst_hotspot_cell, st, 0, 1)->temp;

    if (st->id[0] == 0)
        top = old;
    else if (st->id[0] == st->grid_range[0] - 1)
        bottom = old;
    if (st->id[1] == 0)
        left = old;
    else if (st->id[1] == st->grid_range[1] - 1)
        right = old;

    float new_temp = old + p->Cap_1 * (power + (bottom + top - 2.f * old) * p->Ry_1 +
                                       (right + left - 2.f * old) * p->Rx_1 +
                                       (amb_temp - old) * p->Rz_1);
    stst_hotspot_cell *out = (stst_hotspot_cell *)next;
    out->temp = new_temp;
    out->power = power;
}

/* ---- FDTD (coef material): examples/fdtd/src/Kernel.hpp:80-128, material/CoefResolver.hpp:60-67 ------- */
static void fdtd_tdv(const void *params, size_t i_iteration, stencil_view *st) {
    const stst_fdtd_params *p = (const stst_fdtd_params *)params;
    float current_time = i_iteration * p->dt;
    float wave_progress = (current_time - p->t_0) / p->tau;
    float v = cosf(p->omega * current_time) * expf(-1 * wave_progress * wave_progress);
    st->tdv_f64 = v;
}

static void fdtd_fn(const void *params, const stencil_view *st, void *next) {
    const stst_fdtd_params *p = (const stst_fdtd_params *)params;
    stst_fdtd_cell cell = *NB(stst_fdtd_cell, st, 0, 0);

    float r = st->id[0];
    float c = st->id[1];
    float source_distance_score = r * (r - 2 * p->source_r) + c * (c - 2 * p->source_c);

    const stst_fdtd_cell *centre = NB(stst_fdtd_cell, st, 0, 0);
    float ca = centre->ca, cb = centre->cb, da = centre->da, db = centre->db;

    if (st->subiteration == 0) {
        cell.ex *= ca;
        cell.ex += cb * (centre->hz - NB(stst_fdtd_cell, st, 0, -1)->hz);
        cell.ey *= ca;
        cell.ey += cb * (NB(stst_fdtd_cell, st, -1, 0)->hz - centre->hz);
    } else {
        cell.hz *= da;
        cell.hz += db * (NB(stst_fdtd_cell, st, 0, 1)->ex - centre->ex + centre->ey -
                         NB(stst_fdtd_cell, st, 1, 0)->ey);

        if (source_distance_score <= p->source_distance_bound &&
            st->iteration <= p->cutoff_iteration) {
            float interp_factor;
            if (p->source_radius_squared != 0) {
                float cell_distance_squared =
                    source_distance_score + p->source_c * p->source_c + p->source_r * p->source_r;
                interp_factor = 1.0 - (float)(cell_distance_squared) / p->source_radius_squared;
            } else {
                interp_factor = 1.0;
            }
            float source_amplitude = (float)st->tdv_f64;
            cell.hz += interp_factor * source_amplitude;
        }
        if (st->iteration > p->detect_iteration)
            cell.hz_sum += cell.hz * cell.hz;
    }
    *(stst_fdtd_cell *)next = cell;
}

/* ---- Convection: examples/convection/convection.cpp:94-182 (pseudo-transient), :195-241 (thermal) --- */
#define CV(dr, dc) NB(stst_convection_cell, st, dr, dc)

static void convection_pt_fn(const void *params, const stencil_view *st, void *next) {
    const stst_convection_pt_params *p = (const stst_convection_pt_params *)params;
    stst_convection_cell n = *CV(0, 0);
    size_t x = st->id[0], y = st->id[1];
    size_t nx = p->nx, ny = p->ny;

    if (st->subiteration == 0) {
        if (x < nx && y < ny + 1)
            n.ErrV = CV(0, 0)->Vy;
        if (x < nx && y < ny)
            n.ErrP = CV(0, 0)->Pt;
        if (x < nx && y < ny) {
            double delta_V =
                (CV(1, 0)->Vx - CV(0, 0)->Vx) / p->dx + (CV(0, 1)->Vy - CV(0, 0)->Vy) / p->dy;
            double eta = p->eta0 * (1.0 - p->delta_eta_delta_T * (CV(0, 0)->T + p->deltaT / 2.0));
            n.Pt = CV(0, 0)->Pt - p->delta_tau_iter / p->beta * delta_V;
            n.tau_xx = 2.0 * eta * ((CV(1, 0)->Vx - CV(0, 0)->Vx) / p->dx - (1.0 / 3.0) * delta_V);
            n.tau_yy = 2.0 * eta * ((CV(0, 1)->Vy - CV(0, 0)->Vy) / p->dy - (1.0 / 3.0) * delta_V);
            if (x < nx - 1 && y < ny - 1)
                n.sigma_xy = eta * ((CV(1, 1)->Vx - CV(1, 0)->Vx) / p->dy +
                                    (CV(1, 1)->Vy - CV(0, 1)->Vy) / p->dx);
        }
    } else if (st->subiteration == 1) {
        if (x >= 1 && y >= 1) {
            if (x < (nx + 1) - 1 && y < ny - 1) {
                double Rx = 1.0 / p->rho *
                            ((CV(0, 0)->tau_xx - CV(-1, 0)->tau_xx) / p->dx +
                             (CV(-1, 0)->sigma_xy - CV(-1, -1)->sigma_xy) / p->dy -
                             (CV(0, 0)->Pt - CV(-1, 0)->Pt) / p->dx);
                n.dVxd_tau = p->dampX * CV(0, 0)->dVxd_tau + Rx * p->delta_tau_iter;
                n.Vx = CV(0, 0)->Vx + n.dVxd_tau * p->delta_tau_iter;
            }
            if (x < nx - 1 && y < (ny + 1) - 1) {
                double Ry = 1.0 / p->rho *
                            ((CV(0, 0)->tau_yy - CV(0, -1)->tau_yy) / p->dy +
                             (CV(0, -1)->sigma_xy - CV(-1, -1)->sigma_xy) / p->dx -
                             (CV(0, 0)->Pt - CV(0, -1)->Pt) / p->dy +
                             p->roh0_g_alpha * ((CV(0, -1)->T + CV(0, 0)->T) * 0.5));
                n.dVyd_tau = p->dampY * CV(0, 0)->dVyd_tau + Ry * p->delta_tau_iter;
                n.Vy = CV(0, 0)->Vy + n.dVyd_tau * p->delta_tau_iter;
            }
        }
    } else if (st->subiteration == 2) {
        if (x < nx + 1 && y < ny) {
            if (y == 0)
                n.Vx = CV(0, 1)->Vx;
            if (y == ny - 1)
                n.Vx = CV(0, -1)->Vx;
        }
        if (x < nx && y < ny + 1) {
            if (x == 0)
                n.Vy = CV(1, 0)->Vy;
            if (x == nx - 1)
                n.Vy = CV(-1, 0)->Vy;
        }
        if (x < nx && y < ny + 1)
            n.ErrV = CV(0, 0)->ErrV - n.Vy;
        if (x < nx && y < ny)
            n.ErrP = CV(0, 0)->ErrP - CV(0, 0)->Pt;
    }
    *(stst_convection_cell *)next = n;
}

static void convection_thermal_fn(const void *params, const stencil_view *st, void *next) {
    const stst_convection_thermal_params *p = (const stst_convection_thermal_params *)params;
    stst_convection_cell n = *CV(0, 0);
    size_t x = st->id[0], y = st->id[1];
    size_t nx = p->nx, ny = p->ny;

    if (st->subiteration == 0) {
        if (x > 0 && y > 0 && x < nx - 1 && y < ny - 1) {
            double qTx_top_left = -p->DcT * (CV(0, 0)->T - CV(-1, 0)->T) / p->dx;
            double qTx_top = -p->DcT * (CV(1, 0)->T - CV(0, 0)->T) / p->dx;
            double qTy_top_left = -p->DcT * (CV(0, 0)->T - CV(0, -1)->T) / p->dy;
            double qTy_left = -p->DcT * (CV(0, 1)->T - CV(0, 0)->T) / p->dy;
            double dT_dt =
                -((qTx_top - qTx_top_left) / p->dx + (qTy_left - qTy_top_left) / p->dy);
            if (CV(0, 0)->Vx > 0)
                dT_dt -= CV(0, 0)->Vx * (CV(0, 0)->T - CV(-1, 0)->T) / p->dx;
            if (CV(1, 0)->Vx < 0)
                dT_dt -= CV(1, 0)->Vx * (CV(1, 0)->T - CV(0, 0)->T) / p->dx;
            if (CV(0, 0)->Vy > 0)
                dT_dt -= CV(0, 0)->Vy * (CV(0, 0)->T - CV(0, -1)->T) / p->dy;
            if (CV(0, 1)->Vy < 0)
                dT_dt -= CV(0, 1)->Vy * (CV(0, 1)->T - CV(0, 0)->T) / p->dy;
            n.T = CV(0, 0)->T + dT_dt * p->dt;
        }
    } else if (st->subiteration == 1) {
        if (x == nx - 1 && y < ny)
            n.T = CV(-1, 0)->T;
        if (x == 0 && y < ny)
            n.T = CV(1, 0)->T;
    }
    *(stst_convection_cell *)next = n;
}

/* ---- Self-checking known-answer functor: tests/TransFuncs.hpp:55-104 (FPGATransFunc<radius>) ---------- */
static void kat_tdv(const void *params, size_t i_iteration, stencil_view *st) {
    (void)params;
    st->tdv_size = i_iteration;
}

static void kat_fn(const void *params, const stencil_view *st, void *next) {
    (void)params;
    const int n_subiterations = 2;
    stst_kat_cell n = *NB(stst_kat_cell, st, 0, 0);
    int is_valid = 1;
    for (int r = -st->radius; r <= st->radius; r++) {
        for (int c = -st->radius; c <= st->radius; c++) {
            const stst_kat_cell *old = NB(stst_kat_cell, st, r, c);
            int cell_r = (int)st->id[0] + r;
            int cell_c = (int)st->id[1] + c;
            if (cell_r >= 0 && cell_c >= 0 && (size_t)cell_r < st->grid_range[0] &&
                (size_t)cell_c < st->grid_range[1]) {
                is_valid &= old->r == cell_r;
                is_valid &= old->c == cell_c;
                is_valid &= (size_t)old->i_iteration == st->iteration;
                is_valid &= (size_t)old->i_subiteration == st->subiteration;
                is_valid &= old->status == 0; /* Normal */
            } else {
                is_valid &= old->r == 0 && old->c == 0 && old->i_iteration == 0 &&
                            old->i_subiteration == 0 && old->status == 2; /* Cell::halo() */
            }
        }
    }
    is_valid &= st->tdv_size == st->iteration;

    n.status = is_valid ? 0 : 1;
    if (n.i_subiteration == n_subiterations - 1) {
        n.i_iteration += 1;
        n.i_subiteration = 0;
    } else {
        n.i_subiteration++;
    }
    *(stst_kat_cell *)next = n;
}

static const workload_def workloads[] = {
    {"conway", 1, 1, 1, conway_fn, NULL},
    {"jacobi5", sizeof(float), 1, 1, jacobi5_fn, NULL},
    {"jacobi9", sizeof(float), 1, 1, jacobi9_fn, NULL},
    {"jacobi_r2", sizeof(float), 2, 1, jacobi_star_fn, NULL},
    {"jacobi_r3", sizeof(float), 3, 1, jacobi_star_fn, NULL},
    {"hotspot", sizeof(stst_hotspot_cell), 1, 1, hotspot_fn, NULL},
    {"fdtd", sizeof(stst_fdtd_cell), 1, 2, fdtd_fn, fdtd_tdv},
    {"convection_pt", sizeof(stst_convection_cell), 1, 3, convection_pt_fn, NULL},
    {"convection_thermal", sizeof(stst_convection_cell), 1, 2, convection_thermal_fn, NULL},
    {"kat", sizeof(stst_kat_cell), 1, 2, kat_fn, kat_tdv},
    {"kat_r2", sizeof(stst_kat_cell), 2, 2, kat_fn, kat_tdv},
};

static _Thread_local const char *g_error = "";

const char *oracle_last_error(void) { return g_error; }

const char *oracle_kind(void) { return "port"; }

/* One sweep: reference StencilStream/cpu/StencilUpdate.hpp:199-221. */
/* `rows` x `cols` cells are held; they are rows [row0, row0 + rows) x columns [col0, col0 + cols) of
 * a grid of `global_rows` x `global_cols` cells (row0 = col0 = 0 and global = held extent for a whole
 * grid). */
static void sweep(const workload_def *w, const void *params, const unsigned char *halo,
                  const unsigned char *src, unsigned char *dst, size_t rows, size_t cols,
                  size_t row0, size_t col0, size_t global_rows, size_t global_cols, size_t i_iter,
                  size_t i_subiter) {
    const int radius = w->radius;
    const size_t cb = w->cell_bytes;
    stencil_view proto;
    memset(&proto, 0, sizeof(proto));
    proto.radius = radius;
    proto.grid_range[0] = global_rows;
    proto.grid_range[1] = global_cols;
    proto.iteration = i_iter;
    proto.subiteration = i_subiter;
    if (w->tdv)
        w->tdv(params, i_iter, &proto); /* once per sweep on the host, :197 */

#pragma omp parallel for schedule(static)
    for (long long r = 0; r < (long long)rows; r++) {
        stencil_view st = proto;
        for (size_t c = 0; c < cols; c++) {
            st.id[0] = row0 + (size_t)r;
            st.id[1] = col0 + c;
            for (int rel_r = 0; rel_r < 2 * radius + 1; rel_r++) {
                for (int rel_c = 0; rel_c < 2 * radius + 1; rel_c++) {
                    /* in-grid test exactly as :205-208 (unsigned arithmetic, shifted by radius) */
                    size_t rr = (size_t)r + (size_t)rel_r, cc = c + (size_t)rel_c;
                    if (rr >= (size_t)radius && cc >= (size_t)radius &&
                        rr < rows + (size_t)radius && cc < cols + (size_t)radius) {
                        st.cell[rel_r][rel_c] =
                            src + ((rr - (size_t)radius) * cols + (cc - (size_t)radius)) * cb;
                    } else {
                        st.cell[rel_r][rel_c] = halo;
                    }
                }
            }
            w->fn(params, &st, dst + ((size_t)r * cols + c) * cb);
        }
    }
}

/*
 * Run `n_iterations` iterations of `workload` starting at iteration index `iteration_offset`.
 * cells_in / cells_out: dense row-major arrays of rows x cols cells (may not alias);
 * halo: one cell, or NULL for an all-zero cell. Returns 0 on success.
 * Loop order and ping-pong: reference StencilStream/cpu/StencilUpdate.hpp:110-129.
 */
static int run_rows(const char *workload, const void *params, const void *halo,
                    const void *cells_in, void *cells_out, size_t rows, size_t cols, size_t row0,
                    size_t col0, size_t global_rows, size_t global_cols, size_t iteration_offset,
                    size_t n_iterations) {
    const workload_def *w = NULL;
    for (size_t i = 0; i < sizeof(workloads) / sizeof(workloads[0]); i++)
        if (strcmp(workloads[i].name, workload) == 0)
            w = &workloads[i];
    if (!w) {
        g_error = "unknown workload";
        return -1;
    }
    const size_t bytes = rows * cols * w->cell_bytes;
    unsigned char zero_cell[256];
    memset(zero_cell, 0, sizeof(zero_cell));
    const unsigned char *halo_cell = halo ? (const unsigned char *)halo : zero_cell;

    if (n_iterations == 0 || bytes == 0) {
        memcpy(cells_out, cells_in, bytes);
        return 0;
    }

    unsigned char *a = (unsigned char *)malloc(bytes);
    unsigned char *b = (unsigned char *)malloc(bytes);
    if (!a || !b) {
        free(a);
        free(b);
        g_error = "out of memory";
        return -4;
    }
    const unsigned char *src = (const unsigned char *)cells_in;
    unsigned char *dst = b;
    for (size_t i_iter = 0; i_iter < n_iterations; i_iter++) {
        for (size_t i_sub = 0; i_sub < (size_t)w->n_subiterations; i_sub++) {
            sweep(w, params, halo_cell, src, dst, rows, cols, row0, col0, global_rows, global_cols,
                  iteration_offset + i_iter, i_sub);
            if (i_iter == 0 && i_sub == 0) {
                src = b;
                dst = a;
            } else {
                const unsigned char *t = src;
                src = dst;
                dst = (unsigned char *)t;
            }
        }
    }
    memcpy(cells_out, src, bytes);
    free(a);
    free(b);
    return 0;
}

int oracle_run(const char *workload, const void *params, const void *halo, const void *cells_in,
               void *cells_out, size_t rows, size_t cols, size_t iteration_offset,
               size_t n_iterations) {
    return run_rows(workload, params, halo, cells_in, cells_out, rows, cols, 0, 0, rows, cols,
                    iteration_offset, n_iterations);
}

/*
 * The same loop on a WINDOW of a larger grid: `rows` x `cols` cells that are the rows
 * [row0, row0 + rows) of a grid with `global_rows` rows. Transition functions see global
 * coordinates and the global extent; cells outside the window read as `halo`, which is wrong for
 * window-edge rows that have in-grid neighbours outside the window — after n iterations only rows
 * further than n * n_subiterations * radius from such an edge are exact. Used by the CPU stand-in
 * for a GPU slab (tests/fake_slab.py) to check the partitioner's halo-depth arithmetic.
 */
int oracle_run_window(const char *workload, const void *params, const void *halo,
                      const void *cells_in, void *cells_out, size_t rows, size_t cols, size_t row0,
                      size_t global_rows, size_t iteration_offset, size_t n_iterations) {
    if (row0 + rows > global_rows) {
        g_error = "window exceeds the grid";
        return -2;
    }
    return run_rows(workload, params, halo, cells_in, cells_out, rows, cols, row0, 0, global_rows,
                    cols, iteration_offset, n_iterations);
}

/*
 * The two-dimensional form: `cells_in` holds rows [row0, row0 + rows) x columns [col0, col0 + cols)
 * of a global_rows x global_cols grid. Same exactness rule in both directions; a crop edge that
 * coincides with the grid's border is exact (the halo really is there). Used by the full-size parity
 * tests (tests/window_oracle.py) to check windows of 16384^2 results against a crop that contains the
 * window's whole domain of dependence.
 */
int oracle_run_window2d(const char *workload, const void *params, const void *halo,
                        const void *cells_in, void *cells_out, size_t rows, size_t cols, size_t row0,
                        size_t col0, size_t global_rows, size_t global_cols,
                        size_t iteration_offset, size_t n_iterations) {
    if (row0 + rows > global_rows || col0 + cols > global_cols) {
        g_error = "window exceeds the grid";
        return -2;
    }
    return run_rows(workload, params, halo, cells_in, cells_out, rows, cols, row0, col0, global_rows,
                    global_cols, iteration_offset, n_iterations);
}
