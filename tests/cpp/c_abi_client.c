/* A plain-C client of the two C-ABI headers: proves that include/stst_rt.h and
 * include/stst_workloads.h are valid C11 (not only C++), that every call used here links against the
 * shared libraries, and — on a machine without a CUDA device — that the product path reports an error
 * instead of computing anything. With a device it runs HotSpot 64 x 64 for 8 iterations and prints
 * the centre cell. Built and run by tests/test_abi.py. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <stst_rt.h>
#include <stst_workloads.h>

int main(void) {
    if (stst_workloads_abi_version() != STST_WORKLOADS_ABI_VERSION) {
        printf("ABI version mismatch\n");
        return 1;
    }
    stst_workload_info info;
    if (stst_workload_get_info("hotspot", &info) != STST_OK || info.cell_bytes != sizeof(stst_hotspot_cell)) {
        printf("registry mismatch\n");
        return 1;
    }
    if (stst_workload_get_info("no-such-workload", &info) != STST_ERR_UNKNOWN_WORKLOAD) {
        printf("unknown workload not reported\n");
        return 1;
    }

    int devices = 0;
    const int have_device = stst_device_count(&devices) == 0 && devices > 0;

    const size_t rows = 64, cols = 64;
    stst_grid *grid = NULL;
    int status = stst_grid_create("hotspot", rows, cols, -1, &grid);
    if (!have_device) {
        /* no CPU fallback: creation (or, at the latest, the first transfer) must fail loudly */
        if (status == STST_OK) {
            stst_hotspot_cell cell = {30.0f, 0.5f};
            status = stst_grid_copy_from_host(grid, &cell, sizeof(cell)); /* wrong size AND no device */
        }
        printf("no device: status %d (%s)\n", status, stst_workloads_last_error());
        return status == STST_OK ? 1 : 0;
    }
    if (status != STST_OK) {
        printf("grid_create failed: %s\n", stst_workloads_last_error());
        return 1;
    }

    stst_hotspot_cell *cells = malloc(rows * cols * sizeof(*cells));
    for (size_t i = 0; i < rows * cols; i++) {
        cells[i].temp = 30.0f + (float)(i % 7);
        cells[i].power = 0.25f;
    }
    if (stst_grid_copy_from_host(grid, cells, rows * cols * sizeof(*cells)) != STST_OK ||
        stst_grid_copy_from_host(grid, cells, 8) != STST_ERR_RANGE) {
        printf("copy_from_host: %s\n", stst_workloads_last_error());
        return 1;
    }
    stst_hotspot_params tf = {1.0f, 1.0f, 0.01f, 0.001f};
    stst_hotspot_cell halo = {0.0f, 0.0f};
    stst_update_params params;
    memset(&params, 0, sizeof(params));
    params.transition_function = &tf;
    params.transition_function_bytes = sizeof(tf);
    params.halo_value = &halo;
    params.halo_value_bytes = sizeof(halo);
    params.n_iterations = 8;
    params.blocking = 1;
    params.cuda_device = -1;
    stst_update *update = NULL;
    stst_grid *result = NULL;
    if (stst_update_create("hotspot", &params, &update) != STST_OK ||
        stst_update_apply(update, grid, &result) != STST_OK ||
        stst_grid_copy_to_host(result, cells, rows * cols * sizeof(*cells)) != STST_OK) {
        printf("update failed: %s\n", stst_workloads_last_error());
        return 1;
    }
    stst_field_extent extent = {0, rows, cols};
    double max_temp = 0.0;
    if (stst_grid_max_abs(result, &extent, 1, &max_temp) != STST_OK) {
        printf("max_abs failed: %s\n", stst_workloads_last_error());
        return 1;
    }
    stst_update_stats stats;
    stst_update_get_stats(update, &stats);
    printf("centre temp %.6f, max |temp| %.6f, %zu launches, pass-through planes %#x\n",
           cells[(rows / 2) * cols + cols / 2].temp, max_temp, stats.n_launches, stats.passthrough_planes);
    stst_update_destroy(update);
    stst_grid_destroy(result);
    stst_grid_destroy(grid);
    free(cells);
    return 0;
}
