// blueice_b200 -- device-side schedule of the DMMA K2 kernel (no host round trip per batch).
//
// One CTA buckets the evaluable points (status == 0) of a batch by hypercube cell (counting sort on the
// flat cell index), cuts every cell's points into point groups of at most `unit_points` points (whole
// 8-point m-tiles, sizes balanced inside a cell) and picks the superblock range length, writing
//     group_points [P]           point indices, cell-major
//     groups       [n_groups][4] (first, count, range counter = 0, 0); first / count index group_points
//     header       [8]           n_groups, n_ranges, sb_per_range, n_units, fetch counter (= 0), 0, 0, 0
// The persistent K2 kernel hands out (group, superblock range) pairs: warps stick to a group and fetch its ranges
// from the group's counter.
// Results never depend on the schedule, so the order of points inside a cell
// (decided by atomics) is free.
#include "bi_common.cuh"

#include "bi_plan.cuh"

__global__ void __launch_bounds__(BI_PLAN_THREADS)
k_plan_units(const __grid_constant__ BiPlanDims dims, int n_cells, int64_t n_points,
             const int32_t* __restrict__ cell, const int32_t* __restrict__ status,
             int unit_points, int64_t n_super, int target_units, int full_units,
             int32_t* __restrict__ group_points, int32_t* __restrict__ groups, int32_t* __restrict__ header) {
    extern __shared__ int bi_plan_smem[];
    int* poff = bi_plan_smem;                     // [n_cells + 1] counts -> point offsets
    int* cursor = poff + n_cells + 1;             // [n_cells]     scatter cursors
    int* goff = cursor + n_cells;                 // [n_cells + 1] groups per cell -> group offsets
    __shared__ int carry[33];
    const int tid = threadIdx.x;
    const int D = dims.n_dims;
    const int cd = D > 0 ? D : 1;                 // K1 writes cell[P, max(D, 1)]

    for (int c = tid; c <= n_cells; c += BI_PLAN_THREADS) { poff[c] = 0; goff[c] = 0; if (c < n_cells) cursor[c] = 0; }
    __syncthreads();
    for (int64_t p = tid; p < n_points; p += BI_PLAN_THREADS)
        if (status[p] == 0) atomicAdd(&poff[D ? bi_flat_cell(dims, cell + p * cd) : 0], 1);
    __syncthreads();
    const int tiles_per_unit = unit_points / 8;
    for (int c = tid; c < n_cells; c += BI_PLAN_THREADS) {
        const int n_mt = (poff[c] + 7) >> 3;
        goff[c] = (n_mt + tiles_per_unit - 1) / tiles_per_unit;
    }
    __syncthreads();
    const int n_evaluable = bi_block_exclusive_scan(poff, n_cells, carry);
    if (tid == 0) poff[n_cells] = n_evaluable;
    const int n_groups = bi_block_exclusive_scan(goff, n_cells, carry);
    if (tid == 0) goff[n_cells] = n_groups;
    __syncthreads();
    for (int64_t p = tid; p < n_points; p += BI_PLAN_THREADS) {
        if (status[p] == 0) {
            const int c = D ? bi_flat_cell(dims, cell + p * cd) : 0;
            group_points[poff[c] + atomicAdd(&cursor[c], 1)] = (int32_t)p;
        }
    }
    for (int c = tid; c < n_cells; c += BI_PLAN_THREADS) {
        const int n = poff[c + 1] - poff[c];
        const int u = goff[c + 1] - goff[c];
        if (u > 0) {
            const int n_mt = (n + 7) >> 3;
            const int base = n_mt / u, rem = n_mt - base * u;
            int pos = poff[c];
            const int end = poff[c + 1];
            for (int i = 0; i < u; ++i) {
                int cnt = full_units ? unit_points : 8 * (base + (i < rem ? 1 : 0));
                if (cnt > end - pos) cnt = end - pos;
                groups[4 * (goff[c] + i)] = pos;
                groups[4 * (goff[c] + i) + 1] = cnt;
                groups[4 * (goff[c] + i) + 2] = 0;                 // range counter of the persistent K2 kernel
                groups[4 * (goff[c] + i) + 3] = 0;
                pos += cnt;
            }
        }
    }
    if (tid == 0) {
        // range length: as short as the unit budget allows -- warps stick to a group and stream its ranges
        // back to back, so a short range costs nothing and evens out the tail
        int64_t sb_per = 1, n_ranges = 0, n_units = 0;
        if (n_groups > 0 && n_super > 0) {
            sb_per = (n_super * (int64_t)n_groups + target_units - 1) / target_units;
            if (sb_per < 1) sb_per = 1;
            if (sb_per > n_super) sb_per = n_super;
            n_ranges = (n_super + sb_per - 1) / sb_per;
            n_units = n_ranges * n_groups;
            if (n_units > 0x7fffffff) {                       // keep the unit index in int32
                sb_per = (n_super * (int64_t)n_groups + 0x3fffffff) / 0x40000000;
                n_ranges = (n_super + sb_per - 1) / sb_per;
                n_units = n_ranges * n_groups;
            }
        }
        header[0] = n_groups;
        header[1] = (int32_t)n_ranges;
        header[2] = (int32_t)sb_per;
        header[3] = (int32_t)n_units;
        header[4] = 0;                                         // fetch counter of the persistent kernel
        header[5] = n_evaluable;
        header[6] = 0;
        header[7] = 0;
    }
}

int bi_fill_grid(BiGrid* g, int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host);

extern "C" int64_t bi_plan_max_cells(void) { return BI_PLAN_MAX_CELLS; }

extern "C" int bi_unbinned_plan(int32_t n_dims, const int32_t* n_anchors_host, int64_t n_points,
                                const int32_t* cell_dev, const int32_t* status_dev, int32_t unit_points,
                                int64_t n_events, int32_t target_units, int32_t full_units,
                                int32_t* group_points_dev, int32_t* groups_dev, int32_t* header_dev, void* stream) {
    BI_REQUIRE(n_dims >= 0 && n_dims <= BI_MAX_DIMS, "n_dims=%d outside [0,%d]", n_dims, BI_MAX_DIMS);
    BI_REQUIRE(n_points >= 0 && n_points < 0x7fffffff, "n_points outside [0, 2^31)");
    BI_REQUIRE(unit_points >= 8 && unit_points % 8 == 0, "unit_points must be a positive multiple of 8");
    BI_REQUIRE(target_units >= 1, "target_units < 1");
    BI_REQUIRE(status_dev && group_points_dev && groups_dev && header_dev && (n_dims == 0 || (cell_dev && n_anchors_host)),
               "bi_unbinned_plan: NULL pointer");
    BiPlanDims dims;
    memset(&dims, 0, sizeof(dims));
    dims.n_dims = n_dims;
    int64_t n_cells = 1;
    for (int d = n_dims - 1; d >= 0; --d) {
        BI_REQUIRE(n_anchors_host[d] >= 1, "dimension %d has %d anchors", d, n_anchors_host[d]);
        dims.cells[d] = n_anchors_host[d] > 1 ? n_anchors_host[d] - 1 : 1;
        dims.stride[d] = (int32_t)n_cells;
        n_cells *= dims.cells[d];
        BI_REQUIRE(n_cells <= BI_PLAN_MAX_CELLS, "anchor grid has more than %d hypercube cells", BI_PLAN_MAX_CELLS);
    }
    const size_t smem = (size_t)(3 * n_cells + 2) * sizeof(int);
    static bool configured = false;
    if (!configured) {
        BI_CUDA_CHECK(cudaFuncSetAttribute(k_plan_units, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)((3 * BI_PLAN_MAX_CELLS + 2) * sizeof(int))));
        configured = true;
    }
    k_plan_units<<<1, BI_PLAN_THREADS, smem, (cudaStream_t)stream>>>(
        dims, (int)n_cells, n_points, cell_dev, status_dev, unit_points, bi_num_superblocks(n_events), target_units,
        full_units,
        group_points_dev, groups_dev, header_dev);
    BI_LAUNCH_CHECK();
    return BI_OK;
}
