// blueice_b200 -- analysis-space descriptors shared by the template lookup (K3), the event binning and the
// template-space likelihood kernels: bin edges / centres by value in kernel-parameter space (the library
// allocates nothing), staged into shared memory for the divergent binary searches.
#pragma once
#include <string.h>

#include "bi_common.cuh"

#define BI_MAX_EDGE_POINTS 1024

// per-dim bin centres (linear) or edges (piecewise / histogramdd), passed by value in kernel-parameter
// space (the library allocates nothing) and staged into shared memory for the divergent binary searches
struct BiPoints { double v[BI_MAX_EDGE_POINTS]; };

__device__ __forceinline__ void bi_stage_points(const BiPoints& pts, int n, double* s_pts) {
    for (int k = threadIdx.x; k < n; k += blockDim.x) s_pts[k] = pts.v[k];
    __syncthreads();
}

struct BiSpace {
    int32_t n_space;
    int32_t n_bins[BI_MAX_SPACE_DIMS];
    int32_t stride[BI_MAX_SPACE_DIMS];        // C-order strides of the flattened template
    int32_t offset[BI_MAX_SPACE_DIMS];        // offset of each dim's points in `points`
    int64_t n_cells;                          // prod(n_bins)
};

static inline int bi_fill_space(BiSpace* sp, int32_t n_space, const int32_t* n_bins_host) {
    BI_REQUIRE(n_space >= 1 && n_space <= BI_MAX_SPACE_DIMS, "n_space=%d outside [1,%d]", n_space, BI_MAX_SPACE_DIMS);
    BI_REQUIRE(n_bins_host, "n_bins_host is NULL");
    memset(sp, 0, sizeof(BiSpace));
    sp->n_space = n_space;
    int64_t cells = 1;
    for (int d = 0; d < n_space; ++d) {
        BI_REQUIRE(n_bins_host[d] >= 1, "dimension %d has %d bins", d, n_bins_host[d]);
        sp->n_bins[d] = n_bins_host[d];
        cells *= n_bins_host[d];
    }
    BI_REQUIRE(cells < (1LL << 31), "template has too many bins");
    int stride = 1;
    for (int d = n_space - 1; d >= 0; --d) { sp->stride[d] = stride; stride *= n_bins_host[d]; }
    sp->n_cells = cells;
    return BI_OK;
}

static inline int bi_fill_points(BiSpace* sp, BiPoints* pts, const double* edges_host, bool centres, int* total_out) {
    BI_REQUIRE(edges_host, "edges_host is NULL");
    int in_off = 0, out_off = 0;
    for (int d = 0; d < sp->n_space; ++d) {
        const int nb = sp->n_bins[d];
        const int n_out = centres ? nb : nb + 1;
        BI_REQUIRE(out_off + n_out <= BI_MAX_EDGE_POINTS, "too many bin edges (max %d in total)", BI_MAX_EDGE_POINTS);
        for (int k = 0; k < nb; ++k)
            BI_REQUIRE(edges_host[in_off + k + 1] > edges_host[in_off + k], "bin edges of dimension %d are not increasing", d);
        sp->offset[d] = out_off;
        for (int k = 0; k < n_out; ++k) {
            // multihist bin_centers: 0.5 * (e[1:] + e[:-1])
            pts->v[out_off + k] = centres ? 0.5 * (edges_host[in_off + k + 1] + edges_host[in_off + k])
                                          : edges_host[in_off + k];
        }
        in_off += nb + 1;
        out_off += n_out;
    }
    *total_out = out_off;
    return BI_OK;
}


#ifdef __CUDACC__
// Linear lookup (source.py:235-240): x clipped to the bin-centre range, scipy find_indices rule on the centres
// (`points`, staged per dim at sp.offset[d]).  Returns the flat index of the cell's low corner; cell[d] = -1 and
// y[d] = 0 for a one-bin dimension (index -1 aliases the only bin).
__device__ __forceinline__ int bi_event_cell_linear(const BiSpace& sp, const double* points, const double* coords,
                                                    int64_t ld_coords, int64_t i, int* cell, double* y) {
    int flat0 = 0;
    for (int d = 0; d < sp.n_space; ++d) {
        const double* c = points + sp.offset[d];
        const int n = sp.n_bins[d];
        double x = coords[(int64_t)d * ld_coords + i];
        // np.clip(x, c.min(), c.max()): NaN stays NaN (the reference then raises; the host rejects NaN first)
        if (x < c[0]) x = c[0];
        if (x > c[n - 1]) x = c[n - 1];
        if (n == 1) { cell[d] = -1; y[d] = 0.0; }
        else {
            int k = bi_upper_bound(c, n, x) - 1;
            k = k < 0 ? 0 : (k > n - 2 ? n - 2 : k);
            cell[d] = k;
            y[d] = __ddiv_rn(__dsub_rn(x, c[k]), __dsub_rn(c[k + 1], c[k]));
        }
        const int lo = cell[d] < 0 ? n - 1 : cell[d];
        flat0 += lo * sp.stride[d];
    }
    return flat0;
}

// Piecewise lookup (source.py:242-243, multihist Histdd.lookup): idx = clip(searchsorted(edges, x, 'left') - 1, 0, nbins - 1)
__device__ __forceinline__ int bi_event_bin_piecewise(const BiSpace& sp, const double* points, const double* coords,
                                                      int64_t ld_coords, int64_t i) {
    int flat = 0;
    for (int d = 0; d < sp.n_space; ++d) {
        const double* e = points + sp.offset[d];
        const int nb = sp.n_bins[d];
        const double x = coords[(int64_t)d * ld_coords + i];
        int k = bi_lower_bound(e, nb + 1, x) - 1;
        // np.searchsorted sorts NaN last: NaN -> index n_edges -> clipped to the last bin
        if (x != x) k = nb;
        k = k < 0 ? 0 : (k > nb - 1 ? nb - 1 : k);
        flat += k * sp.stride[d];
    }
    return flat;
}
#endif
