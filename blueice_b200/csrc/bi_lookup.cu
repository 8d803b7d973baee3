// blueice_b200 -- K3: histogram-template lookup and event binning.
//
// bi_hist_lookup replaces HistogramPdfSource.pdf (blueice/source.py:219-246), which
// UnbinnedLogLikelihood.set_data calls once per anchor model and source
// (likelihood.py:557-560 -> model.py:97-99).  One thread per event computes the per-dimension cell
// and fraction ONCE and then walks all T templates (anchors x sources share their bin edges), so the
// event coordinates are read once and every output row is written fully coalesced.
//
//   linear   : x clipped to [first centre, last centre] (source.py:235-238), cell rule of scipy's
//              find_indices on the bin CENTRES, then
//                2-D: evaluate_linear_2d's operation order  r += V00*(1-y0)*(1-y1) ... (left-assoc.)
//                else: generic corner loop  value = value + V[corner] * ((1*t0)*t1...)
//              all with separately rounded multiplies/adds -> BIT-IDENTICAL to scipy.
//   piecewise: multihist Histdd.lookup: idx = clip(searchsorted(edges, x, 'left') - 1, 0, nbins-1).
//
// bi_histogramdd replaces Histdd.add / np.histogramdd (likelihood.py:604-609):
//   bin = searchsorted(edges, x, 'right') - 1; x == last edge -> last bin; outside or NaN -> dropped.
#include "bi_space.cuh"

// `points` holds, per dim, either the bin centres (linear: n_bins values) or the edges (n_bins + 1 values)
__global__ void __launch_bounds__(256)
k_hist_lookup_linear(const __grid_constant__ BiSpace sp, const __grid_constant__ BiPoints pts, int n_points_total,
                     const double* __restrict__ templates, int64_t n_templates,
                     const double* __restrict__ coords, int64_t ld_coords, int64_t n_events,
                     double* __restrict__ out, int64_t ld_out, int32_t* __restrict__ bin_index) {
    __shared__ double s_pts[BI_MAX_EDGE_POINTS];
    bi_stage_points(pts, n_points_total, s_pts);
    const double* points = s_pts;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_events) return;
    const int D = sp.n_space;
    int cell[BI_MAX_SPACE_DIMS];
    double y[BI_MAX_SPACE_DIMS];
    const int flat0 = bi_event_cell_linear(sp, points, coords, ld_coords, i, cell, y);
    if (bin_index) bin_index[i] = flat0;

    if (D == 2) {
        // scipy evaluate_linear_2d (Cython fast path used for 2-D float64 values)
        const int i0 = cell[0] < 0 ? sp.n_bins[0] - 1 : cell[0];
        const int i1 = cell[1] < 0 ? sp.n_bins[1] - 1 : cell[1];
        const int i0p = cell[0] < 0 ? 0 : cell[0] + 1;
        const int i1p = cell[1] < 0 ? 0 : cell[1] + 1;
        const int s0 = sp.stride[0];
        const double y0 = y[0], y1 = y[1];
        const double u0 = __dsub_rn(1.0, y0), u1 = __dsub_rn(1.0, y1);
        for (int64_t t = 0; t < n_templates; ++t) {
            const double* V = templates + t * sp.n_cells;
            double r = 0.0;
            r = __dadd_rn(r, __dmul_rn(__dmul_rn(V[i0 * s0 + i1], u0), u1));
            r = __dadd_rn(r, __dmul_rn(__dmul_rn(V[i0 * s0 + i1p], u0), y1));
            r = __dadd_rn(r, __dmul_rn(__dmul_rn(V[i0p * s0 + i1], y0), u1));
            r = __dadd_rn(r, __dmul_rn(__dmul_rn(V[i0p * s0 + i1p], y0), y1));
            out[t * ld_out + i] = r;
        }
        return;
    }
    // generic corner loop (_rgi.py:520-549): first dim slowest, weight = ((1*t0)*t1)..., value += V*weight
    const int C = 1 << D;
    for (int64_t t = 0; t < n_templates; ++t) {
        const double* V = templates + t * sp.n_cells;
        double acc = 0.0;
        for (int c = 0; c < C; ++c) {
            double w = 1.0;
            int flat = 0;
            for (int d = 0; d < D; ++d) {
                const int bit = (c >> (D - 1 - d)) & 1;
                w = __dmul_rn(w, bit ? y[d] : __dsub_rn(1.0, y[d]));
                int idx = cell[d] + bit;
                if (idx < 0) idx += sp.n_bins[d];
                flat += idx * sp.stride[d];
            }
            acc = __dadd_rn(acc, __dmul_rn(V[flat], w));
        }
        out[t * ld_out + i] = acc;
    }
}

__global__ void __launch_bounds__(256)
k_hist_lookup_piecewise(const __grid_constant__ BiSpace sp, const __grid_constant__ BiPoints pts, int n_points_total,
                        const double* __restrict__ templates, int64_t n_templates,
                        const double* __restrict__ coords, int64_t ld_coords, int64_t n_events,
                        double* __restrict__ out, int64_t ld_out, int32_t* __restrict__ bin_index) {
    __shared__ double s_pts[BI_MAX_EDGE_POINTS];
    bi_stage_points(pts, n_points_total, s_pts);
    const double* points = s_pts;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_events) return;
    const int flat = bi_event_bin_piecewise(sp, points, coords, ld_coords, i);
    if (bin_index) bin_index[i] = flat;
    for (int64_t t = 0; t < n_templates; ++t) out[t * ld_out + i] = templates[t * sp.n_cells + flat];
}

__global__ void __launch_bounds__(256)
k_histogramdd(const __grid_constant__ BiSpace sp, const __grid_constant__ BiPoints pts, int n_points_total,
              const double* __restrict__ coords, int64_t ld_coords, int64_t n_events,
              unsigned long long* __restrict__ counts, int32_t* __restrict__ bin_index) {
    __shared__ double s_pts[BI_MAX_EDGE_POINTS];
    bi_stage_points(pts, n_points_total, s_pts);
    const double* edges = s_pts;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_events) return;
    int flat = 0;
    bool keep = true;
    for (int d = 0; d < sp.n_space; ++d) {
        const double* e = edges + sp.offset[d];
        const int nb = sp.n_bins[d];
        const double x = coords[(int64_t)d * ld_coords + i];
        int k = bi_upper_bound(e, nb + 1, x) - 1;
        if (x == e[nb]) k = nb - 1;                           // right-most edge is inclusive
        if (!(x >= e[0] && x <= e[nb])) keep = false;         // outside or NaN
        flat += (keep ? k : 0) * sp.stride[d];
    }
    if (bin_index) bin_index[i] = keep ? flat : -1;
    if (keep) atomicAdd(&counts[flat], 1ULL);
}

// the same binning for MANY datasets back to back (toys): event e of dataset t (binary search in the offsets)
// increments counts[t, bin]
__global__ void __launch_bounds__(256)
k_histogramdd_toys(const __grid_constant__ BiSpace sp, const __grid_constant__ BiPoints pts, int n_points_total,
                   const double* __restrict__ coords, int64_t ld_coords, int64_t n_events,
                   const int64_t* __restrict__ offsets, int64_t n_datasets,
                   unsigned long long* __restrict__ counts, int64_t ld_counts) {
    __shared__ double s_pts[BI_MAX_EDGE_POINTS];
    bi_stage_points(pts, n_points_total, s_pts);
    const double* edges = s_pts;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_events) return;
    int flat = 0;
    bool keep = true;
    for (int d = 0; d < sp.n_space; ++d) {
        const double* e = edges + sp.offset[d];
        const int nb = sp.n_bins[d];
        const double x = coords[(int64_t)d * ld_coords + i];
        int k = bi_upper_bound(e, nb + 1, x) - 1;
        if (x == e[nb]) k = nb - 1;                           // right-most edge is inclusive
        if (!(x >= e[0] && x <= e[nb])) keep = false;         // outside or NaN
        flat += (keep ? k : 0) * sp.stride[d];
    }
    if (!keep) return;
    int64_t lo = 0, hi = n_datasets;                          // largest t with offsets[t] <= i
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= i) lo = mid; else hi = mid;
    }
    atomicAdd(&counts[lo * ld_counts + flat], 1ULL);
}

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
extern "C" int bi_hist_lookup(const double* templates_dev, int64_t n_templates, int32_t n_space,
                              const int32_t* n_bins_host, const double* edges_host,
                              const double* coords_dev, int64_t ld_coords, int64_t n_events, int32_t method,
                              double* out_dev, int64_t ld_out, int32_t* bin_index_dev, void* stream) {
    BiSpace sp;
    int rc = bi_fill_space(&sp, n_space, n_bins_host);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(method == BI_LOOKUP_LINEAR || method == BI_LOOKUP_PIECEWISE, "unknown lookup method %d", method);
    BI_REQUIRE(n_templates >= 0 && n_events >= 0, "negative size");
    BiPoints pts;
    int total = 0;
    rc = bi_fill_points(&sp, &pts, edges_host, method == BI_LOOKUP_LINEAR, &total);
    if (rc != BI_OK) return rc;
    if (n_events == 0 || n_templates == 0) return BI_OK;
    BI_REQUIRE(templates_dev && coords_dev && out_dev, "bi_hist_lookup: NULL device pointer");
    BI_REQUIRE(ld_coords >= n_events && ld_out >= n_events, "leading dimensions smaller than n_events");
    const int64_t blocks = (n_events + 255) / 256;
    cudaStream_t st = (cudaStream_t)stream;
    if (method == BI_LOOKUP_LINEAR)
        k_hist_lookup_linear<<<(unsigned)blocks, 256, 0, st>>>(sp, pts, total, templates_dev, n_templates, coords_dev,
                                                               ld_coords, n_events, out_dev, ld_out, bin_index_dev);
    else
        k_hist_lookup_piecewise<<<(unsigned)blocks, 256, 0, st>>>(sp, pts, total, templates_dev, n_templates, coords_dev,
                                                                  ld_coords, n_events, out_dev, ld_out, bin_index_dev);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

extern "C" int bi_histogramdd(int32_t n_space, const int32_t* n_bins_host, const double* edges_host,
                              const double* coords_dev, int64_t ld_coords, int64_t n_events,
                              unsigned long long* counts_dev, int32_t* bin_index_dev, void* stream) {
    BiSpace sp;
    int rc = bi_fill_space(&sp, n_space, n_bins_host);
    if (rc != BI_OK) return rc;
    BiPoints pts;
    int total = 0;
    rc = bi_fill_points(&sp, &pts, edges_host, false, &total);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(n_events >= 0, "n_events < 0");
    if (n_events == 0) return BI_OK;
    BI_REQUIRE(coords_dev && counts_dev && ld_coords >= n_events, "bi_histogramdd: bad arguments");
    const int64_t blocks = (n_events + 255) / 256;
    k_histogramdd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(sp, pts, total, coords_dev, ld_coords, n_events,
                                                                      counts_dev, bin_index_dev);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

extern "C" int bi_histogramdd_toys(int32_t n_space, const int32_t* n_bins_host, const double* edges_host,
                                   const double* coords_dev, int64_t ld_coords, int64_t n_events,
                                   const int64_t* dataset_offset_dev, int64_t n_datasets,
                                   unsigned long long* counts_dev, int64_t ld_counts, void* stream) {
    BiSpace sp;
    int rc = bi_fill_space(&sp, n_space, n_bins_host);
    if (rc != BI_OK) return rc;
    BiPoints pts;
    int total = 0;
    rc = bi_fill_points(&sp, &pts, edges_host, false, &total);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(n_events >= 0 && n_datasets >= 0, "negative size");
    if (n_events == 0 || n_datasets == 0) return BI_OK;
    BI_REQUIRE(coords_dev && counts_dev && dataset_offset_dev && ld_coords >= n_events && ld_counts >= sp.n_cells,
               "bi_histogramdd_toys: bad arguments");
    const int64_t blocks = (n_events + 255) / 256;
    k_histogramdd_toys<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(sp, pts, total, coords_dev, ld_coords, n_events,
                                                                           dataset_offset_dev, n_datasets, counts_dev, ld_counts);
    BI_LAUNCH_CHECK();
    return BI_OK;
}
