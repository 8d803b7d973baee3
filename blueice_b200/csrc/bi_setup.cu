// blueice_b200 -- K1: per-point anchor-grid morphing set-up (cells, corner weights, scaled rates).
//
// One thread per parameter point.  All arithmetic that the reference performs in NumPy/SciPy with
// separately rounded multiplies and adds is done here with __dmul_rn/__dadd_rn (no FMA contraction),
// so cells, fractions, weights and mus are BIT-IDENTICAL to the reference:
//   find_indices / _evaluate_linear            scipy/interpolate/_rgi.py:520-549
//   mus = itp(zs)[0]                           blueice/pdf_morphers.py:70, likelihood.py:355
//   mus[s] *= mult; mus *= lt; mus[eff] *= e   blueice/likelihood.py:366-393
//   unphysical-rate test                       blueice/likelihood.py:397-415
//   mu.sum()                                   numpy pairwise sum (n < 8 sequential; n <= 128 eight lanes)
#include "bi_common.cuh"
#include "bi_setup_point.cuh"

__global__ void __launch_bounds__(128)
k_point_setup(const __grid_constant__ BiGrid grid, int n_sources, int64_t n_points,
              const double* __restrict__ zs, const double* __restrict__ rate_mult,
              const double* __restrict__ scale, const double* __restrict__ eff,
              const double* __restrict__ mus_anchor, const __grid_constant__ BiAllowNegative allow,
              int32_t* __restrict__ cell_out, double* __restrict__ frac_out,
              int32_t* __restrict__ corner_out, double* __restrict__ weight_out,
              double* __restrict__ mus_out, double* __restrict__ musum_out,
              int32_t* __restrict__ status_out, int32_t* __restrict__ row_out, double* __restrict__ coef_out,
              double* __restrict__ wterm_out, int32_t* __restrict__ term_source_out) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0 && term_source_out)
        for (int k = 0; k < grid.n_corners * n_sources; ++k) term_source_out[k] = k % n_sources;
    if (p >= n_points) return;
    const int D = grid.n_dims, C = grid.n_corners, S = n_sources;
    const int64_t K = (int64_t)C * S;
    bi_setup_point(grid, S, zs + p * D, rate_mult + p * S, scale ? scale + p : nullptr, eff ? eff + p * S : nullptr,
                   mus_anchor, allow, cell_out + p * D, frac_out + p * D, corner_out + p * C, weight_out + p * C,
                   mus_out + p * S, musum_out + p, status_out + p, row_out ? row_out + p * K : nullptr,
                   row_out ? coef_out + p * K : nullptr, row_out ? wterm_out + p * K : nullptr);
}

// ---------------------------------------------------------------------------------------------
// The same set-up with ONE WARP per point (small batches: minimiser steps, finite-difference batches): the corners,
// the sources and the contraction terms of a point are spread over the lanes, so the latency of a single evaluation is
// a few dependent steps instead of a serial walk over C * S terms.  Every value is formed by the same operations in the
// same order as in k_point_setup (bit-identical outputs).
// ---------------------------------------------------------------------------------------------
#define BI_SETUP_WARPS 4
#ifndef BI_SETUP_WARP_MAX_POINTS
#define BI_SETUP_WARP_MAX_POINTS 8192     /* above: one thread per point (k_point_setup) */
#endif
__global__ void __launch_bounds__(BI_SETUP_WARPS * 32)
k_point_setup_warp(const __grid_constant__ BiGrid grid, int n_sources, int64_t n_points,
                   const double* __restrict__ zs, const double* __restrict__ rate_mult,
                   const double* __restrict__ scale, const double* __restrict__ eff,
                   const double* __restrict__ mus_anchor, const __grid_constant__ BiAllowNegative allow,
                   int32_t* __restrict__ cell_out, double* __restrict__ frac_out,
                   int32_t* __restrict__ corner_out, double* __restrict__ weight_out,
                   double* __restrict__ mus_out, double* __restrict__ musum_out,
                   int32_t* __restrict__ status_out, int32_t* __restrict__ row_out, double* __restrict__ coef_out,
                   double* __restrict__ wterm_out, int32_t* __restrict__ term_source_out) {
    __shared__ int s_corner_all[BI_SETUP_WARPS][1 << BI_MAX_DIMS];
    __shared__ double s_weight_all[BI_SETUP_WARPS][1 << BI_MAX_DIMS];
    __shared__ double s_mu_all[BI_SETUP_WARPS][BI_MAX_SOURCES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* s_corner = s_corner_all[warp];
    double* s_weight = s_weight_all[warp];
    double* s_mu = s_mu_all[warp];
    const int64_t p = (int64_t)blockIdx.x * BI_SETUP_WARPS + warp;
    const int D = grid.n_dims, C = grid.n_corners, S = n_sources;
    if (p == 0 && term_source_out)
        for (int k = lane; k < C * S; k += 32) term_source_out[k] = k % S;
    if (p >= n_points) return;

    // (a) cell and fraction of dimension d by lane d, then known to every lane
    int my_cell = 0;
    double my_frac = 0.0;
    bool out_of_range = false;
    if (lane < D) {
        const double* axis = grid.axes + grid.axis_offset[lane];
        const int n = grid.n_anchors[lane];
        const double z = zs[p * D + lane];
        out_of_range = !(axis[0] <= z && z <= axis[n - 1]);           // likelihood.py:345-346 (NaN fails)
        if (n == 1) { my_cell = -1; my_frac = 0.0; }
        else {
            int c = bi_upper_bound(axis, n, z) - 1;
            c = c < 0 ? 0 : (c > n - 2 ? n - 2 : c);
            my_cell = c;
            my_frac = __ddiv_rn(__dsub_rn(z, axis[c]), __dsub_rn(axis[c + 1], axis[c]));
        }
        cell_out[p * D + lane] = my_cell;
        frac_out[p * D + lane] = my_frac;
    }
    int status = __any_sync(BI_FULL_MASK, out_of_range) ? BI_POINT_OUT_OF_RANGE : BI_POINT_OK;
    int cell[BI_MAX_DIMS];
    double frac[BI_MAX_DIMS];
#pragma unroll
    for (int d = 0; d < BI_MAX_DIMS; ++d) {
        cell[d] = __shfl_sync(BI_FULL_MASK, my_cell, d);
        frac[d] = __shfl_sync(BI_FULL_MASK, my_frac, d);
    }

    // (b) corners over the lanes: first dim slowest, weight = ((1*t_0)*t_1)*...
    for (int c = lane; c < C; c += 32) {
        double w = 1.0;
        int flat = 0;
#pragma unroll
        for (int d = 0; d < BI_MAX_DIMS; ++d) {
            if (d < D) {
                const int bit = (c >> (D - 1 - d)) & 1;
                w = __dmul_rn(w, bit ? frac[d] : __dsub_rn(1.0, frac[d]));
                int idx = cell[d] + bit;
                if (idx < 0) idx += grid.n_anchors[d];
                flat += idx * grid.stride[d];
            }
        }
        corner_out[p * C + c] = flat;
        weight_out[p * C + c] = w;
        s_corner[c] = flat;
        s_weight[c] = w;
    }
    __syncwarp();

    // (c) mus over the lanes: value = 0; value = value + M[corner] * weight over the corners in order, then the scalings
    for (int s = lane; s < S; s += 32) {
        double acc;
        if (D == 0) {
            acc = mus_anchor[s];
        } else {
            acc = 0.0;
            for (int c = 0; c < C; ++c)
                acc = __dadd_rn(acc, __dmul_rn(__ldg(mus_anchor + (int64_t)s_corner[c] * S + s), s_weight[c]));
        }
        acc = __dmul_rn(acc, rate_mult[p * S + s]);
        if (scale) acc = __dmul_rn(acc, scale[p]);
        if (eff) acc = __dmul_rn(acc, eff[p * S + s]);
        s_mu[s] = acc;
        mus_out[p * S + s] = acc;
    }
    __syncwarp();

    // (d) mu.sum() in NumPy's order and the unphysical-rate test (likelihood.py:397-415) by lane 0
    if (lane == 0) {
        const double musum = bi_numpy_sum_small(s_mu, S);
        musum_out[p] = musum;
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
        bool bad = false;
        if (!allow.any) {
            for (int s = 0; s < S; ++s) bad |= !((s_mu[s] >= 0.0) && (s_mu[s] < inf));
        } else {
            bool any_finite = false;
            for (int s = 0; s < S; ++s) any_finite |= (s_mu[s] < inf);
            if (!any_finite || (musum < 0.0)) bad = true;
            for (int s = 0; s < S; ++s)
                if (!(0.0 <= s_mu[s]) && !allow.flag[s]) bad = true;
        }
        if (bad) status |= BI_POINT_UNPHYSICAL;
        status_out[p] = status;
    }

    // (e) contraction terms k = c * S + s over the lanes
    if (row_out) {
        const int K = C * S;
        for (int k = lane; k < K; k += 32) {
            const int c = k / S, s = k - c * S;
            row_out[p * K + k] = s_corner[c] * S + s;
            wterm_out[p * K + k] = s_weight[c];
            coef_out[p * K + k] = __dmul_rn(s_weight[c], s_mu[s]);
        }
    }
}

int bi_fill_grid(BiGrid* g, int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host) {
    BI_REQUIRE(n_dims >= 0 && n_dims <= BI_MAX_DIMS, "n_dims=%d outside [0,%d]", n_dims, BI_MAX_DIMS);
    BI_REQUIRE(n_dims == 0 || (n_anchors_host && axes_host), "anchor grid pointers are NULL");
    memset(g, 0, sizeof(BiGrid));
    g->n_dims = n_dims;
    g->n_corners = 1 << n_dims;
    int total = 0;
    for (int d = 0; d < n_dims; ++d) {
        BI_REQUIRE(n_anchors_host[d] >= 1, "dimension %d has %d anchors", d, n_anchors_host[d]);
        g->n_anchors[d] = n_anchors_host[d];
        g->axis_offset[d] = total;
        total += n_anchors_host[d];
    }
    BI_REQUIRE(total <= BI_MAX_AXIS_POINTS, "sum of anchors per dim = %d exceeds %d", total, BI_MAX_AXIS_POINTS);
    int stride = 1;
    for (int d = n_dims - 1; d >= 0; --d) { g->stride[d] = stride; stride *= n_anchors_host[d]; }
    for (int i = 0; i < total; ++i) g->axes[i] = axes_host[i];
    for (int d = 0; d < n_dims; ++d)
        for (int i = 1; i < g->n_anchors[d]; ++i)
            BI_REQUIRE(g->axes[g->axis_offset[d] + i] > g->axes[g->axis_offset[d] + i - 1],
                       "anchor axis %d is not strictly increasing", d);
    return BI_OK;
}

extern "C" int bi_point_setup(int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host,
                              int32_t n_sources, int64_t n_points,
                              const double* zs_dev, const double* rate_mult_dev, const double* scale_dev,
                              const double* eff_dev, const double* mus_anchor_dev,
                              const uint8_t* allow_negative_host,
                              int32_t* cell_dev, double* frac_dev, int32_t* corner_dev, double* weight_dev,
                              double* mus_dev, double* musum_dev, int32_t* status_dev,
                              int32_t* row_dev, double* coef_dev, double* wterm_dev, int32_t* term_source_dev,
                              void* stream) {
    BiGrid grid;
    int rc = bi_fill_grid(&grid, n_dims, n_anchors_host, axes_host);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(n_sources >= 1 && n_sources <= BI_MAX_SOURCES, "n_sources=%d outside [1,%d]", n_sources, BI_MAX_SOURCES);
    BI_REQUIRE(n_points >= 0, "n_points < 0");
    if (n_points == 0) return BI_OK;
    BI_REQUIRE(rate_mult_dev && mus_anchor_dev && corner_dev && weight_dev && mus_dev && musum_dev && status_dev,
               "bi_point_setup: NULL device pointer");
    BI_REQUIRE(n_dims == 0 || (zs_dev && cell_dev && frac_dev), "bi_point_setup: NULL zs/cell/frac pointer");
    BI_REQUIRE(!row_dev || (coef_dev && wterm_dev && term_source_dev),
               "bi_point_setup: row_dev needs coef_dev, wterm_dev and term_source_dev");
    BiAllowNegative allow;
    memset(&allow, 0, sizeof(allow));
    if (allow_negative_host)
        for (int s = 0; s < n_sources; ++s) { allow.flag[s] = allow_negative_host[s] ? 1 : 0; allow.any |= allow.flag[s]; }
    if (n_points <= BI_SETUP_WARP_MAX_POINTS) {                   // small batches: one warp per point (latency)
        const int64_t wblocks = (n_points + BI_SETUP_WARPS - 1) / BI_SETUP_WARPS;
        k_point_setup_warp<<<(unsigned)wblocks, BI_SETUP_WARPS * 32, 0, (cudaStream_t)stream>>>(
            grid, n_sources, n_points, zs_dev, rate_mult_dev, scale_dev, eff_dev, mus_anchor_dev, allow,
            cell_dev, frac_dev, corner_dev, weight_dev, mus_dev, musum_dev, status_dev, row_dev, coef_dev, wterm_dev,
            term_source_dev);
        BI_LAUNCH_CHECK();
        return BI_OK;
    }
    const int threads = 128;
    const int64_t blocks = (n_points + threads - 1) / threads;
    k_point_setup<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        grid, n_sources, n_points, zs_dev, rate_mult_dev, scale_dev, eff_dev, mus_anchor_dev, allow,
        cell_dev, frac_dev, corner_dev, weight_dev, mus_dev, musum_dev, status_dev, row_dev, coef_dev, wterm_dev,
        term_source_dev);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

// =============================================================================================
// Source-wise interpolation (likelihood.py:113-145,152-169,210-240,534-555): every source is morphed
// over its own sub-grid -- the shape parameters it does not ignore -- by its own RegularGridInterpolator.
// The contraction terms of a point are (source s, corner c_s of that source's sub-grid), s-major.
// =============================================================================================
struct BiSourcewise {
    uint32_t mask[BI_MAX_SOURCES];        // bit d: source uses shape parameter d
    int32_t row_base[BI_MAX_SOURCES];     // first row (sub-anchor 0) of the source in the row matrix / mus_rows
    int32_t term_base[BI_MAX_SOURCES + 1];
};

__global__ void __launch_bounds__(128)
k_point_setup_sw(const __grid_constant__ BiGrid grid, const __grid_constant__ BiSourcewise sw, int n_sources,
                 int64_t n_points, const double* __restrict__ zs, const double* __restrict__ rate_mult,
                 const double* __restrict__ scale, const double* __restrict__ eff,
                 const double* __restrict__ mus_rows, const __grid_constant__ BiAllowNegative allow,
                 int32_t* __restrict__ cell_out, double* __restrict__ frac_out, double* __restrict__ mus_out,
                 double* __restrict__ musum_out, int32_t* __restrict__ status_out, int32_t* __restrict__ row_out,
                 double* __restrict__ coef_out, double* __restrict__ wterm_out, int32_t* __restrict__ term_source_out) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int D = grid.n_dims, S = n_sources;
    const int K = sw.term_base[S];
    if (p == 0)
        for (int s = 0; s < S; ++s)
            for (int k = sw.term_base[s]; k < sw.term_base[s + 1]; ++k) term_source_out[k] = s;
    if (p >= n_points) return;
    int status = BI_POINT_OK;
    int cell[BI_MAX_DIMS];
    double frac[BI_MAX_DIMS];
    for (int d = 0; d < D; ++d) {
        const double* axis = grid.axes + grid.axis_offset[d];
        const int n = grid.n_anchors[d];
        const double z = zs[p * D + d];
        if (!(axis[0] <= z && z <= axis[n - 1])) status |= BI_POINT_OUT_OF_RANGE;     // likelihood.py:345-346
        bi_find_cell(axis, n, z, &cell[d], &frac[d]);
        cell_out[p * D + d] = cell[d];
        frac_out[p * D + d] = frac[d];
    }
    double mu_local[BI_MAX_SOURCES];
    for (int s = 0; s < S; ++s) {
        const uint32_t mask = sw.mask[s];
        int dims[BI_MAX_DIMS], stride[BI_MAX_DIMS], nd = 0;
        for (int d = 0; d < D; ++d)
            if ((mask >> d) & 1u) dims[nd++] = d;
        int st = 1;
        for (int j = nd - 1; j >= 0; --j) { stride[j] = st; st *= grid.n_anchors[dims[j]]; }   // C order over the sub-grid
        const int Cs = 1 << nd;
        double acc = 0.0;
        for (int c = 0; c < Cs; ++c) {
            double w = 1.0;
            int flat = 0;
            for (int j = 0; j < nd; ++j) {
                const int d = dims[j];
                const int bit = (c >> (nd - 1 - j)) & 1;
                const double t = bit ? frac[d] : __dsub_rn(1.0, frac[d]);
                w = __dmul_rn(w, t);
                int idx = cell[d] + bit;
                if (idx < 0) idx += grid.n_anchors[d];
                flat += idx * stride[j];
            }
            const int r = sw.row_base[s] + flat;
            const int64_t k = p * K + sw.term_base[s] + c;
            row_out[k] = r;
            wterm_out[k] = w;
            // mus: RegularGridInterpolator value = value + M * w  (a source without shape parameters keeps its base value)
            acc = nd ? __dadd_rn(acc, __dmul_rn(mus_rows[r], w)) : mus_rows[r];
        }
        acc = __dmul_rn(acc, rate_mult[p * S + s]);
        if (scale) acc = __dmul_rn(acc, scale[p]);
        if (eff) acc = __dmul_rn(acc, eff[p * S + s]);
        mu_local[s] = acc;
        mus_out[p * S + s] = acc;
        for (int c = 0; c < Cs; ++c) {
            const int64_t k = p * K + sw.term_base[s] + c;
            coef_out[k] = __dmul_rn(wterm_out[k], acc);
        }
    }
    const double musum = bi_numpy_sum_small(mu_local, S);
    musum_out[p] = musum;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    bool bad = false;
    if (!allow.any) {
        for (int s = 0; s < S; ++s) bad |= !((mu_local[s] >= 0.0) && (mu_local[s] < inf));
    } else {
        bool any_finite = false;
        for (int s = 0; s < S; ++s) any_finite |= (mu_local[s] < inf);
        if (!any_finite || (musum < 0.0)) bad = true;
        for (int s = 0; s < S; ++s)
            if (!(0.0 <= mu_local[s]) && !allow.flag[s]) bad = true;
    }
    if (bad) status |= BI_POINT_UNPHYSICAL;
    status_out[p] = status;
}

extern "C" int32_t bi_sourcewise_terms(int32_t n_sources, const uint32_t* dim_mask_host) {
    if (n_sources < 1 || n_sources > BI_MAX_SOURCES || !dim_mask_host) return -1;
    int32_t k = 0;
    for (int s = 0; s < n_sources; ++s) k += 1 << __builtin_popcount(dim_mask_host[s]);
    return k;
}

extern "C" int bi_point_setup_sourcewise(int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host,
                                         int32_t n_sources, const uint32_t* dim_mask_host,
                                         const int32_t* row_base_host, int64_t n_points,
                                         const double* zs_dev, const double* rate_mult_dev, const double* scale_dev,
                                         const double* eff_dev, const double* mus_rows_dev,
                                         const uint8_t* allow_negative_host,
                                         int32_t* cell_dev, double* frac_dev, double* mus_dev, double* musum_dev,
                                         int32_t* status_dev, int32_t* row_dev, double* coef_dev, double* wterm_dev,
                                         int32_t* term_source_dev, void* stream) {
    BiGrid grid;
    int rc = bi_fill_grid(&grid, n_dims, n_anchors_host, axes_host);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(n_sources >= 1 && n_sources <= BI_MAX_SOURCES, "n_sources=%d outside [1,%d]", n_sources, BI_MAX_SOURCES);
    BI_REQUIRE(dim_mask_host && row_base_host, "bi_point_setup_sourcewise: NULL descriptor");
    BI_REQUIRE(n_points >= 0, "n_points < 0");
    BiSourcewise sw;
    memset(&sw, 0, sizeof(sw));
    for (int s = 0; s < n_sources; ++s) {
        BI_REQUIRE((dim_mask_host[s] >> n_dims) == 0, "source %d uses a shape parameter >= n_dims", s);
        sw.mask[s] = dim_mask_host[s];
        sw.row_base[s] = row_base_host[s];
        sw.term_base[s + 1] = sw.term_base[s] + (1 << __builtin_popcount(dim_mask_host[s]));
    }
    if (n_points == 0) return BI_OK;
    BI_REQUIRE(rate_mult_dev && mus_rows_dev && mus_dev && musum_dev && status_dev && row_dev && coef_dev && wterm_dev &&
                   term_source_dev, "bi_point_setup_sourcewise: NULL device pointer");
    BI_REQUIRE(n_dims == 0 || (zs_dev && cell_dev && frac_dev), "bi_point_setup_sourcewise: NULL zs/cell/frac pointer");
    BiAllowNegative allow;
    memset(&allow, 0, sizeof(allow));
    if (allow_negative_host)
        for (int s = 0; s < n_sources; ++s) { allow.flag[s] = allow_negative_host[s] ? 1 : 0; allow.any |= allow.flag[s]; }
    const int threads = 128;
    const int64_t blocks = (n_points + threads - 1) / threads;
    k_point_setup_sw<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        grid, sw, n_sources, n_points, zs_dev, rate_mult_dev, scale_dev, eff_dev, mus_rows_dev, allow, cell_dev, frac_dev,
        mus_dev, musum_dev, status_dev, row_dev, coef_dev, wterm_dev, term_source_dev);
    BI_LAUNCH_CHECK();
    return BI_OK;
}
