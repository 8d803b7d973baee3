// blueice_b200 -- K1 for ONE parameter point as a device function (shared by k_point_setup and the single-launch path
// for tiny batches, bi_small.cu).  All pointers are already offset to the point: zs [D], rate_mult [S], scale [1] or
// NULL, eff [S] or NULL; outputs cell / frac [D], corner / weight [C], mus [S], *musum, *status and (optional, all or
// none) row / coef / wterm [C * S].  Arithmetic: see bi_setup.cu.
#pragma once
#include "bi_common.cuh"

struct BiAllowNegative { uint8_t flag[BI_MAX_SOURCES]; int32_t any; };

__device__ __forceinline__ double bi_numpy_sum_small(const double* a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r = __dadd_rn(r, a[i]);
        return r;
    }
    double r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], a[i + k]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

__device__ __forceinline__ void bi_setup_point(const BiGrid& grid, int n_sources,
                                               const double* __restrict__ zs, const double* __restrict__ rate_mult,
                                               const double* __restrict__ scale, const double* __restrict__ eff,
                                               const double* __restrict__ mus_anchor, const BiAllowNegative& allow,
                                               int32_t* __restrict__ cell_out, double* __restrict__ frac_out,
                                               int32_t* __restrict__ corner_out, double* __restrict__ weight_out,
                                               double* __restrict__ mus_out, double* __restrict__ musum_out,
                                               int32_t* __restrict__ status_out, int32_t* __restrict__ row_out,
                                               double* __restrict__ coef_out, double* __restrict__ wterm_out) {
    const int D = grid.n_dims, C = grid.n_corners, S = n_sources;
    int status = BI_POINT_OK;

    int cell[BI_MAX_DIMS];
    double frac[BI_MAX_DIMS];
    for (int d = 0; d < D; ++d) {
        const double* axis = grid.axes + grid.axis_offset[d];
        const int n = grid.n_anchors[d];
        const double z = zs[d];
        // likelihood.py:345-346: `if not minbound <= z <= maxbound: return -inf` (NaN fails)
        if (!(axis[0] <= z && z <= axis[n - 1])) status |= BI_POINT_OUT_OF_RANGE;
        int c; double y;
        if (n == 1) { c = -1; y = 0.0; }
        else {
            c = bi_upper_bound(axis, n, z) - 1;
            c = c < 0 ? 0 : (c > n - 2 ? n - 2 : c);
            y = __ddiv_rn(__dsub_rn(z, axis[c]), __dsub_rn(axis[c + 1], axis[c]));
        }
        cell[d] = c; frac[d] = y;
        cell_out[d] = c;
        frac_out[d] = y;
    }

    // corners, first dim slowest (itertools.product order), weight = ((1*t_0)*t_1)*...
    // (kept in thread-local arrays too: reading the just-written global rows back costs a round trip per term)
    int corner_l[1 << BI_MAX_DIMS];
    double weight_l[1 << BI_MAX_DIMS];
    for (int c = 0; c < C; ++c) {
        double w = 1.0;
        int flat = 0;
        for (int d = 0; d < D; ++d) {
            const int bit = (c >> (D - 1 - d)) & 1;
            const double t = bit ? frac[d] : __dsub_rn(1.0, frac[d]);
            w = __dmul_rn(w, t);
            int idx = cell[d] + bit;
            if (idx < 0) idx += grid.n_anchors[d];   // one-point axis: index -1 aliases the last (= only) anchor
            flat += idx * grid.stride[d];
        }
        corner_out[c] = flat;
        weight_out[c] = w;
        corner_l[c] = flat;
        weight_l[c] = w;
    }

    // mus: value = 0; value = value + M[corner] * weight   (then the three in-place scalings)
    double mu_local[BI_MAX_SOURCES];
    for (int s = 0; s < S; ++s) {
        double acc;
        if (D == 0) {
            acc = mus_anchor[s];
        } else {
            acc = 0.0;
            for (int c = 0; c < C; ++c)
                acc = __dadd_rn(acc, __dmul_rn(__ldg(mus_anchor + (int64_t)corner_l[c] * S + s), weight_l[c]));
        }
        acc = __dmul_rn(acc, rate_mult[s]);
        if (scale) acc = __dmul_rn(acc, scale[0]);
        if (eff) acc = __dmul_rn(acc, eff[s]);
        mu_local[s] = acc;
        mus_out[s] = acc;
    }
    const double musum = bi_numpy_sum_small(mu_local, S);
    *musum_out = musum;

    // contraction terms of K2 (DMMA form): k = c * S + s -> row of the [G * S, ld] anchor tensor, weight, coefficient
    if (row_out) {
        for (int c = 0; c < C; ++c) {
            const int flat = corner_l[c];
            const double w = weight_l[c];
            for (int s = 0; s < S; ++s) {
                row_out[c * S + s] = flat * S + s;
                wterm_out[c * S + s] = w;
                coef_out[c * S + s] = __dmul_rn(w, mu_local[s]);
            }
        }
    }

    // likelihood.py:397-415
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
    *status_out = status;
}
