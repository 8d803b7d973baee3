// blueice_b200 -- device helpers shared by the template-space kernels (bi_template.cu: K5 / K5b, bi_template_bm.cu: the
// bin-major toy sweep): template lookups in scipy's operation order (source.py:219-246, as K3), the reference-semantics
// density of the rare path, pair-group descriptors.
#pragma once
#include "bi_space.cuh"

#ifndef BI_TS_BATCH
#define BI_TS_BATCH 4       /* template rows whose gathers are in flight together (K5) */
#endif
#define BI_TS_THREADS 128
#define BI_TS_WARPS (BI_TS_THREADS / 32)
#define BI_RANGE_LO ((1023 - 126) << 20)
#define BI_RANGE_SPAN (253u << 20)

struct BiTsSpace {
    int32_t n_space;
    int32_t n_corner;                              // 2^n_space lookup corners (linear), 1 (piecewise)
    int64_t corner_off[1 << BI_MAX_SPACE_DIMS];    // element offset of lookup corner c from the low corner, x bin_stride
};

// ---------------------------------------------------------------------------------------------
// template value of one row at one prepared event, scipy's operation order (== k_hist_lookup_linear)
// ---------------------------------------------------------------------------------------------
// Template values of one row at one prepared event.  Linear lookups read a PACKED layout in which element (row, bin)
// holds the bin together with its neighbours along the last (and second-last) dimension, so that the lookup corners come
// with ONE wide load -- K5 is bound by the number of scattered L2 requests, not by bytes:
//   1-D:    [row][bin][2] = (T[b], T[b + 1])                                   one 128-bit load
//   >= 2-D: [row][bin][4] = (T[b], T[b + 1], T[b + s], T[b + s + 1])            one 256-bit load (LDG.E.256) per
//           s = stride of the second-last dimension                             four lookup corners
template <int NS>
__device__ __forceinline__ void bi_ts_gather(const double* __restrict__ V, const BiTsSpace& sp, double (&v)[1 << NS]) {
    if constexpr (NS == 0) {
        v[0] = __ldg(V);                                            // piecewise: the bin's value (plain layout)
    } else if constexpr (NS == 1) {
        const double2 t = __ldg(reinterpret_cast<const double2*>(V));
        v[0] = t.x;
        v[1] = t.y;
    } else {
#pragma unroll
        for (int c = 0; c < (1 << NS); c += 4) {
            const double* q = V + (c ? sp.corner_off[c] : 0);
            asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
                : "=d"(v[c]), "=d"(v[c + 1]), "=d"(v[c + 2]), "=d"(v[c + 3]) : "l"(q));
        }
    }
}
template <int NS>
__device__ __forceinline__ double bi_ts_eval(const double (&v)[1 << NS], const double (&y)[NS > 0 ? NS : 1]) {
    if constexpr (NS == 0) {
        return v[0];
    } else if constexpr (NS == 2) {
        const double u0 = __dsub_rn(1.0, y[0]), u1 = __dsub_rn(1.0, y[1]);
        double r = 0.0;
        r = __dadd_rn(r, __dmul_rn(__dmul_rn(v[0], u0), u1));
        r = __dadd_rn(r, __dmul_rn(__dmul_rn(v[1], u0), y[1]));
        r = __dadd_rn(r, __dmul_rn(__dmul_rn(v[2], y[0]), u1));
        r = __dadd_rn(r, __dmul_rn(__dmul_rn(v[3], y[0]), y[1]));
        return r;
    } else {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < (1 << NS); ++c) {
            double w = 1.0;
#pragma unroll
            for (int d = 0; d < NS; ++d) {
                const int bit = (c >> (NS - 1 - d)) & 1;
                w = __dmul_rn(w, bit ? y[d] : __dsub_rn(1.0, y[d]));
            }
            acc = __dadd_rn(acc, __dmul_rn(v[c], w));
        }
        return acc;
    }
}

template <int NS>
__device__ __forceinline__ double bi_ts_lookup(const double* __restrict__ V, const BiTsSpace& sp, const double (&y)[NS > 0 ? NS : 1]) {
    double v[1 << NS];
    bi_ts_gather<NS>(V, sp, v);
    return bi_ts_eval<NS>(v, y);
}

// density of one event with the reference's semantics (likelihood.py:686-689), rare path
template <int NS>
static __device__ __noinline__ double bi_ts_slow_density(const double* __restrict__ T, const int64_t* rowoff, int64_t base,
                                                         const BiTsSpace& sp, const double (&y)[NS > 0 ? NS : 1], int K, int S,
                                                         const int32_t* __restrict__ term_source,
                                                         const double* __restrict__ wterm,
                                                         const double* __restrict__ mu, double outlier) {
    double acc = 0.0;
    for (int s = 0; s < S; ++s) {
        double ps = 0.0;
        for (int k = 0; k < K; ++k)
            if (term_source[k] == s) ps = fma(bi_ts_lookup<NS>(T + rowoff[k] + base, sp, y), wterm[k], ps);
        const double term = __dmul_rn(mu[s], ps);
        if (term == term) acc = __dadd_rn(acc, term);               // nansum: NaN terms count as 0
    }
    return bi_fix_density(acc, outlier);
}

// pair group descriptor: pairs [first, first + count) of the pair list, all on dataset `dataset`
struct BiTsGroup { int32_t first, count, dataset, pad; };

// bi_template.cu: K5's launcher (pre_dev != NULL: densities already formed, see k_template_partials)
int bi_template_partials_impl(const double* templates_dev, int64_t row_stride, int64_t bin_stride,
                              int32_t n_space, const int32_t* n_bins_host, int32_t method,
                              const int32_t* ev_bin_dev, const double* ev_frac_dev, int64_t ld_frac,
                              const int64_t* dataset_offset_dev, int32_t n_terms, int32_t n_sources,
                              const int32_t* row_dev, const double* coef_dev, const double* wterm_dev,
                              const int32_t* term_source_dev, const double* mus_dev, const int32_t* status_dev,
                              int64_t n_groups, int32_t group_points, const int32_t* groups_dev,
                              const int64_t* unit_offset_dev, const int32_t* unit_group_dev, int64_t n_units,
                              const int32_t* pair_point_dev, const int64_t* pair_partial_offset_dev,
                              double outlier_likelihood, double* partial_dev,
                              const int32_t* group_order_dev, const int32_t* n_ordered_dev, int32_t sb_max,
                              const double* pre_dev, const int32_t* pre_corner_dev, const double* pre_weight_dev,
                              void* stream);
