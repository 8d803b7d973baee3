// blueice_b200 -- shared device/host helpers for the sm_100a likelihood kernels.
//
// Everything that defines the *canonical arithmetic* of the hot path lives here so that the
// streaming kernel (lanes = events), the DMMA kernel (tiles of points x events) and the finalize kernel
// produce bit-identical numbers for the same parameter point.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/blueice_b200.h"

#define BI_FULL_MASK 0xffffffffu

// ---------------------------------------------------------------------------------------------
// error plumbing (thread-local message, no exceptions across the ABI)
// ---------------------------------------------------------------------------------------------
void bi_set_error(const char* fmt, ...);

#define BI_REQUIRE(cond, ...)                         \
    do {                                              \
        if (!(cond)) {                                \
            bi_set_error(__VA_ARGS__);                \
            return BI_ERR_INVALID_ARGUMENT;           \
        }                                             \
    } while (0)

#define BI_CUDA_CHECK(expr)                                                            \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            bi_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return BI_ERR_CUDA;                                                        \
        }                                                                              \
    } while (0)

#define BI_LAUNCH_CHECK() BI_CUDA_CHECK(cudaGetLastError())

// ---------------------------------------------------------------------------------------------
// Anchor grid descriptor passed to kernels by value (small, lives in the parameter bank)
// ---------------------------------------------------------------------------------------------
struct BiGrid {
    int32_t n_dims;
    int32_t n_corners;                    // 2^n_dims
    int32_t n_anchors[BI_MAX_DIMS];
    int32_t stride[BI_MAX_DIMS];          // flat anchor-index stride per dim (C order)
    int32_t axis_offset[BI_MAX_DIMS];     // offset of each axis in `axes`
    double axes[BI_MAX_AXIS_POINTS];
};

// ---------------------------------------------------------------------------------------------
// Index rules (bit-exact vs the reference's third-party numerics)
// ---------------------------------------------------------------------------------------------

// Number of elements of sorted a[0..n) that are <= x  == np.searchsorted(a, x, side='right').
// NaN compares false everywhere -> returns 0 (callers handle NaN before relying on it).
__host__ __device__ inline int bi_upper_bound(const double* a, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Number of elements of sorted a[0..n) that are < x == np.searchsorted(a, x, side='left').
__host__ __device__ inline int bi_lower_bound(const double* a, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// scipy find_indices rule (pinned in tests/test_oracle_pins.py):
//   i = clip(searchsorted(axis, z, 'right') - 1, 0, n - 2),  y = (z - a[i]) / (a[i+1] - a[i]);
//   one-point axis: i = -1 (aliases element 0 for both corners), y = 0.
__host__ __device__ inline void bi_find_cell(const double* axis, int n, double z, int* cell, double* frac) {
    if (n == 1) { *cell = -1; *frac = 0.0; return; }
    int i = bi_upper_bound(axis, n, z) - 1;
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    *cell = i;
    *frac = (z - axis[i]) / (axis[i + 1] - axis[i]);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// Canonical log-sum arithmetic
// ---------------------------------------------------------------------------------------------
#define BI_LN2_HI 6.93147180369123816490e-01 /* 0x3fe62e42fee00000, low 21 mantissa bits zero */
#define BI_LN2_LO 1.90821492927058770002e-10 /* ln2 - BI_LN2_HI */

// p is a normal, positive, finite double  <=>  0x00100000 <= hi32(p) < 0x7ff00000  (one unsigned compare)
__device__ __forceinline__ bool bi_is_normal_positive(double p) {
    return (unsigned)(__double2hiint(p) - 0x00100000) < 0x7fe00000u;
}

// split a normal positive p into mantissa in [1,2) and unbiased exponent
__device__ __forceinline__ void bi_split(double p, double* m, int* e) {
    int hi = __double2hiint(p);
    *e = (hi >> 20) - 1023;
    *m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(p));
}

// L_b = log(M * 2^E) for a block product M in [1, 2^8), exact-integer E
__device__ __forceinline__ double bi_block_log(double M, int E) {
    double e = (double)E;
    return fma(e, BI_LN2_LO, fma(e, BI_LN2_HI, log(M)));
}

// Canonical QUAD of four normal positive densities p0..p3:
//     value = fl( fl(p0*p1) * fl(p2*p3) )   evaluated as if the exponent range were unbounded,
// returned normalised as (m in [1,2), e).  Fast path: the two pair products and the quad product are
// formed directly and the quad product is split once (3 DMUL + ~6 integer ops per 4 events).  If a pair
// or the quad product leaves the normal range, the same value is formed from the four mantissas
// (scaling by powers of two is exact, so both routes give identical bits).
__device__ __forceinline__ void bi_quad_from_mantissas(double p0, double p1, double p2, double p3, double* m, int* e) {
    double m0, m1, m2, m3;
    int e0, e1, e2, e3;
    bi_split(p0, &m0, &e0); bi_split(p1, &m1, &e1); bi_split(p2, &m2, &e2); bi_split(p3, &m3, &e3);
    const double c = __dmul_rn(__dmul_rn(m0, m1), __dmul_rn(m2, m3));       // in [1, 16)
    int ec;
    bi_split(c, m, &ec);
    *e = ((e0 + e1) + (e2 + e3)) + ec;
}

// Per-event density with the reference's semantics (likelihood.py:686-689):
//   p = nansum_s(mu_s * ps_s); if (outlier != 0 && !(p > 0)) p = outlier.
// Fast path: p = fma chain; only if the result is not a normal positive finite number is the
// per-term NaN-dropping path taken (NaN/inf always poison the chain, so the fast result is exact
// whenever it is normal).  `terms` are mu_s * ps_s recomputed by the caller on the slow path.
__device__ __forceinline__ double bi_fix_density(double p_nansum, double outlier) {
    if (outlier != 0.0 && !(p_nansum > 0.0)) return outlier;
    return p_nansum;
}

// u = ((0 + src[first]) + src[first + 256]) + ...  -- one lane of the canonical total; the loads of 32 (then 8) terms are
// issued together (the chain of adds is short, the latency of dependent-looking loads is what costs)
__device__ __forceinline__ double bi_strided_sum(const double* __restrict__ src, int64_t first, int64_t n) {
    double u = 0.0;
    int64_t j = first;
    for (; j + 31 * 256 < n; j += 32 * 256) {
        double t[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) t[i] = src[j + i * 256];
#pragma unroll
        for (int i = 0; i < 32; ++i) u = __dadd_rn(u, t[i]);
    }
    for (; j + 7 * 256 < n; j += 8 * 256) {
        double t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = src[j + i * 256];
#pragma unroll
        for (int i = 0; i < 8; ++i) u = __dadd_rn(u, t[i]);
    }
    for (; j < n; j += 256) u = __dadd_rn(u, src[j]);
    return u;
}

__device__ __forceinline__ double bi_warp_sum_xor(double v) {
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) v += __shfl_xor_sync(BI_FULL_MASK, v, k);
    return v;
}
#endif  // __CUDACC__
