// blueice_b200 -- K4: binned Poisson log-likelihood with the Beeston-Barlow single-source adjustment.
//
// Replaces, for a batch of P parameter points over B bins:
//   pmf / n_model_events interpolation            blueice/likelihood.py:356-357 (scipy _rgi.py:520-549)
//   BinnedLogLikelihood.adjust_expectations       blueice/likelihood.py:618-660
//   beeston_barlow_root1 / root2                  blueice/likelihood.py:693-712
//   BinnedLogLikelihood._compute_likelihood       blueice/likelihood.py:662-675
//   scipy.stats.poisson(lam).logpmf(k)            xlogy(k, lam) - gammaln(k + 1) - lam with scipy's
//                                                 argument checks (pinned in tests/test_oracle_pins.py)
//
// Per bin every operation is performed in the reference's order with separately rounded
// multiplies/adds, so the morphed pmfs, the Beeston-Barlow roots and lambda_b are bit-identical to
// NumPy; only the reductions over bins use the canonical order of DESIGN.md section 4
// (32-bin tree blocks, 512-bin sequential superblocks, 256-lane strided total).
//
// One warp per (point, 512-bin superblock); lane = bin.  Passes:
//   MODE 0  no Beeston-Barlow: lambda_b, log pmf, partial sums
//   MODE 1  BB pass A: t_b = A_b * w_b, partial sums of t (needs sum_b a_b, morphed from per-anchor sums)
//   MODE 2  BB pass B: pmf_i' = t_b / sum t, mu_i' = sum t * p_cal, lambda_b, log pmf, partial sums
#include "bi_common.cuh"

struct BiBinnedArgs {
    const double* pmf_anchor;      // [G, S, ld]
    const double* nm_anchor;       // [G, ld] calibration events of the BB source (or NULL)
    const double* nm_sum_anchor;   // [G] sum over bins of nm_anchor (or NULL)
    const double* observed;        // [B], or [P, obs_stride] when obs_stride != 0 (one dataset per point: toys)
    const double* lgamma_obs;      // same layout
    const int32_t* corner;         // [P, C]
    const double* weight;          // [P, C]
    const double* mus;             // [P, S]
    const int32_t* status;         // [P]
    const double* sum_t;           // [P] (MODE 2)
    double* partial;               // [P, n_chunks]
    int32_t* flags;                // [P]
    int64_t ld, n_bins, n_points, n_chunks, obs_stride;
    int32_t S, C, bb_source;
};

__device__ __forceinline__ double bi_morph_value(const double* __restrict__ base, int64_t row_stride, int64_t off,
                                                 const int32_t* __restrict__ corner, const double* __restrict__ w, int C) {
    if (C == 1) return base[(int64_t)corner[0] * row_stride + off];
    double acc = 0.0;
    for (int c = 0; c < C; ++c)
        acc = __dadd_rn(acc, __dmul_rn(base[(int64_t)corner[c] * row_stride + off], w[c]));
    return acc;
}

// the same morph with the point's corner offsets (corner * row_stride) and weights staged in shared memory: the C loads
// are independent of the accumulation chain, four of them are issued together
__device__ __forceinline__ double bi_morph_staged(const double* __restrict__ base, int64_t off,
                                                  const int64_t* __restrict__ s_off, const double* __restrict__ s_w, int C) {
    if (C == 1) return __ldg(base + s_off[0] + off);
    double acc = 0.0;
    int c = 0;
    for (; c + 4 <= C; c += 4) {
        const double v0 = __ldg(base + s_off[c] + off), v1 = __ldg(base + s_off[c + 1] + off);
        const double v2 = __ldg(base + s_off[c + 2] + off), v3 = __ldg(base + s_off[c + 3] + off);
        acc = __dadd_rn(acc, __dmul_rn(v0, s_w[c]));
        acc = __dadd_rn(acc, __dmul_rn(v1, s_w[c + 1]));
        acc = __dadd_rn(acc, __dmul_rn(v2, s_w[c + 2]));
        acc = __dadd_rn(acc, __dmul_rn(v3, s_w[c + 3]));
    }
    for (; c < C; ++c) acc = __dadd_rn(acc, __dmul_rn(__ldg(base + s_off[c] + off), s_w[c]));
    return acc;
}

// scipy.stats.poisson(lam).logpmf(k) with lgk = gammaln(k + 1) precomputed
__device__ __forceinline__ double bi_poisson_logpmf(double k, double lgk, double lam) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    if (!(lam >= 0.0) || k != k) return nan;
    if (k < 0.0 || k != floor(k)) return -inf;
    const double xl = (k == 0.0) ? 0.0 : __dmul_rn(k, log(lam));     // xlogy
    return __dsub_rn(__dsub_rn(xl, lgk), lam);
}

// Beeston-Barlow roots, operation order of likelihood.py:698-700 / 706-708
__device__ __forceinline__ void bi_bb_roots(double a, double p, double U, double d, double* r1, double* r2) {
    const double U2 = __dmul_rn(U, U), p2 = __dmul_rn(p, p), a2 = __dmul_rn(a, a), d2 = __dmul_rn(d, d);
    const double twoU = __dmul_rn(2.0, U);
    double disc = __dmul_rn(U2, p2);                                        // U**2*p**2
    disc = __dadd_rn(disc, __dmul_rn(__dmul_rn(2.0, U2), p));               // + 2*U**2*p
    disc = __dadd_rn(disc, U2);                                             // + U**2
    disc = __dadd_rn(disc, __dmul_rn(__dmul_rn(twoU, a), p2));              // + 2*U*a*p**2
    disc = __dadd_rn(disc, __dmul_rn(__dmul_rn(twoU, a), p));               // + 2*U*a*p
    disc = __dsub_rn(disc, __dmul_rn(__dmul_rn(twoU, d), p2));              // - 2*U*d*p**2
    disc = __dsub_rn(disc, __dmul_rn(__dmul_rn(twoU, d), p));               // - 2*U*d*p
    disc = __dadd_rn(disc, __dmul_rn(a2, p2));                              // + a**2*p**2
    disc = __dadd_rn(disc, __dmul_rn(__dmul_rn(__dmul_rn(2.0, a), d), p2)); // + 2*a*d*p**2
    disc = __dadd_rn(disc, __dmul_rn(d2, p2));                              // + d**2*p**2
    const double sq = sqrt(disc);
    double lin = __dmul_rn(-U, p);                                          // -U*p
    lin = __dsub_rn(lin, U);                                                // - U
    lin = __dadd_rn(lin, __dmul_rn(a, p));                                  // + a*p
    lin = __dadd_rn(lin, __dmul_rn(d, p));                                  // + d*p
    const double den = __dmul_rn(__dmul_rn(2.0, p), __dadd_rn(p, 1.0));     // 2*p*(p + 1)
    *r1 = __ddiv_rn(__dsub_rn(lin, sq), den);
    *r2 = __ddiv_rn(__dadd_rn(lin, sq), den);
}

template <int MODE>
__global__ void __launch_bounds__(256) k_binned_pass(const __grid_constant__ BiBinnedArgs a) {
    // per warp: the point's corner offsets into the pmf / n_model tensors and its corner weights
    __shared__ int64_t s_off_pmf_all[8][1 << BI_MAX_DIMS], s_off_nm_all[8][1 << BI_MAX_DIMS];
    __shared__ double s_w_all[8][1 << BI_MAX_DIMS];
    int64_t* s_off_pmf = s_off_pmf_all[threadIdx.x >> 5];
    int64_t* s_off_nm = s_off_nm_all[threadIdx.x >> 5];
    double* s_w = s_w_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_tasks = a.n_points * a.n_chunks;
    const int S = a.S, C = a.C, bi = a.bb_source;
    const int64_t row_stride = (int64_t)S * a.ld;

    for (int64_t task = warp_global; task < n_tasks; task += n_warps) {
        // chunk-major task order: warps that run together work on the SAME 512-bin chunk for different points, so the
        // anchor rows of the chunk (C * (S + 1) * 4 kB per hypercube cell) are served from L1 / L2 instead of HBM
        const int64_t j = task / a.n_points;
        const int64_t p = task - j * a.n_points;
        if (a.status[p] != 0) continue;
        const int32_t* corner = a.corner + p * C;
        const double* w = a.weight + p * C;
        const double* mu = a.mus + p * S;
        __syncwarp();
        for (int c = lane; c < C; c += 32) {
            s_off_pmf[c] = (int64_t)corner[c] * row_stride;
            s_off_nm[c] = (int64_t)corner[c] * a.ld;
            s_w[c] = w[c];
        }
        __syncwarp();

        double sum_a = 0.0, p_cal = 0.0, mu_adj = 0.0, sum_t = 0.0;
        if (MODE != 0) {
            // n_model_events[source_i].sum(): morph of the per-anchor bin sums
            sum_a = bi_morph_value(a.nm_sum_anchor, 1, 0, corner, w, C);
            p_cal = __ddiv_rn(mu[bi], sum_a);                                // likelihood.py:645
            if (MODE == 2) { sum_t = a.sum_t[p]; mu_adj = __dmul_rn(sum_t, p_cal); }   // likelihood.py:658
        }

        double s_sum = 0.0;
        int flag = 0;
        for (int it = 0; it < BI_SUPERBLOCK / 32; ++it) {
            const int64_t b0 = j * BI_SUPERBLOCK + (int64_t)it * 32;
            if (b0 >= a.n_bins) break;                                       // warp-uniform
            const int64_t b = b0 + lane;
            double val = 0.0;
            if (b < a.n_bins) {
                const double d = a.observed[p * a.obs_stride + b];
                double t_b = 0.0, pmf_i = 0.0;
                double pm_keep[8];
                if (MODE != 0) {
                    // u_b: sum over sources in order, the BB source contributes pmf_i * 0. (likelihood.py:635-641)
                    double u = 0.0;
                    if (S <= 8) {                                     // keep the morphed pmfs for the lambda_b sum below
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            if (s < S) {
                                const double pm = bi_morph_staged(a.pmf_anchor, (int64_t)s * a.ld + b, s_off_pmf, s_w, C);
                                pm_keep[s] = pm;
                                if (s == bi) pmf_i = pm;
                                const double term = __dmul_rn(pm, s == bi ? 0.0 : mu[s]);
                                u = (s == 0) ? term : __dadd_rn(u, term);
                            }
                        }
                    } else {
                        for (int s = 0; s < S; ++s) {
                            const double pm = bi_morph_staged(a.pmf_anchor, (int64_t)s * a.ld + b, s_off_pmf, s_w, C);
                            if (s == bi) pmf_i = pm;
                            const double term = __dmul_rn(pm, s == bi ? 0.0 : mu[s]);
                            u = (s == 0) ? term : __dadd_rn(u, term);
                        }
                    }
                    const double a_b = bi_morph_staged(a.nm_anchor, b, s_off_nm, s_w, C);
                    const double w_b = __dmul_rn(__ddiv_rn(pmf_i, a_b), sum_a);       // likelihood.py:646
                    double r1, r2;
                    bi_bb_roots(a_b, __dmul_rn(w_b, p_cal), u, d, &r1, &r2);
                    const double special = __ddiv_rn(__dadd_rn(d, a_b), __dadd_rn(1.0, p_cal));  // :652
                    const double A = (u == 0.0) ? special : r2;                       // :653
                    if (!(r1 <= 0.0)) flag |= BI_BB_ROOT1_POSITIVE;                   // :649
                    if (!(0.0 <= A)) flag |= BI_BB_NEGATIVE_A;                        // :655
                    t_b = __dmul_rn(A, w_b);                                          // :656
                }
                if (MODE == 1) {
                    val = t_b;
                } else {
                    // lambda_b = sum_s pmf_s * mu_s in source order (likelihood.py:667-670)
                    double lam = 0.0;
                    if (MODE == 2 && S <= 8) {
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            if (s < S) {
                                const double term = (s == bi) ? __dmul_rn(__ddiv_rn(t_b, sum_t), mu_adj)   // :657-658
                                                              : __dmul_rn(pm_keep[s], mu[s]);
                                lam = (s == 0) ? term : __dadd_rn(lam, term);
                            }
                        }
                    } else {
                        for (int s = 0; s < S; ++s) {
                            double term;
                            if (MODE == 2 && s == bi) {
                                term = __dmul_rn(__ddiv_rn(t_b, sum_t), mu_adj);      // :657-658
                            } else {
                                const double pm = bi_morph_staged(a.pmf_anchor, (int64_t)s * a.ld + b, s_off_pmf, s_w, C);
                                term = __dmul_rn(pm, mu[s]);
                            }
                            lam = (s == 0) ? term : __dadd_rn(lam, term);
                        }
                    }
                    val = bi_poisson_logpmf(d, a.lgamma_obs[p * a.obs_stride + b], lam);   // :674
                }
            }
#pragma unroll
            for (int x = 1; x < 32; x <<= 1) val = __dadd_rn(val, __shfl_xor_sync(BI_FULL_MASK, val, x));
            s_sum = __dadd_rn(s_sum, val);
        }
        if (MODE == 1) {
#pragma unroll
            for (int x = 1; x < 32; x <<= 1) flag |= __shfl_xor_sync(BI_FULL_MASK, flag, x);
            if (lane == 0 && flag) atomicOr(&a.flags[p], flag);
        }
        if (lane == 0) a.partial[p * a.n_chunks + j] = s_sum;
    }
}

// canonical total of [P, n_chunks] partials (same order as k_unbinned_finalize); status != 0 -> fill
__global__ void __launch_bounds__(256)
k_canonical_total(const double* __restrict__ partial, int64_t n_chunks, const int32_t* __restrict__ status,
                  double fill, double* __restrict__ out) {
    __shared__ double warp_tot[8];
    const int64_t p = blockIdx.x;
    const int t = threadIdx.x;
    if (status[p] != 0) { if (t == 0) out[p] = fill; return; }
    double u = 0.0;
    for (int64_t j = t; j < n_chunks; j += 256) u = __dadd_rn(u, partial[p * n_chunks + j]);
#pragma unroll
    for (int x = 1; x < 32; x <<= 1) u = __dadd_rn(u, __shfl_xor_sync(BI_FULL_MASK, u, x));
    if ((t & 31) == 0) warp_tot[t >> 5] = u;
    __syncthreads();
    if (t == 0)
        out[p] = __dadd_rn(__dadd_rn(__dadd_rn(warp_tot[0], warp_tot[1]), __dadd_rn(warp_tot[2], warp_tot[3])),
                           __dadd_rn(__dadd_rn(warp_tot[4], warp_tot[5]), __dadd_rn(warp_tot[6], warp_tot[7])));
}

__global__ void k_binned_mus_adj(const double* __restrict__ mus, const double* __restrict__ sum_t,
                                 const double* __restrict__ nm_sum_anchor, const int32_t* __restrict__ corner,
                                 const double* __restrict__ weight, const int32_t* __restrict__ status,
                                 int S, int C, int bb_source, int64_t n_points, double* __restrict__ mus_adj) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_points) return;
    for (int s = 0; s < S; ++s) mus_adj[p * S + s] = mus[p * S + s];
    if (bb_source >= 0 && status[p] == 0) {
        const double sum_a = bi_morph_value(nm_sum_anchor, 1, 0, corner + p * C, weight + p * C, C);
        const double p_cal = __ddiv_rn(mus[p * S + bb_source], sum_a);
        mus_adj[p * S + bb_source] = __dmul_rn(sum_t[p], p_cal);
    }
}

// pmf grid [S, B] for one point (full_output): morphed, BB source row replaced by t_b / sum_t
__global__ void __launch_bounds__(256)
k_binned_pmfs(const __grid_constant__ BiBinnedArgs a, double* __restrict__ out, int64_t ld_out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.n_bins) return;
    const int S = a.S, C = a.C, bi = a.bb_source;
    const int64_t row_stride = (int64_t)S * a.ld;
    const int32_t* corner = a.corner;
    const double* w = a.weight;
    const double* mu = a.mus;
    double u = 0.0, pmf_i = 0.0;
    for (int s = 0; s < S; ++s) {
        const double pm = bi_morph_value(a.pmf_anchor, row_stride, (int64_t)s * a.ld + b, corner, w, C);
        out[(int64_t)s * ld_out + b] = pm;
        if (bi >= 0) {
            if (s == bi) pmf_i = pm;
            const double term = __dmul_rn(pm, s == bi ? 0.0 : mu[s]);
            u = (s == 0) ? term : __dadd_rn(u, term);
        }
    }
    if (bi >= 0) {
        const double sum_a = bi_morph_value(a.nm_sum_anchor, 1, 0, corner, w, C);
        const double p_cal = __ddiv_rn(mu[bi], sum_a);
        const double a_b = bi_morph_value(a.nm_anchor, a.ld, b, corner, w, C);
        const double w_b = __dmul_rn(__ddiv_rn(pmf_i, a_b), sum_a);
        double r1, r2;
        bi_bb_roots(a_b, __dmul_rn(w_b, p_cal), u, a.observed[b], &r1, &r2);
        const double special = __ddiv_rn(__dadd_rn(a.observed[b], a_b), __dadd_rn(1.0, p_cal));
        const double A = (u == 0.0) ? special : r2;
        out[(int64_t)bi * ld_out + b] = __ddiv_rn(__dmul_rn(A, w_b), a.sum_t[0]);
    }
}

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
extern "C" int64_t bi_binned_scratch_doubles(int64_t n_points, int64_t n_bins) {
    return n_points * (2 * bi_num_superblocks(n_bins) + 1);
}

static int bi_binned_fill(BiBinnedArgs* a, const double* pmf_anchor_dev, const double* n_model_anchor_dev,
                          const double* n_model_sum_anchor_dev, int64_t ld_bins, int64_t n_bins, int32_t n_sources,
                          int32_t n_corners, int32_t bb_source, const double* observed_dev,
                          const double* lgamma_obs_dev, const int32_t* corner_dev, const double* weight_dev,
                          const double* mus_dev, const int32_t* status_dev, int64_t n_points) {
    BI_REQUIRE(n_bins >= 1 && ld_bins >= n_bins, "n_bins=%lld ld_bins=%lld", (long long)n_bins, (long long)ld_bins);
    BI_REQUIRE(n_sources >= 1 && n_sources <= BI_MAX_SOURCES, "n_sources=%d outside [1,%d]", n_sources, BI_MAX_SOURCES);
    BI_REQUIRE(n_corners >= 1 && n_corners <= (1 << BI_MAX_DIMS) && (n_corners & (n_corners - 1)) == 0, "bad n_corners=%d", n_corners);
    BI_REQUIRE(bb_source < n_sources, "bb_source=%d >= n_sources=%d", bb_source, n_sources);
    BI_REQUIRE(pmf_anchor_dev && observed_dev && corner_dev && weight_dev && mus_dev && status_dev, "bi_binned: NULL pointer");
    BI_REQUIRE(bb_source < 0 || (n_model_anchor_dev && n_model_sum_anchor_dev), "Beeston-Barlow needs the n_model tensors");
    memset(a, 0, sizeof(*a));
    a->pmf_anchor = pmf_anchor_dev; a->nm_anchor = n_model_anchor_dev; a->nm_sum_anchor = n_model_sum_anchor_dev;
    a->observed = observed_dev; a->lgamma_obs = lgamma_obs_dev;
    a->corner = corner_dev; a->weight = weight_dev; a->mus = mus_dev; a->status = status_dev;
    a->ld = ld_bins; a->n_bins = n_bins; a->n_points = n_points; a->n_chunks = bi_num_superblocks(n_bins);
    a->S = n_sources; a->C = n_corners; a->bb_source = bb_source;
    return BI_OK;
}

extern "C" int bi_binned_ll_batch(const double* pmf_anchor_dev, const double* n_model_anchor_dev,
                                  const double* n_model_sum_anchor_dev,
                                  int64_t ld_bins, int64_t n_bins, int32_t n_sources, int32_t n_corners,
                                  int32_t bb_source, const double* observed_dev, const double* lgamma_obs_dev,
                                  const int32_t* corner_dev, const double* weight_dev, const double* mus_dev,
                                  const int32_t* status_dev, int64_t n_points, double* scratch_dev,
                                  double* logl_dev, double* mus_adj_dev, int32_t* flags_dev, void* stream) {
    return bi_binned_ll_batch_toys(pmf_anchor_dev, n_model_anchor_dev, n_model_sum_anchor_dev, ld_bins, n_bins, n_sources,
                                   n_corners, bb_source, observed_dev, lgamma_obs_dev, 0, corner_dev, weight_dev, mus_dev,
                                   status_dev, n_points, scratch_dev, logl_dev, mus_adj_dev, flags_dev, stream);
}

extern "C" int bi_binned_ll_batch_toys(const double* pmf_anchor_dev, const double* n_model_anchor_dev,
                                       const double* n_model_sum_anchor_dev,
                                       int64_t ld_bins, int64_t n_bins, int32_t n_sources, int32_t n_corners,
                                       int32_t bb_source, const double* observed_dev, const double* lgamma_obs_dev,
                                       int64_t observed_stride,
                                       const int32_t* corner_dev, const double* weight_dev, const double* mus_dev,
                                       const int32_t* status_dev, int64_t n_points, double* scratch_dev,
                                       double* logl_dev, double* mus_adj_dev, int32_t* flags_dev, void* stream) {
    BI_REQUIRE(n_points >= 0, "n_points < 0");
    BI_REQUIRE(observed_stride == 0 || observed_stride >= n_bins, "observed_stride=%lld smaller than n_bins=%lld",
               (long long)observed_stride, (long long)n_bins);
    if (n_points == 0) return BI_OK;
    BiBinnedArgs a;
    int rc = bi_binned_fill(&a, pmf_anchor_dev, n_model_anchor_dev, n_model_sum_anchor_dev, ld_bins, n_bins, n_sources,
                            n_corners, bb_source, observed_dev, lgamma_obs_dev, corner_dev, weight_dev, mus_dev,
                            status_dev, n_points);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(lgamma_obs_dev && scratch_dev && logl_dev && flags_dev, "bi_binned_ll_batch: NULL pointer");
    a.obs_stride = observed_stride;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_tasks = n_points * a.n_chunks;
    int64_t blocks = (n_tasks + 7) / 8;
    if (blocks > 148 * 8 * 4) blocks = 148 * 8 * 4;
    double* part_a = scratch_dev;
    double* part_b = scratch_dev + n_points * a.n_chunks;
    double* sum_t = scratch_dev + 2 * n_points * a.n_chunks;
    a.flags = flags_dev;
    BI_CUDA_CHECK(cudaMemsetAsync(flags_dev, 0, sizeof(int32_t) * n_points, st));
    const double ninf = -INFINITY;
    if (bb_source < 0) {
        a.partial = part_b;
        k_binned_pass<0><<<(unsigned)blocks, 256, 0, st>>>(a);
    } else {
        a.partial = part_a;
        k_binned_pass<1><<<(unsigned)blocks, 256, 0, st>>>(a);
        k_canonical_total<<<(unsigned)n_points, 256, 0, st>>>(part_a, a.n_chunks, status_dev, 0.0, sum_t);
        a.partial = part_b;
        a.sum_t = sum_t;
        k_binned_pass<2><<<(unsigned)blocks, 256, 0, st>>>(a);
    }
    k_canonical_total<<<(unsigned)n_points, 256, 0, st>>>(part_b, a.n_chunks, status_dev, ninf, logl_dev);
    if (mus_adj_dev) {
        const int64_t nb = (n_points + 127) / 128;
        k_binned_mus_adj<<<(unsigned)nb, 128, 0, st>>>(mus_dev, sum_t, n_model_sum_anchor_dev, corner_dev, weight_dev,
                                                       status_dev, n_sources, n_corners, bb_source, n_points, mus_adj_dev);
    }
    BI_LAUNCH_CHECK();
    return BI_OK;
}

extern "C" int bi_binned_pmfs(const double* pmf_anchor_dev, const double* n_model_anchor_dev,
                              const double* n_model_sum_anchor_dev,
                              int64_t ld_bins, int64_t n_bins, int32_t n_sources, int32_t n_corners,
                              int32_t bb_source, const double* observed_dev,
                              const int32_t* corner_dev, const double* weight_dev, const double* mus_dev,
                              const double* sum_t_dev, double* pmf_out_dev, int64_t ld_out, void* stream) {
    BiBinnedArgs a;
    int32_t dummy_status = 0;
    int rc = bi_binned_fill(&a, pmf_anchor_dev, n_model_anchor_dev, n_model_sum_anchor_dev, ld_bins, n_bins, n_sources,
                            n_corners, bb_source, observed_dev, NULL, corner_dev, weight_dev, mus_dev,
                            &dummy_status, 1);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(pmf_out_dev && ld_out >= n_bins, "bi_binned_pmfs: bad output");
    BI_REQUIRE(bb_source < 0 || sum_t_dev, "bi_binned_pmfs: sum_t_dev needed with Beeston-Barlow");
    a.status = NULL;
    a.sum_t = sum_t_dev;
    const int64_t blocks = (n_bins + 255) / 256;
    k_binned_pmfs<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, pmf_out_dev, ld_out);
    BI_LAUNCH_CHECK();
    return BI_OK;
}
