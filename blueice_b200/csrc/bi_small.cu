// blueice_b200 -- the whole unbinned evaluation of a TINY batch in ONE launch (K1 + K2 + finalize fused).
//
// Replaces LogLikelihoodBase.__call__'s numerics (blueice/likelihood.py:318-427 with :571-573, :678-690) for the latency
// regime: a minimiser or interval search calling ll(**params) thousands of times on a few thousand events
// (inference.py:131-178,332-389; BASELINE config 1).  There the four dependent launches of bi_unbinned_ll_batch cost more
// than the arithmetic.  One CTA per parameter point:
//   phase 0  thread 0: the point set-up (bi_setup_point: cells, weights, mus, contraction terms), tables in shared memory
//   phase 1  warps: units (superblock j, 32-event group g), lane = event: density = fma chain over the terms, canonical
//            pair / quad / oct product tree by shuffles -> per class (m, e) and the reference-semantics fallback log l
//   phase 2  4 lanes per superblock: the sequential products over its 16 groups, classes combined, ONE log
//   phase 3  warp 0: the canonical total of bi_unbinned_finalize
// Every value is formed by the operations of K1 / k_unbinned_mma / k_unbinned_finalize in the same order, so results are
// BIT-IDENTICAL to bi_unbinned_ll_batch's multi-launch path (tested); inputs and outputs may live in pinned host memory
// (the kernel reads / writes them over PCIe directly), which removes the copy nodes around the launch.
#include <string.h>

#include "bi_setup_point.cuh"

#define BI_SMALL_THREADS 256
#define BI_SMALL_MAX_SUPER 16           /* <= 8192 events */
#define BI_SMALL_MAX_TERMS 128
#define BI_RANGE_LO ((1023 - 126) << 20)
#define BI_RANGE_SPAN (253u << 20)

struct BiSmallArgs {
    BiGrid grid;
    BiAllowNegative allow;
    const double* zs; const double* rate_mult; const double* scale; const double* eff;   // device-accessible
    const double* mus_anchor;
    const double* rows;                 // [G * S, ld] anchor tensor
    int64_t ld, n_events, n_points;
    double outlier;
    double* logl; double* logsum; double* musum; int32_t* status;                         // device-accessible
    int32_t n_sources;
};

// reference semantics for one event (likelihood.py:686-689): per source ps = fma chain over its terms of A * w_corner,
// p = nansum_s(mu_s * ps), non-positive / NaN -> outlier (as k_unbinned_mma's and k_template_partials' fallbacks)
static __device__ __noinline__ double bi_small_slow_density(const double* __restrict__ rows, const int64_t* rowoff,
                                                            int64_t i, int K, int S, const double* wterm,
                                                            const double* mu, double outlier) {
    double acc = 0.0;
    for (int s = 0; s < S; ++s) {
        double ps = 0.0;
        for (int k = s; k < K; k += S) ps = fma(rows[rowoff[k] + i], wterm[k], ps);      // term k = corner * S + source
        const double term = __dmul_rn(mu[s], ps);
        if (term == term) acc = __dadd_rn(acc, term);
    }
    return bi_fix_density(acc, outlier);
}

__global__ void __launch_bounds__(BI_SMALL_THREADS) k_unbinned_small(const __grid_constant__ BiSmallArgs a) {
    __shared__ int64_t s_rowoff[BI_SMALL_MAX_TERMS];
    __shared__ double s_coef[BI_SMALL_MAX_TERMS], s_wterm[BI_SMALL_MAX_TERMS];
    __shared__ int32_t s_row[BI_SMALL_MAX_TERMS], s_corner[1 << BI_MAX_DIMS], s_cell[BI_MAX_DIMS];
    __shared__ double s_weight[1 << BI_MAX_DIMS], s_frac[BI_MAX_DIMS], s_mu[BI_MAX_SOURCES];
    __shared__ double s_musum;
    __shared__ int32_t s_status;
    __shared__ double s_m[BI_SMALL_MAX_SUPER][16][4], s_l[BI_SMALL_MAX_SUPER][16][4];
    __shared__ int32_t s_e[BI_SMALL_MAX_SUPER][16][4], s_slow[BI_SMALL_MAX_SUPER];
    __shared__ double s_partial[BI_SMALL_MAX_SUPER];
    const int64_t p = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = a.grid.n_dims, S = a.n_sources, K = a.grid.n_corners * S;
    const int n_super = (int)((a.n_events + BI_SUPERBLOCK - 1) / BI_SUPERBLOCK);

    // ---- phase 0
    if (tid == 0) {
        bi_setup_point(a.grid, S, a.zs + p * D, a.rate_mult + p * S, a.scale ? a.scale + p : nullptr,
                       a.eff ? a.eff + p * S : nullptr, a.mus_anchor, a.allow, s_cell, s_frac, s_corner, s_weight, s_mu,
                       &s_musum, &s_status, s_row, s_coef, s_wterm);
        for (int k = 0; k < K; ++k) s_rowoff[k] = (int64_t)s_row[k] * a.ld;
    }
    if (tid < BI_SMALL_MAX_SUPER) s_slow[tid] = 0;
    __syncthreads();
    const int status = s_status;
    if (status != 0) {                                          // soft failure: -inf (likelihood.py:347,402)
        if (tid == 0) {
            a.logl[p] = -__longlong_as_double(0x7ff0000000000000LL);
            if (a.logsum) a.logsum[p] = 0.0;
            a.musum[p] = s_musum;
            a.status[p] = status;
        }
        return;
    }

    // ---- phase 1: units (j, g)
    const int t_class = (lane >> 1) & 3;
    const unsigned class_mask = 0x03030303u << (2 * t_class);   // lanes of class t: 8n + 2t + {0, 1}
    const int n_units = n_super * 16;
    for (int u = warp; u < n_units; u += BI_SMALL_THREADS / 32) {
        const int j = u >> 4, g = u & 15;
        const int64_t i = (int64_t)j * BI_SUPERBLOCK + g * 32 + lane;
        const bool valid = i < a.n_events;
        double pd = 0.0;
        if (valid) {
            for (int k = 0; k < K; ++k) pd = fma(a.rows[s_rowoff[k] + i], s_coef[k], pd);
        } else {
            pd = 1.0;                                           // events >= N count as p = 1
        }
        const bool in_range = (unsigned)(__double2hiint(pd) - BI_RANGE_LO) < BI_RANGE_SPAN;
        const unsigned bad = ~__ballot_sync(BI_FULL_MASK, in_range);
        double v = __dmul_rn(pd, __shfl_xor_sync(BI_FULL_MASK, pd, 1));
        v = __dmul_rn(v, __shfl_xor_sync(BI_FULL_MASK, v, 8));
        v = __dmul_rn(v, __shfl_xor_sync(BI_FULL_MASK, v, 16));
        double m;
        int e;
        bi_split(v, &m, &e);
        const bool class_bad = (bad & class_mask) != 0;
        if (class_bad) { m = 1.0; e = 0; }
        double l = 0.0;
        if (bad) {                                              // warp-uniform; rare
            if (class_bad && valid)
                l = log(bi_small_slow_density(a.rows, s_rowoff, i, K, S, s_wterm, s_mu, a.outlier));
            l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 1));
            l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 8));
            l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 16));
            if (!class_bad) l = 0.0;
            if (lane == 0) s_slow[j] = 1;
        }
        if ((lane & 0x19) == 0) {                               // lanes 0, 2, 4, 6: the representatives of classes 0..3
            s_m[j][g][t_class] = m;
            s_e[j][g][t_class] = e;
            s_l[j][g][t_class] = l;
        }
    }
    __syncthreads();

    // ---- phase 2: superblock j by lanes 4 j .. 4 j + 3 (class t): M_t = (((1 * m_0) * m_1) ...), E_t, L_t sequential
    if (tid < 4 * n_super) {
        const int j = tid >> 2, t = tid & 3;
        const int n_ev = (int)min((int64_t)BI_SUPERBLOCK, a.n_events - (int64_t)j * BI_SUPERBLOCK);
        const int n_grp = (n_ev + 31) >> 5;                     // k_unbinned_mma walks only the groups that hold events
        double M = 1.0, L = 0.0;
        int E = 0;
        for (int g = 0; g < n_grp; ++g) {
            M = __dmul_rn(M, s_m[j][g][t]);
            E += s_e[j][g][t];
            L = __dadd_rn(L, s_l[j][g][t]);
        }
        const unsigned quad = 0xfu << (4 * (j & 7));             // the 4 lanes of this superblock
        M = __dmul_rn(M, __shfl_xor_sync(quad, M, 1));
        M = __dmul_rn(M, __shfl_xor_sync(quad, M, 2));
        E += __shfl_xor_sync(quad, E, 1);
        E += __shfl_xor_sync(quad, E, 2);
        double part = bi_block_log(M, E);
        if (s_slow[j]) {
            L = __dadd_rn(L, __shfl_xor_sync(quad, L, 1));
            L = __dadd_rn(L, __shfl_xor_sync(quad, L, 2));
            part = __dadd_rn(part, L);
        }
        if (t == 0) s_partial[j] = part;
    }
    __syncthreads();

    // ---- phase 3: canonical total (bi_unbinned_finalize): lane t holds 0 + partial[t], xor butterfly per warp, 8 warps
    if (warp == 0) {
        double u = 0.0;
        if (lane < n_super) u = __dadd_rn(u, s_partial[lane]);
#pragma unroll
        for (int x = 1; x < 32; x <<= 1) u = __dadd_rn(u, __shfl_xor_sync(BI_FULL_MASK, u, x));
        if (lane == 0) {
            const double z = 0.0;                               // warps 1..7 of the 256-thread finalize hold zeros
            const double total = __dadd_rn(__dadd_rn(__dadd_rn(u, z), __dadd_rn(z, z)),
                                           __dadd_rn(__dadd_rn(z, z), __dadd_rn(z, z)));
            a.logl[p] = __dadd_rn(-s_musum, total);             // likelihood.py:690
            if (a.logsum) a.logsum[p] = total;
            a.musum[p] = s_musum;
            a.status[p] = 0;
        }
    }
}

int bi_fill_grid(BiGrid* g, int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host);

// 1 when bi_unbinned_ll_small can evaluate this batch (bi_unbinned_ll_batch uses it then)
extern "C" int32_t bi_unbinned_small_ok(int32_t n_dims, int32_t n_sources, int64_t n_points, int64_t n_events) {
    if (n_dims < 0 || n_dims > BI_MAX_DIMS || n_sources < 1) return 0;
    const int64_t K = ((int64_t)1 << n_dims) * n_sources;
    const int64_t n_super = (n_events + BI_SUPERBLOCK - 1) / BI_SUPERBLOCK;
    return K <= BI_SMALL_MAX_TERMS && n_events > 0 && n_super <= BI_SMALL_MAX_SUPER && n_points >= 1 &&
           n_points * n_super <= 2048;
}

extern "C" int bi_unbinned_ll_small(int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host,
                                    int32_t n_sources, int64_t n_points,
                                    const double* zs, const double* rate_mult, const double* scale, const double* eff,
                                    const double* mus_anchor_dev, const uint8_t* allow_negative_host,
                                    const double* ps_anchor_dev, int64_t ld_events, int64_t n_events,
                                    double outlier_likelihood,
                                    double* logl, double* logsum, double* musum, int32_t* status, void* stream) {
    BI_REQUIRE(bi_unbinned_small_ok(n_dims, n_sources, n_points, n_events),
               "bi_unbinned_ll_small: batch too large (terms <= %d, events <= %d, points x superblocks <= 2048)",
               BI_SMALL_MAX_TERMS, BI_SMALL_MAX_SUPER * BI_SUPERBLOCK);
    BI_REQUIRE(rate_mult && mus_anchor_dev && ps_anchor_dev && logl && musum && status && (n_dims == 0 || zs),
               "bi_unbinned_ll_small: NULL pointer");
    BI_REQUIRE(ld_events >= n_events, "ld_events < n_events");
    BiSmallArgs a;
    BiSmallArgs* ap = &a;
    int rc = bi_fill_grid(&ap->grid, n_dims, n_anchors_host, axes_host);
    if (rc != BI_OK) return rc;
    memset(&ap->allow, 0, sizeof(ap->allow));
    if (allow_negative_host)
        for (int s = 0; s < n_sources; ++s) {
            ap->allow.flag[s] = allow_negative_host[s] ? 1 : 0;
            ap->allow.any |= ap->allow.flag[s];
        }
    ap->zs = zs; ap->rate_mult = rate_mult; ap->scale = scale; ap->eff = eff;
    ap->mus_anchor = mus_anchor_dev; ap->rows = ps_anchor_dev;
    ap->ld = ld_events; ap->n_events = n_events; ap->n_points = n_points; ap->outlier = outlier_likelihood;
    ap->logl = logl; ap->logsum = logsum; ap->musum = musum; ap->status = status; ap->n_sources = n_sources;
    k_unbinned_small<<<(unsigned)n_points, BI_SMALL_THREADS, 0, (cudaStream_t)stream>>>(*ap);
    BI_LAUNCH_CHECK();
    return BI_OK;
}
