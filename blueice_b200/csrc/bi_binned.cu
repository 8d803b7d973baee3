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
#include <stdlib.h>
#include <string.h>

#include "bi_common.cuh"
#include "bi_tma.cuh"

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
    // tiled kernels (k_binned_tile): schedule written by bi_unbinned_plan, per-bin scratch, per-block sums
    const int32_t* group_points;   // [P] point indices, cell-major
    const int32_t* groups;         // [n_groups, 4] (first, count, -, -): points of a group share their hypercube cell
    const int32_t* header;         // header[0] = n_groups
    double* tbuf;                  // [P, ld] t_b = A_b * w_b of the Beeston-Barlow pass (written by MODE 1, read by MODE 2)
    double* blockval;              // [P, 16 * n_super] tree-reduced sums of 32-bin blocks, block k of a point at bi_bv_index(k)
    int64_t n_blocks, n_tiles, tb_ld;
    int32_t tile;                  // bins per tile: 128 or 256
    int32_t store_terms;           // MODE 1: keep all S terms of lambda_b per (point, bin) in tbuf [P, BI_BIN_STORE_S, tb_ld]
    int32_t bv_interleaved;        // layout of blockval rows, see bi_bv_index
    double* pt_const;              // [P, 4] per-point constants of the Beeston-Barlow passes: sum_a, p_cal (k_binned_plan),
                                   // sum_t, mu_adj (k_canonical_total_blocks after pass A)
};

// block sums are stored superblock-interleaved -- block k = 16 j + i at i * n_super + j -- so that the thread that adds the
// 16 blocks of superblock j reads 16 coalesced rows (k_canonical_total_blocks; block-major rows cost one tag lookup per
// value: 13 us per total at 25 000 blocks)
// (a handful of points only: at a scan the 8 blocks of a tile would land in 8 different sectors, and the totals of many
// points run side by side anyway -- there the rows stay block-major)
__host__ __device__ __forceinline__ int64_t bi_bv_index(int64_t k, int64_t n_blocks, int interleaved) {
    if (!interleaved) return k;
    const int64_t n_super = (n_blocks + 15) >> 4;
    return (k & 15) * n_super + (k >> 4);
}
__host__ __device__ __forceinline__ int64_t bi_bv_ld(int64_t n_blocks) { return ((n_blocks + 15) >> 4) << 4; }

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

// the same arithmetic with root 1 left undivided: returns den, *x1 = numerator of root 1 (root1 = x1 / den), *r2 = root 2
__device__ __forceinline__ double bi_bb_roots_lazy(double a, double p, double U, double d, double* x1, double* r2) {
    const double U2 = __dmul_rn(U, U), p2 = __dmul_rn(p, p), a2 = __dmul_rn(a, a), d2 = __dmul_rn(d, d);
    const double twoU = __dmul_rn(2.0, U);
    double disc = __dmul_rn(U2, p2);
    disc = __dadd_rn(disc, __dmul_rn(__dmul_rn(2.0, U2), p));
    disc = __dadd_rn(disc, U2);
    disc = __dadd_rn(disc, __dmul_rn(__dmul_rn(twoU, a), p2));
    disc = __dadd_rn(disc, __dmul_rn(__dmul_rn(twoU, a), p));
    disc = __dsub_rn(disc, __dmul_rn(__dmul_rn(twoU, d), p2));
    disc = __dsub_rn(disc, __dmul_rn(__dmul_rn(twoU, d), p));
    disc = __dadd_rn(disc, __dmul_rn(a2, p2));
    disc = __dadd_rn(disc, __dmul_rn(__dmul_rn(__dmul_rn(2.0, a), d), p2));
    disc = __dadd_rn(disc, __dmul_rn(d2, p2));
    const double sq = sqrt(disc);
    double lin = __dmul_rn(-U, p);
    lin = __dsub_rn(lin, U);
    lin = __dadd_rn(lin, __dmul_rn(a, p));
    lin = __dadd_rn(lin, __dmul_rn(d, p));
    const double den = __dmul_rn(__dmul_rn(2.0, p), __dadd_rn(p, 1.0));
    *x1 = __dsub_rn(lin, sq);
    *r2 = __ddiv_rn(__dadd_rn(lin, sq), den);
    return den;
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

// ---------------------------------------------------------------------------------------------
// Tiled form: the anchor rows of a bin tile live in shared memory, staged ONCE per (tile, point group) by 1-D TMA bulk
// copies, and every point of the group (points that share their hypercube cell, bi_unbinned_plan) is evaluated on them.
// One CTA per task (tile-major, so the cells of a scan find the tile's anchors in L2); inside, one warp per
// (point, 32-bin block), lane = bin.  The per-bin arithmetic is exactly that of k_binned_pass (reference operation
// order, separately rounded); the sum over the bins of a block is the same xor tree, and the 16 block sums of a
// 512-bin superblock are added sequentially by k_canonical_total_blocks -- so results are bit-identical to k_binned_pass.
//   MODE 0  no Beeston-Barlow: stages S rows per corner
//   MODE 1  BB pass A: stages S + 1 rows per corner (pmfs + calibration counts); t_b -> tbuf, block sums of t
//   MODE 2  BB pass B: stages the S - 1 pmf rows of the other sources; t_b from tbuf; block sums of log pmf
// ---------------------------------------------------------------------------------------------
#define BI_BIN_TILE 256                 /* largest tile: 8 blocks of 32 bins (the kernel runs 128- or 256-bin tiles) */
#define BI_BIN_STORE_S 8                /* stored-terms form of the Beeston-Barlow passes: up to 8 sources */
#define BI_BIN_GROUP_POINTS 32          /* points per group */
#define BI_BIN_GCACHE 32                /* groups whose cell anchors k_binned_tile keeps in shared memory ... */
#define BI_BIN_GCACHE_C 16              /* ... when the morph has at most this many corners */

// byte offset of the staged rows inside the dynamic shared memory of k_binned_tile (tables first, rows 128-byte aligned)
__host__ __device__ inline int bi_bin_rows_offset(int C, int S) {
    return (128 + (BI_BIN_GROUP_POINTS * (C + S + 4)) * 8 + (BI_BIN_GROUP_POINTS + C) * 4 + 127) & ~127;
}

// morph of staged row kind k at tile offset off: rows[(k * C + c) * NB + off] over the corners c (compile-time stride:
// the corner loop costs no address arithmetic), accumulated like bi_morph_value
template <int CT, int NB>   // CT = compile-time corner count (1, 2, 4, 8) or 0: run-time C
__device__ __forceinline__ double bi_morph_smem(const double* __restrict__ rows_k, const double* __restrict__ w, int C) {
    const int n_c = CT ? CT : C;
    if (n_c == 1) return rows_k[0];
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < n_c; ++c) acc = __dadd_rn(acc, __dmul_rn(rows_k[c * NB], w[c]));
    return acc;
}

// dynamic shared memory of k_binned_tile: two stages (the next task's tile is fetched while the current one is evaluated),
// each = tables (bi_bin_rows_offset bytes: mbarrier, weights, mus, per-point constants, point indices, corners) + rows
__host__ __device__ inline int bi_bin_stage_bytes(int C, int S, int n_kinds, int NB) {
    return bi_bin_rows_offset(C, S) + C * n_kinds * NB * 8;
}

template <int MODE, int CT, int NB, int NT, int NSTAGE>
__device__ __forceinline__ void bi_binned_tile_body(const BiBinnedArgs& a, unsigned char* bi_bin_smem) {
    const int S = a.S, C = CT ? CT : a.C, bi = a.bb_source;
    const int n_kinds = MODE == 0 ? S : (MODE == 1 ? S + 1 : S - 1);          // staged row kinds (x C corners each)
    const int n_rows = C * n_kinds;
    const int stage_bytes = bi_bin_stage_bytes(C, S, n_kinds, NB);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_groups = a.header[0];
    const int64_t n_tasks = a.n_tiles * n_groups;
    if (tid < NSTAGE) bi_mbar_init(reinterpret_cast<uint64_t*>(bi_bin_smem + tid * stage_bytes), 1);
    // a few groups (a handful of points): their (first, count) and the anchors of their cell are kept in shared memory --
    // otherwise every task starts with three dependent global loads (group -> lead point -> corners) before its TMA
    // copies can be issued, 2 us per 80 kB tile at one point
    __shared__ int32_t s_ginfo[BI_BIN_GCACHE][2];
    __shared__ int32_t s_ganchor[BI_BIN_GCACHE][BI_BIN_GCACHE_C];
    const bool gcache = n_groups <= BI_BIN_GCACHE && C <= BI_BIN_GCACHE_C;
    if (gcache) {
        for (int i = tid; i < n_groups * C; i += NT) {
            const int g = i / C, c = i - g * C;
            const int first = a.groups[4 * g];
            if (c == 0) { s_ginfo[g][0] = first; s_ginfo[g][1] = a.groups[4 * g + 1]; }
            s_ganchor[g][c] = a.corner[(int64_t)a.group_points[first] * C + c];
        }
    }
    __syncthreads();

    // stage a task: tables by ordinary stores, rows by TMA bulk copies that complete on the stage's mbarrier
    auto stage = [&](int64_t task, int buf) {
        unsigned char* base = bi_bin_smem + buf * stage_bytes;
        uint64_t* bar = reinterpret_cast<uint64_t*>(base);
        double* s_w = reinterpret_cast<double*>(base + 128);
        double* s_mu = s_w + BI_BIN_GROUP_POINTS * C;
        double* s_pt = s_mu + BI_BIN_GROUP_POINTS * S;
        int32_t* s_pidx = reinterpret_cast<int32_t*>(s_pt + BI_BIN_GROUP_POINTS * 4);
        double* s_rows = reinterpret_cast<double*>(base + bi_bin_rows_offset(C, S));
        const int64_t tile = task / n_groups;
        const int g = (int)(task - tile * n_groups);
        const int first = gcache ? s_ginfo[g][0] : a.groups[4 * g], count = gcache ? s_ginfo[g][1] : a.groups[4 * g + 1];
        const int64_t bin0 = tile * NB;
        const int tile_bins = (int)min((int64_t)NB, a.ld - bin0);            // even (ld is)
        if (tid == 0) bi_mbar_expect_tx(bar, (unsigned)(n_rows * tile_bins * 8));
        const int64_t lead = gcache ? 0 : a.group_points[first];             // the group's cell = its first point's
        // a bulk copy is issued lane by lane (uniform-register operands): row r goes to lane r / n_warps of warp r % n_warps,
        // so the copies of a tile are spread over all warps instead of serialised in the first one or two
#ifndef BI_BINNED_SPREAD_COPIES
#define BI_BINNED_SPREAD_COPIES 1
#endif
        for (int r = BI_BINNED_SPREAD_COPIES ? lane * (NT / 32) + warp : tid; r < n_rows; r += NT) {
            const int k = r / C, c = r - k * C;
            const int64_t anchor = gcache ? s_ganchor[g][c] : a.corner[lead * C + c];
            const double* src;
            if (MODE == 1 && k == S) {
                src = a.nm_anchor + anchor * a.ld + bin0;
            } else {
                const int s = (MODE == 2 && k >= bi) ? k + 1 : k;            // MODE 2 skips the BB source's row
                src = a.pmf_anchor + (anchor * S + s) * a.ld + bin0;
            }
            bi_bulk_g2s(s_rows + (int64_t)r * NB, src, (unsigned)(tile_bins * 8), bar);
        }
        if (tid < count) s_pidx[tid] = a.group_points[first + tid];
        for (int i = tid; i < count * C; i += NT) {
            const int q = i / C, c = i - q * C;
            s_w[q * C + c] = a.weight[(int64_t)a.group_points[first + q] * C + c];
        }
        for (int i = tid; i < count * S; i += NT) {
            const int q = i / S, s2 = i - q * S;
            s_mu[q * S + s2] = a.mus[(int64_t)a.group_points[first + q] * S + s2];
        }
        if (MODE != 0 && tid < 4 * count) {                   // sum_a, p_cal, sum_t, mu_adj (likelihood.py:645,658)
            const int64_t p = a.group_points[first + (tid >> 2)];
            s_pt[tid] = a.pt_const[4 * p + (tid & 3)];
        }
    };

    unsigned parity[2] = {0, 0};
    int buf = 0;
    if (NSTAGE == 2 && (int64_t)blockIdx.x < n_tasks) stage(blockIdx.x, 0);
    for (int64_t task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        // NSTAGE = 2: the other stage was released by the __syncthreads that closed the previous iteration: prefetch the
        // next task.  NSTAGE = 1: one stage, filled now (the co-resident CTA evaluates meanwhile)
        if (NSTAGE == 2) { if (task + gridDim.x < n_tasks) stage(task + gridDim.x, buf ^ 1); }
        else stage(task, 0);
        unsigned char* base = bi_bin_smem + buf * stage_bytes;
        uint64_t* bar = reinterpret_cast<uint64_t*>(base);
        const double* s_w = reinterpret_cast<const double*>(base + 128);
        const double* s_mu = s_w + BI_BIN_GROUP_POINTS * C;
        const double* s_pt = s_mu + BI_BIN_GROUP_POINTS * S;
        const int32_t* s_pidx = reinterpret_cast<const int32_t*>(s_pt + BI_BIN_GROUP_POINTS * 4);
        const double* s_rows = reinterpret_cast<const double*>(base + bi_bin_rows_offset(C, S));
        const int64_t tile = task / n_groups;
        const int g = (int)(task - tile * n_groups);
        const int count = gcache ? s_ginfo[g][1] : a.groups[4 * g + 1];
        const int64_t bin0 = tile * NB;
        __syncthreads();                       // this stage's tables (written by other threads) are visible
        bi_mbar_wait(bar, parity[buf]);
        parity[buf] ^= 1;

        // ---- compute: items = (point q of the group, PAIR of 32-bin blocks kk, kk + half): the two blocks are independent
        // dependency chains the compiler interleaves (the FP64 chains of one block alone leave the pipe idle)
        const int n_blk = (int)min((int64_t)(NB / 32), (a.n_bins - bin0 + 31) / 32);
        // a handful of points: one block per item, so that every warp of the CTA has work (a single point is 8 blocks)
        const bool pairs = count * ((n_blk + 1) >> 1) >= NT / 32;
        const int half = pairs ? (n_blk + 1) >> 1 : n_blk;
        for (int item = warp; item < count * half; item += NT / 32) {
            const int q = item / half, kk = item - q * half;
            const int64_t p = s_pidx[q];
            double wr[CT ? CT : 1];
            const double* w = s_w + q * C;
            if (CT) {
#pragma unroll
                for (int c = 0; c < (CT ? CT : 1); ++c) wr[c] = w[c];
                w = wr;
            }
            const double* mu = s_mu + q * S;
            double val[2] = {0.0, 0.0};
            int flag = 0;
            const double* obs_p = a.observed + p * a.obs_stride;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int k = kk + j * half;
                const int off = k * 32 + lane;
                const int64_t b = bin0 + off;
                if ((j == 0 || pairs) && k < n_blk && b < a.n_bins) {
                    const double d = obs_p[b];
                    const double* r0 = s_rows + off;
                    if (MODE == 1) {
                        const double sum_a = s_pt[q * 4], p_cal = s_pt[q * 4 + 1];
                        // u_b: sum over sources in order, the BB source contributes pmf_i * 0. (likelihood.py:635-641)
                        double u = 0.0, pmf_i = 0.0;
                        for (int s = 0; s < S; ++s) {
                            const double pm = bi_morph_smem<CT, NB>(r0 + s * C * NB, w, C);
                            if (s == bi) pmf_i = pm;
                            const double term = __dmul_rn(pm, s == bi ? 0.0 : mu[s]);
                            u = (s == 0) ? term : __dadd_rn(u, term);
                            if (a.store_terms && s != bi)                              // the term lambda_b takes (:667-670)
                                a.tbuf[(p * BI_BIN_STORE_S + s) * a.tb_ld + b] = __dmul_rn(pm, mu[s]);
                        }
                        const double a_b = bi_morph_smem<CT, NB>(r0 + S * C * NB, w, C);
                        const double w_b = __dmul_rn(__ddiv_rn(pmf_i, a_b), sum_a);   // likelihood.py:646
                        double x1, r2;
                        const double den = bi_bb_roots_lazy(a_b, __dmul_rn(w_b, p_cal), u, d, &x1, &r2);
                        // A = where(u == 0, (d + a) / (1 + p_cal), root2)  (:652-653); the special case is rare
                        const double A = (u == 0.0) ? __ddiv_rn(__dadd_rn(d, a_b), __dadd_rn(1.0, p_cal)) : r2;
                        // assert all(root1 <= 0) (:649): root1 = x1 / den; finite x1 <= 0 < den (or x1 >= 0 > den) gives a
                        // quotient that is <= 0 (or -0) without dividing; anything else (equal signs, zero / NaN den) is divided
                        if (!(fabs(x1) <= 1.7976931348623157e308 && fabs(den) <= 1.7976931348623157e308 &&
                              ((x1 <= 0.0 && den > 0.0) || (x1 >= 0.0 && den < 0.0)))) {
                            if (!(__ddiv_rn(x1, den) <= 0.0)) flag |= BI_BB_ROOT1_POSITIVE;
                        }
                        if (!(0.0 <= A)) flag |= BI_BB_NEGATIVE_A;                    // :655
                        val[j] = __dmul_rn(A, w_b);                                   // :656
                        if (a.store_terms) a.tbuf[(p * BI_BIN_STORE_S + bi) * a.tb_ld + b] = val[j];
                        else a.tbuf[p * a.tb_ld + b] = val[j];
                    } else {
                        // lambda_b = sum_s pmf_s * mu_s in source order (likelihood.py:667-670)
                        double lam = 0.0;
                        for (int s = 0; s < S; ++s) {
                            double term;
                            if (MODE == 2 && s == bi) {
                                term = __dmul_rn(__ddiv_rn(a.tbuf[p * a.tb_ld + b], s_pt[q * 4 + 2]), s_pt[q * 4 + 3]);   // :657-658
                            } else {
                                const int k_row = (MODE == 2 && s > bi) ? s - 1 : s;
                                term = __dmul_rn(bi_morph_smem<CT, NB>(r0 + k_row * C * NB, w, C), mu[s]);
                            }
                            lam = (s == 0) ? term : __dadd_rn(lam, term);
                        }
                        val[j] = bi_poisson_logpmf(d, a.lgamma_obs[p * a.obs_stride + b], lam);   // :674
                    }
                }
            }
#pragma unroll
            for (int x = 1; x < 32; x <<= 1) {
                val[0] = __dadd_rn(val[0], __shfl_xor_sync(BI_FULL_MASK, val[0], x));
                val[1] = __dadd_rn(val[1], __shfl_xor_sync(BI_FULL_MASK, val[1], x));
            }
            if (MODE == 1) {
#pragma unroll
                for (int x = 1; x < 32; x <<= 1) flag |= __shfl_xor_sync(BI_FULL_MASK, flag, x);
                if (lane == 0 && flag) atomicOr(&a.flags[p], flag);
            }
            if (lane == 0) {
                a.blockval[p * bi_bv_ld(a.n_blocks) + bi_bv_index((bin0 >> 5) + kk, a.n_blocks, a.bv_interleaved)] = val[0];
                if (pairs && kk + half < n_blk)
                    a.blockval[p * bi_bv_ld(a.n_blocks) + bi_bv_index((bin0 >> 5) + kk + half, a.n_blocks, a.bv_interleaved)] = val[1];
            }
        }
        __syncthreads();                       // every warp is done with this stage: it may be refilled
        if (NSTAGE == 2) buf ^= 1;
    }
}

template <int MODE, int CT, int NB, int NT, int NSTAGE>
__global__ void __launch_bounds__(NT, 1024 / NT) k_binned_tile(const __grid_constant__ BiBinnedArgs a) {
    extern __shared__ __align__(128) unsigned char bi_bin_smem_dyn[];
    bi_binned_tile_body<MODE, CT, NB, NT, NSTAGE>(a, bi_bin_smem_dyn);
}

// schedule of k_binned_tile for a chunk of <= 1024 points: the evaluable points sorted by hypercube cell (key = anchor
// index of the cell's first corner) and cut into groups of <= BI_BIN_GROUP_POINTS points of one cell.  One CTA, O(P^2)
// rank counting in shared memory (P <= 1024: a few microseconds, no global atomics, deterministic order).
#define BI_BIN_CHUNK_MAX 1024
__global__ void __launch_bounds__(BI_BIN_CHUNK_MAX)
k_binned_plan(const int32_t* __restrict__ corner, const int32_t* __restrict__ status, int C, int n_points,
              int32_t* __restrict__ group_points, int32_t* __restrict__ groups, int32_t* __restrict__ header,
              int32_t* __restrict__ flags, const double* __restrict__ nm_sum_anchor, const double* __restrict__ weight,
              const double* __restrict__ mus, int S, int bb_source, double* __restrict__ pt_const) {
    __shared__ int key[BI_BIN_CHUNK_MAX];
    __shared__ int skey[BI_BIN_CHUNK_MAX];
    __shared__ int gflag[BI_BIN_CHUNK_MAX];
    const int t = threadIdx.x;
    const int none = 0x7fffffff;
    key[t] = (t < n_points && status[t] == 0) ? corner[(int64_t)t * C] : none;
    if (t < n_points) flags[t] = 0;
    if (t < n_points && bb_source >= 0) {
        // n_model_events[source_i].sum(): morph of the per-anchor bin sums; p_cal = mu_i / that (likelihood.py:645)
        const double sum_a = bi_morph_value(nm_sum_anchor, 1, 0, corner + (int64_t)t * C, weight + (int64_t)t * C, C);
        pt_const[4 * t] = sum_a;
        pt_const[4 * t + 1] = __ddiv_rn(mus[(int64_t)t * S + bb_source], sum_a);
        pt_const[4 * t + 2] = 0.0;
        pt_const[4 * t + 3] = 0.0;
    }
    __syncthreads();
    const int my = key[t];
    int rank = 0, run_start = 0;
    for (int j = 0; j < n_points; ++j) {
        const int kj = key[j];
        rank += (kj < my) || (kj == my && j < t);
        run_start += (kj < my);
    }
    if (t < n_points && my != none) { group_points[rank] = t; skey[rank] = my; gflag[rank] = ((rank - run_start) % BI_BIN_GROUP_POINTS) == 0; }
    __syncthreads();
    int n_eval = 0;
    for (int j = 0; j < n_points; ++j) n_eval += key[j] != none;
    if (t < n_eval && gflag[t]) {
        int gi = 0;
        for (int j = 0; j < t; ++j) gi += gflag[j];
        int cnt = 1;
        while (t + cnt < n_eval && !gflag[t + cnt]) ++cnt;
        groups[4 * gi] = t; groups[4 * gi + 1] = cnt; groups[4 * gi + 2] = skey[t]; groups[4 * gi + 3] = 0;
    }
    if (t == 0) {
        int ng = 0;
        for (int j = 0; j < n_eval; ++j) ng += gflag[j];
        header[0] = ng; header[1] = n_eval;
    }
}

// Beeston-Barlow pass B from the terms pass A stored (few points: the evaluation is HBM-bound and re-reading the anchor
// rows would cost more than 16 S bytes per (point, bin)): lambda_b = sum over s in source order of the stored terms, the
// BB source's being (t_b / sum_t) * mu_adj (likelihood.py:657-658,667-670); one warp per (point, 32-bin block)
// the stored terms of the warp's next item travel during the reduction of the current one
__device__ __forceinline__ void bi_binned_passb_stored_body(const BiBinnedArgs& a, double* blockval) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_items = a.n_points * a.n_blocks;
    const int S = a.S, bi = a.bb_source;
    double v[BI_BIN_STORE_S], d = 0.0, lg = 0.0, sum_t = 0.0, mu_adj = 0.0;
    // (point, block) of an item: 32-bit division whenever the item count allows (a 64-bit one costs ~100 instructions,
    // and the kernel is bound by instruction issue)
    const bool small = n_items < (1LL << 31);
    const unsigned nb32 = (unsigned)a.n_blocks;
    auto split = [&](int64_t item, int64_t* p, int64_t* k) {
        if (small) { const unsigned q = (unsigned)item / nb32; *p = q; *k = (unsigned)item - q * nb32; }
        else { *p = item / a.n_blocks; *k = item - *p * a.n_blocks; }
    };
    auto load_item = [&](int64_t item) {
        int64_t p, k;
        split(item < n_items ? item : 0, &p, &k);
        const int64_t b = k * 32 + lane;
        if (item < n_items && b < a.n_bins) {
#pragma unroll
            for (int s = 0; s < BI_BIN_STORE_S; ++s)
                if (s < S) v[s] = a.tbuf[(p * BI_BIN_STORE_S + s) * a.tb_ld + b];
            d = a.observed[p * a.obs_stride + b];
            lg = a.lgamma_obs[p * a.obs_stride + b];
            sum_t = a.pt_const[4 * p + 2];
            mu_adj = a.pt_const[4 * p + 3];                        // sum_t * p_cal (likelihood.py:658)
        }
    };
    load_item(warp_global);
    for (int64_t item = warp_global; item < n_items; item += n_warps) {
        int64_t p, k;
        split(item, &p, &k);
        const int64_t b = k * 32 + lane;
        double val = 0.0;
        const bool live = a.status[p] == 0;
        if (live && b < a.n_bins) {
            double lam = 0.0;
#pragma unroll
            for (int s = 0; s < BI_BIN_STORE_S; ++s) {
                if (s < S) {
                    const double term = (s == bi) ? __dmul_rn(__ddiv_rn(v[s], sum_t), mu_adj) : v[s];
                    lam = (s == 0) ? term : __dadd_rn(lam, term);
                }
            }
            val = bi_poisson_logpmf(d, lg, lam);
        }
        load_item(item + n_warps);
#pragma unroll
        for (int x = 1; x < 32; x <<= 1) val = __dadd_rn(val, __shfl_xor_sync(BI_FULL_MASK, val, x));
        if (live && lane == 0) blockval[p * bi_bv_ld(a.n_blocks) + bi_bv_index(k, a.n_blocks, a.bv_interleaved)] = val;
    }
}
__global__ void __launch_bounds__(256) k_binned_passb_stored(const __grid_constant__ BiBinnedArgs a) {
    bi_binned_passb_stored_body(a, a.blockval);
}

// canonical total from per-block sums: superblock j = ((0 + v_16j) + v_16j+1) + ... (the sequential sum k_binned_pass
// forms in its loop), then the 256-lane strided total of k_canonical_total
// One CTA of 1024 threads per point.  The 16-block superblock sums are formed by all threads at once (16 loads in flight
// each; a minimiser step has ONE point, and 256 threads walking 1563 superblocks one after the other cost 16 us of load
// latency per total), kept in shared memory, and added by lanes 0..255 in the canonical strided order.
#define BI_TOTAL_THREADS 1024
#define BI_TOTAL_SMEM_SUPERS 4096
__global__ void __launch_bounds__(BI_TOTAL_THREADS)
k_canonical_total_blocks(const double* __restrict__ blockval, int64_t n_blocks, int interleaved,
                         const int32_t* __restrict__ status, double fill, double* __restrict__ out, double* pt_const) {
    __shared__ double warp_tot[8];
    __shared__ double s_sj[BI_TOTAL_SMEM_SUPERS];
    const int64_t p = blockIdx.x;
    const int t = threadIdx.x;
    if (status[p] != 0) { if (t == 0) out[p] = fill; return; }
    const double* v = blockval + p * bi_bv_ld(n_blocks);
    const int64_t n_super = (n_blocks + 15) / 16;
    auto super_sum = [&](int64_t j) {                              // ((0 + v_16j) + v_16j+1) + ...; block 16 j + i at i * n_super + j
        const int64_t k0 = j * 16, k1 = min(k0 + 16, n_blocks);
        double sj = 0.0;
        if (k1 - k0 == 16) {                                       // all 16 loads in flight before the sequential adds
            double x[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = interleaved ? v[k * n_super + j] : v[k0 + k];
#pragma unroll
            for (int k = 0; k < 16; ++k) sj = __dadd_rn(sj, x[k]);
        } else {
            for (int64_t k = k0; k < k1; ++k) sj = __dadd_rn(sj, interleaved ? v[(k - k0) * n_super + j] : v[k]);
        }
        return sj;
    };
    double u = 0.0;
    for (int64_t base = 0; base < n_super; base += BI_TOTAL_SMEM_SUPERS) {       // BI_TOTAL_SMEM_SUPERS is a multiple of 256
        const int64_t n_here = min((int64_t)BI_TOTAL_SMEM_SUPERS, n_super - base);
        for (int64_t j = t; j < n_here; j += BI_TOTAL_THREADS) s_sj[j] = super_sum(base + j);
        __syncthreads();
        if (t < 256)
            for (int64_t j = t; j < n_here; j += 256) u = __dadd_rn(u, s_sj[j]);
        __syncthreads();
    }
    if (t < 256) {
#pragma unroll
        for (int x = 1; x < 32; x <<= 1) u = __dadd_rn(u, __shfl_xor_sync(BI_FULL_MASK, u, x));
        if ((t & 31) == 0) warp_tot[t >> 5] = u;
    }
    __syncthreads();
    if (t == 0) {
        const double total = __dadd_rn(__dadd_rn(__dadd_rn(warp_tot[0], warp_tot[1]), __dadd_rn(warp_tot[2], warp_tot[3])),
                                       __dadd_rn(__dadd_rn(warp_tot[4], warp_tot[5]), __dadd_rn(warp_tot[6], warp_tot[7])));
        out[p] = total;
        if (pt_const) {                                            // after pass A: sum_t and mu_adj = sum_t * p_cal (:658)
            pt_const[4 * p + 2] = total;
            pt_const[4 * p + 3] = __dmul_rn(total, pt_const[4 * p + 1]);
        }
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
// points evaluated per pass: bounds the per-bin scratch (t_b of the Beeston-Barlow pass) to 2 GiB and fits the planner
static int64_t bi_bin_chunk_points(int64_t n_points, int64_t n_bins) {
    const int64_t tb_ld = (n_bins + 63) / 64 * 64;
    int64_t pc = ((int64_t)1 << 28) / tb_ld;
    if (pc > BI_BIN_CHUNK_MAX) pc = BI_BIN_CHUNK_MAX;
    if (pc < BI_BIN_GROUP_POINTS) pc = BI_BIN_GROUP_POINTS;
    return n_points < pc ? n_points : pc;
}

// stored-terms form of the two Beeston-Barlow passes (few points, HBM-bound): all points in one pass, S <= 8 terms kept
static bool bi_bin_store_terms(int64_t n_points, int64_t n_bins) {
    const int64_t tb_ld = (n_bins + 63) / 64 * 64;
    const char* e = getenv("BI_BINNED_STORE_LOG2");                  // experiments: scratch limit of the stored-terms form
    const int lg = e ? atoi(e) : 26;
    return n_points * BI_BIN_STORE_S * tb_ld <= ((int64_t)1 << lg);
}

// scratch layout (doubles): sum_t [P] | point constants [PC, 4] | block sums A [PC, n_blocks] | block sums B [PC, n_blocks] | t_b [PC, tb_ld] (or the
// stored terms [P, 8, tb_ld]) | schedule (int32: group_points [PC], groups [(PC + 1) * 4], header [8]); PC = points per pass
extern "C" int64_t bi_binned_scratch_doubles(int64_t n_points, int64_t n_bins) {
    if (n_points <= 0 || n_bins <= 0) return 0;
    const int64_t pc = bi_bin_chunk_points(n_points, n_bins);
    const int64_t n_blocks = (n_bins + 31) / 32, tb_ld = (n_bins + 63) / 64 * 64;
    const int64_t per_bin = bi_bin_store_terms(n_points, n_bins) ? n_points * BI_BIN_STORE_S * tb_ld : pc * tb_ld;
    return n_points + 4 * pc + 2 * pc * bi_bv_ld(n_blocks) + per_bin + (5 * pc + 12 + 1) / 2 + 8;
}
extern "C" int64_t bi_binned_sum_t_offset(int64_t n_points, int64_t n_bins) { (void)n_points; (void)n_bins; return 0; }

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

template <int MODE, int NB, int NT, int NSTAGE>
static int bi_binned_launch_tile_nb(const BiBinnedArgs& a, int smem, int grid, cudaStream_t st) {
#define BI_BIN_LAUNCH(CT)                                                                                          \
    do {                                                                                                           \
        BI_CUDA_CHECK(cudaFuncSetAttribute(k_binned_tile<MODE, CT, NB, NT, NSTAGE>,                                \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                    \
        k_binned_tile<MODE, CT, NB, NT, NSTAGE><<<grid, NT, smem, st>>>(a);                                        \
    } while (0)
    switch (a.C) {
        case 1: BI_BIN_LAUNCH(1); break;
        case 2: BI_BIN_LAUNCH(2); break;
        case 4: BI_BIN_LAUNCH(4); break;
        case 8: BI_BIN_LAUNCH(8); break;
        default: BI_BIN_LAUNCH(0); break;
    }
#undef BI_BIN_LAUNCH
    return BI_OK;
}

// 512-thread CTAs, two per SM (64 registers), one stage each: while one CTA waits for its tile the other evaluates.
// Measured on B200 (config 3, 256-point scan / single point): this shape 9.9 ms / 0.20 ms; a two-stage prefetch pipeline in
// one 1024-thread CTA per SM 11.5 / 0.26; 256-thread CTAs 12.7 / 0.19; 128-bin tiles 10.8 / 0.25.  Round 2, single point
// with the launch sequence replayed as a graph (0.097 ms): two-stage variants (128 x 128, 256 x 256, 128 x 256 threads)
// 0.127-0.129 ms; one stage of 128 bins with 128 / 256 threads and up to 8 / 4 CTAs per SM 0.110 / 0.096 ms.
template <int MODE>
static int bi_binned_launch_tile(const BiBinnedArgs& a, int smem, int grid, cudaStream_t st) {
    if (a.tile == 128) return bi_binned_launch_tile_nb<MODE, 128, 512, 1>(a, smem, grid, st);
    // a handful of points: the evaluation streams the tensors once (HBM-bound) and a tile holds 8 (point, block) items
    // at most per point -- 256-thread CTAs measured faster there (pass A at P = 1: 63 us against 77 us)
    if (a.n_points <= 8) return bi_binned_launch_tile_nb<MODE, 256, 256, 1>(a, smem, grid, st);
    return bi_binned_launch_tile_nb<MODE, 256, 512, 1>(a, smem, grid, st);
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
    BiBinnedArgs a0;
    int rc = bi_binned_fill(&a0, pmf_anchor_dev, n_model_anchor_dev, n_model_sum_anchor_dev, ld_bins, n_bins, n_sources,
                            n_corners, bb_source, observed_dev, lgamma_obs_dev, corner_dev, weight_dev, mus_dev,
                            status_dev, n_points);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(lgamma_obs_dev && scratch_dev && logl_dev && flags_dev, "bi_binned_ll_batch: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int S = n_sources, C = n_corners;
    const int64_t pc = bi_bin_chunk_points(n_points, n_bins);
    const int64_t n_blocks = (n_bins + 31) / 32, tb_ld = (n_bins + 63) / 64 * 64;
    const bool store_terms = bb_source >= 0 && S <= BI_BIN_STORE_S && bi_bin_store_terms(n_points, n_bins) && pc == n_points;
    double* sum_t_all = scratch_dev;
    double* pt_const = scratch_dev + n_points;
    double* bva = pt_const + 4 * pc;
    double* bvb = bva + pc * bi_bv_ld(n_blocks);
    double* tbuf = bvb + pc * bi_bv_ld(n_blocks);
    int32_t* plan = reinterpret_cast<int32_t*>(
        tbuf + (bi_bin_store_terms(n_points, n_bins) ? n_points * BI_BIN_STORE_S * tb_ld : pc * tb_ld));
    int32_t* group_points = plan;
    int32_t* groups = plan + ((pc + 3) & ~(int64_t)3);
    int32_t* header = groups + 4 * (pc + 1);
    // tiled path: the staged rows of a tile must fit in shared memory, every row segment must be a valid bulk copy
    const char* legacy_str = getenv("BI_BINNED_LEGACY");
    const bool legacy_env = legacy_str != nullptr && legacy_str[0] == '1';
    const char* tile_str = getenv("BI_BINNED_TILE");                         // experiments: force 128 / 256
    int tile = tile_str ? atoi(tile_str) : 256;                              // measured: 256-bin tiles win at P = 1 and at P = 256
    if (tile != 128 && tile != 256) tile = 256;
    if (bi_bin_stage_bytes(C, S, S + 1, tile) > 110 * 1024) tile = 128;      // two CTAs per SM
    const int smem_full = bi_bin_stage_bytes(C, S, S + 1, tile);
    const bool tiled = !legacy_env && smem_full <= 220 * 1024 && ld_bins % 2 == 0 &&
                       ((uintptr_t)pmf_anchor_dev & 15) == 0 && (bb_source < 0 || ((uintptr_t)n_model_anchor_dev & 15) == 0);
    if (!tiled) BI_CUDA_CHECK(cudaMemsetAsync(flags_dev, 0, sizeof(int32_t) * n_points, st));
    const double ninf = -INFINITY;
    for (int64_t p0 = 0; p0 < n_points; p0 += pc) {
        const int64_t np = n_points - p0 < pc ? n_points - p0 : pc;
        BiBinnedArgs a = a0;
        a.n_points = np;
        a.corner = corner_dev + p0 * C; a.weight = weight_dev + p0 * C; a.mus = mus_dev + p0 * S;
        a.status = status_dev + p0; a.flags = flags_dev + p0;
        a.observed = observed_dev + p0 * observed_stride; a.lgamma_obs = lgamma_obs_dev + p0 * observed_stride;
        a.obs_stride = observed_stride;
        double* sum_t = sum_t_all + p0;
        if (tiled) {
            a.group_points = group_points; a.groups = groups; a.header = header;
            a.tbuf = tbuf; a.n_blocks = n_blocks; a.tile = tile; a.n_tiles = (n_bins + tile - 1) / tile;
            a.tb_ld = tb_ld;
            a.store_terms = store_terms ? 1 : 0;
            a.bv_interleaved = np <= BI_BIN_GROUP_POINTS ? 1 : 0;
            a.pt_const = pt_const;
            k_binned_plan<<<1, BI_BIN_CHUNK_MAX, 0, st>>>(a.corner, a.status, C, (int)np, group_points, groups, header, a.flags,
                                                         a.nm_sum_anchor, a.weight, a.mus, S, bb_source, pt_const);
            auto launch = [&](int mode, double* blockval) -> int {
                a.blockval = blockval;
                const int n_kinds = mode == 0 ? S : (mode == 1 ? S + 1 : S - 1);
                const int smem = bi_bin_stage_bytes(C, S, n_kinds, tile);
                const int per_sm = 2 * (smem + 1024) <= 227 * 1024 ? 2 : 1;
                const int grid = 148 * per_sm;
                if (mode == 0) return bi_binned_launch_tile<0>(a, smem, grid, st);
                if (mode == 1) return bi_binned_launch_tile<1>(a, smem, grid, st);
                return bi_binned_launch_tile<2>(a, smem, grid, st);
            };
            if (bb_source < 0) {
                if ((rc = launch(0, bvb)) != BI_OK) return rc;
            } else {
                if ((rc = launch(1, bva)) != BI_OK) return rc;
                k_canonical_total_blocks<<<(unsigned)np, BI_TOTAL_THREADS, 0, st>>>(bva, n_blocks, a.bv_interleaved, a.status, 0.0, sum_t, pt_const);
                a.sum_t = sum_t;
                if (store_terms) {
                    a.blockval = bvb;
                    int64_t blocks = (np * n_blocks + 7) / 8;
                    if (blocks > 148 * 8) blocks = 148 * 8;         // one wave of 8 CTAs per SM
                    k_binned_passb_stored<<<(unsigned)blocks, 256, 0, st>>>(a);
                } else if ((rc = launch(2, bvb)) != BI_OK) {
                    return rc;
                }
            }
            k_canonical_total_blocks<<<(unsigned)np, BI_TOTAL_THREADS, 0, st>>>(bvb, n_blocks, a.bv_interleaved, a.status, ninf, logl_dev + p0, nullptr);
        } else {
            const int64_t n_tasks = np * a.n_chunks;
            int64_t blocks = (n_tasks + 7) / 8;
            if (blocks > 148 * 8 * 4) blocks = 148 * 8 * 4;
            if (bb_source < 0) {
                a.partial = bvb;
                k_binned_pass<0><<<(unsigned)blocks, 256, 0, st>>>(a);
            } else {
                a.partial = bva;
                k_binned_pass<1><<<(unsigned)blocks, 256, 0, st>>>(a);
                k_canonical_total<<<(unsigned)np, 256, 0, st>>>(bva, a.n_chunks, a.status, 0.0, sum_t);
                a.partial = bvb;
                a.sum_t = sum_t;
                k_binned_pass<2><<<(unsigned)blocks, 256, 0, st>>>(a);
            }
            k_canonical_total<<<(unsigned)np, 256, 0, st>>>(bvb, a.n_chunks, a.status, ninf, logl_dev + p0);
        }
    }
    if (mus_adj_dev) {
        const int64_t nb = (n_points + 127) / 128;
        k_binned_mus_adj<<<(unsigned)nb, 128, 0, st>>>(mus_dev, sum_t_all, n_model_sum_anchor_dev, corner_dev, weight_dev,
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
