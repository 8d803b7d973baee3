// blueice_b200 -- K5: template-space unbinned likelihood (fused template lookup + morph + mixture + log + reduce)
// over MANY datasets (toy Monte Carlos) or one dataset too large for a dense anchor tensor.
//
// The anchor-tensor form (K3 -> K2) materialises A[G, S, N] = HistogramPdfSource.pdf of every anchor model at every
// event (blueice/likelihood.py:557-562) and contracts it per point.  That tensor is 3 TB for 1e6 toys x 1e3 events x
// 125 anchors x 3 sources and for 1e8 events x 625 anchors x 6 sources.  Here the per-event values are never stored:
//     A[row, i] = lookup(template[row], x_i)          -- source.py:219-246, scipy's operation order, as in K3
//     f_i       = fma chain over k of A[row_k, i] * coef_k      -- as in K2 (bi_unbinned_mma.cuh)
// evaluated on the fly from the L2-resident templates [n_rows, n_bins] and the prepared events
// (low-corner bin + per-dimension fractions, computed once per dataset by bi_template_prepare_events).  Every
// intermediate is formed by the same operations in the same order as K3 + K2, and the log-sum goes through the same
// canonical product tree, so a (dataset, point) pair evaluates BIT-IDENTICALLY to
// set_data(dataset) + ll(point) on the anchor-tensor engine.
//
// Work: a pair group = up to BI_TS_GROUP_POINTS points evaluated on one dataset that share their row list (same
// hypercube cell): the template values of an event are gathered once and contracted with every point of the group.
// A unit = (pair group, superblock of 512 events); one warp per unit, lane = event.
#include <stdlib.h>

#include "bi_space.cuh"
#include "bi_plan.cuh"
#include "bi_ts.cuh"


// ---------------------------------------------------------------------------------------------
// event preparation (once per dataset): coordinates -> low-corner bin + fractions
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_template_prepare(const __grid_constant__ BiSpace sp, const __grid_constant__ BiPoints pts, int n_points_total, int linear,
                   const double* __restrict__ coords, int64_t ld_coords, int64_t n_events,
                   int32_t* __restrict__ ev_bin, double* __restrict__ ev_frac, int64_t ld_frac) {
    __shared__ double s_pts[BI_MAX_EDGE_POINTS];
    bi_stage_points(pts, n_points_total, s_pts);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_events) return;
    if (linear) {
        int cell[BI_MAX_SPACE_DIMS];
        double y[BI_MAX_SPACE_DIMS];
        ev_bin[i] = bi_event_cell_linear(sp, s_pts, coords, ld_coords, i, cell, y);
        for (int d = 0; d < sp.n_space; ++d) ev_frac[(int64_t)d * ld_frac + i] = y[d];
    } else {
        ev_bin[i] = bi_event_bin_piecewise(sp, s_pts, coords, ld_coords, i);
    }
}


// ---------------------------------------------------------------------------------------------
// the kernel: one warp per unit (pair group, superblock), lane = event
// ---------------------------------------------------------------------------------------------
template <int NP, int NS, bool PRE>
__global__ void __launch_bounds__(BI_TS_THREADS)
k_template_partials(const double* __restrict__ T, int64_t row_stride, int64_t bin_stride, const __grid_constant__ BiTsSpace sp,
                    const int32_t* __restrict__ ev_bin, const double* __restrict__ ev_frac, int64_t ld_frac,
                    const int64_t* __restrict__ dataset_offset, int K, int S,
                    const int32_t* __restrict__ row, const double* __restrict__ coef, const double* __restrict__ wterm,
                    const int32_t* __restrict__ term_source, const double* __restrict__ mus,
                    const int32_t* __restrict__ status,
                    int64_t n_groups, const BiTsGroup* __restrict__ groups, const int64_t* __restrict__ unit_offset,
                    const int32_t* __restrict__ unit_group, int64_t n_units,
                    const int32_t* __restrict__ pair_point, const int64_t* __restrict__ pair_partial_offset,
                    double outlier, double* __restrict__ partial,
                    const int32_t* __restrict__ group_order, const int32_t* __restrict__ n_ordered, int sb_max,
                    const double* __restrict__ pre, const int32_t* __restrict__ pre_corner,
                    const double* __restrict__ pre_weight) {
    // pre (NP = 1 only): densities p_i already formed by the bin-major pass (bi_template_bm.cu), in event order -- the
    // gather + contraction loop is skipped, the range test, the canonical tree and the rare path are the code below.
    // The rows / weights of the rare path then come from K1's corner / weight outputs [P, C] (pre_corner, pre_weight:
    // row_k = corner[k / S] * S + k % S, wterm_k = weight[k / S]) when a unit first needs them; row / coef / wterm and
    // the prepared events are not read otherwise.
    constexpr bool use_pre = PRE;
    constexpr int NY = NS > 0 ? NS : 1;
    extern __shared__ __align__(16) unsigned char bi_ts_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per warp: row offsets [K] (int64) + coefficients [NP][K]
    int64_t* rowoff = reinterpret_cast<int64_t*>(bi_ts_smem) + (size_t)warp * (size_t)K * (1 + NP);
    double* coef_s = reinterpret_cast<double*>(rowoff + K);
    const int t_class = (lane >> 1) & 3;
    const int64_t n_warps = (int64_t)gridDim.x * BI_TS_WARPS;

    // ordered mode (toy Monte Carlos): the groups are walked in hypercube-cell order (group_order, written by
    // k_ts_order_groups), every group's superblocks back to back -- co-resident warps then gather from the SAME few
    // template rows, which stay in L2 (in toy order the 120 MB of packed templates thrash it: 7x the algorithmic DRAM bytes)
    const int64_t n_iter = group_order ? (int64_t)n_ordered[0] * sb_max : n_units;
    for (int64_t u = (int64_t)blockIdx.x * BI_TS_WARPS + warp; u < n_iter; u += n_warps) {
        // ---- unit -> (pair group, superblock)
        int64_t g, sb;
        if (group_order) {
            g = group_order[u / sb_max];
            sb = u - (u / sb_max) * sb_max;
            if (sb >= unit_offset[g + 1] - unit_offset[g]) continue;
        } else {
            if (unit_group) g = unit_group[u];
            else {                                                  // largest g with unit_offset[g] <= u
                int64_t lo = 0, hi = n_groups;
                while (hi - lo > 1) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (unit_offset[mid] <= u) lo = mid; else hi = mid;
                }
                g = lo;
            }
            sb = u - unit_offset[g];
        }
        const BiTsGroup gp = groups[g];
        const int np = gp.count < NP ? gp.count : NP;
        // live points of the group (status 0); the row list comes from the first live one
        int32_t point[NP];
        unsigned live = 0;
        int lead = -1;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            point[q] = q < np ? pair_point[gp.first + q] : 0;
            if (q < np && status[point[q]] == 0) {
                live |= 1u << q;
                if (lead < 0) lead = q;
            }
        }
        if (!live) continue;                                        // their results are -inf (finalize)
        __syncwarp();
        if (!use_pre) {
            for (int k = lane; k < K; k += 32) {
                rowoff[k] = (int64_t)row[(int64_t)point[lead] * K + k] * row_stride;
#pragma unroll
                for (int q = 0; q < NP; ++q) coef_s[q * K + k] = coef[(int64_t)point[(live >> q) & 1u ? q : lead] * K + k];
            }
        }
        __syncwarp();
        bool rows_ready = !use_pre;

        const int64_t ev_begin = dataset_offset[gp.dataset] + sb * BI_SUPERBLOCK;
        const int64_t left = dataset_offset[gp.dataset + 1] - ev_begin;
        const int n_ev = left < BI_SUPERBLOCK ? (int)left : BI_SUPERBLOCK;

        double M[NP], Lslow[NP];
        int E[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) { M[q] = 1.0; E[q] = 0; Lslow[q] = 0.0; }
        bool any_slow = false;

        // pre mode: the superblock's densities are all in flight before the tree starts (16 coalesced loads per lane)
        double pp[BI_SUPERBLOCK / BI_EVENT_BLOCK];
        if (use_pre) {
#pragma unroll
            for (int i = 0; i < BI_SUPERBLOCK / BI_EVENT_BLOCK; ++i)
                pp[i] = (i * BI_EVENT_BLOCK + lane < n_ev) ? pre[ev_begin + i * BI_EVENT_BLOCK + lane] : 1.0;
        }
#pragma unroll 1
        for (int e0 = 0; e0 < n_ev; e0 += BI_EVENT_BLOCK) {
            const bool valid = e0 + lane < n_ev;
            const int64_t ev = ev_begin + e0 + lane;
            int64_t base = 0;
            double y[NY];
#pragma unroll
            for (int d = 0; d < NY; ++d) y[d] = 0.0;
            if (!use_pre) {
                base = valid ? (int64_t)ev_bin[ev] * bin_stride : 0;
#pragma unroll
                for (int d = 0; d < NY; ++d) y[d] = (NS > 0 && valid) ? ev_frac[(int64_t)d * ld_frac + ev] : 0.0;
            }

            double p[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) p[q] = 0.0;
            // BI_TS_BATCH rows at a time: all their gathers are issued before the first value is used (the kernel is
            // bound by the latency / throughput of scattered L2 loads); the contraction keeps the term order
            constexpr int KB = (1 << NS) <= 4 ? BI_TS_BATCH : (BI_TS_BATCH / 2 > 0 ? BI_TS_BATCH / 2 : 1);
            int k = 0;
            if (use_pre) {
                p[0] = 1.0;
#pragma unroll
                for (int i = 0; i < BI_SUPERBLOCK / BI_EVENT_BLOCK; ++i)
                    if (i * BI_EVENT_BLOCK == e0) p[0] = pp[i];
                k = K;
            }
#pragma unroll 1
            for (; k + KB <= K; k += KB) {
                double v[KB][1 << NS];
#pragma unroll
                for (int i = 0; i < KB; ++i) bi_ts_gather<NS>(T + rowoff[k + i] + base, sp, v[i]);
#pragma unroll
                for (int i = 0; i < KB; ++i) {
                    const double r = bi_ts_eval<NS>(v[i], y);
#pragma unroll
                    for (int q = 0; q < NP; ++q) p[q] = fma(r, coef_s[q * K + k + i], p[q]);
                }
            }
#pragma unroll 1
            for (; k < K; ++k) {
                const double r = bi_ts_lookup<NS>(T + rowoff[k] + base, sp, y);
#pragma unroll
                for (int q = 0; q < NP; ++q) p[q] = fma(r, coef_s[q * K + k], p[q]);
            }
            // range test + canonical tree: pair (events 2j, 2j+1) -> quad (octets 0,1 / 2,3) -> oct; lanes of one class
            // agree.  Every tree level runs over all points before the next one (no serialisation behind the shuffles).
            const unsigned class_mask = 0x03030303u << (2 * t_class);       // lanes of class t: 8n + 2t + {0, 1}
            unsigned bad_any = 0;
            unsigned bad[NP];
            double v[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                if (!valid) p[q] = 1.0;                             // events >= N count as p = 1
                const bool in_range = (unsigned)(__double2hiint(p[q]) - BI_RANGE_LO) < BI_RANGE_SPAN;
                bad[q] = ~__ballot_sync(BI_FULL_MASK, in_range);
                if ((live >> q) & 1u) bad_any |= bad[q];
            }
#pragma unroll
            for (int q = 0; q < NP; ++q) v[q] = __dmul_rn(p[q], __shfl_xor_sync(BI_FULL_MASK, p[q], 1));
#pragma unroll
            for (int q = 0; q < NP; ++q) v[q] = __dmul_rn(v[q], __shfl_xor_sync(BI_FULL_MASK, v[q], 8));
#pragma unroll
            for (int q = 0; q < NP; ++q) v[q] = __dmul_rn(v[q], __shfl_xor_sync(BI_FULL_MASK, v[q], 16));
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                double m;
                int e;
                bi_split(v[q], &m, &e);
                if (bad[q] & class_mask) { m = 1.0; e = 0; }
                M[q] = __dmul_rn(M[q], m);
                E[q] += e;
            }
            if (bad_any) {                                          // warp-uniform; rare: reference-semantics fallback
                if (use_pre) {
                    if (!rows_ready) {
                        const int64_t pt = point[0];
                        const int C = K / S;
                        for (int k2 = lane; k2 < K; k2 += 32) {
                            rowoff[k2] = ((int64_t)pre_corner[pt * C + k2 / S] * S + k2 % S) * row_stride;
                            coef_s[k2] = pre_weight[pt * C + k2 / S];
                        }
                        __syncwarp();
                        rows_ready = true;
                    }
                    base = valid ? (int64_t)ev_bin[ev] * bin_stride : 0;
#pragma unroll
                    for (int d = 0; d < NY; ++d) y[d] = (NS > 0 && valid) ? ev_frac[(int64_t)d * ld_frac + ev] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    if (bad[q] && ((live >> q) & 1u)) {
                        const bool class_bad = (bad[q] & class_mask) != 0;
                        double l = 0.0;
                        if (class_bad && valid) {
                            const int64_t pt = point[q];
                            l = log(bi_ts_slow_density<NS>(T, rowoff, base, sp, y, K, S, term_source,
                                                           use_pre ? coef_s : wterm + pt * K, mus + pt * S, outlier));
                        }
                        l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 1));
                        l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 8));
                        l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 16));
                        if (class_bad) Lslow[q] = __dadd_rn(Lslow[q], l);
                        any_slow = true;
                    }
                }
            }
        }
        // ---- close the superblock: M = (M_0 * M_1) * (M_2 * M_3), E = sum, L = (L_0 + L_1) + (L_2 + L_3)
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            double m = M[q];
            m = __dmul_rn(m, __shfl_xor_sync(BI_FULL_MASK, m, 2));
            m = __dmul_rn(m, __shfl_xor_sync(BI_FULL_MASK, m, 4));
            int e = E[q];
            e += __shfl_xor_sync(BI_FULL_MASK, e, 2);
            e += __shfl_xor_sync(BI_FULL_MASK, e, 4);
            double L = bi_block_log(m, e);
            if (any_slow) {
                double l = Lslow[q];
                l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 2));
                l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 4));
                L = __dadd_rn(L, l);
            }
            if (lane == 0 && ((live >> q) & 1u)) partial[pair_partial_offset[gp.first + q] + sb] = L;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K5b, mixture form: morph the TEMPLATES per point, then look the events up in the morphed mixture template.
//     Tmix_q[b] = fma chain over k of T[row_{q,k}, b] * coef_{q,k}          (k_template_mix, once per point)
//     f_i       = lookup(Tmix_q, x_i)                                        (k_mixture_partials)
// The lookup is linear in the template values, so this equals the reference's sum_s mu_s * sum_c w_c * lookup(T_{c,s}, x_i)
// in exact arithmetic; in floating point the operation order differs (about 1e-16 relative per event, far inside the
// 1e-9 * N contract) -- this family is deterministic and batch-shape independent, but not bit-identical to K2 / K5a.
// Cost per point-event: one lookup instead of K; with events sorted by bin (the engine does that once per dataset)
// the 2^n_space template loads are warp-uniform and the kernel streams the prepared events at HBM speed.
// Requires finite templates (the reference's per-source nansum cannot be reproduced from a mixture); a density that is
// not a normal positive number takes  p = (outlier != 0 && !(p > 0)) ? outlier : p  (likelihood.py:686-689).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_template_mix(const double* __restrict__ T, int64_t row_stride, int64_t bin_stride, int64_t n_bins, int K,
               const int32_t* __restrict__ row, const double* __restrict__ coef, const int32_t* __restrict__ status,
               const int32_t* __restrict__ pair_point, int64_t n_pairs, double* __restrict__ tmix,
               int pack, int64_t off1, int64_t off2) {
  for (int64_t q = blockIdx.y; q < n_pairs; q += gridDim.y) {
    const int64_t pt = pair_point ? pair_point[q] : q;
    if (status[pt] != 0) continue;
    const int32_t* rw = row + pt * K;
    const double* cf = coef + pt * K;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_bins; b += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        int k = 0;
        for (; k + 32 <= K; k += 32) {                               // 32 independent loads in flight, then the chain
            double t[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __ldg(T + (int64_t)rw[k + i] * row_stride + b * bin_stride);
#pragma unroll
            for (int i = 0; i < 32; ++i) acc = fma(t[i], cf[k + i], acc);
        }
        for (; k + 8 <= K; k += 8) {
            double t[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = __ldg(T + (int64_t)rw[k + i] * row_stride + b * bin_stride);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc = fma(t[i], cf[k + i], acc);
        }
        for (; k < K; ++k) acc = fma(__ldg(T + (int64_t)rw[k] * row_stride + b * bin_stride), cf[k], acc);
        // packed output (as K5's templates): element (q, b) = (m[b], m[b + off1], m[b + off2], m[b + off2 + off1]);
        // every bin stores its value into the (up to four) elements it belongs to
        double* dst = tmix + q * n_bins * pack;
        dst[b * pack] = acc;
        if (pack >= 2 && b - off1 >= 0) dst[(b - off1) * pack + 1] = acc;
        if (pack == 4) {
            if (b - off2 >= 0) dst[(b - off2) * 4 + 2] = acc;
            if (b - off2 - off1 >= 0) dst[(b - off2 - off1) * 4 + 3] = acc;
        }
    }
  }
}

// lookup in a mixture template with PRE-MULTIPLIED corner weights (this family's own operation order):
//     w_c = ((1 * t_0) * t_1) ...  per event, r = fma chain over the corners c of V[c] * w_c  (all terms positive)
template <int NS>
__device__ __forceinline__ void bi_mix_weights(const double (&y)[NS > 0 ? NS : 1], double (&w)[1 << NS]) {
#pragma unroll
    for (int c = 0; c < (1 << NS); ++c) {
        double v = 1.0;
#pragma unroll
        for (int d = 0; d < NS; ++d) v = __dmul_rn(v, ((c >> (NS - 1 - d)) & 1) ? y[d] : __dsub_rn(1.0, y[d]));
        w[c] = v;
    }
}
// the packed element(s) of the event's low-corner bin: one wide load per four lookup corners
template <int NS>
__device__ __forceinline__ void bi_mix_gather(const double* __restrict__ V, const BiTsSpace& sp, double (&v)[1 << NS]) {
    if constexpr (NS == 0) {
        v[0] = __ldg(V);
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
// corners c ascending: r = v[0] * w[0], then fma
template <int NS>
__device__ __forceinline__ double bi_mix_eval(const double (&v)[1 << NS], const double (&w)[1 << NS]) {
    double r = __dmul_rn(v[0], w[0]);
#pragma unroll
    for (int c = 1; c < (1 << NS); ++c) r = fma(v[c], w[c], r);
    return r;
}
template <int NS>
__device__ __forceinline__ double bi_mix_lookup(const double* __restrict__ V, const BiTsSpace& sp, const double (&w)[1 << NS]) {
    double v[1 << NS];
    bi_mix_gather<NS>(V, sp, v);
    return bi_mix_eval<NS>(v, w);
}

// One canonical group (32 events) of one superblock per HALF-WARP: lane l16 of the half owns the adjacent events
// 2 l16, 2 l16 + 1 (one canonical pair), so the pair product is formed in the lane and the tree needs two shuffle
// levels; the two halves of a warp walk two consecutive superblocks in lock step.
// FULL: both superblocks hold 512 events (no masking).
template <int NP, int NS, bool FULL, bool FG>
__device__ __forceinline__ void bi_mix_group(const double* __restrict__ Vbase, int64_t n_bins, int np, const BiTsSpace& sp,
                                             const int (&bin)[2], const double (&y)[2][NS > 0 ? NS : 1], int n_left, int l16,
                                             unsigned half_shift, unsigned live, double outlier, double (&M)[NP], int (&E)[NP],
                                             double (&Lslow)[NP], bool& any_slow) {
    // FG: the group is full (np == NP): no guards, one basic block.
    // Vbase: mixture template of the group's first pair (pair q: Vbase + q * n_bins); np <= NP pairs in the group (slots
    // q >= np are skipped; a dead pair's row is unwritten memory whose results are ignored);
    // n_left: events of this half's superblock from this group's first event on (may be <= 0)
    constexpr int PACK = NS == 0 ? 1 : (NS == 1 ? 2 : 4);
    const int64_t e0 = (int64_t)bin[0] * PACK, e1 = (int64_t)bin[1] * PACK;      // packed elements of the two events
    const bool valid0 = FULL || 2 * l16 < n_left, valid1 = FULL || 2 * l16 + 1 < n_left;
    double w0[1 << NS], w1[1 << NS];
    bi_mix_weights<NS>(y[0], w0);
    bi_mix_weights<NS>(y[1], w1);
    double v[NP];
    unsigned class_bad_mask = 0, bad_any = 0;                              // bit q: this lane's class of pair q left the fast range
    const unsigned class_mask = (0x1111u << (l16 & 3)) << half_shift;     // lanes of class t in this half
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        if (FG || q < np) {
            const double* V = Vbase + (int64_t)q * n_bins;
            double p0 = bi_mix_lookup<NS>(V + e0, sp, w0);
            double p1 = bi_mix_lookup<NS>(V + e1, sp, w1);
            if (!FULL) {
                if (!valid0) p0 = 1.0;                                     // events >= N count as p = 1
                if (!valid1) p1 = 1.0;
            }
            const bool ok = ((unsigned)(__double2hiint(p0) - BI_RANGE_LO) < BI_RANGE_SPAN) &&
                            ((unsigned)(__double2hiint(p1) - BI_RANGE_LO) < BI_RANGE_SPAN);
            const unsigned bad = ~__ballot_sync(BI_FULL_MASK, ok);
            if ((live >> q) & 1u) bad_any |= bad;
            if (bad & class_mask) class_bad_mask |= 1u << q;
            v[q] = __dmul_rn(p0, p1);                                      // pair
        }
    }
#pragma unroll
    for (int q = 0; q < NP; ++q)
        if (FG || q < np) v[q] = __dmul_rn(v[q], __shfl_xor_sync(BI_FULL_MASK, v[q], 4));    // quad: octets (0,1), (2,3)
#pragma unroll
    for (int q = 0; q < NP; ++q)
        if (FG || q < np) v[q] = __dmul_rn(v[q], __shfl_xor_sync(BI_FULL_MASK, v[q], 8));    // oct
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        if (FG || q < np) {
            double m;
            int e;
            bi_split(v[q], &m, &e);
            if ((class_bad_mask >> q) & 1u) { m = 1.0; e = 0; }
            M[q] = __dmul_rn(M[q], m);
            E[q] += e;
        }
    }
    if (bad_any) {                                                         // rare, warp-uniform: the densities are formed again
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            if ((FG || q < np) && ((live >> q) & 1u)) {
                const double* V = Vbase + (int64_t)q * n_bins;
                double p0 = bi_mix_lookup<NS>(V + e0, sp, w0);
                double p1 = bi_mix_lookup<NS>(V + e1, sp, w1);
                if (!valid0) p0 = 1.0;
                if (!valid1) p1 = 1.0;
                const bool ok = ((unsigned)(__double2hiint(p0) - BI_RANGE_LO) < BI_RANGE_SPAN) &&
                                ((unsigned)(__double2hiint(p1) - BI_RANGE_LO) < BI_RANGE_SPAN);
                const unsigned bad = ~__ballot_sync(BI_FULL_MASK, ok);
                if (bad) {
                    const bool class_bad = (bad & class_mask) != 0;
                    double l = 0.0;
                    if (class_bad) l = __dadd_rn(log(bi_fix_density(p0, outlier)), log(bi_fix_density(p1, outlier)));
                    l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 4));
                    l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 8));
                    if (class_bad) Lslow[q] = __dadd_rn(Lslow[q], l);
                    any_slow = true;
                }
            }
        }
    }
}

// prepared events of one group for this lane: two adjacent events (vector loads when the pair is 8 / 16-byte aligned)
template <int NS>
__device__ __forceinline__ void bi_mix_load(const int32_t* __restrict__ ev_bin, const double* __restrict__ ev_frac,
                                            int64_t ld_frac, int64_t ev, int n_left2, bool vec, int (&bin)[2],
                                            double (&y)[2][NS > 0 ? NS : 1]) {
    // ev: this lane's first event; n_left2: events left from ev on (<= 0: none)
    if (vec && n_left2 >= 2) {
        const int2 b = *reinterpret_cast<const int2*>(ev_bin + ev);
        bin[0] = b.x; bin[1] = b.y;
#pragma unroll
        for (int d = 0; d < NS; ++d) {
            const double2 f = *reinterpret_cast<const double2*>(ev_frac + (int64_t)d * ld_frac + ev);
            y[0][d] = f.x; y[1][d] = f.y;
        }
    } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const bool ok = h < n_left2;
            bin[h] = ok ? ev_bin[ev + h] : 0;
#pragma unroll
            for (int d = 0; d < NS; ++d) y[h][d] = ok ? ev_frac[(int64_t)d * ld_frac + ev + h] : 0.0;
        }
    }
    if (NS == 0) y[0][0] = y[1][0] = 0.0;
}

// unit = (pair group, PAIR of consecutive superblocks): half-warp h walks superblock 2 * unit_sb + h
#ifndef BI_MIX_MINCTAS
#define BI_MIX_MINCTAS 3      /* 168 registers: measured best for the grouped instantiations (1: 246 regs, 4: spills) */
#endif
template <int NP, int NS>
__global__ void __launch_bounds__(BI_TS_THREADS, (NP > 1 ? BI_MIX_MINCTAS : 5))      // NP = 1: <= 96 registers, 20 warps / SM
k_mixture_partials(const double* __restrict__ tmix, int64_t n_bins /* doubles per mixture row: bins x pack */,
                   const __grid_constant__ BiTsSpace sp,
                   const int32_t* __restrict__ ev_bin, const double* __restrict__ ev_frac, int64_t ld_frac,
                   const int64_t* __restrict__ dataset_offset, const int32_t* __restrict__ status,
                   int64_t n_groups, const BiTsGroup* __restrict__ groups, const int64_t* __restrict__ unit_offset,
                   const int32_t* __restrict__ unit_group, int64_t n_units,
                   const int32_t* __restrict__ pair_point, const int64_t* __restrict__ pair_partial_offset,
                   double outlier, double* __restrict__ partial) {
    constexpr int NY = NS > 0 ? NS : 1;
    constexpr int CH = NP == 1 ? 2 : 1;                              // groups per prefetch chunk
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, l16 = lane & 15;
    const unsigned half_shift = 16u * half;
    const int64_t n_warps = (int64_t)gridDim.x * BI_TS_WARPS;

    for (int64_t u = (int64_t)blockIdx.x * BI_TS_WARPS + warp; u < n_units; u += n_warps) {
        int64_t g;
        if (unit_group) g = unit_group[u];
        else {
            int64_t lo = 0, hi = n_groups;
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (unit_offset[mid] <= u) lo = mid; else hi = mid;
            }
            g = lo;
        }
        const BiTsGroup gp = groups[g];
        const int np = gp.count < NP ? gp.count : NP;
        unsigned live = 0;
#pragma unroll
        for (int q = 0; q < NP; ++q)
            if (q < np && status[pair_point ? pair_point[gp.first + q] : gp.first + q] == 0) live |= 1u << q;
        if (!live) continue;
        const double* Vbase = tmix + (int64_t)gp.first * n_bins;    // pair q of the group: Vbase + q * n_bins

        const int64_t ds_begin = dataset_offset[gp.dataset], ds_end = dataset_offset[gp.dataset + 1];
        const int64_t sb = 2 * (u - unit_offset[g]) + half;          // this half's superblock
        const int64_t ev_begin = ds_begin + sb * BI_SUPERBLOCK;
        const int64_t left = ds_end - ev_begin;                      // <= 0: this half has no superblock
        const int n_ev = left < BI_SUPERBLOCK ? (left > 0 ? (int)left : 0) : BI_SUPERBLOCK;
        const bool all_full = __all_sync(BI_FULL_MASK, n_ev == BI_SUPERBLOCK);
        const int n_max = max(n_ev, __shfl_xor_sync(BI_FULL_MASK, n_ev, 16));
        const bool vec = ((ev_begin | ld_frac) & 1) == 0;            // pairs 8-byte (bins) / 16-byte (fractions) aligned

        double M[NP], Lslow[NP];
        int E[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) { M[q] = 1.0; E[q] = 0; Lslow[q] = 0.0; }
        bool any_slow = false;

        int bin[CH][2], bin_next[CH][2];
        double y[CH][2][NY], y_next[CH][2][NY];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int e = j * BI_EVENT_BLOCK + 2 * l16;
            bi_mix_load<NS>(ev_bin, ev_frac, ld_frac, ev_begin + e, n_ev - e, vec, bin[j], y[j]);
        }
#pragma unroll 1
        for (int c0 = 0; c0 < n_max; c0 += CH * BI_EVENT_BLOCK) {
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int e = c0 + (CH + j) * BI_EVENT_BLOCK + 2 * l16;
                bi_mix_load<NS>(ev_bin, ev_frac, ld_frac, ev_begin + e, n_ev - e, vec, bin_next[j], y_next[j]);
            }
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int e0 = c0 + j * BI_EVENT_BLOCK;
                if (all_full && np == NP)
                    bi_mix_group<NP, NS, true, true>(Vbase, n_bins, np, sp, bin[j], y[j], BI_SUPERBLOCK, l16, half_shift, live, outlier, M, E, Lslow, any_slow);
                else if (all_full)
                    bi_mix_group<NP, NS, true, false>(Vbase, n_bins, np, sp, bin[j], y[j], BI_SUPERBLOCK, l16, half_shift, live, outlier, M, E, Lslow, any_slow);
                else if (e0 < n_max)
                    bi_mix_group<NP, NS, false, false>(Vbase, n_bins, np, sp, bin[j], y[j], n_ev - e0, l16, half_shift, live, outlier, M, E, Lslow, any_slow);
            }
#pragma unroll
            for (int j = 0; j < CH; ++j) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    bin[j][h] = bin_next[j][h];
#pragma unroll
                    for (int d = 0; d < NY; ++d) y[j][h][d] = y_next[j][h][d];
                }
            }
        }
        // ---- close the superblocks: M = (M_0 * M_1) * (M_2 * M_3), E = sum, L = (L_0 + L_1) + (L_2 + L_3)
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            double m = M[q];
            m = __dmul_rn(m, __shfl_xor_sync(BI_FULL_MASK, m, 1));
            m = __dmul_rn(m, __shfl_xor_sync(BI_FULL_MASK, m, 2));
            int e = E[q];
            e += __shfl_xor_sync(BI_FULL_MASK, e, 1);
            e += __shfl_xor_sync(BI_FULL_MASK, e, 2);
            double L = bi_block_log(m, e);
            if (any_slow) {
                double l = Lslow[q];
                l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 1));
                l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 2));
                L = __dadd_rn(L, l);
            }
            if (l16 == 0 && n_ev > 0 && ((live >> q) & 1u)) partial[pair_partial_offset[gp.first + q] + sb] = L;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K5b on the FP64 tensor pipe: groups of up to 8 * MT points (the finite-difference batch of a minimiser step, a scan).
// With the events sorted by bin, the 8 events of an octet almost always share their low-corner bin, and the lookup of
// 8 points x 8 events is ONE 8x8x4 contraction over the lookup corners:
//     f[q, i] = sum_c Tmix_q[bin, c] * w_c(x_i)       A = 8 points x 4 corners (one 8-byte load per lane from the packed
//                                                     mixture rows), B = 4 corner weights x 8 events, D = densities
// DMMA accumulates as the fma chain of bi_mix_eval (fma(v0, w0, +0) = fl(v0 * w0), then corners ascending; proved on the
// device, profiles/microbench/dmma_probe_b200.log), and its D fragment hands lane (g, t) the events 8n + 2t, 8n + 2t + 1
// of point g -- the canonical class layout -- so the product tree runs inside the lane and every result is BIT-IDENTICAL
// to k_mixture_partials<1 | 8, NS>.  An octet whose events straddle a bin edge takes 8 contractions (one per event's
// bin, its own column kept).  Per 32 events x 8 points: 4 loads + 4 DMMA per corner quartet instead of 16 256-bit
// gathers per lane; the prepared events are loaded coalesced, 4 groups (2.5 kB per warp) ahead, and handed to the
// fragment lanes by shuffles (shared by the MT m-tiles of the warp).  A chunk of 4 groups whose octets all share
// their bins (the common case) runs branch-free with the A fragments of the next group in flight.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bi_ts_dmma(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

#define BI_MIXM_CHUNK 4        /* groups (of 32 events) per prefetch chunk */
#ifndef BI_MIXM_MINCTAS
#define BI_MIXM_MINCTAS 4
#endif
#ifndef BI_MIXM_MINCTAS2
#define BI_MIXM_MINCTAS2 3
#endif

template <int NS>
struct BiMixm {
    static constexpr int NY = NS > 0 ? NS : 1;
    static constexpr int PACK = NS == 0 ? 1 : (NS == 1 ? 2 : 4);      // doubles per bin of a packed mixture row
    static constexpr int KS = NS <= 2 ? 1 : (1 << NS) / 4;            // corner quartets = DMMA k-steps
    static constexpr int CORNERS = 1 << NS;                           // k >= CORNERS of the only quartet: zero padding
};

template <int NS>
struct BiMixChunk {
    int bin[BI_MIXM_CHUNK];
    double y[BiMixm<NS>::NY][BI_MIXM_CHUNK];
};

// lane's events c0 + 32 i + lane of the superblock (events >= n_ev: bin 0, fractions 0; their densities count as 1)
template <int NS>
__device__ __forceinline__ void bi_mixm_load(const int32_t* __restrict__ ev_bin, const double* __restrict__ ev_frac,
                                             int64_t ld_frac, int64_t ev_begin, int c0, int n_ev, int lane, BiMixChunk<NS>& ch) {
#pragma unroll
    for (int i = 0; i < BI_MIXM_CHUNK; ++i) {
        const int e = c0 + 32 * i + lane;
        const bool ok = e < n_ev;
        ch.bin[i] = ok ? __ldg(ev_bin + ev_begin + e) : 0;
#pragma unroll
        for (int d = 0; d < NS; ++d) ch.y[d][i] = ok ? __ldg(ev_frac + (int64_t)d * ld_frac + ev_begin + e) : 0.0;
        if (NS == 0) ch.y[0][i] = 0.0;
    }
}

// (class, group) fallback: the canonical tree over log(p_i) of the class's 8 events (noinline: out of the hot loop)
static __device__ __noinline__ double bi_mixm_slow(double p00, double p01, double p10, double p11, double p20, double p21,
                                                   double p30, double p31, double outlier) {
    const double l0 = __dadd_rn(log(bi_fix_density(p00, outlier)), log(bi_fix_density(p01, outlier)));
    const double l1 = __dadd_rn(log(bi_fix_density(p10, outlier)), log(bi_fix_density(p11, outlier)));
    const double l2 = __dadd_rn(log(bi_fix_density(p20, outlier)), log(bi_fix_density(p21, outlier)));
    const double l3 = __dadd_rn(log(bi_fix_density(p30, outlier)), log(bi_fix_density(p31, outlier)));
    return __dadd_rn(__dadd_rn(l0, l1), __dadd_rn(l2, l3));
}

// B fragments.  The corner weights  w_c = ((1 * x_0) * x_1) ...  (x_d = the event's fraction or fl(1 - fraction) along
// dimension d) are formed ONCE per event by the lane that loaded it -- NS subtractions and the corner products, all 2^NS
// corners of the event in one lane -- and staged in shared memory as [group of the chunk][quartet j][event][4 corners];
// fragment lane (g, t) then reads the weight of corner 4 j + t at event 8 n + g with one conflict-free LDS.64 per octet
// and quartet.  (Round 1 handed the fractions to the fragment lanes by shuffles and formed the weights there: 16 SHFL +
// 12 FP64 operations per group and a shuffle -> subtract -> multiply chain in front of every DMMA.)
template <int NS>
struct BiMixmW {
    static constexpr int GROUP_DOUBLES = BiMixm<NS>::KS * 32 * 4;               // weights of one 32-event group
    static constexpr int WARP_DOUBLES = BI_MIXM_CHUNK * GROUP_DOUBLES;          // one chunk per warp
    static constexpr int SMEM_BYTES = BI_TS_WARPS * WARP_DOUBLES * 8;
};
template <int NS>
__device__ __forceinline__ void bi_mixm_store_weights(const BiMixChunk<NS>& ch, double* wbuf, int lane) {
    constexpr int KS = BiMixm<NS>::KS, CORNERS = BiMixm<NS>::CORNERS, NY = BiMixm<NS>::NY;
#pragma unroll
    for (int i = 0; i < BI_MIXM_CHUNK; ++i) {
        double up[NY], dn[NY];
#pragma unroll
        for (int dd = 0; dd < NY; ++dd) {
            up[dd] = ch.y[dd][i];
            dn[dd] = __dsub_rn(1.0, up[dd]);
        }
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            double v4[4];
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) {
                const int c = (NS <= 1) ? (tt & 1) : 4 * j + tt;
                double v = 1.0;
#pragma unroll
                for (int dd = 0; dd < NS; ++dd) {
                    const double f = ((c >> (NS - 1 - dd)) & 1) ? up[dd] : dn[dd];
                    v = dd == 0 ? f : __dmul_rn(v, f);
                }
                v4[tt] = tt >= CORNERS ? 0.0 : v;                   // fewer than 4 lookup corners: zero padding
            }
            double2* dst = reinterpret_cast<double2*>(wbuf + ((size_t)(i * KS + j) * 32 + lane) * 4);
            dst[0] = make_double2(v4[0], v4[1]);
            dst[1] = make_double2(v4[2], v4[3]);
        }
    }
}
// weights of the group at wgrp for this fragment lane: corner 4 j + t at event 8 n + g
template <int NS>
__device__ __forceinline__ void bi_mixm_weights(const double* wgrp, int g, int t, double (&w)[4][BiMixm<NS>::KS]) {
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int j = 0; j < BiMixm<NS>::KS; ++j) w[n][j] = wgrp[((j * 32) + 8 * n + g) * 4 + t];
}

// densities of this lane's class -> canonical product tree -> (M, E) of the superblock; returns true when a density left
// the fast range (the class then contributes (m, e) = (1, 0) and its events go through bi_mixm_slow)
template <bool FULL>
__device__ __forceinline__ bool bi_mixm_tree(double (&d)[4][2], int n_left, int t, double& M, int& E) {
    if (!FULL) {
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int e = 8 * n + 2 * t;
            if (e >= n_left) d[n][0] = 1.0;                           // events >= N count as p = 1
            if (e + 1 >= n_left) d[n][1] = 1.0;
        }
    }
    unsigned tmax = (unsigned)(__double2hiint(d[0][0]) - BI_RANGE_LO);
    tmax = max(tmax, (unsigned)(__double2hiint(d[0][1]) - BI_RANGE_LO));
#pragma unroll
    for (int n = 1; n < 4; ++n) {
        tmax = max(tmax, (unsigned)(__double2hiint(d[n][0]) - BI_RANGE_LO));
        tmax = max(tmax, (unsigned)(__double2hiint(d[n][1]) - BI_RANGE_LO));
    }
    const double q0 = __dmul_rn(__dmul_rn(d[0][0], d[0][1]), __dmul_rn(d[1][0], d[1][1]));
    const double q1 = __dmul_rn(__dmul_rn(d[2][0], d[2][1]), __dmul_rn(d[3][0], d[3][1]));
    double m;
    int e;
    bi_split(__dmul_rn(q0, q1), &m, &e);
    const bool bad = tmax >= BI_RANGE_SPAN;
    if (bad) { m = 1.0; e = 0; }
    M = __dmul_rn(M, m);
    E += e;
    return bad;
}
// rare: the class's 8 events through the log tree
__device__ __forceinline__ void bi_mixm_slow_apply(const double (&d)[4][2], double outlier, double& L, bool& any_slow) {
    L = __dadd_rn(L, bi_mixm_slow(d[0][0], d[0][1], d[1][0], d[1][1], d[2][0], d[2][1], d[3][0], d[3][1], outlier));
    any_slow = true;
}

// A fragments of a group whose four octets each share one bin: element t of the packed bin of point slot g
template <int NS, int MT>
__device__ __forceinline__ void bi_mixm_load_a(const double* const (&Vq)[MT], bool a_zero, const BiTsSpace& sp, int bin,
                                               double (&a)[MT][4][BiMixm<NS>::KS]) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const int64_t e = (int64_t)__shfl_sync(BI_FULL_MASK, bin, 8 * n) * BiMixm<NS>::PACK;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int j = 0; j < BiMixm<NS>::KS; ++j) a[mt][n][j] = a_zero ? 0.0 : __ldg(Vq[mt] + e + sp.corner_off[4 * j]);
    }
}

// one full group with shared bins per octet.  DEFER = false: the (rare) slow path runs right after each m-tile.
// DEFER = true: branch-free -- bit mt of the result flags an m-tile whose class left the fast range; the caller runs
// bi_mixm_group_redo for such groups afterwards, in group order (the order in which L accumulates).
template <int NS, int MT, bool DEFER>
__device__ __forceinline__ unsigned bi_mixm_group_fast(const double (&a)[MT][4][BiMixm<NS>::KS], const double* wgrp,
                                                       int g, int t, bool a_zero, const bool (&live_me)[MT], double outlier,
                                                       double (&M)[MT], int (&E)[MT], double (&L)[MT], bool& any_slow) {
    double w[4][BiMixm<NS>::KS];
    bi_mixm_weights<NS>(wgrp, g, t, w);
    unsigned flags = 0;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        double d[4][2];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            d[n][0] = d[n][1] = 0.0;
#pragma unroll
            for (int j = 0; j < BiMixm<NS>::KS; ++j) bi_ts_dmma(d[n][0], d[n][1], a[mt][n][j], w[n][j]);
        }
        const bool bad = bi_mixm_tree<true>(d, 32, t, M[mt], E[mt]);
        if (DEFER) flags |= (bad && live_me[mt]) ? 1u << mt : 0u;
        else if (bad && live_me[mt]) bi_mixm_slow_apply(d, outlier, L[mt], any_slow);
    }
    return flags;
}
// the densities of a flagged group again (warp-collective), then the log tree in the lanes that flagged it
template <int NS, int MT>
__device__ __forceinline__ void bi_mixm_group_redo(const double (&a)[MT][4][BiMixm<NS>::KS], const double* wgrp,
                                                   int g, int t, bool a_zero, unsigned flags, double outlier, double (&L)[MT],
                                                   bool& any_slow) {
    double w[4][BiMixm<NS>::KS];
    bi_mixm_weights<NS>(wgrp, g, t, w);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        double d[4][2];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            d[n][0] = d[n][1] = 0.0;
#pragma unroll
            for (int j = 0; j < BiMixm<NS>::KS; ++j) bi_ts_dmma(d[n][0], d[n][1], a[mt][n][j], w[n][j]);
        }
        if ((flags >> mt) & 1u) bi_mixm_slow_apply(d, outlier, L[mt], any_slow);
    }
}

// any group: octets that straddle a bin edge take one contraction per event; events >= n_left count as p = 1
template <int NS, int MT>
__device__ __forceinline__ void bi_mixm_group_any(const double* const (&Vq)[MT], const BiTsSpace& sp, int bin,
                                                  const double* wgrp, int n_left, int g, int t, bool a_zero,
                                                  const bool (&live_me)[MT], double outlier, double (&M)[MT], int (&E)[MT],
                                                  double (&L)[MT], bool& any_slow) {
    constexpr int KS = BiMixm<NS>::KS, PACK = BiMixm<NS>::PACK;
    const int first = __shfl_sync(BI_FULL_MASK, bin, threadIdx.x & 24);
    const unsigned neq = __ballot_sync(BI_FULL_MASK, bin != first);   // octets whose 8 events do not share one bin
    double w[4][KS];
    bi_mixm_weights<NS>(wgrp, g, t, w);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        double d[4][2];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            d[n][0] = d[n][1] = 0.0;
            if (((neq >> (8 * n)) & 0xffu) == 0) {                    // warp-uniform
                const int64_t e = (int64_t)__shfl_sync(BI_FULL_MASK, bin, 8 * n) * PACK;
#pragma unroll
                for (int j = 0; j < KS; ++j)
                    bi_ts_dmma(d[n][0], d[n][1], a_zero ? 0.0 : __ldg(Vq[mt] + e + sp.corner_off[4 * j]), w[n][j]);
            } else if constexpr (KS <= 2) {                           // event 8n + jj: its own bin, its own column
                double a8[8][KS];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {                      // the eight loads in flight together
                    const int64_t e = (int64_t)__shfl_sync(BI_FULL_MASK, bin, 8 * n + jj) * PACK;
#pragma unroll
                    for (int j = 0; j < KS; ++j) a8[jj][j] = a_zero ? 0.0 : __ldg(Vq[mt] + e + sp.corner_off[4 * j]);
                }
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    double t0 = 0.0, t1 = 0.0;
#pragma unroll
                    for (int j = 0; j < KS; ++j) bi_ts_dmma(t0, t1, a8[jj][j], w[n][j]);
                    if (2 * t == jj) d[n][0] = t0;
                    if (2 * t + 1 == jj) d[n][1] = t1;
                }
            } else {
#pragma unroll 1
                for (int jj = 0; jj < 8; ++jj) {
                    const int64_t e = (int64_t)__shfl_sync(BI_FULL_MASK, bin, 8 * n + jj) * PACK;
                    double t0 = 0.0, t1 = 0.0;
#pragma unroll
                    for (int j = 0; j < KS; ++j)
                        bi_ts_dmma(t0, t1, a_zero ? 0.0 : __ldg(Vq[mt] + e + sp.corner_off[4 * j]), w[n][j]);
                    if (2 * t == jj) d[n][0] = t0;
                    if (2 * t + 1 == jj) d[n][1] = t1;
                }
            }
        }
        const bool bad = n_left >= 32 ? bi_mixm_tree<true>(d, 32, t, M[mt], E[mt]) : bi_mixm_tree<false>(d, n_left, t, M[mt], E[mt]);
        if (bad && live_me[mt]) bi_mixm_slow_apply(d, outlier, L[mt], any_slow);
    }
}

// unit = (pair group, pair of consecutive superblocks), as k_mixture_partials: the warp walks the two superblocks in turn;
// lane (g, t) owns class t of the point slots g, g + 8, ... of the group
template <int NS, int MT>
__global__ void __launch_bounds__(BI_TS_THREADS, (MT == 1 ? BI_MIXM_MINCTAS : BI_MIXM_MINCTAS2))
k_mixture_partials_mma(const double* __restrict__ tmix, int64_t n_bins /* doubles per mixture row: bins x pack */,
                       const __grid_constant__ BiTsSpace sp,
                       const int32_t* __restrict__ ev_bin, const double* __restrict__ ev_frac, int64_t ld_frac,
                       const int64_t* __restrict__ dataset_offset, const int32_t* __restrict__ status,
                       int64_t n_groups, const BiTsGroup* __restrict__ groups, const int64_t* __restrict__ unit_offset,
                       const int32_t* __restrict__ unit_group, int64_t n_units,
                       const int32_t* __restrict__ pair_point, const int64_t* __restrict__ pair_partial_offset,
                       double outlier, double* __restrict__ partial) {
    static_assert(NS >= 0 && NS <= 4, "piecewise lookups or linear lookups in 1..4 dimensions");
    constexpr int KS = BiMixm<NS>::KS;
    extern __shared__ __align__(16) double bi_mixm_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int64_t n_warps = (int64_t)gridDim.x * BI_TS_WARPS;
    double* wbuf = bi_mixm_smem + (size_t)warp * BiMixmW<NS>::WARP_DOUBLES;      // corner weights of this warp's chunk
    const bool a_zero = BiMixm<NS>::CORNERS < 4 && t >= BiMixm<NS>::CORNERS;   // fewer than 4 lookup corners: zero padding

    for (int64_t u = (int64_t)blockIdx.x * BI_TS_WARPS + warp; u < n_units; u += n_warps) {
        int64_t gi;
        if (unit_group) gi = unit_group[u];
        else {
            int64_t lo = 0, hi = n_groups;
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (unit_offset[mid] <= u) lo = mid; else hi = mid;
            }
            gi = lo;
        }
        const BiTsGroup gp = groups[gi];
        const int np = gp.count < 8 * MT ? gp.count : 8 * MT;
        int64_t pair[MT];
        bool live_me[MT];
        const double* Vq[MT];
        bool any_live = false;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int slot = 8 * mt + g;
            pair[mt] = gp.first + (slot < np ? slot : np - 1);       // slots >= np repeat the last pair; results unused
            live_me[mt] = slot < np && status[pair_point ? pair_point[pair[mt]] : pair[mt]] == 0;
            any_live |= live_me[mt];
            Vq[mt] = tmix + pair[mt] * n_bins + (a_zero ? 0 : t);
        }
        if (!__any_sync(BI_FULL_MASK, any_live)) continue;
        const int64_t ds_begin = dataset_offset[gp.dataset], ds_end = dataset_offset[gp.dataset + 1];

#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int64_t sb = 2 * (u - unit_offset[gi]) + h;
            const int64_t ev_begin = ds_begin + sb * BI_SUPERBLOCK;
            const int64_t left = ds_end - ev_begin;
            if (left <= 0) break;
            const int n_ev = left < BI_SUPERBLOCK ? (int)left : BI_SUPERBLOCK;
            double M[MT], L[MT];
            int E[MT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) { M[mt] = 1.0; L[mt] = 0.0; E[mt] = 0; }
            bool any_slow = false;
            BiMixChunk<NS> cur, nxt;
            bi_mixm_load<NS>(ev_bin, ev_frac, ld_frac, ev_begin, 0, n_ev, lane, cur);
#pragma unroll 1
            for (int c0 = 0; c0 < n_ev; c0 += 32 * BI_MIXM_CHUNK) {
                bi_mixm_load<NS>(ev_bin, ev_frac, ld_frac, ev_begin, c0 + 32 * BI_MIXM_CHUNK, n_ev, lane, nxt);
                __syncwarp();                                         // the previous chunk's fragment reads are done
                bi_mixm_store_weights<NS>(cur, wbuf, lane);
                __syncwarp();
                bool ne = false;
#pragma unroll
                for (int i = 0; i < BI_MIXM_CHUNK; ++i)
                    ne |= cur.bin[i] != __shfl_sync(BI_FULL_MASK, cur.bin[i], lane & 24);
                if (n_ev - c0 >= 32 * BI_MIXM_CHUNK && !__any_sync(BI_FULL_MASK, ne)) {
                    // every octet of the chunk shares its bin: branch-free, the next group's A fragments in flight
                    const int b0 = __shfl_sync(BI_FULL_MASK, cur.bin[0], 0);
                    bool one = true;
#pragma unroll
                    for (int i = 0; i < BI_MIXM_CHUNK; ++i) one &= cur.bin[i] == b0;
                    double a[MT][4][KS], an[MT][4][KS];
                    if (__all_sync(BI_FULL_MASK, one)) {
                        // the whole chunk lies in ONE bin (dense datasets): one A fragment per m-tile for its 16 octets
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                            for (int j = 0; j < KS; ++j) {
                                const double v = a_zero ? 0.0 : __ldg(Vq[mt] + (int64_t)b0 * BiMixm<NS>::PACK + sp.corner_off[4 * j]);
#pragma unroll
                                for (int n = 0; n < 4; ++n) a[mt][n][j] = v;
                            }
                        unsigned flags = 0;                                   // bit i * MT + mt: group i, m-tile mt left the fast range
#pragma unroll
                        for (int i = 0; i < BI_MIXM_CHUNK; ++i)               // one branch-free block: the groups interleave
                            flags |= bi_mixm_group_fast<NS, MT, true>(a, wbuf + i * BiMixmW<NS>::GROUP_DOUBLES, g, t, a_zero,
                                                                      live_me, outlier, M, E, L, any_slow) << (i * MT);
                        if (__any_sync(BI_FULL_MASK, flags != 0)) {           // rare: log trees, in group order
#pragma unroll 1
                            for (int i = 0; i < BI_MIXM_CHUNK; ++i) {
                                const unsigned f = (flags >> (i * MT)) & ((1u << MT) - 1u);
                                if (!__any_sync(BI_FULL_MASK, f != 0)) continue;
                                bi_mixm_group_redo<NS, MT>(a, wbuf + i * BiMixmW<NS>::GROUP_DOUBLES, g, t, a_zero, f, outlier, L,
                                                           any_slow);
                            }
                        }
                    } else {
                        bi_mixm_load_a<NS, MT>(Vq, a_zero, sp, cur.bin[0], a);
#pragma unroll
                        for (int i = 0; i < BI_MIXM_CHUNK; ++i) {
                            if (i + 1 < BI_MIXM_CHUNK) bi_mixm_load_a<NS, MT>(Vq, a_zero, sp, cur.bin[i + 1], an);
                            bi_mixm_group_fast<NS, MT, false>(a, wbuf + i * BiMixmW<NS>::GROUP_DOUBLES, g, t, a_zero, live_me,
                                                              outlier, M, E, L, any_slow);
                            if (i + 1 < BI_MIXM_CHUNK) {
#pragma unroll
                                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                                    for (int n = 0; n < 4; ++n)
#pragma unroll
                                        for (int j = 0; j < KS; ++j) a[mt][n][j] = an[mt][n][j];
                            }
                        }
                    }
                } else {
#pragma unroll 1
                    for (int i = 0; i < BI_MIXM_CHUNK; ++i) {
                        const int e0 = c0 + 32 * i;
                        if (e0 >= n_ev) break;
                        int bin_i = cur.bin[0];
#pragma unroll
                        for (int k = 1; k < BI_MIXM_CHUNK; ++k)
                            if (i == k) bin_i = cur.bin[k];
                        const double* wgrp = wbuf + i * BiMixmW<NS>::GROUP_DOUBLES;
                        const bool shared = !__any_sync(BI_FULL_MASK, bin_i != __shfl_sync(BI_FULL_MASK, bin_i, lane & 24));
                        if (shared && n_ev - e0 >= 32) {                  // the group's octets share their bins
                            double a[MT][4][KS];
                            bi_mixm_load_a<NS, MT>(Vq, a_zero, sp, bin_i, a);
                            bi_mixm_group_fast<NS, MT, false>(a, wgrp, g, t, a_zero, live_me, outlier, M, E, L, any_slow);
                        } else {
                            bi_mixm_group_any<NS, MT>(Vq, sp, bin_i, wgrp, n_ev - e0, g, t, a_zero, live_me, outlier, M, E, L, any_slow);
                        }
                    }
                }
                cur = nxt;
            }
            // ---- close the superblock: M = (M_0 * M_1) * (M_2 * M_3), E = sum, L = (L_0 + L_1) + (L_2 + L_3)
            // (the four lanes of a row hold identical (m, e, l) after the butterflies: lane t evaluates the log of m-tile t,
            // so a warp runs ONE log stream per superblock whatever MT is)
            static_assert(MT <= 4, "one class lane per m-tile");
            const bool slow = __any_sync(BI_FULL_MASK, any_slow);
            double m_mine = 1.0, l_mine = 0.0;
            int e_mine = 0;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                double m = M[mt];
                m = __dmul_rn(m, __shfl_xor_sync(BI_FULL_MASK, m, 1));
                m = __dmul_rn(m, __shfl_xor_sync(BI_FULL_MASK, m, 2));
                int e = E[mt];
                e += __shfl_xor_sync(BI_FULL_MASK, e, 1);
                e += __shfl_xor_sync(BI_FULL_MASK, e, 2);
                double l = 0.0;
                if (slow) {
                    l = L[mt];
                    l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 1));
                    l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 2));
                }
                if (MT == 1 || t == mt) { m_mine = m; e_mine = e; l_mine = l; }
            }
            double r = bi_block_log(m_mine, e_mine);
            if (slow) r = __dadd_rn(r, l_mine);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
                if (t == mt && live_me[mt]) partial[pair_partial_offset[pair[mt]] + sb] = r;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// ragged finalize: logl[q] = -musum[point] + total(partials of pair q), the canonical total of
// k_unbinned_finalize (256 strided lanes, xor butterfly per 32, pairwise over the 8 warp totals) by ONE warp
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_template_finalize(const double* __restrict__ partial, const int64_t* __restrict__ pair_partial_offset,
                    const int32_t* __restrict__ pair_point, const double* __restrict__ musum,
                    const int32_t* __restrict__ status, int64_t n_pairs, double* __restrict__ logl,
                    double* __restrict__ logsum) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= n_pairs) return;
    const int64_t pt = pair_point ? pair_point[q] : q;
    if (status[pt] != 0) {
        if (lane == 0) {
            logl[q] = -__longlong_as_double(0x7ff0000000000000LL);
            if (logsum) logsum[q] = 0.0;
        }
        return;
    }
    const double* src = partial + pair_partial_offset[q];
    const int64_t n = pair_partial_offset[q + 1] - pair_partial_offset[q];
    double w[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) {                                   // virtual warp v of the 256-thread finalize
        double u = bi_strided_sum(src, v * 32 + lane, n);
#pragma unroll
        for (int x = 1; x < 32; x <<= 1) u = __dadd_rn(u, __shfl_xor_sync(BI_FULL_MASK, u, x));
        w[v] = u;
    }
    if (lane == 0) {
        const double total = __dadd_rn(__dadd_rn(__dadd_rn(w[0], w[1]), __dadd_rn(w[2], w[3])),
                                       __dadd_rn(__dadd_rn(w[4], w[5]), __dadd_rn(w[6], w[7])));
        logl[q] = __dadd_rn(-musum[pt], total);
        if (logsum) logsum[q] = total;
    }
}

// the same total with one 256-thread CTA per pair (k_unbinned_finalize's own shape): pairs with many partials
__global__ void __launch_bounds__(256)
k_template_finalize_cta(const double* __restrict__ partial, const int64_t* __restrict__ pair_partial_offset,
                        const int32_t* __restrict__ pair_point, const double* __restrict__ musum,
                        const int32_t* __restrict__ status, double* __restrict__ logl, double* __restrict__ logsum) {
    __shared__ double warp_tot[8];
    const int64_t q = blockIdx.x;
    const int t = threadIdx.x;
    const int64_t pt = pair_point ? pair_point[q] : q;
    if (status[pt] != 0) {                                        // block-uniform
        if (t == 0) {
            logl[q] = -__longlong_as_double(0x7ff0000000000000LL);
            if (logsum) logsum[q] = 0.0;
        }
        return;
    }
    const double* src = partial + pair_partial_offset[q];
    const int64_t n = pair_partial_offset[q + 1] - pair_partial_offset[q];
    double u = bi_strided_sum(src, t, n);
#pragma unroll
    for (int x = 1; x < 32; x <<= 1) u = __dadd_rn(u, __shfl_xor_sync(BI_FULL_MASK, u, x));
    if ((t & 31) == 0) warp_tot[t >> 5] = u;
    __syncthreads();
    if (t == 0) {
        const double total = __dadd_rn(__dadd_rn(__dadd_rn(warp_tot[0], warp_tot[1]), __dadd_rn(warp_tot[2], warp_tot[3])),
                                       __dadd_rn(__dadd_rn(warp_tot[4], warp_tot[5]), __dadd_rn(warp_tot[6], warp_tot[7])));
        logl[q] = __dadd_rn(-musum[pt], total);
        if (logsum) logsum[q] = total;
    }
}

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
extern "C" int bi_template_prepare_events(int32_t n_space, const int32_t* n_bins_host, const double* edges_host,
                                          int32_t method, const double* coords_dev, int64_t ld_coords,
                                          int64_t n_events, int32_t* ev_bin_dev, double* ev_frac_dev, int64_t ld_frac,
                                          void* stream) {
    BiSpace sp;
    int rc = bi_fill_space(&sp, n_space, n_bins_host);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(method == BI_LOOKUP_LINEAR || method == BI_LOOKUP_PIECEWISE, "unknown lookup method %d", method);
    BiPoints pts;
    int total = 0;
    rc = bi_fill_points(&sp, &pts, edges_host, method == BI_LOOKUP_LINEAR, &total);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(n_events >= 0, "n_events < 0");
    if (n_events == 0) return BI_OK;
    BI_REQUIRE(coords_dev && ev_bin_dev && ld_coords >= n_events, "bi_template_prepare_events: bad arguments");
    BI_REQUIRE(method == BI_LOOKUP_PIECEWISE || (ev_frac_dev && ld_frac >= n_events),
               "bi_template_prepare_events: ev_frac_dev / ld_frac needed for the linear method");
    const int64_t blocks = (n_events + 255) / 256;
    k_template_prepare<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        sp, pts, total, method == BI_LOOKUP_LINEAR, coords_dev, ld_coords, n_events, ev_bin_dev, ev_frac_dev, ld_frac);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

// groups of one pair each (toys) bucketed by the hypercube cell of their point: counting sort in ONE CTA (shared-memory
// histogram over the cells, scan, scatter).  group_order [n_groups]: live groups (status 0), cell-major; n_ordered[0]: their
// number.  The order inside a cell is decided by atomics -- results never depend on the schedule.
__global__ void __launch_bounds__(BI_PLAN_THREADS)
k_ts_order_groups(const __grid_constant__ BiPlanDims dims, int n_cells, int64_t n_groups,
                  const int32_t* __restrict__ cell, const int32_t* __restrict__ status,
                  const int32_t* __restrict__ pair_point, const BiTsGroup* __restrict__ groups,
                  int32_t* __restrict__ group_order, int32_t* __restrict__ n_ordered) {
    extern __shared__ int bi_order_smem[];
    int* off = bi_order_smem;                      // [n_cells + 1]
    int* cursor = off + n_cells + 1;               // [n_cells]
    __shared__ int carry[33];
    const int tid = threadIdx.x, D = dims.n_dims;
    for (int c = tid; c <= n_cells; c += BI_PLAN_THREADS) { off[c] = 0; if (c < n_cells) cursor[c] = 0; }
    __syncthreads();
    for (int64_t g = tid; g < n_groups; g += BI_PLAN_THREADS) {
        const int32_t pt = pair_point[groups[g].first];
        if (status[pt] == 0) atomicAdd(&off[bi_flat_cell(dims, cell + (int64_t)pt * D)], 1);
    }
    __syncthreads();
    const int total = bi_block_exclusive_scan(off, n_cells, carry);
    for (int64_t g = tid; g < n_groups; g += BI_PLAN_THREADS) {
        const int32_t pt = pair_point[groups[g].first];
        if (status[pt] == 0) {
            const int c = bi_flat_cell(dims, cell + (int64_t)pt * D);
            group_order[off[c] + atomicAdd(&cursor[c], 1)] = (int32_t)g;
        }
    }
    if (tid == 0) n_ordered[0] = total;
}

template <int NP, int NS, bool PRE>
static int bi_ts_launch(const double* T, int64_t row_stride, int64_t bin_stride, const BiTsSpace& sp,
                        const int32_t* ev_bin, const double* ev_frac, int64_t ld_frac, const int64_t* dataset_offset,
                        int K, int S, const int32_t* row, const double* coef, const double* wterm,
                        const int32_t* term_source, const double* mus, const int32_t* status, int64_t n_groups,
                        const BiTsGroup* groups, const int64_t* unit_offset, const int32_t* unit_group, int64_t n_units,
                        const int32_t* pair_point, const int64_t* pair_partial_offset, double outlier, double* partial,
                        const int32_t* group_order, const int32_t* n_ordered, int sb_max, const double* pre,
                        const int32_t* pre_corner, const double* pre_weight, cudaStream_t st) {
    const int smem = BI_TS_WARPS * K * (1 + NP) * 8;
    static int per_sm_cached[BI_TS_MAX_TERMS + 1] = {0};
    static int sms = 0;
    if (!per_sm_cached[K]) {
        int dev = 0, per_sm = 0;
        BI_CUDA_CHECK(cudaGetDevice(&dev));
        BI_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        // the cap is per kernel, not per launch: always raise it to what the largest term count needs
        BI_CUDA_CHECK(cudaFuncSetAttribute(k_template_partials<NP, NS, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           BI_TS_WARPS * BI_TS_MAX_TERMS * (1 + NP) * 8));
        BI_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_template_partials<NP, NS, PRE>, BI_TS_THREADS, smem));
        BI_REQUIRE(per_sm >= 1, "k_template_partials<%d,%d> does not fit on this device", NP, NS);
        per_sm_cached[K] = per_sm;
    }
    int64_t blocks = (int64_t)sms * per_sm_cached[K];
    const int64_t needed = (n_units + BI_TS_WARPS - 1) / BI_TS_WARPS;
    if (blocks > needed) blocks = needed;
    k_template_partials<NP, NS, PRE><<<(unsigned)blocks, BI_TS_THREADS, smem, st>>>(
        T, row_stride, bin_stride, sp, ev_bin, ev_frac, ld_frac, dataset_offset, K, S, row, coef, wterm, term_source, mus,
        status, n_groups, groups, unit_offset, unit_group, n_units, pair_point, pair_partial_offset, outlier, partial,
        group_order, n_ordered, sb_max, pre, pre_corner, pre_weight);
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        bi_set_error("k_template_partials<%d,%d> launch failed: %s (blocks=%lld, smem=%d, K=%d, units=%lld)", NP, NS,
                     cudaGetErrorString(err), (long long)blocks, smem, K, (long long)n_units);
        return BI_ERR_CUDA;
    }
    return BI_OK;
}

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
                                    void* stream) {
    BiSpace space;
    int rc = bi_fill_space(&space, n_space, n_bins_host);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(method == BI_LOOKUP_LINEAR || method == BI_LOOKUP_PIECEWISE, "unknown lookup method %d", method);
    BI_REQUIRE(n_terms >= 1 && n_terms <= BI_TS_MAX_TERMS, "n_terms=%d outside [1,%d]", n_terms, BI_TS_MAX_TERMS);
    BI_REQUIRE(n_sources >= 1 && n_sources <= BI_MAX_SOURCES, "n_sources=%d outside [1,%d]", n_sources, BI_MAX_SOURCES);
    BI_REQUIRE(group_points == 1 || group_points == BI_TS_GROUP_POINTS, "group_points must be 1 or %d", BI_TS_GROUP_POINTS);
    BI_REQUIRE(n_groups >= 0 && n_units >= 0, "negative size");
    if (n_groups == 0 || n_units == 0) return BI_OK;
    BI_REQUIRE(templates_dev && ev_bin_dev && dataset_offset_dev && term_source_dev &&
                   mus_dev && status_dev && groups_dev && unit_offset_dev && pair_point_dev && pair_partial_offset_dev &&
                   partial_dev,
               "bi_template_partials: NULL device pointer");
    BI_REQUIRE(pre_dev ? (group_points == 1 && pre_corner_dev && pre_weight_dev) : (row_dev && coef_dev && wterm_dev),
               "bi_template_partials: NULL term pointer");
    BI_REQUIRE(method == BI_LOOKUP_PIECEWISE || ev_frac_dev, "bi_template_partials: ev_frac_dev is NULL");
    const int pack = method == BI_LOOKUP_PIECEWISE ? 1 : (n_space == 1 ? 2 : 4);
    BI_REQUIRE(((uintptr_t)templates_dev & (8 * pack - 1)) == 0 && (row_stride % pack) == 0 && (bin_stride % pack) == 0,
               "bi_template_partials: the lookup reads %d-double packed template elements (aligned, strides multiples of %d)",
               pack, pack);
    BiTsSpace sp;
    memset(&sp, 0, sizeof(sp));
    const int ns = method == BI_LOOKUP_LINEAR ? n_space : 0;
    sp.n_space = ns;
    sp.n_corner = 1 << ns;
    for (int c = 0; c < sp.n_corner; ++c) {
        int64_t off = 0;
        for (int d = 0; d < ns; ++d)
            if (((c >> (ns - 1 - d)) & 1) && space.n_bins[d] > 1) off += space.stride[d];
        sp.corner_off[c] = off * bin_stride;
    }
    const BiTsGroup* groups = reinterpret_cast<const BiTsGroup*>(groups_dev);
    cudaStream_t st = (cudaStream_t)stream;
#define BI_TS_CASE(NPV, NSV)                                                                                            \
    if (group_points == NPV && ns == NSV && (NPV == 1) && pre_dev)                                                      \
        return bi_ts_launch<1, NSV, true>(templates_dev, row_stride, bin_stride, sp, ev_bin_dev, ev_frac_dev, ld_frac,  \
                                      dataset_offset_dev, n_terms, n_sources, row_dev, coef_dev, wterm_dev,             \
                                      term_source_dev, mus_dev, status_dev, n_groups, groups, unit_offset_dev,          \
                                      unit_group_dev, n_units, pair_point_dev, pair_partial_offset_dev,                 \
                                      outlier_likelihood, partial_dev, group_order_dev, n_ordered_dev, sb_max,         \
                                      pre_dev, pre_corner_dev, pre_weight_dev, st);                                     \
    if (group_points == NPV && ns == NSV)                                                                               \
        return bi_ts_launch<NPV, NSV, false>(templates_dev, row_stride, bin_stride, sp, ev_bin_dev, ev_frac_dev, ld_frac,      \
                                      dataset_offset_dev, n_terms, n_sources, row_dev, coef_dev, wterm_dev,             \
                                      term_source_dev, mus_dev, status_dev, n_groups, groups, unit_offset_dev,          \
                                      unit_group_dev, n_units, pair_point_dev, pair_partial_offset_dev,                 \
                                      outlier_likelihood, partial_dev, group_order_dev, n_ordered_dev, sb_max,         \
                                      pre_dev, pre_corner_dev, pre_weight_dev, st);
    BI_TS_CASE(1, 0) BI_TS_CASE(1, 1) BI_TS_CASE(1, 2) BI_TS_CASE(1, 3) BI_TS_CASE(1, 4)
    BI_TS_CASE(BI_TS_GROUP_POINTS, 0) BI_TS_CASE(BI_TS_GROUP_POINTS, 1) BI_TS_CASE(BI_TS_GROUP_POINTS, 2)
    BI_TS_CASE(BI_TS_GROUP_POINTS, 3) BI_TS_CASE(BI_TS_GROUP_POINTS, 4)
#undef BI_TS_CASE
    bi_set_error("bi_template_partials: unsupported configuration");
    return BI_ERR_UNSUPPORTED;
}

extern "C" int bi_template_partials(const double* templates_dev, int64_t row_stride, int64_t bin_stride,
                                    int32_t n_space, const int32_t* n_bins_host, int32_t method,
                                    const int32_t* ev_bin_dev, const double* ev_frac_dev, int64_t ld_frac,
                                    const int64_t* dataset_offset_dev, int32_t n_terms, int32_t n_sources,
                                    const int32_t* row_dev, const double* coef_dev, const double* wterm_dev,
                                    const int32_t* term_source_dev, const double* mus_dev, const int32_t* status_dev,
                                    int64_t n_groups, int32_t group_points, const int32_t* groups_dev,
                                    const int64_t* unit_offset_dev, const int32_t* unit_group_dev, int64_t n_units,
                                    const int32_t* pair_point_dev, const int64_t* pair_partial_offset_dev,
                                    double outlier_likelihood, double* partial_dev, void* stream) {
    return bi_template_partials_impl(templates_dev, row_stride, bin_stride, n_space, n_bins_host, method, ev_bin_dev,
                                     ev_frac_dev, ld_frac, dataset_offset_dev, n_terms, n_sources, row_dev, coef_dev,
                                     wterm_dev, term_source_dev, mus_dev, status_dev, n_groups, group_points, groups_dev,
                                     unit_offset_dev, unit_group_dev, n_units, pair_point_dev, pair_partial_offset_dev,
                                     outlier_likelihood, partial_dev, nullptr, nullptr, 1, nullptr, nullptr, nullptr, stream);
}

extern "C" int bi_template_finalize(const double* partial_dev, const int64_t* pair_partial_offset_dev,
                                    const int32_t* pair_point_dev, const double* musum_dev, const int32_t* status_dev,
                                    int64_t n_pairs, int64_t max_partials, double* logl_dev, double* logsum_dev,
                                    void* stream) {
    BI_REQUIRE(n_pairs >= 0, "n_pairs < 0");
    if (n_pairs == 0) return BI_OK;
    BI_REQUIRE(pair_partial_offset_dev && musum_dev && status_dev && logl_dev, "bi_template_finalize: NULL pointer");
    if (max_partials > 256 && n_pairs < (1LL << 31)) {            // few long pairs: one CTA each (same summation order)
        k_template_finalize_cta<<<(unsigned)n_pairs, 256, 0, (cudaStream_t)stream>>>(
            partial_dev, pair_partial_offset_dev, pair_point_dev, musum_dev, status_dev, logl_dev, logsum_dev);
        BI_LAUNCH_CHECK();
        return BI_OK;
    }
    const int64_t blocks = (n_pairs + 7) / 8;
    k_template_finalize<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        partial_dev, pair_partial_offset_dev, pair_point_dev, musum_dev, status_dev, n_pairs, logl_dev, logsum_dev);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

// ---------------------------------------------------------------------------------------------
// C-ABI of the mixture form
// ---------------------------------------------------------------------------------------------
// packing of the mixture templates (as K5's template layout): doubles per bin and the bin offsets of the neighbours
static void bi_ts_packing(const BiSpace& space, int method, int* pack, int64_t* off1, int64_t* off2) {
    *pack = 1; *off1 = 0; *off2 = 0;
    if (method != BI_LOOKUP_LINEAR) return;
    const int ns = space.n_space;
    *off1 = space.n_bins[ns - 1] > 1 ? 1 : 0;
    if (ns == 1) { *pack = 2; return; }
    *pack = 4;
    *off2 = space.n_bins[ns - 2] > 1 ? space.stride[ns - 2] : 0;
}

extern "C" int bi_template_mix(const double* templates_dev, int64_t row_stride, int64_t bin_stride,
                               int32_t n_space, const int32_t* n_bins_host, int32_t method,
                               int32_t n_terms, const int32_t* row_dev, const double* coef_dev,
                               const int32_t* status_dev, const int32_t* pair_point_dev, int64_t n_pairs,
                               double* tmix_dev, void* stream) {
    BiSpace space;
    int rc = bi_fill_space(&space, n_space, n_bins_host);
    if (rc != BI_OK) return rc;
    const int64_t n_bins = space.n_cells;
    int pack;
    int64_t off1, off2;
    bi_ts_packing(space, method, &pack, &off1, &off2);
    BI_REQUIRE(n_terms >= 1 && n_bins >= 1 && n_pairs >= 0, "bi_template_mix: bad sizes");
    if (n_pairs == 0) return BI_OK;
    BI_REQUIRE(templates_dev && row_dev && coef_dev && status_dev && tmix_dev, "bi_template_mix: NULL device pointer");
    int64_t bx = (n_bins + 63) / 64;
    if (bx > 4096) bx = 4096;
    dim3 grid((unsigned)bx, (unsigned)(n_pairs < 65535 ? n_pairs : 65535));
    k_template_mix<<<grid, 64, 0, (cudaStream_t)stream>>>(templates_dev, row_stride, bin_stride, n_bins, n_terms, row_dev,
                                                         coef_dev, status_dev, pair_point_dev, n_pairs, tmix_dev, pack, off1, off2);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

template <int NP, int NS>
static int bi_mix_launch(const double* tmix, int64_t n_bins, const BiTsSpace& sp, const int32_t* ev_bin,
                         const double* ev_frac, int64_t ld_frac, const int64_t* dataset_offset, const int32_t* status,
                         int64_t n_groups, const BiTsGroup* groups, const int64_t* unit_offset, const int32_t* unit_group,
                         int64_t n_units, const int32_t* pair_point, const int64_t* pair_partial_offset, double outlier,
                         double* partial, cudaStream_t st) {
    static int resident = 0;
    if (!resident) {
        int dev = 0, sms = 0, per_sm = 0;
        BI_CUDA_CHECK(cudaGetDevice(&dev));
        BI_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        BI_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mixture_partials<NP, NS>, BI_TS_THREADS, 0));
        BI_REQUIRE(per_sm >= 1, "k_mixture_partials<%d,%d> does not fit on this device", NP, NS);
        resident = sms * per_sm;
    }
    int64_t blocks = resident;
    const int64_t needed = (n_units + BI_TS_WARPS - 1) / BI_TS_WARPS;
    if (blocks > needed) blocks = needed;
    k_mixture_partials<NP, NS><<<(unsigned)blocks, BI_TS_THREADS, 0, st>>>(
        tmix, n_bins, sp, ev_bin, ev_frac, ld_frac, dataset_offset, status, n_groups, groups, unit_offset, unit_group,
        n_units, pair_point, pair_partial_offset, outlier, partial);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

template <int NS, int MT>
static int bi_mixm_launch(const double* tmix, int64_t n_bins, const BiTsSpace& sp, const int32_t* ev_bin,
                          const double* ev_frac, int64_t ld_frac, const int64_t* dataset_offset, const int32_t* status,
                          int64_t n_groups, const BiTsGroup* groups, const int64_t* unit_offset, const int32_t* unit_group,
                          int64_t n_units, const int32_t* pair_point, const int64_t* pair_partial_offset, double outlier,
                          double* partial, cudaStream_t st) {
    static int resident = 0;
    if (!resident) {
        int dev = 0, sms = 0, per_sm = 0;
        BI_CUDA_CHECK(cudaGetDevice(&dev));
        BI_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        BI_CUDA_CHECK(cudaFuncSetAttribute(k_mixture_partials_mma<NS, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           BiMixmW<NS>::SMEM_BYTES));
        BI_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mixture_partials_mma<NS, MT>, BI_TS_THREADS,
                                                                    BiMixmW<NS>::SMEM_BYTES));
        BI_REQUIRE(per_sm >= 1, "k_mixture_partials_mma<%d,%d> does not fit on this device", NS, MT);
        resident = sms * per_sm;
    }
    int64_t blocks = resident;
    const int64_t needed = (n_units + BI_TS_WARPS - 1) / BI_TS_WARPS;
    if (blocks > needed) blocks = needed;
    k_mixture_partials_mma<NS, MT><<<(unsigned)blocks, BI_TS_THREADS, BiMixmW<NS>::SMEM_BYTES, st>>>(
        tmix, n_bins, sp, ev_bin, ev_frac, ld_frac, dataset_offset, status, n_groups, groups, unit_offset, unit_group,
        n_units, pair_point, pair_partial_offset, outlier, partial);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

// grouped linear lookups run on the tensor-pipe kernel (bit-identical results); BI_MIX_MMA=0 in the environment keeps
// the gather kernel for A/B measurements
static bool bi_mix_use_mma() {
    static int use = -1;
    if (use < 0) {
        const char* v = getenv("BI_MIX_MMA");
        use = (v && v[0] == '0') ? 0 : 1;
    }
    return use != 0;
}

extern "C" int bi_mixture_partials(const double* tmix_dev, int32_t n_space, const int32_t* n_bins_host, int32_t method,
                                   const int32_t* ev_bin_dev, const double* ev_frac_dev, int64_t ld_frac,
                                   const int64_t* dataset_offset_dev, const int32_t* status_dev,
                                   int64_t n_groups, int32_t group_points, const int32_t* groups_dev,
                                   const int64_t* unit_offset_dev, const int32_t* unit_group_dev, int64_t n_units,
                                   const int32_t* pair_point_dev, const int64_t* pair_partial_offset_dev,
                                   double outlier_likelihood, double* partial_dev, void* stream) {
    BiSpace space;
    int rc = bi_fill_space(&space, n_space, n_bins_host);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(method == BI_LOOKUP_LINEAR || method == BI_LOOKUP_PIECEWISE, "unknown lookup method %d", method);
    BI_REQUIRE(group_points == 1 || group_points == BI_MIX_GROUP_POINTS ||
                   (group_points == BI_MIX_GROUP_POINTS_WIDE && bi_mix_use_mma()),
               "group_points must be 1, %d or %d", BI_MIX_GROUP_POINTS, BI_MIX_GROUP_POINTS_WIDE);
    BI_REQUIRE(n_groups >= 0 && n_units >= 0, "negative size");
    if (n_groups == 0 || n_units == 0) return BI_OK;
    BI_REQUIRE(tmix_dev && ev_bin_dev && dataset_offset_dev && status_dev && groups_dev && unit_offset_dev &&
                   pair_partial_offset_dev && partial_dev,
               "bi_mixture_partials: NULL device pointer");
    BI_REQUIRE(method == BI_LOOKUP_PIECEWISE || ev_frac_dev, "bi_mixture_partials: ev_frac_dev is NULL");
    BiTsSpace sp;
    memset(&sp, 0, sizeof(sp));
    const int ns = method == BI_LOOKUP_LINEAR ? n_space : 0;
    sp.n_space = ns;
    sp.n_corner = 1 << ns;
    for (int c = 0; c < sp.n_corner; ++c) {
        int64_t off = 0;
        for (int d = 0; d < ns; ++d)
            if (((c >> (ns - 1 - d)) & 1) && space.n_bins[d] > 1) off += space.stride[d];
        sp.corner_off[c] = off * (ns == 0 ? 1 : (ns == 1 ? 2 : 4));       // packed mixture rows (bi_template_mix)
    }
    const int64_t row_doubles = space.n_cells * (ns == 0 ? 1 : (ns == 1 ? 2 : 4));
    BI_REQUIRE(((uintptr_t)tmix_dev & 31) == 0, "bi_mixture_partials: tmix_dev must be 32-byte aligned");
    const BiTsGroup* groups = reinterpret_cast<const BiTsGroup*>(groups_dev);
    cudaStream_t st = (cudaStream_t)stream;
#define BI_MIX_CASE(NPV, NSV)                                                                                          \
    if (group_points == NPV && ns == NSV)                                                                              \
        return bi_mix_launch<NPV, NSV>(tmix_dev, row_doubles, sp, ev_bin_dev, ev_frac_dev, ld_frac,                    \
                                       dataset_offset_dev, status_dev, n_groups, groups, unit_offset_dev,              \
                                       unit_group_dev, n_units, pair_point_dev, pair_partial_offset_dev,               \
                                       outlier_likelihood, partial_dev, st);
#define BI_MIXM_CASE(NSV)                                                                                              \
    if (group_points > 1 && ns == NSV && bi_mix_use_mma()) {                                                           \
        if (group_points == BI_MIX_GROUP_POINTS)                                                                       \
            return bi_mixm_launch<NSV, 1>(tmix_dev, row_doubles, sp, ev_bin_dev, ev_frac_dev, ld_frac,                 \
                                          dataset_offset_dev, status_dev, n_groups, groups, unit_offset_dev,           \
                                          unit_group_dev, n_units, pair_point_dev, pair_partial_offset_dev,            \
                                          outlier_likelihood, partial_dev, st);                                        \
        return bi_mixm_launch<NSV, 2>(tmix_dev, row_doubles, sp, ev_bin_dev, ev_frac_dev, ld_frac, dataset_offset_dev, \
                                      status_dev, n_groups, groups, unit_offset_dev, unit_group_dev, n_units,          \
                                      pair_point_dev, pair_partial_offset_dev, outlier_likelihood, partial_dev, st);   \
    }
    BI_MIXM_CASE(0) BI_MIXM_CASE(1) BI_MIXM_CASE(2) BI_MIXM_CASE(3) BI_MIXM_CASE(4)
#undef BI_MIXM_CASE
    BI_MIX_CASE(1, 0) BI_MIX_CASE(1, 1) BI_MIX_CASE(1, 2) BI_MIX_CASE(1, 3) BI_MIX_CASE(1, 4)
    BI_MIX_CASE(BI_MIX_GROUP_POINTS, 0) BI_MIX_CASE(BI_MIX_GROUP_POINTS, 1) BI_MIX_CASE(BI_MIX_GROUP_POINTS, 2)
    BI_MIX_CASE(BI_MIX_GROUP_POINTS, 3) BI_MIX_CASE(BI_MIX_GROUP_POINTS, 4)
#undef BI_MIX_CASE
    bi_set_error("bi_mixture_partials: unsupported configuration");
    return BI_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------
// The whole template-space evaluation in ONE call: K1 -> (template morph) -> K5 / K5b -> ragged finalize, four or
// five launches back to back with no host round trip in between (the counterpart of bi_unbinned_ll_batch).
// ---------------------------------------------------------------------------------------------
static inline int64_t bi_ts_align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

struct BiTemplateWorkspace { int64_t cell, frac, corner, weight, mus, row, coef, wterm, term_source, partial, tmix, order, total; };

static BiTemplateWorkspace bi_template_layout(int32_t D, int32_t S, int64_t P, int64_t n_partials, int64_t n_pairs,
                                              int64_t n_bins, int32_t mixture) {
    const int64_t C = (int64_t)1 << D, Dd = D > 0 ? D : 1, K = C * S;
    BiTemplateWorkspace w;
    int64_t o = 0;
    w.cell = o;        o += bi_ts_align256(P * Dd * 4);
    w.frac = o;        o += bi_ts_align256(P * Dd * 8);
    w.corner = o;      o += bi_ts_align256(P * C * 4);
    w.weight = o;      o += bi_ts_align256(P * C * 8);
    w.mus = o;         o += bi_ts_align256(P * S * 8);
    w.row = o;         o += bi_ts_align256(P * K * 4);
    w.coef = o;        o += bi_ts_align256(P * K * 8);
    w.wterm = o;       o += bi_ts_align256(P * K * 8);
    w.term_source = o; o += bi_ts_align256(K * 4);
    w.partial = o;     o += bi_ts_align256((n_partials > 0 ? n_partials : 1) * 8);
    w.tmix = o;        o += mixture ? bi_ts_align256(n_pairs * n_bins * 8 * 4) : 0;     // packed: up to 4 doubles per bin
    w.order = o;       o += mixture ? 0 : bi_ts_align256((n_pairs + 8) * 4);            // cell-major group order (toys)
    w.total = o;
    return w;
}

extern "C" int64_t bi_template_workspace_bytes(int32_t n_dims, int32_t n_sources, int64_t n_points, int64_t n_partials,
                                               int64_t n_pairs, int64_t n_template_bins, int32_t mixture) {
    if (n_dims < 0 || n_dims > BI_MAX_DIMS || n_sources < 1 || n_points < 0 || n_partials < 0 || n_pairs < 0 ||
        n_template_bins < 1)
        return -1;
    return bi_template_layout(n_dims, n_sources, n_points, n_partials, n_pairs, n_template_bins, mixture).total;
}

extern "C" int bi_template_ll_batch(int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host,
                                    int32_t n_sources, int64_t n_points,
                                    const double* zs_dev, const double* rate_mult_dev, const double* scale_dev,
                                    const double* eff_dev, const double* mus_anchor_dev, const uint8_t* allow_negative_host,
                                    const double* templates_dev, int64_t row_stride, int64_t bin_stride,
                                    int32_t n_space, const int32_t* n_bins_host, int32_t method, int32_t mixture,
                                    const int32_t* ev_bin_dev, const double* ev_frac_dev, int64_t ld_frac,
                                    const int64_t* dataset_offset_dev,
                                    int64_t n_groups, int32_t group_points, const int32_t* groups_dev,
                                    const int64_t* unit_offset_dev, const int32_t* unit_group_dev, int64_t n_units,
                                    const int32_t* pair_point_dev, const int64_t* pair_partial_offset_dev,
                                    int64_t n_pairs, int64_t n_partials, int64_t max_partials,
                                    double outlier_likelihood, void* workspace_dev, int64_t workspace_bytes,
                                    double* logl_dev, double* logsum_dev, double* musum_dev, int32_t* status_dev,
                                    void* stream) {
    BI_REQUIRE(n_points >= 0 && n_pairs >= 0, "negative size");
    if (n_points == 0 || n_pairs == 0) return BI_OK;
    BiSpace space;
    int rc = bi_fill_space(&space, n_space, n_bins_host);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(n_dims >= 0 && n_dims <= BI_MAX_DIMS, "n_dims=%d outside [0,%d]", n_dims, BI_MAX_DIMS);
    BI_REQUIRE(n_sources >= 1 && n_sources <= BI_MAX_SOURCES, "n_sources=%d outside [1,%d]", n_sources, BI_MAX_SOURCES);
    const int32_t K = (1 << n_dims) * n_sources;
    const BiTemplateWorkspace w = bi_template_layout(n_dims, n_sources, n_points, n_partials, n_pairs, space.n_cells, mixture);
    BI_REQUIRE(workspace_dev && workspace_bytes >= w.total, "workspace too small: %lld < %lld bytes",
               (long long)workspace_bytes, (long long)w.total);
    BI_REQUIRE(((uintptr_t)workspace_dev & 255) == 0, "workspace_dev must be 256-byte aligned");
    BI_REQUIRE(logl_dev && musum_dev && status_dev, "bi_template_ll_batch: NULL output pointer");
    char* base = (char*)workspace_dev;
    int32_t* row = (int32_t*)(base + w.row);
    double* coef = (double*)(base + w.coef);
    double* wterm = (double*)(base + w.wterm);
    int32_t* term_source = (int32_t*)(base + w.term_source);
    double* mus = (double*)(base + w.mus);
    double* partial = (double*)(base + w.partial);
    rc = bi_point_setup(n_dims, n_anchors_host, axes_host, n_sources, n_points, zs_dev, rate_mult_dev, scale_dev, eff_dev,
                        mus_anchor_dev, allow_negative_host, (int32_t*)(base + w.cell), (double*)(base + w.frac),
                        (int32_t*)(base + w.corner), (double*)(base + w.weight), mus, musum_dev, status_dev, row, coef,
                        wterm, term_source, stream);
    if (rc != BI_OK) return rc;
    if (n_units > 0) {
        if (mixture) {
            double* tmix = (double*)(base + w.tmix);
            rc = bi_template_mix(templates_dev, row_stride, bin_stride, n_space, n_bins_host, method, K, row, coef,
                                 status_dev, pair_point_dev, n_pairs, tmix, stream);
            if (rc != BI_OK) return rc;
            rc = bi_mixture_partials(tmix, n_space, n_bins_host, method, ev_bin_dev, ev_frac_dev, ld_frac, dataset_offset_dev,
                                     status_dev, n_groups, group_points, groups_dev, unit_offset_dev, unit_group_dev,
                                     n_units, pair_point_dev, pair_partial_offset_dev, outlier_likelihood, partial, stream);
        } else {
            // many single-pair groups (toy Monte Carlos): walk them in hypercube-cell order (BI_TS_ORDER=0: toy order)
            const int32_t* group_order = nullptr;
            const int32_t* n_ordered = nullptr;
            const char* order_env = getenv("BI_TS_ORDER");
            if (group_points == 1 && n_groups >= 4096 && n_groups <= n_pairs && n_dims > 0 && max_partials >= 1 &&
                max_partials < 64 && !(order_env && order_env[0] == '0')) {
                BiPlanDims dims;
                memset(&dims, 0, sizeof(dims));
                dims.n_dims = n_dims;
                int64_t n_cells = 1;
                for (int d = n_dims - 1; d >= 0; --d) {
                    dims.cells[d] = n_anchors_host[d] > 1 ? n_anchors_host[d] - 1 : 1;
                    dims.stride[d] = (int32_t)n_cells;
                    n_cells *= dims.cells[d];
                }
                if (n_cells > 1 && n_cells <= BI_PLAN_MAX_CELLS) {
                    int32_t* order = (int32_t*)(base + w.order);
                    static bool configured = false;
                    if (!configured) {
                        BI_CUDA_CHECK(cudaFuncSetAttribute(k_ts_order_groups, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                           (int)((2 * BI_PLAN_MAX_CELLS + 2) * sizeof(int))));
                        configured = true;
                    }
                    k_ts_order_groups<<<1, BI_PLAN_THREADS, (size_t)(2 * n_cells + 2) * sizeof(int), (cudaStream_t)stream>>>(
                        dims, (int)n_cells, n_groups, (const int32_t*)(base + w.cell), status_dev, pair_point_dev,
                        reinterpret_cast<const BiTsGroup*>(groups_dev), order + 8, order);
                    BI_LAUNCH_CHECK();
                    group_order = order + 8;
                    n_ordered = order;
                }
            }
            rc = bi_template_partials_impl(templates_dev, row_stride, bin_stride, n_space, n_bins_host, method, ev_bin_dev,
                                           ev_frac_dev, ld_frac, dataset_offset_dev, K, n_sources, row, coef, wterm,
                                           term_source, mus, status_dev, n_groups, group_points, groups_dev, unit_offset_dev,
                                           unit_group_dev, n_units, pair_point_dev, pair_partial_offset_dev,
                                           outlier_likelihood, partial, group_order, n_ordered, (int32_t)max_partials,
                                           nullptr, nullptr, nullptr, stream);
        }
        if (rc != BI_OK) return rc;
    }
    return bi_template_finalize(partial, pair_partial_offset_dev, pair_point_dev, musum_dev, status_dev, n_pairs,
                                max_partials, logl_dev, logsum_dev, stream);
}

// ---------------------------------------------------------------------------------------------
// A toy sweep (one parameter point per dataset) with the densities formed bin-major (bi_template_bm.cu), in ONE call:
// K1 -> point records -> k_bm_density -> k_template_partials on those densities (range test, canonical tree, rare path)
// -> ragged finalize.  Same workspace as bi_template_ll_batch (mixture = 0); results bit-identical to it.
// ---------------------------------------------------------------------------------------------
extern "C" int bi_template_ll_toys_bm(int32_t n_dims, const int32_t* n_anchors_host, const double* axes_host,
                                      int32_t n_sources, int64_t n_points,
                                      const double* zs_dev, const double* rate_mult_dev, const double* scale_dev,
                                      const double* eff_dev, const double* mus_anchor_dev, const uint8_t* allow_negative_host,
                                      const double* templates_dev, int64_t row_stride, int64_t bin_stride,
                                      const double* templates_bm_dev, int64_t n_rows,
                                      int32_t n_space, const int32_t* n_bins_host,
                                      const int32_t* ev_bin_dev, const double* ev_frac_dev, int64_t ld_frac,
                                      const int64_t* dataset_offset_dev,
                                      const int32_t* task_bin_dev, const int64_t* task_start_dev,
                                      const int32_t* task_count_dev, int64_t n_tasks,
                                      const int32_t* bm_toy_dev, const int32_t* bm_src_dev, const double* bm_frac_dev,
                                      int64_t ld_bm,
                                      int64_t n_groups, const int32_t* groups_dev, const int64_t* unit_offset_dev,
                                      const int32_t* unit_group_dev, int64_t n_units, const int32_t* pair_point_dev,
                                      const int64_t* pair_partial_offset_dev, int64_t n_partials, int64_t max_partials,
                                      double outlier_likelihood, void* workspace_dev, int64_t workspace_bytes,
                                      double* record_dev, double* density_dev,
                                      double* logl_dev, double* logsum_dev, double* musum_dev, int32_t* status_dev,
                                      void* stream) {
    BI_REQUIRE(n_points >= 0, "negative size");
    if (n_points == 0) return BI_OK;
    BI_REQUIRE(n_groups == n_points, "bi_template_ll_toys_bm: one pair per dataset and point (got %lld groups for %lld points)",
               (long long)n_groups, (long long)n_points);
    BiSpace space;
    int rc = bi_fill_space(&space, n_space, n_bins_host);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(n_dims >= 0 && n_dims <= BI_MAX_DIMS, "n_dims=%d outside [0,%d]", n_dims, BI_MAX_DIMS);
    BI_REQUIRE(n_sources >= 1 && n_sources <= BI_MAX_SOURCES, "n_sources=%d outside [1,%d]", n_sources, BI_MAX_SOURCES);
    const int32_t K = (1 << n_dims) * n_sources;
    const BiTemplateWorkspace w = bi_template_layout(n_dims, n_sources, n_points, n_partials, n_points, space.n_cells, 0);
    BI_REQUIRE(workspace_dev && workspace_bytes >= w.total, "workspace too small: %lld < %lld bytes",
               (long long)workspace_bytes, (long long)w.total);
    BI_REQUIRE(((uintptr_t)workspace_dev & 255) == 0, "workspace_dev must be 256-byte aligned");
    BI_REQUIRE(logl_dev && musum_dev && status_dev, "bi_template_ll_toys_bm: NULL output pointer");
    char* base = (char*)workspace_dev;
    int32_t* term_source = (int32_t*)(base + w.term_source);
    double* mus = (double*)(base + w.mus);
    double* partial = (double*)(base + w.partial);
    rc = bi_point_setup(n_dims, n_anchors_host, axes_host, n_sources, n_points, zs_dev, rate_mult_dev, scale_dev, eff_dev,
                        mus_anchor_dev, allow_negative_host, (int32_t*)(base + w.cell), (double*)(base + w.frac),
                        (int32_t*)(base + w.corner), (double*)(base + w.weight), mus, musum_dev, status_dev, nullptr, nullptr,
                        nullptr, term_source, stream);           // no per-point term lists: the records carry the inputs
    if (rc != BI_OK) return rc;
    if (n_units > 0) {
        rc = bi_template_bm_density(templates_bm_dev, n_rows, n_space, n_dims, n_anchors_host, n_sources, n_points,
                                    (const int32_t*)(base + w.cell), (const double*)(base + w.frac), mus, status_dev,
                                    task_bin_dev, task_start_dev, task_count_dev, n_tasks, bm_toy_dev, bm_src_dev,
                                    bm_frac_dev, ld_bm, record_dev, density_dev, stream);
        if (rc != BI_OK) return rc;
        rc = bi_template_partials_impl(templates_dev, row_stride, bin_stride, n_space, n_bins_host, BI_LOOKUP_LINEAR,
                                       ev_bin_dev, ev_frac_dev, ld_frac, dataset_offset_dev, K, n_sources, nullptr, nullptr,
                                       nullptr, term_source, mus, status_dev, n_groups, 1, groups_dev, unit_offset_dev,
                                       unit_group_dev, n_units, pair_point_dev, pair_partial_offset_dev,
                                       outlier_likelihood, partial, nullptr, nullptr, 1, density_dev,
                                       (const int32_t*)(base + w.corner), (const double*)(base + w.weight), stream);
        if (rc != BI_OK) return rc;
    }
    return bi_template_finalize(partial, pair_partial_offset_dev, pair_point_dev, musum_dev, status_dev, n_points,
                                max_partials, logl_dev, logsum_dev, stream);
}
