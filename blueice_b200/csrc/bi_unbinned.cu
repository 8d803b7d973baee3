// blueice_b200 -- K2: fused anchor-morph + mixture density + log + reduce for the unbinned likelihood.
//
// Replaces, for a batch of P parameter points over N events (blueice/likelihood.py:355-356,678-690 and
// scipy/interpolate/_rgi.py:520-549):
//     ps[s,i]  = sum_c  A[corner_c, s, i] * w_c          (multilinear morph of the per-event pdf tensor)
//     p_i      = nansum_s mu_s * ps[s,i];  p_i <- outlier_likelihood where !(p_i > 0)
//     partial  = sum_i log p_i                            (per 512-event superblock, canonical order)
//
// k_unbinned_stream: lanes = events (16-byte loads straight from HBM/L2), one warp per (point, superblock).
// HBM-bound when few points share the data.  It is the general kernel (any number of sources, up to 32 corners);
// batches within the limits of the DMMA kernel (bi_unbinned_mma.cuh) run there and give BIT-IDENTICAL
// per-superblock partials for the same point.
//
// Canonical arithmetic.  For each block of 32 consecutive events:
//   quad k (4 events) = fl(fl(p0*p1)*fl(p2*p3)) = m_k * 2^e_k (unbounded-exponent semantics);
//   block (8 quads)   = binary-tree product M_b of the quad mantissas, E_b = sum e_k;
//   a block with a density that is not a normal positive double contributes M_b = 1, E_b = 0 and
//   instead adds the same tree over log(p_i) to slow_j (sequentially, in block order);
//   superblock (16 blocks) = binary-tree product M_j of the M_b, E_j = sum E_b;
//   S_j = fma(E_j, LN2_LO, fma(E_j, LN2_HI, log(M_j))) + slow_j        -> ONE log per 512 events
// events >= N count as p = 1 (exact identity).
#include "bi_common.cuh"


// ---------------------------------------------------------------------------------------------
// slow per-event density: exact nansum semantics + outlier replacement (likelihood.py:686-689)
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ double bi_slow_density_global(const double* __restrict__ A, int64_t ld, int S, int C,
                                                      const int32_t* __restrict__ corner_p,
                                                      const double* __restrict__ weight_p,
                                                      const double* __restrict__ mu, int64_t ev, double outlier) {
    double acc = 0.0;
    for (int s = 0; s < S; ++s) {
        double ps = 0.0;
        for (int c = 0; c < C; ++c)
            ps = fma(A[((int64_t)corner_p[c] * S + s) * ld + ev], weight_p[c], ps);
        const double t = __dmul_rn(mu[s], ps);
        if (t == t) acc = __dadd_rn(acc, t);       // nansum: NaN terms count as 0
    }
    return bi_fix_density(acc, outlier);
}

// =============================================================================================
// Streaming kernel: one warp per (point, superblock) task, lane = 2 consecutive events per iteration
// =============================================================================================
template <int C>
__global__ void __launch_bounds__(256, (C <= 8) ? 2 : 1)
k_unbinned_stream(const double* __restrict__ A, int64_t ld, int64_t N, int S,
                  const int32_t* __restrict__ point_index, int64_t n_points, int64_t n_super,
                  const int32_t* __restrict__ corner, const double* __restrict__ weight,
                  const double* __restrict__ mus, const int32_t* __restrict__ status,
                  double outlier, double* __restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_tasks = n_points * n_super;
    constexpr int CH = (C < 8) ? C : 8;                       // corners per load batch
    constexpr int SG = (C >= 8) ? 1 : (8 / C);                // sources per load batch (SG * CH <= 8 loads in flight)

    for (int64_t task = warp_global; task < n_tasks; task += n_warps) {
        const int64_t k = task / n_super;
        const int64_t j = task - k * n_super;
        const int64_t p = point_index ? (int64_t)point_index[k] : k;
        if (status[p] != 0) continue;                         // warp-uniform

        const int64_t ev0 = j * BI_SUPERBLOCK;
        double w[C];
        const double* ptr[C];                                 // this lane's first event, source 0
#pragma unroll
        for (int c = 0; c < C; ++c) {
            w[c] = weight[p * C + c];
            ptr[c] = A + (int64_t)corner[p * C + c] * S * ld + ev0 + 2 * lane;
        }
        const double* mu = mus + p * S;

        double l1 = 1.0, l2 = 1.0, l3 = 1.0, m_super = 1.0, slow_sum = 0.0;
        int e_super = 0;
        int64_t left = N - ev0;
        if (left > BI_SUPERBLOCK) left = BI_SUPERBLOCK;
        const int n_iter = (int)((left + 63) >> 6);
#pragma unroll 1
        for (int it = 0; it < n_iter; ++it) {
            const int64_t e = ev0 + (int64_t)it * 64 + 2 * lane;
            const bool in_ld = e < ld;                        // ld even -> e + 1 < ld as well
            double p0 = 0.0, p1 = 0.0;
            for (int s0 = 0; s0 < S; s0 += SG) {
                double acc0[SG], acc1[SG];
#pragma unroll
                for (int c0 = 0; c0 < C; c0 += CH) {
                    double2 v[SG][CH];
#pragma unroll
                    for (int g = 0; g < SG; ++g) {
                        if (s0 + g < S) {                     // warp-uniform
                            const int64_t off = (int64_t)(s0 + g) * ld + it * 64;
#pragma unroll
                            for (int c = 0; c < CH; ++c)
                                v[g][c] = in_ld ? __ldg(reinterpret_cast<const double2*>(ptr[c0 + c] + off))
                                                : make_double2(0.0, 0.0);
                        }
                    }
#pragma unroll
                    for (int g = 0; g < SG; ++g) {
                        if (s0 + g < S) {
                            double ps0 = (c0 == 0) ? 0.0 : acc0[g], ps1 = (c0 == 0) ? 0.0 : acc1[g];
#pragma unroll
                            for (int c = 0; c < CH; ++c) {
                                ps0 = fma(v[g][c].x, w[c0 + c], ps0);
                                ps1 = fma(v[g][c].y, w[c0 + c], ps1);
                            }
                            acc0[g] = ps0; acc1[g] = ps1;
                        }
                    }
                }
#pragma unroll
                for (int g = 0; g < SG; ++g) {
                    if (s0 + g < S) {
                        const double m = mu[s0 + g];
                        p0 = fma(m, acc0[g], p0);
                        p1 = fma(m, acc1[g], p1);
                    }
                }
            }
            if (e >= N) p0 = 1.0;
            if (e + 1 >= N) p1 = 1.0;
            bool ok0 = bi_is_normal_positive(p0);
            if (!ok0) { p0 = bi_slow_density_global(A, ld, S, C, corner + p * C, weight + p * C, mu, e, outlier); ok0 = bi_is_normal_positive(p0); }
            bool ok1 = bi_is_normal_positive(p1);
            if (!ok1) { p1 = bi_slow_density_global(A, ld, S, C, corner + p * C, weight + p * C, mu, e + 1, outlier); ok1 = bi_is_normal_positive(p1); }
            const unsigned bad = __ballot_sync(BI_FULL_MASK, !(ok0 && ok1));
            // a block (half warp) with an abnormal density contributes 1 to the product and its log tree to slow_sum
            const bool half_bad = ((lane < 16) ? (bad & 0xffffu) : (bad >> 16)) != 0;

            // canonical quad = the two events of this lane and of lane ^ 1
            const double f0 = half_bad ? 1.0 : p0, f1 = half_bad ? 1.0 : p1;
            const double q2 = __dmul_rn(f0, f1);
            const double q2o = __shfl_xor_sync(BI_FULL_MASK, q2, 1);
            const double q4 = __dmul_rn(q2, q2o);
            const bool direct_ok = bi_is_normal_positive(q2) && bi_is_normal_positive(q2o) && bi_is_normal_positive(q4);
            double q; int E;
            if (__ballot_sync(BI_FULL_MASK, !direct_ok) == 0) {
                bi_split(q4, &q, &E);
            } else {                                          // warp-uniform, rare: go through the mantissas
                const double g0 = __shfl_xor_sync(BI_FULL_MASK, f0, 1), g1 = __shfl_xor_sync(BI_FULL_MASK, f1, 1);
                if (lane & 1) bi_quad_from_mantissas(g0, g1, f0, f1, &q, &E);
                else bi_quad_from_mantissas(f0, f1, g0, g1, &q, &E);
            }
            // binary tree over the 8 quads of each block (xor 2, 4, 8) and over the two blocks (xor 16)
#pragma unroll
            for (int x = 2; x < 32; x <<= 1) {
                q = __dmul_rn(q, __shfl_xor_sync(BI_FULL_MASK, q, x));
                E += __shfl_xor_sync(BI_FULL_MASK, E, x);
            }
            e_super += E;
            if (bad) {                                        // warp-uniform, rare
                double l = __dadd_rn(log(p0), log(p1));
#pragma unroll
                for (int x = 1; x < 16; x <<= 1) l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, x));
                const double la = __shfl_sync(BI_FULL_MASK, l, 0), lb = __shfl_sync(BI_FULL_MASK, l, 16);
                if (bad & 0xffffu) slow_sum = __dadd_rn(slow_sum, la);
                if (bad >> 16) slow_sum = __dadd_rn(slow_sum, lb);
            }
            // binary counter over the 8 iterations = upper levels of the superblock product tree
            double v = q;
            if (it & 1) {
                v = __dmul_rn(l1, v);
                if (it & 2) {
                    v = __dmul_rn(l2, v);
                    if (it & 4) m_super = __dmul_rn(l3, v); else l3 = v;
                } else l2 = v;
            } else l1 = v;
        }
        if (n_iter < BI_SUPERBLOCK / 64) {                    // superblock ends early: missing blocks count as 1
            double v = 1.0;
            if (n_iter & 1) v = __dmul_rn(l1, v);
            if (n_iter & 2) v = __dmul_rn(l2, v);
            if (n_iter & 4) v = __dmul_rn(l3, v);
            m_super = v;
        }
        if (lane == 0) partial[p * n_super + j] = __dadd_rn(bi_block_log(m_super, e_super), slow_sum);
    }
}

// =============================================================================================
// Finalize: canonical total over the superblock partials, one CTA per point
// =============================================================================================
__global__ void __launch_bounds__(256)
k_unbinned_finalize(const double* __restrict__ partial, int64_t n_super, const double* __restrict__ musum,
                    const int32_t* __restrict__ status, double* __restrict__ logl, double* __restrict__ logsum) {
    __shared__ double warp_tot[8];
    const int64_t p = blockIdx.x;
    const int t = threadIdx.x;
    if (status[p] != 0) {                                     // block-uniform
        if (t == 0) {
            logl[p] = -__longlong_as_double(0x7ff0000000000000LL);
            if (logsum) logsum[p] = 0.0;
        }
        return;
    }
    double u = bi_strided_sum(partial + p * n_super, t, n_super);
#pragma unroll
    for (int x = 1; x < 32; x <<= 1) u = __dadd_rn(u, __shfl_xor_sync(BI_FULL_MASK, u, x));
    if ((t & 31) == 0) warp_tot[t >> 5] = u;
    __syncthreads();
    if (t == 0) {
        const double total = __dadd_rn(__dadd_rn(__dadd_rn(warp_tot[0], warp_tot[1]), __dadd_rn(warp_tot[2], warp_tot[3])),
                                       __dadd_rn(__dadd_rn(warp_tot[4], warp_tot[5]), __dadd_rn(warp_tot[6], warp_tot[7])));
        logl[p] = __dadd_rn(-musum[p], total);               // likelihood.py:690
        if (logsum) logsum[p] = total;                       // sum_i log p_i alone (event-sharded evaluation)
    }
}

// the same total with ONE WARP per point (lane l plays threads l, l + 32, ... of the 256-thread finalize): for short
// partial lists the cost of k_unbinned_finalize is the scheduling of P 256-thread CTAs, not the sums
__global__ void __launch_bounds__(256)
k_unbinned_finalize_warp(const double* __restrict__ partial, int64_t n_super, const double* __restrict__ musum,
                         const int32_t* __restrict__ status, int64_t n_points, double* __restrict__ logl,
                         double* __restrict__ logsum) {
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n_points) return;
    if (status[p] != 0) {
        if (lane == 0) {
            logl[p] = -__longlong_as_double(0x7ff0000000000000LL);
            if (logsum) logsum[p] = 0.0;
        }
        return;
    }
    const double* src = partial + p * n_super;
    double w[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        double u = bi_strided_sum(src, v * 32 + lane, n_super);
#pragma unroll
        for (int x = 1; x < 32; x <<= 1) u = __dadd_rn(u, __shfl_xor_sync(BI_FULL_MASK, u, x));
        w[v] = u;
    }
    if (lane == 0) {
        const double total = __dadd_rn(__dadd_rn(__dadd_rn(w[0], w[1]), __dadd_rn(w[2], w[3])),
                                       __dadd_rn(__dadd_rn(w[4], w[5]), __dadd_rn(w[6], w[7])));
        logl[p] = __dadd_rn(-musum[p], total);               // likelihood.py:690
        if (logsum) logsum[p] = total;
    }
}

// =============================================================================================
// ps[S, N] for one point, reference operation order (value = value + V * w), full_output=True
// =============================================================================================
__global__ void __launch_bounds__(256)
k_unbinned_ps(const double* __restrict__ A, int64_t ld, int64_t N, int S, int C,
              const int32_t* __restrict__ corner, const double* __restrict__ weight,
              double* __restrict__ out, int64_t ld_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    for (int s = 0; s < S; ++s) {
        double acc;
        if (C == 1) {
            acc = A[((int64_t)corner[0] * S + s) * ld + i];
        } else {
            acc = 0.0;
            for (int c = 0; c < C; ++c)
                acc = __dadd_rn(acc, __dmul_rn(A[((int64_t)corner[c] * S + s) * ld + i], weight[c]));
        }
        out[(int64_t)s * ld_out + i] = acc;
    }
}

// =============================================================================================
// C-ABI
// =============================================================================================
static int bi_check_tensor(const double* A, int64_t ld, int64_t N, int32_t S, int32_t C) {
    BI_REQUIRE(N >= 0, "n_events < 0");
    BI_REQUIRE(N == 0 || A, "ps_anchor_dev is NULL");
    BI_REQUIRE(ld >= N && (ld % 2) == 0, "ld_events=%lld must be even and >= n_events=%lld", (long long)ld, (long long)N);
    BI_REQUIRE(((uintptr_t)A & 15) == 0, "ps_anchor_dev must be 16-byte aligned");
    BI_REQUIRE(S >= 1 && S <= BI_MAX_SOURCES, "n_sources=%d outside [1,%d]", S, BI_MAX_SOURCES);
    BI_REQUIRE(C >= 1 && C <= (1 << BI_MAX_DIMS) && (C & (C - 1)) == 0, "n_corners=%d is not a power of two <= %d", C, 1 << BI_MAX_DIMS);
    return BI_OK;
}

template <int C>
static int bi_launch_stream(const double* A, int64_t ld, int64_t N, int32_t S, const int32_t* point_index,
                            int64_t P, int64_t n_super, const int32_t* corner, const double* weight,
                            const double* mus, const int32_t* status, double outlier, double* partial,
                            cudaStream_t st) {
    const int64_t n_tasks = P * n_super;
    int64_t blocks = (n_tasks + 7) / 8;                       // 8 warps per CTA
    const int64_t max_blocks = 148 * 8 * 4;                   // a few waves of resident CTAs; warps grid-stride beyond
    if (blocks > max_blocks) blocks = max_blocks;
    k_unbinned_stream<C><<<(unsigned)blocks, 256, 0, st>>>(A, ld, N, S, point_index, P, n_super, corner, weight,
                                                           mus, status, outlier, partial);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

extern "C" int bi_unbinned_partials_stream(const double* ps_anchor_dev, int64_t ld_events, int64_t n_events,
                                           int32_t n_sources, int32_t n_corners,
                                           const int32_t* point_index_dev, int64_t n_points,
                                           const int32_t* corner_dev, const double* weight_dev,
                                           const double* mus_dev, const int32_t* status_dev,
                                           double outlier_likelihood, double* partial_dev, void* stream) {
    int rc = bi_check_tensor(ps_anchor_dev, ld_events, n_events, n_sources, n_corners);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(n_points >= 0, "n_points < 0");
    const int64_t n_super = bi_num_superblocks(n_events);
    if (n_points == 0 || n_super == 0) return BI_OK;
    BI_REQUIRE(corner_dev && weight_dev && mus_dev && status_dev && partial_dev, "bi_unbinned_partials_stream: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
#define BI_STREAM_CASE(CC)                                                                                    \
    case CC:                                                                                                  \
        return bi_launch_stream<CC>(ps_anchor_dev, ld_events, n_events, n_sources, point_index_dev, n_points, \
                                    n_super, corner_dev, weight_dev, mus_dev, status_dev, outlier_likelihood, \
                                    partial_dev, st);
    switch (n_corners) {
        BI_STREAM_CASE(1)
        BI_STREAM_CASE(2)
        BI_STREAM_CASE(4)
        BI_STREAM_CASE(8)
        BI_STREAM_CASE(16)
        BI_STREAM_CASE(32)
    }
#undef BI_STREAM_CASE
    bi_set_error("unsupported n_corners=%d", n_corners);
    return BI_ERR_UNSUPPORTED;
}

extern "C" int bi_unbinned_finalize(const double* partial_dev, int64_t n_super, const double* musum_dev,
                                    const int32_t* status_dev, int64_t n_points, double* logl_dev,
                                    double* logsum_dev, void* stream) {
    BI_REQUIRE(n_points >= 0 && n_super >= 0, "negative size");
    if (n_points == 0) return BI_OK;
    BI_REQUIRE(musum_dev && status_dev && logl_dev && (n_super == 0 || partial_dev), "bi_unbinned_finalize: NULL pointer");
    if (n_super <= 1024 && n_points >= 64)                       // many short lists: one warp per point
        k_unbinned_finalize_warp<<<(unsigned)((n_points + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
            partial_dev, n_super, musum_dev, status_dev, n_points, logl_dev, logsum_dev);
    else
        k_unbinned_finalize<<<(unsigned)n_points, 256, 0, (cudaStream_t)stream>>>(partial_dev, n_super, musum_dev,
                                                                                  status_dev, logl_dev, logsum_dev);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

extern "C" int bi_unbinned_ps(const double* ps_anchor_dev, int64_t ld_events, int64_t n_events,
                              int32_t n_sources, int32_t n_corners, const int32_t* corner_dev,
                              const double* weight_dev, double* ps_out_dev, int64_t ld_out, void* stream) {
    int rc = bi_check_tensor(ps_anchor_dev, ld_events, n_events, n_sources, n_corners);
    if (rc != BI_OK) return rc;
    if (n_events == 0) return BI_OK;
    BI_REQUIRE(corner_dev && weight_dev && ps_out_dev && ld_out >= n_events, "bi_unbinned_ps: bad arguments");
    const int64_t blocks = (n_events + 255) / 256;
    k_unbinned_ps<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ps_anchor_dev, ld_events, n_events, n_sources,
                                                                      n_corners, corner_dev, weight_dev, ps_out_dev, ld_out);
    BI_LAUNCH_CHECK();
    return BI_OK;
}
