// blueice_b200 -- K2: fused anchor-morph + mixture density + log + reduce for the unbinned likelihood.
//
// Replaces, for a batch of P parameter points over N events (blueice/likelihood.py:355-356,678-690 and
// scipy/interpolate/_rgi.py:520-549):
//     ps[s,i]  = sum_c  A[corner_c, s, i] * w_c          (multilinear morph of the per-event pdf tensor)
//     p_i      = nansum_s mu_s * ps[s,i];  p_i <- outlier_likelihood where !(p_i > 0)
//     partial  = sum_i log p_i                            (per 512-event superblock, canonical order)
//
// Two kernels produce BIT-IDENTICAL per-superblock partials for the same point:
//   k_unbinned_stream  : lanes = events (16-byte loads straight from HBM/L2), one warp per
//                        (point, superblock).  HBM-bound when few points share the data.
//   k_unbinned_grouped : threads = points that share one hypercube cell; the cell's C*S slabs of an
//                        event tile are staged in shared memory by TMA bulk copies
//                        (cp.async.bulk + mbarrier, multi-stage) and broadcast to all threads.
//                        FP64-pipe-bound for profile scans / toy batches.
//
// Canonical arithmetic.  For each block of 32 consecutive events:
//   fast:  p_i = m_i * 2^e_i;  M = adjacent-pair binary-tree product of the m_i;  E = sum e_i;
//          L_b = fma(E, LN2_LO, fma(E, LN2_HI, log(M)))          -> ONE log per 32 events
//   slow (some p_i not a normal positive double): L_b = same tree shape over log(p_i)
// superblock S_j = sequential sum of its 16 L_b; events >= N count as p = 1.
#include "bi_common.cuh"

#define BI_GROUP_THREADS 128

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

__device__ __noinline__ double bi_slow_density_tile(const double* tile, int T, int S, int C, int ev,
                                                    const double* __restrict__ weight_p,
                                                    const double* __restrict__ mu, double outlier) {
    double acc = 0.0;
    for (int s = 0; s < S; ++s) {
        double ps = 0.0;
        for (int c = 0; c < C; ++c) ps = fma(tile[(size_t)(c * S + s) * T + ev], weight_p[c], ps);
        const double t = __dmul_rn(mu[s], ps);
        if (t == t) acc = __dadd_rn(acc, t);
    }
    return bi_fix_density(acc, outlier);
}

// =============================================================================================
// Streaming kernel: one warp per (point, superblock) task, lane = 2 consecutive events
// =============================================================================================
template <int C>
__global__ void __launch_bounds__(256)
k_unbinned_stream(const double* __restrict__ A, int64_t ld, int64_t N, int S,
                  const int32_t* __restrict__ point_index, int64_t n_points, int64_t n_super,
                  const int32_t* __restrict__ corner, const double* __restrict__ weight,
                  const double* __restrict__ mus, const int32_t* __restrict__ status,
                  double outlier, double* __restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_tasks = n_points * n_super;

    for (int64_t task = warp_global; task < n_tasks; task += n_warps) {
        const int64_t k = task / n_super;
        const int64_t j = task - k * n_super;
        const int64_t p = point_index ? (int64_t)point_index[k] : k;
        if (status[p] != 0) continue;                         // warp-uniform

        double w[C];
        const double* base[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            w[c] = weight[p * C + c];
            base[c] = A + (int64_t)corner[p * C + c] * S * ld;
        }
        const double* mu = mus + p * S;

        double s_sum = 0.0;
        const int64_t ev0 = j * BI_SUPERBLOCK;
#pragma unroll 1
        for (int it = 0; it < BI_SUPERBLOCK / 64; ++it) {
            const int64_t chunk = ev0 + (int64_t)it * 64;
            if (chunk >= N) break;                            // warp-uniform
            const int64_t e = chunk + 2 * lane;
            const bool in_ld = e < ld;                        // ld even -> e + 1 < ld as well
            double p0 = 0.0, p1 = 0.0;
            for (int s = 0; s < S; ++s) {
                constexpr int CH = C < 8 ? C : 8;              // corners loaded per batch (all in flight)
                double ps0 = 0.0, ps1 = 0.0;
#pragma unroll
                for (int c0 = 0; c0 < C; c0 += CH) {
                    double2 v[CH];
#pragma unroll
                    for (int c = 0; c < CH; ++c)
                        v[c] = in_ld ? __ldg(reinterpret_cast<const double2*>(base[c0 + c] + (int64_t)s * ld + e))
                                     : make_double2(0.0, 0.0);
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        ps0 = fma(v[c].x, w[c0 + c], ps0);
                        ps1 = fma(v[c].y, w[c0 + c], ps1);
                    }
                }
                const double m = mu[s];
                p0 = fma(m, ps0, p0);
                p1 = fma(m, ps1, p1);
            }
            if (e >= N) p0 = 1.0;
            if (e + 1 >= N) p1 = 1.0;
            bool ok0 = bi_is_normal_positive(p0);
            if (!ok0) { p0 = bi_slow_density_global(A, ld, S, C, corner + p * C, weight + p * C, mu, e, outlier); ok0 = bi_is_normal_positive(p0); }
            bool ok1 = bi_is_normal_positive(p1);
            if (!ok1) { p1 = bi_slow_density_global(A, ld, S, C, corner + p * C, weight + p * C, mu, e + 1, outlier); ok1 = bi_is_normal_positive(p1); }
            const unsigned bad = __ballot_sync(BI_FULL_MASK, !(ok0 && ok1));

            double m0, m1; int e0, e1;
            bi_split(ok0 ? p0 : 1.0, &m0, &e0);
            bi_split(ok1 ? p1 : 1.0, &m1, &e1);
            double q = __dmul_rn(m0, m1);
            int E = e0 + e1;
#pragma unroll
            for (int x = 1; x < 16; x <<= 1) {
                q = __dmul_rn(q, __shfl_xor_sync(BI_FULL_MASK, q, x));
                E += __shfl_xor_sync(BI_FULL_MASK, E, x);
            }
            double L = bi_block_log(q, E);
            if (bad) {                                        // warp-uniform, rare
                double l = __dadd_rn(log(p0), log(p1));
#pragma unroll
                for (int x = 1; x < 16; x <<= 1) l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, x));
                const unsigned half_bad = (lane < 16) ? (bad & 0xffffu) : (bad >> 16);
                if (half_bad) L = l;
            }
            const double la = __shfl_sync(BI_FULL_MASK, L, 0);
            const double lb = __shfl_sync(BI_FULL_MASK, L, 16);
            s_sum = __dadd_rn(s_sum, la);
            s_sum = __dadd_rn(s_sum, lb);
        }
        if (lane == 0) partial[p * n_super + j] = s_sum;
    }
}

// =============================================================================================
// Grouped kernel: threads = points of one hypercube cell, event tiles staged by TMA bulk copies
// =============================================================================================
__device__ __forceinline__ uint32_t bi_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void bi_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bi_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bi_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bi_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool bi_mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bi_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a byte-count bug would otherwise hang the GPU; after ~2 s of polling the kernel traps
// (a reported CUDA error) instead of spinning forever.
__device__ __forceinline__ void bi_mbar_wait(uint64_t* bar, unsigned parity) {
    if (bi_mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!bi_mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bi_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     bi_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(bi_smem_u32(bar))
                 : "memory");
}

template <int C, int Q>
struct BiGroupRegs {
    double w[Q][C];
    double mu[Q][BI_GROUP_MAX_SOURCES];
};

// densities of 4 consecutive events (tile offsets e..e+3) for the Q points of this thread
template <int C, int Q>
__device__ __forceinline__ void bi_quad_density(const double* __restrict__ tile, int T, int S, int e,
                                                const BiGroupRegs<C, Q>& g, double (&p)[Q][4]) {
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) p[q][k] = 0.0;
#pragma unroll
    for (int s = 0; s < BI_GROUP_MAX_SOURCES; ++s) {
        if (s >= S) break;
        double ps[Q][4];
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
            for (int k = 0; k < 4; ++k) ps[q][k] = 0.0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const double* row = tile + (size_t)(c * S + s) * T + e;
            const double2 a = *reinterpret_cast<const double2*>(row);
            const double2 b = *reinterpret_cast<const double2*>(row + 2);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                ps[q][0] = fma(a.x, g.w[q][c], ps[q][0]);
                ps[q][1] = fma(a.y, g.w[q][c], ps[q][1]);
                ps[q][2] = fma(b.x, g.w[q][c], ps[q][2]);
                ps[q][3] = fma(b.y, g.w[q][c], ps[q][3]);
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
            for (int k = 0; k < 4; ++k) p[q][k] = fma(g.mu[q][s], ps[q][k], p[q][k]);
    }
}

// One canonical block (32 events at tile offset e0) for the Q points of this thread.
// n_valid = number of real events in the tile (events at tile offset >= n_valid count as p = 1).
template <int C, int Q>
__device__ __forceinline__ void bi_group_block(const double* __restrict__ tile, int T, int S, int e0, int n_valid,
                                               const BiGroupRegs<C, Q>& g, const double* const (&wp)[Q],
                                               const double* const (&mp)[Q], double outlier, double (&L)[Q]) {
    double l3[Q], l4[Q], l5[Q], M[Q];
    int E[Q];
    bool slow[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) { E[q] = 0; slow[q] = false; l3[q] = l4[q] = l5[q] = M[q] = 1.0; }

#pragma unroll 1
    for (int o = 0; o < 8; ++o) {
        const int e = e0 + 4 * o;
        double p[Q][4];
        bi_quad_density<C, Q>(tile, T, S, e, g, p);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            double m[4];
            int ex[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double pk = (e + k < n_valid) ? p[q][k] : 1.0;
                if (!bi_is_normal_positive(pk)) {
                    pk = bi_slow_density_tile(tile, T, S, C, e + k, wp[q], mp[q], outlier);
                    if (!bi_is_normal_positive(pk)) { slow[q] = true; pk = 1.0; }
                }
                bi_split(pk, &m[k], &ex[k]);
            }
            double v = __dmul_rn(__dmul_rn(m[0], m[1]), __dmul_rn(m[2], m[3]));
            E[q] += (ex[0] + ex[1]) + (ex[2] + ex[3]);
            // binary-counter merge = adjacent-pair tree over quads
            if (o & 1) {
                v = __dmul_rn(l3[q], v);
                if (o & 2) {
                    v = __dmul_rn(l4[q], v);
                    if (o & 4) M[q] = __dmul_rn(l5[q], v); else l5[q] = v;
                } else l4[q] = v;
            } else l3[q] = v;
        }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) L[q] = bi_block_log(M[q], E[q]);

    // rare: redo flagged points as a tree sum of logs (same tree shape)
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        if (!slow[q]) continue;
        double s3 = 0.0, s4 = 0.0, s5 = 0.0, tot = 0.0;
        for (int o = 0; o < 8; ++o) {
            double lg[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int e = e0 + 4 * o + k;
                double pk = 1.0;
                if (e < n_valid) pk = bi_slow_density_tile(tile, T, S, C, e, wp[q], mp[q], outlier);
                lg[k] = log(pk);
            }
            double v = __dadd_rn(__dadd_rn(lg[0], lg[1]), __dadd_rn(lg[2], lg[3]));
            if (o & 1) {
                v = __dadd_rn(s3, v);
                if (o & 2) {
                    v = __dadd_rn(s4, v);
                    if (o & 4) tot = __dadd_rn(s5, v); else s5 = v;
                } else s4 = v;
            } else s3 = v;
        }
        L[q] = tot;
    }
}

template <int C, int Q>
__device__ __forceinline__ void bi_group_run(const double* __restrict__ A, int64_t ld, int64_t N, int S, int T,
                                             int n_stages, const int32_t* __restrict__ group_points, int4 wk,
                                             int64_t n_super, const int32_t* __restrict__ corner,
                                             const double* __restrict__ weight, const double* __restrict__ mus,
                                             double outlier, double* __restrict__ partial,
                                             double* smem_tiles, uint64_t* full_bar) {
    const int tid = threadIdx.x;
    const int first = wk.x, count = wk.y;
    const int64_t sb_begin = wk.z, sb_end = wk.w;
    const int64_t ev_begin = sb_begin * BI_SUPERBLOCK;
    int64_t ev_end = sb_end * BI_SUPERBLOCK;
    if (ev_end > N) ev_end = N;
    const int n_tiles = (int)((ev_end - ev_begin + T - 1) / T);
    const size_t tile_doubles = (size_t)C * S * T;

    // ---- per-thread point registers ----
    BiGroupRegs<C, Q> g;
    int64_t pidx[Q];
    bool active[Q];
    const double* wp[Q];
    const double* mp[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int k = tid + q * BI_GROUP_THREADS;
        active[q] = k < count;
        pidx[q] = group_points[first + (active[q] ? k : 0)];
        wp[q] = weight + pidx[q] * C;
        mp[q] = mus + pidx[q] * S;
#pragma unroll
        for (int c = 0; c < C; ++c) g.w[q][c] = weight[pidx[q] * C + c];
#pragma unroll
        for (int s = 0; s < BI_GROUP_MAX_SOURCES; ++s) g.mu[q][s] = (s < S) ? mus[pidx[q] * S + s] : 0.0;
    }

    // ---- producer (warp 0): TMA bulk copies of the C*S slab tiles of one event tile ----
    const int64_t lead = group_points[first];
    auto issue = [&](int tile_idx) {
        const int st = tile_idx % n_stages;
        const int64_t ev = ev_begin + (int64_t)tile_idx * T;
        int64_t n_ld = ld - ev;
        if (n_ld > T) n_ld = T;
        const unsigned bytes = (unsigned)(n_ld * sizeof(double));
        const int lane = tid & 31;
        if (lane == 0) bi_mbar_expect_tx(&full_bar[st], bytes * (unsigned)(C * S));
        __syncwarp();
        double* dst = smem_tiles + (size_t)st * tile_doubles;
        for (int k = lane; k < C * S; k += 32) {
            const int c = k / S, s = k - c * S;
            const double* src = A + ((int64_t)corner[lead * C + c] * S + s) * ld + ev;
            bi_bulk_g2s(dst + (size_t)k * T, src, bytes, &full_bar[st]);
        }
    };

    if (tid < 32) {
        for (int t = 0; t < n_stages && t < n_tiles; ++t) issue(t);
    }

    double s_sum[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) s_sum[q] = 0.0;
    const bool warp_has_work = (tid & ~31) < count;           // warps beyond the point count only sync
    const int tiles_per_super = BI_SUPERBLOCK / T;
    int64_t sb = sb_begin;
    int in_super = 0;

    for (int t = 0; t < n_tiles; ++t) {
        const int st = t % n_stages;
        bi_mbar_wait(&full_bar[st], (unsigned)((t / n_stages) & 1));
        const double* tile = smem_tiles + (size_t)st * tile_doubles;
        const int64_t ev = ev_begin + (int64_t)t * T;
        const int n_valid = (int)((N - ev) < (int64_t)T ? (N - ev) : (int64_t)T);
        const int n_blocks = (n_valid + BI_EVENT_BLOCK - 1) / BI_EVENT_BLOCK;
        for (int b = 0; warp_has_work && b < n_blocks; ++b) {
            double L[Q];
            bi_group_block<C, Q>(tile, T, S, b * BI_EVENT_BLOCK, n_valid, g, wp, mp, outlier, L);
#pragma unroll
            for (int q = 0; q < Q; ++q) s_sum[q] = __dadd_rn(s_sum[q], L[q]);
        }
        ++in_super;
        if (in_super == tiles_per_super || t == n_tiles - 1) {
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                if (active[q]) partial[pidx[q] * n_super + sb] = s_sum[q];
                s_sum[q] = 0.0;
            }
            ++sb;
            in_super = 0;
        }
        __syncthreads();                                      // every thread is done reading stage `st`
        if (tid < 32 && t + n_stages < n_tiles) issue(t + n_stages);
    }
}

template <int C>
__global__ void __launch_bounds__(BI_GROUP_THREADS)
k_unbinned_grouped(const double* __restrict__ A, int64_t ld, int64_t N, int S, int T, int n_stages,
                   const int32_t* __restrict__ group_points, const int4* __restrict__ work, int64_t n_super,
                   const int32_t* __restrict__ corner, const double* __restrict__ weight,
                   const double* __restrict__ mus, double outlier, double* __restrict__ partial) {
    extern __shared__ __align__(128) unsigned char bi_smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bi_smem);                    // [n_stages] (<= 8)
    double* smem_tiles = reinterpret_cast<double*>(bi_smem + 128);
    if (threadIdx.x == 0) {
        for (int i = 0; i < n_stages; ++i) bi_mbar_init(&full_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int4 wk = work[blockIdx.x];
    if (wk.y <= BI_GROUP_THREADS)
        bi_group_run<C, 1>(A, ld, N, S, T, n_stages, group_points, wk, n_super, corner, weight, mus, outlier,
                           partial, smem_tiles, full_bar);
    else
        bi_group_run<C, 2>(A, ld, N, S, T, n_stages, group_points, wk, n_super, corner, weight, mus, outlier,
                           partial, smem_tiles, full_bar);
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
    double u = 0.0;
    for (int64_t j = t; j < n_super; j += 256) u = __dadd_rn(u, partial[p * n_super + j]);
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

// tile size and stage count of the grouped kernel for a given C*S (host-side policy, also exported for tests)
static void bi_group_config(int C, int S, int* T, int* n_stages, size_t* smem_bytes) {
    const int cs = C * S;
    int t = 128;
    while (t > 32 && (size_t)cs * t * 8 > 32 * 1024) t >>= 1;
    const size_t tile_bytes = (size_t)cs * t * 8;
    int stages = (int)((160 * 1024) / tile_bytes);
    if (stages > 4) stages = 4;
    if (stages < 2) stages = 2;
    *T = t;
    *n_stages = stages;
    *smem_bytes = 128 + tile_bytes * stages;
}

template <int C>
static int bi_launch_grouped(const double* A, int64_t ld, int64_t N, int32_t S, const int32_t* group_points,
                             const int32_t* work, int64_t n_work, int64_t n_super, const int32_t* corner,
                             const double* weight, const double* mus, double outlier, double* partial,
                             cudaStream_t st) {
    int T, n_stages;
    size_t smem;
    bi_group_config(C, S, &T, &n_stages, &smem);
    BI_REQUIRE(smem <= 200 * 1024, "grouped kernel: C*S=%d needs %zu bytes of shared memory", C * S, smem);
    BI_CUDA_CHECK(cudaFuncSetAttribute(k_unbinned_grouped<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_unbinned_grouped<C><<<(unsigned)n_work, BI_GROUP_THREADS, smem, st>>>(
        A, ld, N, S, T, n_stages, group_points, reinterpret_cast<const int4*>(work), n_super, corner, weight, mus,
        outlier, partial);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

extern "C" int bi_unbinned_partials_grouped(const double* ps_anchor_dev, int64_t ld_events, int64_t n_events,
                                            int32_t n_sources, int32_t n_corners,
                                            const int32_t* group_points_dev, const int32_t* work_dev, int64_t n_work,
                                            const int32_t* corner_dev, const double* weight_dev,
                                            const double* mus_dev, const int32_t* status_dev,
                                            double outlier_likelihood, double* partial_dev, void* stream) {
    (void)status_dev;
    int rc = bi_check_tensor(ps_anchor_dev, ld_events, n_events, n_sources, n_corners);
    if (rc != BI_OK) return rc;
    BI_REQUIRE(n_work >= 0, "n_work < 0");
    const int64_t n_super = bi_num_superblocks(n_events);
    if (n_work == 0 || n_super == 0) return BI_OK;
    BI_REQUIRE(n_sources <= BI_GROUP_MAX_SOURCES, "grouped kernel supports at most %d sources (got %d)", BI_GROUP_MAX_SOURCES, n_sources);
    BI_REQUIRE(n_corners <= 16, "grouped kernel supports at most 16 corners (got %d)", n_corners);
    BI_REQUIRE(group_points_dev && work_dev && corner_dev && weight_dev && mus_dev && partial_dev,
               "bi_unbinned_partials_grouped: NULL pointer");
    BI_REQUIRE(((uintptr_t)work_dev & 15) == 0, "work_dev must be 16-byte aligned");
    BI_REQUIRE(ld_events % 2 == 0, "ld_events must be even");
    cudaStream_t st = (cudaStream_t)stream;
#define BI_GROUP_CASE(CC)                                                                                      \
    case CC:                                                                                                   \
        return bi_launch_grouped<CC>(ps_anchor_dev, ld_events, n_events, n_sources, group_points_dev, work_dev, \
                                     n_work, n_super, corner_dev, weight_dev, mus_dev, outlier_likelihood,      \
                                     partial_dev, st);
    switch (n_corners) {
        BI_GROUP_CASE(1)
        BI_GROUP_CASE(2)
        BI_GROUP_CASE(4)
        BI_GROUP_CASE(8)
        BI_GROUP_CASE(16)
    }
#undef BI_GROUP_CASE
    bi_set_error("unsupported n_corners=%d", n_corners);
    return BI_ERR_UNSUPPORTED;
}

extern "C" int bi_unbinned_finalize(const double* partial_dev, int64_t n_super, const double* musum_dev,
                                    const int32_t* status_dev, int64_t n_points, double* logl_dev,
                                    double* logsum_dev, void* stream) {
    BI_REQUIRE(n_points >= 0 && n_super >= 0, "negative size");
    if (n_points == 0) return BI_OK;
    BI_REQUIRE(musum_dev && status_dev && logl_dev && (n_super == 0 || partial_dev), "bi_unbinned_finalize: NULL pointer");
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
