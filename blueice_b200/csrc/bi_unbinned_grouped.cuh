// blueice_b200 -- K2 grouped kernel (threads = points of one hypercube cell), templated on <C, S>.
//
// One CTA = one work item: up to 256 points that share a hypercube cell x a range of 512-event
// superblocks.  4 warps, 64 points each (lane l owns points l and l + 32 of its warp).  The cell's
// C*S slab tiles of T events are staged in shared memory by TMA bulk copies (cp.async.bulk, SASS
// UBLKCP) through a STAGES-deep ring.  "full" mbarriers signal arrival of the bytes; a stage is
// released with a shared-memory counter and THE LAST WARP TO RELEASE IT ISSUES THE NEXT TILE into it,
// so nobody ever waits for a free stage and there is no CTA-wide barrier and no idle producer warp.
// Consumers read the tile with broadcast LDS.128 (all lanes read the same two events of one slab) at
// compile-time offsets, so the inner loop is DFMA + LDS only:
//     ps_s[k] = fma(A[c,s,e+k], w_c, ps_s[k])   (C*S*4 per quad and point)
//     p[k]    = fma(mu_s, ps_s[k], p[k])
// followed by the canonical quad product (bi_common.cuh): 3 DMUL + one mantissa/exponent split per
// 4 events, one combined normality predicate, and ONE log per 512-event superblock (the product tree
// of the 16 block products lives in shared memory).
#pragma once
#include "bi_common.cuh"

#define BI_GROUP_WARPS 4
#define BI_GROUP_THREADS (BI_GROUP_WARPS * 32)
#define BI_SUPER_LEVELS 4          /* 16 blocks per superblock = 2^4 */

template <int C, int S>
struct BiGroupCfg {
    static constexpr int CS = C * S;
    static constexpr int T = CS <= 16 ? 128 : (CS <= 32 ? 64 : 32);          // events per tile
    static constexpr int STAGES = (CS * T * 8 <= 16384) ? 4 : 3;
    static constexpr int TILE_DOUBLES = CS * T;
    // ring of tiles + per-thread superblock product tree (BI_SUPER_LEVELS x 2 points x 128 threads doubles)
    static constexpr int LEVEL_DOUBLES = BI_SUPER_LEVELS * 2 * BI_GROUP_THREADS;
    static constexpr int SMEM_BYTES = 256 + (STAGES * TILE_DOUBLES + LEVEL_DOUBLES) * 8;
    static constexpr int MIN_BLOCKS = C <= 4 ? 4 : (C == 8 ? 3 : 2);   // 128 threads: 4 CTAs -> 128 regs
};

// ---------------------------------------------------------------------------------------------
// mbarrier / TMA helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bi_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bi_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bi_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bi_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bi_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bi_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bi_smem_u32(bar)) : "memory");
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

// ---------------------------------------------------------------------------------------------
// slow paths (rare; noinline keeps them out of the hot loop's register allocation)
// ---------------------------------------------------------------------------------------------
static __device__ __noinline__ double bi_slow_density_tile(const double* tile, int T, int S, int C, int ev,
                                                    const double* __restrict__ weight_p,
                                                    const double* __restrict__ mu, double outlier) {
    double acc = 0.0;
    for (int s = 0; s < S; ++s) {
        double ps = 0.0;
        for (int c = 0; c < C; ++c) ps = fma(tile[(size_t)(c * S + s) * T + ev], weight_p[c], ps);
        const double t = __dmul_rn(mu[s], ps);
        if (t == t) acc = __dadd_rn(acc, t);       // nansum: NaN terms count as 0
    }
    return bi_fix_density(acc, outlier);
}

struct BiQuad { double m; int e; int slow; };

// Quad whose direct products left the normal range or that contains an abnormal density.
// ev: tile offset of the quad's first event; n_valid: events of the tile that exist.
static __device__ __noinline__ BiQuad bi_quad_slow(double p0, double p1, double p2, double p3, const double* tile, int T,
                                            int S, int C, int ev, int n_valid, const double* __restrict__ weight_p,
                                            const double* __restrict__ mu, double outlier) {
    double f[4] = {p0, p1, p2, p3};
    BiQuad r;
    r.slow = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (ev + k >= n_valid) { f[k] = 1.0; continue; }
        if (!bi_is_normal_positive(f[k])) {
            f[k] = bi_slow_density_tile(tile, T, S, C, ev + k, weight_p, mu, outlier);
            if (!bi_is_normal_positive(f[k])) { r.slow = 1; f[k] = 1.0; }
        }
    }
    const double q2a = __dmul_rn(f[0], f[1]), q2b = __dmul_rn(f[2], f[3]);
    const double q4 = __dmul_rn(q2a, q2b);
    if (bi_is_normal_positive(q2a) && bi_is_normal_positive(q2b) && bi_is_normal_positive(q4))
        bi_split(q4, &r.m, &r.e);
    else
        bi_quad_from_mantissas(f[0], f[1], f[2], f[3], &r.m, &r.e);
    return r;
}

// block (32 events at tile offset e0) recomputed as the binary tree SUM of log(p_i) -- same tree shape
static __device__ __noinline__ double bi_block_slow(const double* tile, int T, int S, int C, int e0, int n_valid,
                                             const double* __restrict__ weight_p, const double* __restrict__ mu,
                                             double outlier) {
    double s3 = 0.0, s4 = 0.0, s5 = 0.0, tot = 0.0;
    for (int o = 0; o < 8; ++o) {
        double lg[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = e0 + 4 * o + k;
            double pk = 1.0;
            if (e < n_valid) pk = bi_slow_density_tile(tile, T, S, C, e, weight_p, mu, outlier);
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
    return tot;
}

// ---------------------------------------------------------------------------------------------
// hot path
// ---------------------------------------------------------------------------------------------
template <int C, int S, int Q>
struct BiGroupRegs {
    double w[Q][C];
    double mu[Q][S];
};

// densities of 4 consecutive events (tile offset of `row0` already includes the event offset)
template <int C, int S, int Q>
__device__ __forceinline__ void bi_quad_density(const double* __restrict__ row0, const BiGroupRegs<C, S, Q>& g,
                                                double (&p)[Q][4]) {
    constexpr int T = BiGroupCfg<C, S>::T;
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) p[q][k] = 0.0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        double ps[Q][4];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const double2 a = *reinterpret_cast<const double2*>(row0 + (c * S + s) * T);
            const double2 b = *reinterpret_cast<const double2*>(row0 + (c * S + s) * T + 2);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                if (c == 0) {
                    // fma(v, w, 0.0) == round(v * w): keep the fused form so the first term rounds like the others
                    ps[q][0] = fma(a.x, g.w[q][c], 0.0);
                    ps[q][1] = fma(a.y, g.w[q][c], 0.0);
                    ps[q][2] = fma(b.x, g.w[q][c], 0.0);
                    ps[q][3] = fma(b.y, g.w[q][c], 0.0);
                } else {
                    ps[q][0] = fma(a.x, g.w[q][c], ps[q][0]);
                    ps[q][1] = fma(a.y, g.w[q][c], ps[q][1]);
                    ps[q][2] = fma(b.x, g.w[q][c], ps[q][2]);
                    ps[q][3] = fma(b.y, g.w[q][c], ps[q][3]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
            for (int k = 0; k < 4; ++k) p[q][k] = fma(g.mu[q][s], ps[q][k], p[q][k]);
    }
}

// One canonical block (32 events at tile offset e0) for the Q points of this thread:
// fast result (M in [1, 2^8), E) or, if some density of the block is abnormal, M = 1, E = 0 and the
// block's tree sum of logs in `Lslow` (0 otherwise).
template <int C, int S, int Q, bool TAIL>
__device__ __forceinline__ void bi_group_block(const double* __restrict__ tile, int e0, int n_valid,
                                               const BiGroupRegs<C, S, Q>& g, const double* const (&wp)[Q],
                                               const double* const (&mp)[Q], double outlier,
                                               double (&M)[Q], int (&E)[Q], double (&Lslow)[Q]) {
    constexpr int T = BiGroupCfg<C, S>::T;
    double l4[Q], l5[Q];
    int slow[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) { E[q] = 0; slow[q] = 0; l4[q] = l5[q] = M[q] = 1.0; }

#pragma unroll 1
    for (int h = 0; h < 4; ++h) {                         // two quads (8 events) per iteration
        double v2[Q];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = e0 + 8 * h + 4 * u;
            double p[Q][4];
            bi_quad_density<C, S, Q>(tile + e, g, p);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const double q2a = __dmul_rn(p[q][0], p[q][1]), q2b = __dmul_rn(p[q][2], p[q][3]);
                const double q4 = __dmul_rn(q2a, q2b);
                // all seven values normal positive  <=>  max over (hi32 - 0x00100000) as unsigned < 0x7fe00000
                unsigned t = (unsigned)(__double2hiint(p[q][0]) - 0x00100000);
                t = max(t, (unsigned)(__double2hiint(p[q][1]) - 0x00100000));
                t = max(t, (unsigned)(__double2hiint(p[q][2]) - 0x00100000));
                t = max(t, (unsigned)(__double2hiint(p[q][3]) - 0x00100000));
                t = max(t, (unsigned)(__double2hiint(q2a) - 0x00100000));
                t = max(t, (unsigned)(__double2hiint(q2b) - 0x00100000));
                t = max(t, (unsigned)(__double2hiint(q4) - 0x00100000));
                double m;
                int qe;
                bool direct = t < 0x7fe00000u;
                if (TAIL) direct = direct && (e + 3 < n_valid);
                if (direct) {
                    bi_split(q4, &m, &qe);
                } else {
                    const BiQuad r = bi_quad_slow(p[q][0], p[q][1], p[q][2], p[q][3], tile, T, S, C, e, n_valid,
                                                  wp[q], mp[q], outlier);
                    m = r.m; qe = r.e; slow[q] |= r.slow;
                }
                E[q] += qe;
                v2[q] = (u == 0) ? m : __dmul_rn(v2[q], m);
            }
        }
        // binary-counter merge over pairs of quads = the canonical binary tree over the 8 quads
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            double v = v2[q];
            if (h & 1) {
                v = __dmul_rn(l4[q], v);
                if (h & 2) M[q] = __dmul_rn(l5[q], v); else l5[q] = v;
            } else l4[q] = v;
        }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        Lslow[q] = 0.0;
        if (slow[q]) {
            Lslow[q] = bi_block_slow(tile, T, S, C, e0, n_valid, wp[q], mp[q], outlier);
            M[q] = 1.0;
            E[q] = 0;
        }
    }
}

// Superblock product tree over the 16 block products of a 512-event superblock, kept in shared memory
// ([level][point slot][thread], conflict-free): binary counter on the block index.
template <int Q>
__device__ __forceinline__ void bi_super_merge(double* lv, int blk, const double (&Mb)[Q], double (&out)[Q]) {
    // lv: this thread's base in the level array; element (level, q) at lv[(level * 2 + q) * BI_GROUP_THREADS]
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        double v = Mb[q];
        bool done = false;
#pragma unroll
        for (int b = 0; b < BI_SUPER_LEVELS; ++b) {
            if (!done) {
                double* slot = lv + (b * 2 + q) * BI_GROUP_THREADS;
                if (blk & (1 << b)) v = __dmul_rn(*slot, v);
                else { *slot = v; done = true; }
            }
        }
        out[q] = v;       // complete product iff blk == 15
    }
}

// product of the pending levels when the superblock ends early (missing blocks count as 1.0, exactly)
template <int Q>
__device__ __forceinline__ void bi_super_flush(const double* lv, int n_blocks_done, double (&out)[Q]) {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        double v = 1.0;
#pragma unroll
        for (int b = 0; b < BI_SUPER_LEVELS; ++b)
            if (n_blocks_done & (1 << b)) v = __dmul_rn(lv[(b * 2 + q) * BI_GROUP_THREADS], v);
        out[q] = v;
    }
}

// producer step, executed by one whole warp: arm the full barrier and issue the C*S bulk copies of tile t
template <int C, int S>
__device__ __forceinline__ void bi_group_issue(const double* __restrict__ A, int64_t ld, int64_t ev_begin, int t,
                                               const int32_t* __restrict__ corner_lead, double* smem_tiles,
                                               uint64_t* full_bar) {
    using Cfg = BiGroupCfg<C, S>;
    constexpr int T = Cfg::T;
    const int lane = threadIdx.x & 31;
    const int st = t % Cfg::STAGES;
    const int64_t ev = ev_begin + (int64_t)t * T;
    int64_t n_ld = ld - ev;
    if (n_ld > T) n_ld = T;
    const unsigned bytes = (unsigned)(n_ld * sizeof(double));
    if (lane == 0) {
        // order the generic-proxy reads of this stage (all warps are done with it) before the async-proxy writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bi_mbar_expect_tx(&full_bar[st], bytes * (unsigned)Cfg::CS);
    }
    __syncwarp();
    double* dst = smem_tiles + (size_t)st * Cfg::TILE_DOUBLES;
    for (int k = lane; k < Cfg::CS; k += 32) {
        const int c = k / S, s = k - c * S;
        const double* src = A + ((int64_t)corner_lead[c] * S + s) * ld + ev;
        bi_bulk_g2s(dst + (size_t)k * T, src, bytes, &full_bar[st]);
    }
}

template <int C, int S, int Q>
__device__ __forceinline__ void bi_group_consume(const double* __restrict__ A, int64_t ld, int64_t N,
                                                 const int32_t* __restrict__ group_points, int first, int count,
                                                 int n_cwarps, int64_t sb_begin, int64_t ev_begin, int n_tiles,
                                                 int64_t n_super, const int32_t* __restrict__ corner,
                                                 const double* __restrict__ weight, const double* __restrict__ mus,
                                                 double outlier, double* __restrict__ partial, double* smem_tiles,
                                                 double* smem_levels, uint64_t* full_bar, int* release_cnt) {
    using Cfg = BiGroupCfg<C, S>;
    constexpr int T = Cfg::T;
    constexpr int BLOCKS_PER_TILE = T / BI_EVENT_BLOCK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    BiGroupRegs<C, S, Q> g;
    int64_t pidx[Q];
    bool active[Q];
    const double* wp[Q];
    const double* mp[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int k = warp * 64 + q * 32 + lane;
        active[q] = k < count;
        pidx[q] = group_points[first + (active[q] ? k : warp * 64)];
        wp[q] = weight + pidx[q] * C;
        mp[q] = mus + pidx[q] * S;
#pragma unroll
        for (int c = 0; c < C; ++c) g.w[q][c] = wp[q][c];
#pragma unroll
        for (int s = 0; s < S; ++s) g.mu[q][s] = mp[q][s];
    }
    const int32_t* corner_lead = corner + (int64_t)group_points[first] * C;
    double* lv = smem_levels + threadIdx.x;

    int e_super[Q];
    double slow_sum[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) { e_super[q] = 0; slow_sum[q] = 0.0; }
    int64_t sb = sb_begin;
    int blk = 0, st = 0;                                    // block index inside the current superblock
    unsigned parity = 0;

    for (int t = 0; t < n_tiles; ++t) {
        bi_mbar_wait(&full_bar[st], parity);
        const double* tile = smem_tiles + (size_t)st * Cfg::TILE_DOUBLES;
        const int64_t remaining = N - (ev_begin + (int64_t)t * T);
        const bool full_tile = remaining >= T;
        const int n_valid = full_tile ? T : (int)remaining;
        const int n_blocks = full_tile ? BLOCKS_PER_TILE : (n_valid + BI_EVENT_BLOCK - 1) / BI_EVENT_BLOCK;
        double Msuper[Q];
#pragma unroll 1
        for (int b = 0; b < n_blocks; ++b) {
            double Mb[Q], Ls[Q];
            int Eb[Q];
            if (full_tile) bi_group_block<C, S, Q, false>(tile, b * BI_EVENT_BLOCK, T, g, wp, mp, outlier, Mb, Eb, Ls);
            else bi_group_block<C, S, Q, true>(tile, b * BI_EVENT_BLOCK, n_valid, g, wp, mp, outlier, Mb, Eb, Ls);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                e_super[q] += Eb[q];
                slow_sum[q] = __dadd_rn(slow_sum[q], Ls[q]);
            }
            bi_super_merge<Q>(lv, blk, Mb, Msuper);
            ++blk;
        }
        // release the stage; the last warp to do so refills it with tile t + STAGES
        __syncwarp();
        int last = 0;
        if (lane == 0) last = (atomicAdd(&release_cnt[st], 1) == n_cwarps - 1);
        last = __shfl_sync(BI_FULL_MASK, last, 0);
        if (last) {
            if (lane == 0) release_cnt[st] = 0;
            if (t + Cfg::STAGES < n_tiles)
                bi_group_issue<C, S>(A, ld, ev_begin, t + Cfg::STAGES, corner_lead, smem_tiles, full_bar);
        }
        if (blk == BI_SUPERBLOCK / BI_EVENT_BLOCK || t == n_tiles - 1) {
            if (blk != BI_SUPERBLOCK / BI_EVENT_BLOCK) bi_super_flush<Q>(lv, blk, Msuper);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const double sj = __dadd_rn(bi_block_log(Msuper[q], e_super[q]), slow_sum[q]);
                if (active[q]) partial[pidx[q] * n_super + sb] = sj;
                e_super[q] = 0;
                slow_sum[q] = 0.0;
            }
            ++sb;
            blk = 0;
        }
        if (++st == Cfg::STAGES) { st = 0; parity ^= 1u; }
    }
}

template <int C, int S>
__global__ void __launch_bounds__(BI_GROUP_THREADS, BiGroupCfg<C, S>::MIN_BLOCKS)
k_unbinned_grouped(const double* __restrict__ A, int64_t ld, int64_t N,
                   const int32_t* __restrict__ group_points, const int4* __restrict__ work, int64_t n_super,
                   const int32_t* __restrict__ corner, const double* __restrict__ weight,
                   const double* __restrict__ mus, double outlier, double* __restrict__ partial) {
    using Cfg = BiGroupCfg<C, S>;
    constexpr int T = Cfg::T;
    extern __shared__ __align__(128) unsigned char bi_smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bi_smem);                   // [STAGES]
    int* release_cnt = reinterpret_cast<int*>(bi_smem + 64);                     // [STAGES]
    double* smem_tiles = reinterpret_cast<double*>(bi_smem + 256);
    double* smem_levels = smem_tiles + Cfg::STAGES * Cfg::TILE_DOUBLES;

    const int4 wk = work[blockIdx.x];
    const int first = wk.x, count = wk.y;
    const int64_t sb_begin = wk.z, sb_end = wk.w;
    const int64_t ev_begin = sb_begin * BI_SUPERBLOCK;
    int64_t ev_end = sb_end * BI_SUPERBLOCK;
    if (ev_end > N) ev_end = N;
    const int n_tiles = (int)((ev_end - ev_begin + T - 1) / T);
    const int n_cwarps = min(BI_GROUP_WARPS, (count + 63) / 64);
    const int warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        for (int i = 0; i < Cfg::STAGES; ++i) {
            bi_mbar_init(&full_bar[i], 1);
            release_cnt[i] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp >= n_cwarps) return;                           // warps without points

    if (warp == 0) {                                        // prologue: fill the ring
        const int32_t* corner_lead = corner + (int64_t)group_points[first] * C;
        for (int t = 0; t < Cfg::STAGES && t < n_tiles; ++t)
            bi_group_issue<C, S>(A, ld, ev_begin, t, corner_lead, smem_tiles, full_bar);
    }
    // 64 points per warp; the last warp may hold 32 or fewer
    if (count - warp * 64 > 32)
        bi_group_consume<C, S, 2>(A, ld, N, group_points, first, count, n_cwarps, sb_begin, ev_begin, n_tiles, n_super,
                                  corner, weight, mus, outlier, partial, smem_tiles, smem_levels, full_bar, release_cnt);
    else
        bi_group_consume<C, S, 1>(A, ld, N, group_points, first, count, n_cwarps, sb_begin, ev_begin, n_tiles, n_super,
                                  corner, weight, mus, outlier, partial, smem_tiles, smem_levels, full_bar, release_cnt);
}

template <int C, int S>
static int bi_launch_grouped_cs(const double* A, int64_t ld, int64_t N, const int32_t* group_points,
                                const int32_t* work, int64_t n_work, int64_t n_super, const int32_t* corner,
                                const double* weight, const double* mus, double outlier, double* partial,
                                cudaStream_t st) {
    using Cfg = BiGroupCfg<C, S>;
    static bool configured = false;
    if (!configured) {
        BI_CUDA_CHECK(cudaFuncSetAttribute(k_unbinned_grouped<C, S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg::SMEM_BYTES));
        configured = true;
    }
    k_unbinned_grouped<C, S><<<(unsigned)n_work, BI_GROUP_THREADS, Cfg::SMEM_BYTES, st>>>(
        A, ld, N, group_points, reinterpret_cast<const int4*>(work), n_super, corner, weight, mus, outlier, partial);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

// one translation unit per C instantiates S = 1..8 (bi_grouped_c*.cu) and exports this dispatcher
template <int C>
static int bi_dispatch_grouped_s(int S, const double* A, int64_t ld, int64_t N, const int32_t* group_points,
                                 const int32_t* work, int64_t n_work, int64_t n_super, const int32_t* corner,
                                 const double* weight, const double* mus, double outlier, double* partial,
                                 cudaStream_t st) {
#define BI_S_CASE(SS)                                                                                             \
    case SS:                                                                                                      \
        return bi_launch_grouped_cs<C, SS>(A, ld, N, group_points, work, n_work, n_super, corner, weight, mus,    \
                                           outlier, partial, st);
    switch (S) {
        BI_S_CASE(1) BI_S_CASE(2) BI_S_CASE(3) BI_S_CASE(4) BI_S_CASE(5) BI_S_CASE(6) BI_S_CASE(7) BI_S_CASE(8)
    }
#undef BI_S_CASE
    bi_set_error("grouped kernel: unsupported n_sources=%d", S);
    return BI_ERR_UNSUPPORTED;
}

#define BI_DEFINE_GROUPED_TU(CC)                                                                                  \
    int bi_grouped_launch_c##CC(int S, const double* A, int64_t ld, int64_t N, const int32_t* group_points,       \
                                const int32_t* work, int64_t n_work, int64_t n_super, const int32_t* corner,      \
                                const double* weight, const double* mus, double outlier, double* partial,         \
                                cudaStream_t st) {                                                                \
        return bi_dispatch_grouped_s<CC>(S, A, ld, N, group_points, work, n_work, n_super, corner, weight, mus,   \
                                         outlier, partial, st);                                                   \
    }
