// blueice_b200 -- K2: fused anchor-morph + mixture density + log + reduce on the FP64 tensor path (DMMA).
//
// Replaces, for a batch of P parameter points over N events (blueice/likelihood.py:355-356,678-690 and
// scipy/interpolate/_rgi.py:520-549):
//     f(theta_p, x_i) = sum_k coef_{p,k} * A[row_{p,k}, i],  coef = fl(w * mu)      (morph x mixture as ONE contraction;
//                       term k = (corner c, source s) of the full anchor grid, or (source s, corner c_s of
//                       that source's own sub-grid) with source-wise interpolation, likelihood.py:210-240)
//     partial[p, j]   = sum_{i in superblock j} log f(theta_p, x_i)
//
// The contraction over the K terms of 8 points x 8 events is one chain of
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4): A operand = the points' coefficients (registers, loaded once
// per work unit), B operand = the event tile (shared memory, staged by TMA bulk copies through a
// per-warp mbarrier ring), D = densities, 2 consecutive events of one point per thread.  DMMA and DFMA
// share one pipe on sm_100a (37.1 TFLOP/s either way, profiles/microbench), but one DMMA replaces 8 warp
// DFMAs, needs no broadcast LDS and keeps the coefficients of 8 points in ONE register pair per k-step,
// so the kernel is bound by that pipe instead of by instruction issue.  DMMA accumulates as a
// sequential fma chain in k order (bit-exact, profiles/microbench/dmma_probe_b200.log).
//
// Work: a point group = up to 8*MT points that share one hypercube cell.  Persistent warps pick a group, load its
// coefficients once, then fetch superblock ranges of that group from the group's counter until none is left
// (the TMA tile stream runs across range boundaries), then move to another group.  Every warp owns its ring,
// its barriers and its TMA issue -- no CTA-wide sync.
//
// Canonical arithmetic -- what every result is defined by, independent of how the
// batch is cut into units:
//   p_i        = fma chain over k = 0..K-1 of A[k,i] * coef_k starting from 0, coef_k = fl(w_c * mu_s)
//   group      = 32 consecutive events (4 octets); class t = 0..3 owns events 8n + 2t, 8n + 2t + 1 of octet n
//   (t, group) : pair_n = fl(p * p'), quad = fl(pair_0 * pair_1), fl(pair_2 * pair_3), oct = fl(quad * quad')
//                -> (m, e) = mantissa / exponent of oct.          [needs every p in [2^-126, 2^127)]
//                otherwise (m, e) = (1, 0) and l = the same tree over log(p_i) with p_i re-evaluated
//                with the reference's nansum / outlier semantics (likelihood.py:686-689)
//   superblock = 16 groups: M_t = (((1 * m_0) * m_1) ...), E_t = sum e, L_t = ((0 + l_0) + l_1) ...
//                M = (M_0 * M_1) * (M_2 * M_3), E = sum E_t, L = (L_0 + L_1) + (L_2 + L_3)
//                partial = fma(E, LN2_LO, fma(E, LN2_HI, log(M))) + L           -> ONE log per 512 events
//   events >= N count as p = 1 (exact identity).
#pragma once
#include <cuda.h>          // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no libcuda link)

#include "bi_common.cuh"
#include "bi_tma.cuh"

#define BI_GROUP_EVENTS 32
#define BI_GROUPS_PER_SUPER (BI_SUPERBLOCK / BI_GROUP_EVENTS)
// p in [2^-126, 2^127)  <=>  (unsigned)(hi32(p) - BI_RANGE_LO) < BI_RANGE_SPAN
#define BI_RANGE_LO ((1023 - 126) << 20)
#define BI_RANGE_SPAN (253u << 20)

// K4 = k-steps of 4 terms (instantiated: 1..8, 12, 16, 24, 32 -> up to 128 terms, rows zero-padded to 4 * K4)
// tuning knobs of the K <= 8 instantiations (overridable at build time for experiments)
#ifndef BI_MMA_MT_SMALL
#define BI_MMA_MT_SMALL 4      /* measured on B200: 4 or 6 m-tiles per warp beat 8 by 4 % at K = 8 (config 2) */
#endif
#ifndef BI_MMA_T_SMALL
#define BI_MMA_T_SMALL 128
#endif
#ifndef BI_MMA_MINCTAS_SMALL
#define BI_MMA_MINCTAS_SMALL 3
#endif
#ifndef BI_MMA_STAGES_SMALL
#define BI_MMA_STAGES_SMALL 2
#endif
#ifndef BI_MMA_GROUP_DEFER
#define BI_MMA_GROUP_DEFER 1    /* K <= 8: the groups of a full tile run as one branch-free block */
#endif
#ifndef BI_MMA_WARPS_SMALL
#define BI_MMA_WARPS_SMALL 4
#endif
template <int K4>
struct BiMmaCfg {
    static constexpr int MT = K4 <= 2 ? BI_MMA_MT_SMALL : (K4 <= 4 ? 4 : (K4 <= 16 ? 2 : 1));   // 8-point m-tiles per unit
    static constexpr int KP = 4 * K4;                              // rows incl. zero padding
    static constexpr int T = K4 <= 2 ? BI_MMA_T_SMALL : (K4 <= 4 ? 64 : 32);  // events per tile (row copies of T*8 bytes)
    static constexpr int RS = T + 4;                               // row stride: B-fragment loads hit 16 distinct banks
    static constexpr int STAGES = K4 <= 2 ? BI_MMA_STAGES_SMALL : K4 <= 4 ? 2 : (K4 <= 8 ? 3 : 2);
    static constexpr int WARPS = K4 <= 2 ? BI_MMA_WARPS_SMALL : K4 <= 8 ? 4 : (K4 <= 16 ? 2 : 1); // work units (warps) per CTA
    static constexpr int THREADS = WARPS * 32;
    static constexpr int STAGE_DOUBLES = KP * RS;
    static constexpr int RING_DOUBLES = STAGES * STAGE_DOUBLES;
    static constexpr int BREG = K4 <= 4;                           // B fragments of a whole group live in registers
    static constexpr int ROWREG = K4 <= 8;                         // K <= 32: lane k keeps the source pointer of row k
    static constexpr int HEADER_BYTES = 256;                       // mbarriers [WARPS][STAGES]
    static constexpr int SLOW_DOUBLES = MT * THREADS;              // L_t accumulators (touched on the slow path only)
    static constexpr int SMEM_BYTES = HEADER_BYTES + (WARPS * RING_DOUBLES + SLOW_DOUBLES) * 8;
    static constexpr int MIN_CTAS = K4 <= 2 ? BI_MMA_MINCTAS_SMALL : K4 <= 4 ? 3 : (K4 <= 8 ? 2 : (K4 <= 16 ? 3 : (K4 <= 24 ? 4 : 3)));
};

// ---------------------------------------------------------------------------------------------
// mbarrier / TMA / DMMA helpers
// ---------------------------------------------------------------------------------------------
// tiled TMA copy of one whole event tile: box [T + 4 events][S sources][2 anchors per shape parameter ...] of the
// anchor tensor viewed as [ld][S][n_D]...[n_1] -> K = S * 2^D dense rows of T + 4 doubles (SASS: UTMALDG)
__device__ __forceinline__ void bi_tensor_g2s(void* dst_smem, const CUtensorMap* tmap, int rank, int c0, int c2, int c3,
                                              int c4, uint64_t* bar) {
    const uint32_t dst = bi_smem_u32(dst_smem), mb = bi_smem_u32(bar);
    const uint64_t tm = reinterpret_cast<uint64_t>(tmap);
    if (rank == 4)
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(dst), "l"(tm), "r"(c0), "r"(0), "r"(c2), "r"(c3), "r"(mb) : "memory");
    else if (rank == 3)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(dst), "l"(tm), "r"(c0), "r"(0), "r"(c2), "r"(mb) : "memory");
    else if (rank == 5)
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(dst), "l"(tm), "r"(c0), "r"(0), "r"(c2), "r"(c3), "r"(c4), "r"(mb) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst), "l"(tm), "r"(c0), "r"(0), "r"(mb) : "memory");
}
// D(8x8) += A(8x4, row) * B(4x8, col); thread (g = lane / 4, t = lane % 4) holds A[g][t], B[t][g], D[g][2t], D[g][2t+1]
__device__ __forceinline__ void bi_dmma(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------
// slow path (rare; noinline keeps it out of the hot loop's register allocation)
// ---------------------------------------------------------------------------------------------
// density of one event with the reference's semantics (likelihood.py:686-689); `col` points at term row 0 of that
// event, rows `rs` apart; term k belongs to source term_source[k] and carries the morph weight wterm[k]
static __device__ __noinline__ double bi_slow_density_rows(const double* col, int rs, int K, int S,
                                                           const int32_t* __restrict__ term_source,
                                                           const double* __restrict__ wterm,
                                                           const double* __restrict__ mu, double outlier) {
    double acc = 0.0;
    for (int s = 0; s < S; ++s) {
        double ps = 0.0;
        for (int k = 0; k < K; ++k)
            if (term_source[k] == s) ps = fma(col[k * rs], wterm[k], ps);
        const double term = __dmul_rn(mu[s], ps);
        if (term == term) acc = __dadd_rn(acc, term);       // nansum: NaN terms count as 0
    }
    return bi_fix_density(acc, outlier);
}

// (t, group) fallback: the canonical tree over log(p_i) of the class's 8 events; grp points at row 0 of
// the group's first event, n_valid = events of the group that exist (others count as p = 1)
static __device__ __noinline__ double bi_slow_group(const double* grp, int rs, int K, int S, int t, int n_valid,
                                                    const int32_t* __restrict__ term_source,
                                                    const double* __restrict__ wterm,
                                                    const double* __restrict__ mu, double outlier) {
    double quad[2];
    for (int h = 0; h < 2; ++h) {
        double pr[2];
        for (int n = 0; n < 2; ++n) {
            const int e = 8 * (2 * h + n) + 2 * t;
            const double p0 = (e < n_valid) ? bi_slow_density_rows(grp + e, rs, K, S, term_source, wterm, mu, outlier) : 1.0;
            const double p1 = (e + 1 < n_valid) ? bi_slow_density_rows(grp + e + 1, rs, K, S, term_source, wterm, mu, outlier) : 1.0;
            pr[n] = __dadd_rn(log(p0), log(p1));
        }
        quad[h] = __dadd_rn(pr[0], pr[1]);
    }
    return __dadd_rn(quad[0], quad[1]);
}

// warp-private producer cursor (shared memory): the range whose tiles are being issued, the next tile of it,
// and the range the producer has fetched ahead of the consumer (one range of lookahead)
struct BiProducer { int range, tile, n_tiles, pending, c2, c3, c4, pad; };   // c2..c4: cell coordinates (tensor-map TMA)

// ---------------------------------------------------------------------------------------------
// producer step (whole warp): arm the stage's barrier and issue the K row copies of tile `tile_idx`.
// K <= 32: src_row = this lane's row (k = lane) at event 0; otherwise rows are looked up in row_lead[K].
// ---------------------------------------------------------------------------------------------
template <int K4>
__device__ __forceinline__ void bi_mma_issue(const double* __restrict__ A, const double* __restrict__ src_row,
                                             const int32_t* __restrict__ row_lead, int64_t ld, int64_t ev_begin,
                                             int tile_idx, int st, int K, double* ring, uint64_t* full_bar,
                                             const CUtensorMap* tmap, int tmap_rank, const BiProducer* prod, int lane) {
    using Cfg = BiMmaCfg<K4>;
    const int64_t ev = ev_begin + (int64_t)tile_idx * Cfg::T;
    double* dst = ring + (size_t)st * Cfg::STAGE_DOUBLES;
    if (tmap_rank) {
        // ONE tiled TMA instruction per tile; out-of-range columns are zero-filled and count as transferred bytes
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bi_mbar_expect_tx(&full_bar[st], (unsigned)(K * Cfg::RS * sizeof(double)));
            bi_tensor_g2s(dst, tmap, tmap_rank, (int)ev, prod->c2, prod->c3, prod->c4, &full_bar[st]);
        }
        __syncwarp();
        return;
    }
    const int64_t left = ld - ev;
    const unsigned bytes = (unsigned)(left < Cfg::T ? left : Cfg::T) * (unsigned)sizeof(double);
    if (lane == 0) {
        // order this warp's generic-proxy reads of the stage before the async-proxy writes that refill it
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bi_mbar_expect_tx(&full_bar[st], bytes * (unsigned)K);
    }
    __syncwarp();
    if (Cfg::ROWREG) {
        if (lane < K) bi_bulk_g2s(dst + (size_t)lane * Cfg::RS, src_row + ev, bytes, &full_bar[st]);
    } else {
        for (int k = lane; k < K; k += 32)
            bi_bulk_g2s(dst + (size_t)k * Cfg::RS, A + (int64_t)row_lead[k] * ld + ev, bytes, &full_bar[st]);
    }
}

// ---------------------------------------------------------------------------------------------
// one canonical group (32 events at tile offset e0) for the NMT m-tiles of this warp
// ---------------------------------------------------------------------------------------------
// densities of the 32 events x 8 points of (group, m-tile): 4 octets x K4 chained DMMA.8x8x4
template <int K4>
__device__ __forceinline__ void bi_mma_tile(const double* __restrict__ bcol, const double (&b)[4][K4], const double (&a)[K4],
                                            double (&d)[4][2]) {
    using Cfg = BiMmaCfg<K4>;
#pragma unroll
    for (int n = 0; n < 4; ++n) d[n][0] = d[n][1] = 0.0;
#pragma unroll
    for (int kk = 0; kk < K4; ++kk)
#pragma unroll
        for (int n = 0; n < 4; ++n)
            bi_dmma(d[n][0], d[n][1], a[kk], Cfg::BREG ? b[n][kk] : bcol[4 * kk * Cfg::RS + 8 * n]);
}

// NMT (compile time) m-tiles in use: the body is ONE branch-free basic block, so ptxas interleaves the DMMAs
// of the next m-tile with the product-tree epilogue of the previous one.
// TAIL: the group holds events >= N (they count as p = 1).
// Returns the m-tiles (bit mt) whose class left the fast range in this group: their (m, e) stay neutral here and the
// caller runs bi_mma_group_slow for them -- right away, or, for the groups of a full tile, after the last group, so that
// the whole tile is ONE basic block and the product-tree tail of a group overlaps the loads and DMMAs of the next.
template <int K4, int NMT, bool TAIL>
__device__ __forceinline__ unsigned bi_mma_group_fast(const double* __restrict__ tile, int e0, int n_valid,
                                                      unsigned active_mask, const double (&a)[NMT][K4],
                                                      double (&M)[NMT], int (&E)[NMT], int lane) {
    using Cfg = BiMmaCfg<K4>;
    const int g = lane >> 2, t = lane & 3;
    // B fragment of octet n, k-step kk: tile[(4 kk + t) * RS + e0 + 8 n + g] (rows >= K are zero)
    const double* bcol = tile + t * Cfg::RS + e0 + g;
    double b[4][K4];
    if (Cfg::BREG) {
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int kk = 0; kk < K4; ++kk) b[n][kk] = bcol[4 * kk * Cfg::RS + 8 * n];
    }
    unsigned bad = 0;
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt) {
        double d[4][2];
        bi_mma_tile<K4>(bcol, b, a[mt], d);
        if (TAIL) {
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const int e = e0 + 8 * n + 2 * t;
                if (e >= n_valid) d[n][0] = 1.0;
                if (e + 1 >= n_valid) d[n][1] = 1.0;
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
        const double oct = __dmul_rn(q0, q1);
        // branch-free: split unconditionally, neutralise (m, e) = (1, 0) when the class leaves the fast range
        double m;
        int e;
        bi_split(oct, &m, &e);
        if (tmax >= BI_RANGE_SPAN) {
            m = 1.0;
            e = 0;
            bad |= 1u << mt;
        }
        M[mt] = __dmul_rn(M[mt], m);
        E[mt] += e;
    }
    return bad & active_mask;
}

// rare: reference-semantics fallback per (t, group) for the m-tiles flagged in `bad`.  Called behind a warp vote (a
// lane-divergent `if (bad)` around the group made every group pay a reconvergence barrier).
template <int K4, int NMT>
__device__ __forceinline__ void bi_mma_group_slow(const double* __restrict__ tile, int e0, int nv, int K, int S, unsigned bad,
                                                  const int32_t* __restrict__ slot_point,
                                                  const int32_t* __restrict__ term_source, const double* __restrict__ wterm,
                                                  const double* __restrict__ mus, double outlier, double* slow_acc,
                                                  bool& slow_any, int lane) {
    using Cfg = BiMmaCfg<K4>;
    const int g = lane >> 2, t = lane & 3;
    for (int mt = 0; mt < NMT; ++mt) {
        if ((bad >> mt) & 1u) {
            const int64_t p = slot_point[mt * 8 + g];
            const double l = bi_slow_group(tile + e0, Cfg::RS, K, S, t, nv, term_source, wterm + p * K, mus + p * S, outlier);
            double* acc = slow_acc + mt * Cfg::THREADS;
            *acc = __dadd_rn(*acc, l);
            slow_any = true;
        }
    }
}

// one group, rare path right behind it
template <int K4, int NMT, bool TAIL>
__device__ __forceinline__ void bi_mma_group(const double* __restrict__ tile, int e0, int n_valid, int K, int S,
                                             unsigned active_mask, const double (&a)[NMT][K4],
                                             const int32_t* __restrict__ slot_point,
                                             const int32_t* __restrict__ term_source, const double* __restrict__ wterm,
                                             const double* __restrict__ mus, double outlier, double* slow_acc,
                                             bool& slow_any, double (&M)[NMT], int (&E)[NMT], int lane) {
    const unsigned bad = bi_mma_group_fast<K4, NMT, TAIL>(tile, e0, n_valid, active_mask, a, M, E, lane);
    if (__any_sync(BI_FULL_MASK, bad != 0)) {
        const int nv = TAIL ? (n_valid - e0 < BI_GROUP_EVENTS ? n_valid - e0 : BI_GROUP_EVENTS) : BI_GROUP_EVENTS;
        bi_mma_group_slow<K4, NMT>(tile, e0, nv, K, S, bad, slot_point, term_source, wterm, mus, outlier, slow_acc, slow_any, lane);
    }
}

// ---------------------------------------------------------------------------------------------
// one work unit with NMT m-tiles: coefficients -> registers, then the tile loop
// ---------------------------------------------------------------------------------------------
template <int K4>
__device__ __forceinline__ int bi_range_tiles(int range, int sb_per, int64_t n_super, int64_t N) {
    const int64_t sb0 = (int64_t)range * sb_per;
    int64_t sb1 = sb0 + sb_per;
    if (sb1 > n_super) sb1 = n_super;
    int64_t ev1 = sb1 * BI_SUPERBLOCK;
    if (ev1 > N) ev1 = N;
    return (int)((ev1 - sb0 * BI_SUPERBLOCK + BiMmaCfg<K4>::T - 1) / BiMmaCfg<K4>::T);
}

// next superblock range of this point group (-1: none left); warp-collective
__device__ __forceinline__ int bi_fetch_range(int* counter, int n_ranges, int lane) {
    int r = 0;
    if (lane == 0) r = atomicAdd(counter, 1);
    r = __shfl_sync(BI_FULL_MASK, r, 0);
    return r < n_ranges ? r : -1;
}

// issue the next tile of the warp's tile stream, fetching the following range when the current one is
// exhausted (warp-collective, all values warp-uniform)
template <int K4>
__device__ __forceinline__ void bi_mma_produce(BiProducer* prod, int* counter, int n_ranges, int sb_per, int64_t n_super,
                                               const double* __restrict__ A, const double* __restrict__ src_row,
                                               const int32_t* __restrict__ row_lead, int64_t ld, int64_t N, int K,
                                               double* ring, uint64_t* full_bar, const CUtensorMap* tmap, int tmap_rank,
                                               unsigned& issued, int lane) {
    using Cfg = BiMmaCfg<K4>;
    int pr = prod->range;
    if (pr < 0) return;
    int pt = prod->tile;
    if (pt == prod->n_tiles) {
        if (prod->pending >= 0) return;                           // lookahead of one range only
        pr = bi_fetch_range(counter, n_ranges, lane);
        pt = 0;
        __syncwarp();
        if (lane == 0) {
            prod->range = pr;
            prod->pending = pr;
            prod->n_tiles = pr >= 0 ? bi_range_tiles<K4>(pr, sb_per, n_super, N) : 0;
        }
        __syncwarp();
        if (pr < 0) return;
    }
    bi_mma_issue<K4>(A, src_row, row_lead, ld, (int64_t)pr * sb_per * BI_SUPERBLOCK, pt, (int)(issued % Cfg::STAGES), K,
                     ring, full_bar, tmap, tmap_rank, prod, lane);
    ++issued;
    if (lane == 0) prod->tile = pt + 1;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// one point group with NMT m-tiles: coefficients -> registers once, then superblock ranges of the group are
// fetched from its counter until none is left (the TMA tile stream runs across range boundaries)
// ---------------------------------------------------------------------------------------------
template <int K4, int NMT>
__device__ __forceinline__ void bi_mma_unit(const double* __restrict__ A, const double* __restrict__ src_row,
                                            const int32_t* __restrict__ row_lead, int64_t ld, int64_t N, int K, int S,
                                            const int32_t* __restrict__ slot_point, unsigned active_mask, int64_t lead,
                                            int* counter, int n_ranges, int sb_per, int64_t n_super,
                                            const double* __restrict__ coef, const int32_t* __restrict__ term_source,
                                            const double* __restrict__ wterm, const double* __restrict__ mus,
                                            double outlier, double* __restrict__ partial,
                                            double* ring, uint64_t* full_bar, double* slow_acc, BiProducer* prod,
                                            const CUtensorMap* tmap, int tmap_rank,
                                            unsigned& issued, unsigned& consumed, int lane) {
    using Cfg = BiMmaCfg<K4>;
    constexpr int T = Cfg::T;
    const int g = lane >> 2, t = lane & 3;

    int cur = bi_fetch_range(counter, n_ranges, lane);
    if (cur < 0) return;
    if (lane == 0) {
        prod->range = cur;
        prod->tile = 0;
        prod->n_tiles = bi_range_tiles<K4>(cur, sb_per, n_super, N);
        prod->pending = -1;
    }
    __syncwarp();
    // the first tiles are in flight while the coefficients are gathered
    for (int i = 0; i < Cfg::STAGES; ++i)
        bi_mma_produce<K4>(prod, counter, n_ranges, sb_per, n_super, A, src_row, row_lead, ld, N, K, ring, full_bar, tmap, tmap_rank, issued, lane);

    double a[NMT][K4];
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt) {
        const int64_t p = ((active_mask >> mt) & 1u) ? (int64_t)slot_point[mt * 8 + g] : lead;
#pragma unroll
        for (int kk = 0; kk < K4; ++kk) {
            const int k = 4 * kk + t;
            a[mt][kk] = (k < K) ? coef[p * K + k] : 0.0;
        }
    }

    double M[NMT];
    int E[NMT];
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt) { M[mt] = 1.0; E[mt] = 0; }
    bool slow_any = false;
    constexpr int TILES_PER_SUPER = BI_SUPERBLOCK / T, GROUPS_PER_TILE = T / BI_GROUP_EVENTS;
    constexpr bool GROUP_DEFER = K4 <= 2 && BI_MMA_GROUP_DEFER;
    // the point whose partial this lane stores at a superblock close (lane t of row g: m-tiles t, t + 4), read once per unit
    int64_t p_store[(NMT + 3) / 4];
#pragma unroll
    for (int r = 0; r < (NMT + 3) / 4; ++r) {
        const int mt_mine = 4 * r + t;
        p_store[r] = (mt_mine < NMT && ((active_mask >> mt_mine) & 1u)) ? (int64_t)slot_point[mt_mine * 8 + g] : -1;
    }

  while (cur >= 0) {
    const int64_t sb_begin = (int64_t)cur * sb_per;
    int64_t sb_end = sb_begin + sb_per;
    if (sb_end > n_super) sb_end = n_super;
    for (int64_t sb = sb_begin; sb < sb_end; ++sb) {
        const int64_t left = N - sb * BI_SUPERBLOCK;
        const int n_ev = left < BI_SUPERBLOCK ? (int)left : BI_SUPERBLOCK;      // events of this superblock that exist
#pragma unroll 1
        for (int ti = 0; ti < TILES_PER_SUPER && ti * T < n_ev; ++ti) {
            const int st = (int)(consumed % Cfg::STAGES);
            bi_mbar_wait(&full_bar[st], (consumed / Cfg::STAGES) & 1u);
            const double* tile = ring + (size_t)st * Cfg::STAGE_DOUBLES;
            const int n_valid = n_ev - ti * T;                                   // may exceed T
            if (n_valid >= T && GROUP_DEFER) {
                // full tile: its groups run as one branch-free block, the rare path afterwards in group order (the order
                // in which the L_t accumulate)
                unsigned bad[GROUPS_PER_TILE], bad_any = 0;
#pragma unroll
                for (int gi = 0; gi < GROUPS_PER_TILE; ++gi) {
                    bad[gi] = bi_mma_group_fast<K4, NMT, false>(tile, gi * BI_GROUP_EVENTS, T, active_mask, a, M, E, lane);
                    bad_any |= bad[gi];
                }
                if (__any_sync(BI_FULL_MASK, bad_any != 0)) {
#pragma unroll
                    for (int gi = 0; gi < GROUPS_PER_TILE; ++gi)
                        if (__any_sync(BI_FULL_MASK, bad[gi] != 0))
                            bi_mma_group_slow<K4, NMT>(tile, gi * BI_GROUP_EVENTS, BI_GROUP_EVENTS, K, S, bad[gi], slot_point,
                                                       term_source, wterm, mus, outlier, slow_acc, slow_any, lane);
                }
            } else if (n_valid >= T) {
#pragma unroll 1
                for (int gi = 0; gi < GROUPS_PER_TILE; ++gi)
                    bi_mma_group<K4, NMT, false>(tile, gi * BI_GROUP_EVENTS, T, K, S, active_mask, a, slot_point, term_source,
                                                 wterm, mus, outlier, slow_acc, slow_any, M, E, lane);
            } else {
#pragma unroll 1
                for (int e0 = 0; e0 < n_valid; e0 += BI_GROUP_EVENTS) {
                    if (e0 + BI_GROUP_EVENTS <= n_valid)
                        bi_mma_group<K4, NMT, false>(tile, e0, n_valid, K, S, active_mask, a, slot_point, term_source, wterm,
                                                     mus, outlier, slow_acc, slow_any, M, E, lane);
                    else
                        bi_mma_group<K4, NMT, true>(tile, e0, n_valid, K, S, active_mask, a, slot_point, term_source, wterm,
                                                    mus, outlier, slow_acc, slow_any, M, E, lane);
                }
            }
            // this warp is done with the stage: refill it with the next tile of the stream
            __syncwarp();
            ++consumed;
            bi_mma_produce<K4>(prod, counter, n_ranges, sb_per, n_super, A, src_row, row_lead, ld, N, K, ring, full_bar,
                               tmap, tmap_rank, issued, lane);
        }
        // ---- close the superblock: combine the four classes, one log per point
        const bool any_slow = __any_sync(BI_FULL_MASK, slow_any);
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt) {
            double m = M[mt];
            m = __dmul_rn(m, __shfl_xor_sync(BI_FULL_MASK, m, 1));
            m = __dmul_rn(m, __shfl_xor_sync(BI_FULL_MASK, m, 2));
            int e = E[mt];
            e += __shfl_xor_sync(BI_FULL_MASK, e, 1);
            e += __shfl_xor_sync(BI_FULL_MASK, e, 2);
            M[mt] = m;
            E[mt] = e;
        }
        // lane t evaluates the log of m-tiles t, t + 4 (the four lanes of a row hold identical values)
#pragma unroll
        for (int r = 0; r < (NMT + 3) / 4; ++r) {
            double m = 1.0;
            int e = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (4 * r + q < NMT && t == q) { m = M[4 * r + q]; e = E[4 * r + q]; }
            double L = bi_block_log(m, e);
            if (any_slow) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (4 * r + q < NMT) {
                        double l = slow_acc[(4 * r + q) * Cfg::THREADS];
                        l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 1));
                        l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 2));
                        if (t == q) L = __dadd_rn(L, l);
                        slow_acc[(4 * r + q) * Cfg::THREADS] = 0.0;
                    }
                }
            }
            if (p_store[r] >= 0) partial[p_store[r] * n_super + sb] = L;
        }
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt) { M[mt] = 1.0; E[mt] = 0; }
        slow_any = false;
    }
    // ---- next range: the one the producer has already fetched (its first tiles are in flight)
    if (prod->pending < 0 && prod->range >= 0)
        bi_mma_produce<K4>(prod, counter, n_ranges, sb_per, n_super, A, src_row, row_lead, ld, N, K, ring, full_bar, tmap, tmap_rank, issued, lane);
    cur = prod->pending;
    __syncwarp();
    if (cur >= 0) {
        if (lane == 0) prod->pending = -1;
        __syncwarp();
        while (issued - consumed < (unsigned)Cfg::STAGES && prod->range >= 0 &&
               !(prod->tile == prod->n_tiles && prod->pending >= 0)) {
            const unsigned before = issued;
            bi_mma_produce<K4>(prod, counter, n_ranges, sb_per, n_super, A, src_row, row_lead, ld, N, K, ring, full_bar,
                               tmap, tmap_rank, issued, lane);
            if (issued == before) break;
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// kernel (persistent): every warp fetches work units u = (point group u % n_groups, superblock range
// u / n_groups) from the schedule header until all n_units are taken
// ---------------------------------------------------------------------------------------------
template <int K4>
__global__ void __launch_bounds__(BiMmaCfg<K4>::THREADS, BiMmaCfg<K4>::MIN_CTAS)
k_unbinned_mma(const double* __restrict__ A, int64_t ld, int64_t N, int K, int S,
               const int32_t* __restrict__ group_points, int4* groups, int32_t* header,
               int64_t n_super, const int32_t* __restrict__ row, const double* __restrict__ coef,
               const double* __restrict__ wterm, const int32_t* __restrict__ term_source,
               const double* __restrict__ mus, double outlier, double* __restrict__ partial,
               const __grid_constant__ CUtensorMap tmap, int tmap_rank, const int32_t* __restrict__ cell, int n_dims) {
    using Cfg = BiMmaCfg<K4>;
    constexpr int MT = Cfg::MT;
    extern __shared__ __align__(128) unsigned char bi_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bi_smem) + warp * Cfg::STAGES;
    double* ring = reinterpret_cast<double*>(bi_smem + Cfg::HEADER_BYTES) + (size_t)warp * Cfg::RING_DOUBLES;
    double* slow_acc = reinterpret_cast<double*>(bi_smem + Cfg::HEADER_BYTES) + (size_t)Cfg::WARPS * Cfg::RING_DOUBLES +
                       threadIdx.x;
    const int n_groups = header[0], n_ranges = header[1], sb_per = header[2];
    BiProducer* prod = reinterpret_cast<BiProducer*>(bi_smem + 128) + warp;

    if (lane == 0) {
        for (int i = 0; i < Cfg::STAGES; ++i) bi_mbar_init(&full_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) slow_acc[mt * Cfg::THREADS] = 0.0;
    // rows K..KP-1 of every stage are never written by the TMA copies: zero them once (0 * coef 0 adds nothing)
    for (int i = lane; i < (Cfg::KP - K) * Cfg::RS * Cfg::STAGES; i += 32) {
        const int st_i = i / ((Cfg::KP - K) * Cfg::RS), r = i - st_i * (Cfg::KP - K) * Cfg::RS;
        ring[(size_t)st_i * Cfg::STAGE_DOUBLES + K * Cfg::RS + r] = 0.0;
    }
    __syncwarp();
    unsigned issued = 0, consumed = 0;                             // ring position carries over from group to group

    for (;;) {
        // a point group with superblock ranges left: start at a ticket (spreads the warps over the groups), then scan
        int start = 0;
        if (lane == 0) start = (int)((unsigned)atomicAdd(&header[4], 1) % (unsigned)n_groups);
        start = __shfl_sync(BI_FULL_MASK, start, 0);
        int grp = -1;
        for (int base = 0; base < n_groups && grp < 0; base += 32) {      // 32 groups per probe, one per lane
            int cand = start + base + lane;
            if (cand >= n_groups) cand -= n_groups;
            const bool open = base + lane < n_groups && *reinterpret_cast<volatile int*>(&groups[cand].z) < n_ranges;
            const unsigned vote = __ballot_sync(BI_FULL_MASK, open);
            if (vote) grp = __shfl_sync(BI_FULL_MASK, cand, __ffs(vote) - 1);
        }
        if (grp < 0) break;
        const int4 gp = groups[grp];
        const int n_pts = gp.y;
        const int n_mt = (n_pts + 7) >> 3;
        int* counter = const_cast<int*>(&groups[grp].z);

        // slots: point of (m-tile mt, row g); slots beyond n_pts replay the group's first point
        const int32_t* slot_point = group_points + gp.x;
        unsigned active_mask = 0;                                  // bit mt: this lane's slot of m-tile mt is live
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
            if (mt * 8 + g < n_pts) active_mask |= 1u << mt;
        const int64_t lead = slot_point[0];
        const int32_t* row_lead = row + lead * K;                  // every point of the group has these rows
        const double* src_row = A;
        if (Cfg::ROWREG && lane < K && !tmap_rank) src_row = A + (int64_t)row_lead[lane] * ld;
        if (tmap_rank && lane == 0) {
            // tensor coordinates of the group's hypercube cell, last shape parameter first (one-point axis: cell -1 -> 0)
            int c[3] = {0, 0, 0};
            for (int d = 0; d < n_dims; ++d) {
                const int v = cell[lead * n_dims + (n_dims - 1 - d)];
                c[d] = v < 0 ? 0 : v;
            }
            prod->c2 = c[0];
            prod->c3 = c[1];
            prod->c4 = c[2];
        }
        __syncwarp();

#define BI_MMA_UNIT(NN)                                                                                              \
    case NN:                                                                                                         \
        if (NN <= MT)                                                                                                \
            bi_mma_unit<K4, (NN <= MT ? NN : 1)>(A, src_row, row_lead, ld, N, K, S, slot_point, active_mask, lead,   \
                                                 counter, n_ranges, sb_per, n_super, coef, term_source, wterm, mus,  \
                                                 outlier, partial, ring, full_bar, slow_acc, prod, &tmap, tmap_rank, \
                                                 issued, consumed, lane);                                            \
        break;
        switch (n_mt) {
            BI_MMA_UNIT(1) BI_MMA_UNIT(2) BI_MMA_UNIT(3) BI_MMA_UNIT(4)
            BI_MMA_UNIT(5) BI_MMA_UNIT(6) BI_MMA_UNIT(7) BI_MMA_UNIT(8)
        }
#undef BI_MMA_UNIT
    }
}

// resident CTAs of the persistent kernel on the current device
template <int K4>
static int bi_mma_grid(int* blocks) {
    static int cached = 0;
    if (!cached) {
        int dev = 0, sms = 0, per_sm = 0;
        BI_CUDA_CHECK(cudaGetDevice(&dev));
        BI_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        BI_CUDA_CHECK(cudaFuncSetAttribute(k_unbinned_mma<K4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           BiMmaCfg<K4>::SMEM_BYTES));
        BI_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_unbinned_mma<K4>, BiMmaCfg<K4>::THREADS,
                                                                    BiMmaCfg<K4>::SMEM_BYTES));
        BI_REQUIRE(per_sm >= 1, "k_unbinned_mma<%d> does not fit on this device", K4);
        cached = sms * per_sm;
    }
    *blocks = cached;
    return BI_OK;
}

template <int K4>
static int bi_launch_mma(const double* A, int64_t ld, int64_t N, int K, int S, const int32_t* group_points,
                         int32_t* groups, int32_t* header, int64_t n_super, const int32_t* row, const double* coef,
                         const double* wterm, const int32_t* term_source, const double* mus, double outlier,
                         double* partial, const CUtensorMap& tmap, int tmap_rank, const int32_t* cell, int n_dims,
                         cudaStream_t st) {
    using Cfg = BiMmaCfg<K4>;
    int blocks = 0;
    int rc = bi_mma_grid<K4>(&blocks);
    if (rc != BI_OK) return rc;
    k_unbinned_mma<K4><<<(unsigned)blocks, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(
        A, ld, N, K, S, group_points, reinterpret_cast<int4*>(groups), header, n_super, row, coef, wterm,
        term_source, mus, outlier, partial, tmap, tmap_rank, cell, n_dims);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

// events per tile row of the K4 instantiation (box inner size of the tensor map)
static int bi_mma_row_stride(int k4) {
    switch (k4) {
        case 1: return BiMmaCfg<1>::RS;  case 2: return BiMmaCfg<2>::RS;  case 3: return BiMmaCfg<3>::RS;
        case 4: return BiMmaCfg<4>::RS;  case 5: return BiMmaCfg<5>::RS;  case 6: return BiMmaCfg<6>::RS;
        case 7: return BiMmaCfg<7>::RS;  case 8: return BiMmaCfg<8>::RS;  case 12: return BiMmaCfg<12>::RS;
        case 16: return BiMmaCfg<16>::RS; case 24: return BiMmaCfg<24>::RS; default: return BiMmaCfg<32>::RS;
    }
}
