// blueice_b200 -- K2 for LONG contractions (K = C*S > 128 terms): the DMMA kernel with a K-chunk loop.
//
// Same replacement and same canonical arithmetic as bi_unbinned_mma.cuh (blueice/likelihood.py:355-356,678-690 and
// scipy/interpolate/_rgi.py:520-549):  f(theta_p, x_i) = fma chain over k = 0..K-1 of A[row_k, i] * coef_{p,k},
// 32-event groups -> pair / quad / oct product tree per class -> (m, e), 16 groups per superblock, ONE log per 512 events;
// the reference-semantics fallback (nansum, outlier replacement, likelihood.py:686-689) where a density leaves the fast range.
//
// What differs is the data movement.  With hundreds of rows per event the per-warp tile ring of k_unbinned_mma has room for
// one 8-point m-tile per warp and one warp per CTA: every 8 points stream all K rows from L2 again.  Here a CTA owns a unit
// = (point group of <= 64 points of one hypercube cell, superblock range) and its 4 consumer warps (2 m-tiles = 16 points
// each) SHARE the event tiles.  A tile = 64 events; the contraction is cut into chunks of 32 terms, and a fifth, producer
// warp brings one chunk per ring stage with 1-D TMA bulk copies: the 32 rows x 64 events of the tile (B operand, lane r
// copies row r) and the 64 points x 32 coefficients of the chunk (A operand) -- ONE copy, because k_wide_pack_coef has
// laid the coefficients out chunk-major in the order of the schedule ([chunk][slot][36], zero beyond K), so the rows of a
// unit's points are one contiguous block.  Both operands are then conflict-free LDS.64 (row pitches = 32 B mod 128 B),
// nothing the consumers need comes from global memory, a consumer thread holds little more than its accumulators
// (128 registers), and with a 2-stage ring three CTAs = 12 consumer warps share an SM.
// The DMMA accumulators of a tile (2 m-tiles x 8 octets) stay in registers across the chunks, so the chain over k is the
// same sequential fma chain (DMMA accumulates in k order, profiles/microbench/dmma_probe_b200.log) however K is cut.
// History (profiles/r2_wide_kernel.md): coefficients prefetched from L2 into registers, 2 CTAs / SM: 18 TFLOP/s at
// K = 160; coefficient rows copied point by point by one producer warp: the producer (a bulk copy is issued lane by lane)
// could not keep up; by three producer warps, 2 CTAs / SM: 23.5-27 TFLOP/s.
//
// Units are handed out statically (unit u = blockIdx.x + j * gridDim.x of the device schedule, bi_plan.cu), so the
// producer runs ahead across unit boundaries without talking to the consumers.
#include <stdlib.h>

#include "bi_unbinned_mma.cuh"

#define BI_WIDE_WARPS 4                 /* consumer warps per CTA */
#define BI_WIDE_MT 2                    /* 8-point m-tiles per consumer warp */
#define BI_WIDE_POINTS (8 * BI_WIDE_WARPS * BI_WIDE_MT)
#define BI_WIDE_TE 64                   /* events per tile (2 canonical groups) */
#define BI_WIDE_RS (BI_WIDE_TE + 4)     /* B row stride: the four k-lanes of a fragment load hit distinct banks */
#define BI_WIDE_KC 32                   /* terms per chunk */
#define BI_WIDE_AS (BI_WIDE_KC + 4)     /* A row stride (one row per point) */
#ifndef BI_WIDE_STAGES
#define BI_WIDE_STAGES 2
#endif
#define BI_WIDE_PRODUCERS 1
#define BI_WIDE_THREADS ((BI_WIDE_WARPS + BI_WIDE_PRODUCERS) * 32)
#define BI_WIDE_B_DOUBLES (BI_WIDE_KC * BI_WIDE_RS)
#define BI_WIDE_A_DOUBLES (BI_WIDE_POINTS * BI_WIDE_AS)
#define BI_WIDE_STAGE_DOUBLES (BI_WIDE_B_DOUBLES + BI_WIDE_A_DOUBLES)
#define BI_WIDE_HEADER_BYTES 128        /* full[STAGES], empty[STAGES] */
#define BI_WIDE_SMEM_BYTES (BI_WIDE_HEADER_BYTES + BI_WIDE_STAGES * BI_WIDE_STAGE_DOUBLES * 8)
#define BI_WIDE_OCTETS (BI_WIDE_TE / 8)

__device__ __forceinline__ void bi_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bi_smem_u32(bar)) : "memory");
}

// density of one event with the reference's semantics (likelihood.py:686-689), rows read from global memory (by the time
// a group turns out to need it the early chunks of its tile have left shared memory); as bi_slow_density_rows
static __device__ __noinline__ double bi_wide_slow_density(const double* __restrict__ A, int64_t ld,
                                                           const int32_t* __restrict__ row_lead, int64_t ev, int K, int S,
                                                           const int32_t* __restrict__ term_source,
                                                           const double* __restrict__ wterm,
                                                           const double* __restrict__ mu, double outlier) {
    double acc = 0.0;
    for (int s = 0; s < S; ++s) {
        double ps = 0.0;
        for (int k = 0; k < K; ++k)
            if (term_source[k] == s) ps = fma(A[(int64_t)row_lead[k] * ld + ev], wterm[k], ps);
        const double term = __dmul_rn(mu[s], ps);
        if (term == term) acc = __dadd_rn(acc, term);       // nansum: NaN terms count as 0
    }
    return bi_fix_density(acc, outlier);
}

// (t, group) fallback: the canonical tree over log(p_i) of the class's 8 events (as bi_slow_group); ev0 = first event of
// the group, n_valid = events of the group that exist (others count as p = 1)
static __device__ __noinline__ double bi_wide_slow_group(const double* __restrict__ A, int64_t ld,
                                                         const int32_t* __restrict__ row_lead, int64_t ev0, int K, int S,
                                                         int t, int n_valid, const int32_t* __restrict__ term_source,
                                                         const double* __restrict__ wterm,
                                                         const double* __restrict__ mu, double outlier) {
    double quad[2];
    for (int h = 0; h < 2; ++h) {
        double pr[2];
        for (int n = 0; n < 2; ++n) {
            const int e = 8 * (2 * h + n) + 2 * t;
            const double p0 = (e < n_valid) ? bi_wide_slow_density(A, ld, row_lead, ev0 + e, K, S, term_source, wterm, mu, outlier) : 1.0;
            const double p1 = (e + 1 < n_valid) ? bi_wide_slow_density(A, ld, row_lead, ev0 + e + 1, K, S, term_source, wterm, mu, outlier) : 1.0;
            pr[n] = __dadd_rn(log(p0), log(p1));
        }
        quad[h] = __dadd_rn(pr[0], pr[1]);
    }
    return __dadd_rn(quad[0], quad[1]);
}

// coefficients in the order the K-chunk kernel reads them: out[chunk][slot][36] = coef[group_points[slot]][32 chunk + j]
// for j < 32 and terms below K, 0 otherwise (pad columns, terms beyond K, slots beyond the evaluable points: a unit always
// copies 64 rows).  One warp per (chunk, slot).
__global__ void __launch_bounds__(256)
k_wide_pack_coef(const double* __restrict__ coef, const int32_t* __restrict__ group_points,
                 const int32_t* __restrict__ header, int K, int64_t n_slots, int n_chunks, double* __restrict__ out) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_slots * n_chunks) return;
    const int ch = (int)(w / n_slots);
    const int64_t slot = w - (int64_t)ch * n_slots;
    const int n_eval = header[5];
    const int k = ch * BI_WIDE_KC + lane;
    double v = 0.0;
    if (slot < n_eval && k < K) v = coef[(int64_t)group_points[slot] * K + k];
    double* dst = out + w * BI_WIDE_AS;
    dst[lane] = v;
    if (lane < BI_WIDE_AS - BI_WIDE_KC) dst[BI_WIDE_KC + lane] = 0.0;
}

// the DMMAs of one chunk: d[mt][n] += sum over the chunk's k-steps of A(point rows, terms 4 kk + t) x B(term rows 4 kk + t,
// octet n); acol[mt] / bcol point at this lane's first elements.  FULL: all 32 terms of the chunk exist; otherwise k-steps
// beyond K are skipped and terms beyond K (stale shared memory) read as 0
// NMT: m-tiles of this warp that hold points (a warp with one live m-tile skips the DMMAs of the other: few-point batches)
template <bool FULL, int NMT>
__device__ __forceinline__ void bi_wide_chunk(const double* __restrict__ bcol, const double* const (&acol)[BI_WIDE_MT],
                                              double (&d)[BI_WIDE_MT][BI_WIDE_OCTETS][2], int k_left, int t) {
#pragma unroll
    for (int kk = 0; kk < BI_WIDE_KC / 4; ++kk) {
        if (!FULL && 4 * kk >= k_left) break;
        double a[NMT], b[BI_WIDE_OCTETS];
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt) a[mt] = acol[mt][4 * kk];        // packed with zeros beyond K
#pragma unroll
        for (int n = 0; n < BI_WIDE_OCTETS; ++n) {
            b[n] = bcol[4 * kk * BI_WIDE_RS + 8 * n];
            if (!FULL && 4 * kk + t >= k_left) b[n] = 0.0;
        }
#pragma unroll
        for (int n = 0; n < BI_WIDE_OCTETS; ++n)
#pragma unroll
            for (int mt = 0; mt < NMT; ++mt) bi_dmma(d[mt][n][0], d[mt][n][1], a[mt], b[n]);
    }
}

__global__ void __maxnreg__(128)
k_unbinned_mma_wide(const double* __restrict__ A, int64_t ld, int64_t N, int K, int S,
                    const int32_t* __restrict__ group_points, const int4* __restrict__ groups,
                    const int32_t* __restrict__ header, int64_t n_super, const int32_t* __restrict__ row,
                    const double* __restrict__ coef_chunks, int64_t n_slots, const double* __restrict__ wterm,
                    const int32_t* __restrict__ term_source, const double* __restrict__ mus, double outlier,
                    double* __restrict__ partial) {
    extern __shared__ __align__(128) unsigned char bi_wide_smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bi_wide_smem);
    uint64_t* empty_bar = full_bar + BI_WIDE_STAGES;
    double* ring = reinterpret_cast<double*>(bi_wide_smem + BI_WIDE_HEADER_BYTES);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_groups = header[0], sb_per = header[2], n_units = header[3];
    const int n_chunks = (K + BI_WIDE_KC - 1) / BI_WIDE_KC;

    if (threadIdx.x == 0) {
        for (int i = 0; i < BI_WIDE_STAGES; ++i) {
            bi_mbar_init(&full_bar[i], BI_WIDE_PRODUCERS);
            bi_mbar_init(&empty_bar[i], BI_WIDE_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int st = 0;                 // ring position and its phase bit: the same sequence of (unit, superblock, tile, chunk) on
    unsigned phase = 0;         // the producer and on every consumer warp

    if (warp >= BI_WIDE_WARPS) {
        // ---------------- producer warp: lane r copies term row r of the chunk, lane 0 also the unit's coefficient block
        bool first_lap = true;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int grp = u % n_groups, range = u / n_groups;
            const int4 gp = groups[grp];
            const int64_t lead = group_points[gp.x];
            const int32_t* row_lead = row + lead * K;                     // every point of the group has these rows
            const double* a_src = coef_chunks + (int64_t)gp.x * BI_WIDE_AS;   // + chunk * n_slots * AS
            // coefficient rows of the m-tiles that hold points (a group of 32 points copies half the block)
            const unsigned a_bytes = (unsigned)(((gp.y + 7) & ~7) < BI_WIDE_POINTS ? ((gp.y + 7) & ~7) : BI_WIDE_POINTS) *
                                     (unsigned)(BI_WIDE_AS * sizeof(double));
            const int64_t sb_begin = (int64_t)range * sb_per;
            const int64_t sb_end = sb_begin + sb_per < n_super ? sb_begin + sb_per : n_super;
            for (int64_t sb = sb_begin; sb < sb_end; ++sb) {
                const int64_t left = N - sb * BI_SUPERBLOCK;
                const int n_ev = left < BI_SUPERBLOCK ? (int)left : BI_SUPERBLOCK;
                for (int ti = 0; ti * BI_WIDE_TE < n_ev; ++ti) {
                    const int64_t ev = sb * BI_SUPERBLOCK + (int64_t)ti * BI_WIDE_TE;
                    const int64_t cols = ld - ev < BI_WIDE_TE ? ld - ev : BI_WIDE_TE;      // even (ld is)
                    const unsigned b_bytes = (unsigned)cols * (unsigned)sizeof(double);
                    for (int ch = 0; ch < n_chunks; ++ch) {
                        if (!first_lap) bi_mbar_wait(&empty_bar[st], phase ^ 1u);          // the stage's previous use is consumed
                        const int n_terms = K - ch * BI_WIDE_KC < BI_WIDE_KC ? K - ch * BI_WIDE_KC : BI_WIDE_KC;
                        double* stage = ring + (size_t)st * BI_WIDE_STAGE_DOUBLES;
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        if (lane == 0) {
                            bi_mbar_expect_tx(&full_bar[st], b_bytes * (unsigned)n_terms + a_bytes);
                            bi_bulk_g2s(stage + BI_WIDE_B_DOUBLES, a_src + (int64_t)ch * n_slots * BI_WIDE_AS, a_bytes,
                                        &full_bar[st]);
                        }
                        __syncwarp();
                        if (lane < n_terms)
                            bi_bulk_g2s(stage + (size_t)lane * BI_WIDE_RS,
                                        A + (int64_t)row_lead[ch * BI_WIDE_KC + lane] * ld + ev, b_bytes, &full_bar[st]);
                        if (++st == BI_WIDE_STAGES) { st = 0; phase ^= 1u; first_lap = false; }
                    }
                }
            }
        }
        return;
    }

    // ---------------- consumer warps
    const int g = lane >> 2, t = lane & 3;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int grp = u % n_groups, range = u / n_groups;
        const int4 gp = groups[grp];
        const int n_pts = gp.y;
        const int32_t* slot_point = group_points + gp.x;
        const int64_t lead = slot_point[0];
        const int32_t* row_lead = row + lead * K;
        // m-tiles of this warp: two adjacent ones, or -- a group of at most 8 * WARPS points -- one per warp, so that a
        // half-filled unit keeps all four warps issuing (a warp alone issues a DMMA every ~32 cycles, the pipe takes one
        // every 16)
        const bool spread = n_pts <= 8 * BI_WIDE_WARPS;
        const int tile0 = spread ? warp : warp * BI_WIDE_MT;
        const bool warp_active = tile0 * 8 < n_pts;                       // warp-uniform
        const bool two_tiles = !spread && (tile0 + 1) * 8 < n_pts;        // warp-uniform: the second m-tile holds points
        bool live[BI_WIDE_MT];
        int64_t p_slot[BI_WIDE_MT];
        int a_off[BI_WIDE_MT];                                            // this lane's first coefficient inside a stage
#pragma unroll
        for (int mt = 0; mt < BI_WIDE_MT; ++mt) {
            const int idx = (tile0 + mt) * 8 + g;
            live[mt] = idx < n_pts && (mt == 0 || !spread);
            p_slot[mt] = live[mt] ? (int64_t)slot_point[idx] : lead;      // dead slots replay the group's first point
            a_off[mt] = BI_WIDE_B_DOUBLES + idx * BI_WIDE_AS + t;
        }

        const int64_t sb_begin = (int64_t)range * sb_per;
        const int64_t sb_end = sb_begin + sb_per < n_super ? sb_begin + sb_per : n_super;
        for (int64_t sb = sb_begin; sb < sb_end; ++sb) {
            const int64_t left = N - sb * BI_SUPERBLOCK;
            const int n_ev = left < BI_SUPERBLOCK ? (int)left : BI_SUPERBLOCK;
            double M[BI_WIDE_MT], L[BI_WIDE_MT];
            int E[BI_WIDE_MT];
#pragma unroll
            for (int mt = 0; mt < BI_WIDE_MT; ++mt) { M[mt] = 1.0; L[mt] = 0.0; E[mt] = 0; }

            for (int ti = 0; ti * BI_WIDE_TE < n_ev; ++ti) {
                double d[BI_WIDE_MT][BI_WIDE_OCTETS][2];
#pragma unroll
                for (int mt = 0; mt < BI_WIDE_MT; ++mt)
#pragma unroll
                    for (int n = 0; n < BI_WIDE_OCTETS; ++n) d[mt][n][0] = d[mt][n][1] = 0.0;

#pragma unroll 1
                for (int ch = 0; ch < n_chunks; ++ch) {
                    bi_mbar_wait(&full_bar[st], phase);
                    if (warp_active) {
                        const double* stage = ring + (size_t)st * BI_WIDE_STAGE_DOUBLES;
                        const double* bcol = stage + t * BI_WIDE_RS + g;
                        const double* acol[BI_WIDE_MT];
#pragma unroll
                        for (int mt = 0; mt < BI_WIDE_MT; ++mt) acol[mt] = stage + a_off[mt];
                        const int k_left = K - ch * BI_WIDE_KC;
                        if (two_tiles) {
                            if (k_left >= BI_WIDE_KC) bi_wide_chunk<true, 2>(bcol, acol, d, k_left, t);
                            else bi_wide_chunk<false, 2>(bcol, acol, d, k_left, t);
                        } else {
                            if (k_left >= BI_WIDE_KC) bi_wide_chunk<true, 1>(bcol, acol, d, k_left, t);
                            else bi_wide_chunk<false, 1>(bcol, acol, d, k_left, t);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) bi_mbar_arrive(&empty_bar[st]);        // this warp is done with the stage
                    if (++st == BI_WIDE_STAGES) { st = 0; phase ^= 1u; }
                }

                if (!warp_active) continue;
                // ---- the tile's two canonical groups: product tree per class, rare path from global memory
                const int n_valid = n_ev - ti * BI_WIDE_TE;               // may exceed the tile
#pragma unroll
                for (int h = 0; h < BI_WIDE_TE / BI_GROUP_EVENTS; ++h) {
                    if (h * BI_GROUP_EVENTS >= n_valid) break;            // warp-uniform: a group of p = 1 changes nothing
                    unsigned bad = 0;
#pragma unroll
                    for (int mt = 0; mt < BI_WIDE_MT; ++mt) {
                        double (&dd)[BI_WIDE_OCTETS][2] = d[mt];
#pragma unroll
                        for (int n = 4 * h; n < 4 * h + 4; ++n) {         // events >= N count as p = 1
                            const int e = 8 * n + 2 * t;
                            if (e >= n_valid) dd[n][0] = 1.0;
                            if (e + 1 >= n_valid) dd[n][1] = 1.0;
                        }
                        unsigned tmax = 0;
#pragma unroll
                        for (int n = 4 * h; n < 4 * h + 4; ++n) {
                            tmax = max(tmax, (unsigned)(__double2hiint(dd[n][0]) - BI_RANGE_LO));
                            tmax = max(tmax, (unsigned)(__double2hiint(dd[n][1]) - BI_RANGE_LO));
                        }
                        const double q0 = __dmul_rn(__dmul_rn(dd[4 * h][0], dd[4 * h][1]), __dmul_rn(dd[4 * h + 1][0], dd[4 * h + 1][1]));
                        const double q1 = __dmul_rn(__dmul_rn(dd[4 * h + 2][0], dd[4 * h + 2][1]), __dmul_rn(dd[4 * h + 3][0], dd[4 * h + 3][1]));
                        const double oct = __dmul_rn(q0, q1);
                        double m;
                        int e;
                        bi_split(oct, &m, &e);
                        if (tmax >= BI_RANGE_SPAN) {
                            m = 1.0;
                            e = 0;
                            if (live[mt]) bad |= 1u << mt;
                        }
                        M[mt] = __dmul_rn(M[mt], m);
                        E[mt] += e;
                    }
                    if (__any_sync(BI_FULL_MASK, bad != 0)) {
                        const int nv = n_valid - h * BI_GROUP_EVENTS < BI_GROUP_EVENTS ? n_valid - h * BI_GROUP_EVENTS : BI_GROUP_EVENTS;
                        const int64_t ev0 = sb * BI_SUPERBLOCK + (int64_t)ti * BI_WIDE_TE + h * BI_GROUP_EVENTS;
#pragma unroll
                        for (int mt = 0; mt < BI_WIDE_MT; ++mt)
                            if ((bad >> mt) & 1u)
                                L[mt] = __dadd_rn(L[mt], bi_wide_slow_group(A, ld, row_lead, ev0, K, S, t, nv, term_source,
                                                                             wterm + p_slot[mt] * K, mus + p_slot[mt] * S, outlier));
                    }
                }
            }

            // ---- close the superblock: combine the four classes, one log per point
            if (warp_active) {
#pragma unroll
                for (int mt = 0; mt < BI_WIDE_MT; ++mt) {
                    double m = M[mt];
                    m = __dmul_rn(m, __shfl_xor_sync(BI_FULL_MASK, m, 1));
                    m = __dmul_rn(m, __shfl_xor_sync(BI_FULL_MASK, m, 2));
                    int e = E[mt];
                    e += __shfl_xor_sync(BI_FULL_MASK, e, 1);
                    e += __shfl_xor_sync(BI_FULL_MASK, e, 2);
                    double l = L[mt];
                    l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 1));
                    l = __dadd_rn(l, __shfl_xor_sync(BI_FULL_MASK, l, 2));
                    const double total = __dadd_rn(bi_block_log(m, e), l);
                    if (live[mt] && t == 0) partial[p_slot[mt] * n_super + sb] = total;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// launcher (called by bi_unbinned_partials_mma for long contractions)
// ---------------------------------------------------------------------------------------------
// doubles of the chunk-major coefficient copy for a batch of n_points points
int64_t bi_mma_wide_scratch_doubles(int K, int64_t n_points) {
    const int64_t n_chunks = (K + BI_WIDE_KC - 1) / BI_WIDE_KC;
    return n_chunks * (n_points + BI_WIDE_POINTS) * BI_WIDE_AS;
}

int bi_launch_mma_wide(const double* A, int64_t ld, int64_t N, int K, int S, const int32_t* group_points,
                       int32_t* groups, int32_t* header, int64_t n_super, const int32_t* row, const double* coef,
                       const double* wterm, const int32_t* term_source, const double* mus, double outlier,
                       double* partial, int64_t n_points, double* coef_chunks, cudaStream_t st) {
    BI_REQUIRE(coef_chunks && n_points >= 1, "bi_unbinned_partials_mma: contractions of %d terms need n_points and "
               "coef_chunks_dev (bi_mma_coef_chunks_doubles(n_terms, n_points) doubles)", K);
    BI_REQUIRE(((uintptr_t)coef_chunks & 15) == 0, "coef_chunks_dev must be 16-byte aligned");
    static int blocks = 0;
    if (!blocks) {
        int dev = 0, sms = 0, per_sm = 0;
        BI_CUDA_CHECK(cudaGetDevice(&dev));
        BI_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        BI_CUDA_CHECK(cudaFuncSetAttribute(k_unbinned_mma_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, BI_WIDE_SMEM_BYTES));
        BI_CUDA_CHECK(cudaFuncSetAttribute(k_unbinned_mma_wide, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           cudaSharedmemCarveoutMaxShared));
        BI_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_unbinned_mma_wide, BI_WIDE_THREADS, BI_WIDE_SMEM_BYTES));
        BI_REQUIRE(per_sm >= 1, "k_unbinned_mma_wide does not fit on this device");
        blocks = sms * per_sm;
        if (getenv("BI_WIDE_VERBOSE")) fprintf(stderr, "k_unbinned_mma_wide: %d CTAs per SM, %d bytes of shared memory\n", per_sm, BI_WIDE_SMEM_BYTES);
    }
    const int n_chunks = (K + BI_WIDE_KC - 1) / BI_WIDE_KC;
    const int64_t n_slots = n_points + BI_WIDE_POINTS;
    const int64_t pack_warps = n_slots * n_chunks;
    k_wide_pack_coef<<<(unsigned)((pack_warps * 32 + 255) / 256), 256, 0, st>>>(coef, group_points, header, K, n_slots, n_chunks,
                                                                               coef_chunks);
    BI_LAUNCH_CHECK();
    k_unbinned_mma_wide<<<(unsigned)blocks, BI_WIDE_THREADS, BI_WIDE_SMEM_BYTES, st>>>(
        A, ld, N, K, S, group_points, reinterpret_cast<const int4*>(groups), header, n_super, row, coef_chunks, n_slots,
        wterm, term_source, mus, outlier, partial);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

int bi_mma_wide_unit_points(void) { return BI_WIDE_POINTS; }
