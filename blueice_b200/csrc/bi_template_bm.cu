// blueice_b200 -- K5c: the densities of a toy-Monte-Carlo sweep formed BIN-MAJOR (one parameter point per toy).
//
// K5 (bi_template.cu) walks the toys one by one and gathers, per event, the K = corners x sources template rows of the
// toy's hypercube cell from L2: 2^n_space * K scattered 8-byte values per event.  ncu (profiles/r2_k5_toys_ncu.md): the
// L1 data pipe is saturated by those gathers (one 32-byte sector per row and event), 9.8 ms per sweep of 1e8 events.
// Here the events of ALL toys are sorted by their low-corner bin once per toy set (plumbing, engine.set_datasets), and a
// CTA takes (bin, chunk of <= 2048 events of that bin):
//   * the bin's packed lookup corners of ALL G * S template rows come with ONE TMA bulk copy from a bin-major copy of the
//     templates [bin][row][pack] (12 kB at 125 anchors x 3 sources) -- every template byte is read once per chunk;
//   * the chunk's events are bucketed by the hypercube cell of their toy's point (shared-memory counting sort), so the
//     lanes of a warp read the SAME rows: broadcast shared-memory loads instead of 32 scattered sectors;
//   * per event only the toy's record travels: fractions, mus and cell of its point, 8 * (1 + D + S) bytes (two sectors);
//     weights and coefficients are re-formed from it exactly as K1 does (w = ((1 * t_0) * t_1) ..., coef = fl(w * mu)).
// Every density is formed by the operations of K5 in K5's order -- lookup in scipy's operation order, fma chain over
// k = c * S + s -- and written to p[event] in toy order; the range test, the canonical product tree, the rare
// reference-semantics path and the finalize are K5's own code (k_template_partials with `pre`), so a sweep is
// BIT-IDENTICAL to K5 and to set_data + ll on the anchor-tensor engine.
//
// Replaces, per toy: blueice/likelihood.py:531-562 (set_data -> score_events), model.py:97-99, source.py:225-240 and
// the density of likelihood.py:678-690.
#include <stdlib.h>
#include <string.h>

#include "bi_plan.cuh"
#include "bi_tma.cuh"
#include "bi_ts.cuh"

#ifndef BI_BM_THREADS
#define BI_BM_THREADS 256
#endif
#ifndef BI_BM_CHUNK
#define BI_BM_CHUNK 2048                 /* events per task */
#endif
#define BI_BM_PER_THREAD (BI_BM_CHUNK / BI_BM_THREADS)
#define BI_BM_MAX_SOURCES 8
#ifndef BI_BM_MIN_CTAS
#define BI_BM_MIN_CTAS 3
#endif
#define BI_BM_MAX_DIMS 4
#define BI_BM_MAX_CELLS 4096
#define BI_BM_MAX_ROW_BYTES (160 * 1024)

struct BiBmArgs {
    const double* tbm;                   // [n_bins][n_rows][pack] bin-major packed templates
    const int32_t* task_bin;             // [n_tasks]
    const int64_t* task_start;           // [n_tasks] first event (bin-sorted position)
    const int32_t* task_count;           // [n_tasks] <= BI_BM_CHUNK
    const int32_t* bm_toy;               // [N] dataset (= point) of the bin-sorted event
    const int32_t* bm_src;               // [N] position of the event in toy order
    const double* bm_frac;               // [NS][ld_bm] lookup fractions, bin-sorted
    const double* rec;                   // [P][R] point records: {int2(base anchor, cell id)}, frac [D], mus [S]
    double* pbuf;                        // [N] densities, toy order
    int64_t n_tasks, ld_bm;
    int32_t n_rows, pack, S, R, n_cells, row_bytes, stages;
    int32_t corner_delta[1 << BI_BM_MAX_DIMS];   // anchor-index offset of morph corner c from corner 0
};

// point records (one thread per point): the inputs of the per-event contraction, 32-byte aligned
struct BiBmStrides { int32_t v[BI_MAX_DIMS]; };

__global__ void __launch_bounds__(256)
k_bm_pack(const __grid_constant__ BiPlanDims dims, const __grid_constant__ BiBmStrides anchor_stride, int S, int R, int64_t n_points,
          const int32_t* __restrict__ cell, const double* __restrict__ frac, const double* __restrict__ mus,
          const int32_t* __restrict__ status, int n_cells, double* __restrict__ rec) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_points) return;
    const int D = dims.n_dims;
    int base = 0;
    for (int d = 0; d < D; ++d) {
        const int c = cell[p * D + d];
        base += (c < 0 ? 0 : c) * anchor_stride.v[d];                 // one-point axis: cell -1 aliases anchor 0
    }
    const int cid = status[p] == 0 ? bi_flat_cell(dims, cell + p * D) : n_cells;      // n_cells: not evaluated
    double* r = rec + p * R;
    r[0] = __hiloint2double(cid, base);
    for (int d = 0; d < D; ++d) r[1 + d] = frac[p * D + d];
    for (int s = 0; s < S; ++s) r[1 + D + s] = mus[p * S + s];
    for (int i = 1 + D + S; i < R; ++i) r[i] = 0.0;
}

// exclusive scan of v[0..n) in place over one CTA of BI_BM_THREADS threads; returns nothing (v[n] is not touched)
static __device__ void bi_bm_scan(int* v, int n, int* carry) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + BI_BM_THREADS - 1) / BI_BM_THREADS;
    const int lo = min(tid * per, n), hi = min(lo + per, n);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += v[i];
    int incl = sum;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
        const int o = __shfl_up_sync(BI_FULL_MASK, incl, k);
        if (lane >= k) incl += o;
    }
    if (lane == 31) carry[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < BI_BM_THREADS / 32 ? carry[lane] : 0;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const int o = __shfl_up_sync(BI_FULL_MASK, w, k);
            if (lane >= k) w += o;
        }
        if (lane < BI_BM_THREADS / 32) carry[lane] = w;
    }
    __syncthreads();
    int run = incl - sum + (warp ? carry[warp - 1] : 0);
    for (int i = lo; i < hi; ++i) { const int x = v[i]; v[i] = run; run += x; }
    __syncthreads();
}

// shared memory: [stages] row stages (row_bytes each, 128-byte aligned) | event slots: y [NS][CHUNK], toy, src |
// hist [n_cells + 2] | carry [32] | mbarriers [2]
template <int NS, int D, int S>
__global__ void __launch_bounds__(BI_BM_THREADS, BI_BM_MIN_CTAS)
k_bm_density(const __grid_constant__ BiBmArgs a) {
    constexpr int C = 1 << D;
    constexpr int PACK = NS == 1 ? 2 : 4;
    constexpr int R = (1 + D + S + 3) / 4 * 4;
    extern __shared__ __align__(128) unsigned char bi_bm_smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int row_stage = (a.row_bytes + 127) & ~127;
    double* s_rows = reinterpret_cast<double*>(bi_bm_smem);
    double* s_y = reinterpret_cast<double*>(bi_bm_smem + (size_t)a.stages * row_stage);            // [NS][CHUNK]
    int32_t* s_toy = reinterpret_cast<int32_t*>(s_y + NS * BI_BM_CHUNK);
    int32_t* s_src = s_toy + BI_BM_CHUNK;
    int* s_hist = s_src + BI_BM_CHUNK;                                                             // [n_cells + 2]
    int* s_carry = s_hist + a.n_cells + 2;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_carry + 32) + 7) & ~(uintptr_t)7);
    const int n_cells = a.n_cells;

    if (tid < 2) bi_mbar_init(&s_bar[tid], 1);
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    auto fetch_rows = [&](int64_t task, int buf) {               // one thread: the bin's rows by ONE bulk copy
        bi_mbar_expect_tx(&s_bar[buf], (unsigned)a.row_bytes);
        bi_bulk_g2s(reinterpret_cast<unsigned char*>(s_rows) + (size_t)buf * row_stage,
                    a.tbm + (int64_t)a.task_bin[task] * a.n_rows * a.pack, (unsigned)a.row_bytes, &s_bar[buf]);
    };
    // the record of one point: {int2(base anchor, cell id)}, frac [D], mus [S] -- R / 4 256-bit loads
    auto load_record = [&](int toy, double (&rb)[R]) {
        const double* rp = a.rec + (int64_t)toy * R;
#pragma unroll
        for (int i = 0; i < R / 4; ++i)
            asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
                : "=d"(rb[4 * i]), "=d"(rb[4 * i + 1]), "=d"(rb[4 * i + 2]), "=d"(rb[4 * i + 3]) : "l"(rp + 4 * i));
    };
    unsigned parity[2] = {0, 0};
    int buf = 0;
    if (tid == 0 && (int64_t)blockIdx.x < a.n_tasks) fetch_rows(blockIdx.x, 0);

    for (int64_t task = blockIdx.x; task < a.n_tasks; task += gridDim.x) {
        const int64_t start = a.task_start[task];
        const int count = a.task_count[task];
        // two stages: the next task's rows travel while this one is evaluated (the stage was released by the barrier that
        // closed the previous iteration); one stage (large row sets): fetched after the previous task is done
        if (a.stages == 2) {
            if (tid == 0 && task + gridDim.x < a.n_tasks) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                fetch_rows(task + gridDim.x, buf ^ 1);
            }
        } else if (task != (int64_t)blockIdx.x && tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            fetch_rows(task, 0);
        }
        // ---- phase 1: bucket the chunk's events by the hypercube cell of their toy's point
        for (int c = tid; c < n_cells + 2; c += BI_BM_THREADS) s_hist[c] = 0;
        __syncthreads();
        int toy[BI_BM_PER_THREAD], cid[BI_BM_PER_THREAD], pos[BI_BM_PER_THREAD];
#pragma unroll
        for (int i = 0; i < BI_BM_PER_THREAD; ++i) {
            const int j = tid + BI_BM_THREADS * i;
            toy[i] = j < count ? __ldg(a.bm_toy + start + j) : -1;
        }
#pragma unroll
        for (int i = 0; i < BI_BM_PER_THREAD; ++i)
            cid[i] = toy[i] >= 0 ? __double2hiint(__ldg(a.rec + (int64_t)toy[i] * R)) : n_cells + 1;
#pragma unroll
        for (int i = 0; i < BI_BM_PER_THREAD; ++i) {
            // warp-aggregated counter: one shared-memory atomic per distinct cell of the warp
            const unsigned peers = __match_any_sync(BI_FULL_MASK, cid[i]);
            const int leader = __ffs(peers) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&s_hist[cid[i]], __popc(peers));
            base = __shfl_sync(BI_FULL_MASK, base, leader);
            pos[i] = base + __popc(peers & ((1u << lane) - 1u));
        }
        __syncthreads();
        bi_bm_scan(s_hist, n_cells + 2, s_carry);                // s_hist[c] = first slot of cell c; [n_cells] = live events
        // ---- phase 2: the events into their slots
#pragma unroll
        for (int i = 0; i < BI_BM_PER_THREAD; ++i) {
            const int j = tid + BI_BM_THREADS * i;
            if (toy[i] >= 0) {
                const int slot = s_hist[cid[i]] + pos[i];
                s_toy[slot] = toy[i];
                s_src[slot] = __ldg(a.bm_src + start + j);
#pragma unroll
                for (int d = 0; d < NS; ++d) s_y[d * BI_BM_CHUNK + slot] = __ldg(a.bm_frac + (int64_t)d * a.ld_bm + start + j);
            }
        }
        __syncthreads();
        const int n_live = s_hist[n_cells];
        bi_mbar_wait(&s_bar[buf], parity[buf]);
        parity[buf] ^= 1;
        const double* rows = reinterpret_cast<const double*>(reinterpret_cast<const unsigned char*>(s_rows) + (size_t)buf * row_stage);

        // ---- phase 3: one event per thread and step; lanes of a warp mostly share their cell (broadcast row loads).
        // The record of the thread's next event and the rows of the next morph corner are in flight while one corner's
        // S rows are evaluated.
        double rb[R];
        if (tid < n_live) load_record(s_toy[tid], rb);
        for (int q = tid; q < n_live; q += BI_BM_THREADS) {
            double y[NS];
#pragma unroll
            for (int d = 0; d < NS; ++d) y[d] = s_y[d * BI_BM_CHUNK + q];
            const int out = s_src[q];
            const int base_anchor = __double2loint(rb[0]);
            double f1[D], f0[D], mu[S];
#pragma unroll
            for (int d = 0; d < D; ++d) { f1[d] = rb[1 + d]; f0[d] = __dsub_rn(1.0, f1[d]); }
#pragma unroll
            for (int s = 0; s < S; ++s) mu[s] = rb[1 + D + s];
            if (q + BI_BM_THREADS < n_live) load_record(s_toy[q + BI_BM_THREADS], rb);
            double v[2][S][PACK];
            auto load_corner = [&](int c, double (&vc)[S][PACK]) {
                const double* rc = rows + (size_t)(base_anchor + a.corner_delta[c]) * (S * PACK);
#pragma unroll
                for (int s = 0; s < S; ++s) {
#pragma unroll
                    for (int h = 0; h < PACK; h += 2) {
                        const double2 t = *reinterpret_cast<const double2*>(rc + s * PACK + h);
                        vc[s][h] = t.x;
                        vc[s][h + 1] = t.y;
                    }
                }
            };
            load_corner(0, v[0]);
            double p = 0.0;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                if (c + 1 < C) load_corner(c + 1, v[(c + 1) & 1]);
                double w = ((c >> (D - 1)) & 1) ? f1[0] : f0[0];                   // fl(1 * t_0) = t_0
#pragma unroll
                for (int d = 1; d < D; ++d) w = __dmul_rn(w, ((c >> (D - 1 - d)) & 1) ? f1[d] : f0[d]);
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    double vv[1 << NS];
#pragma unroll
                    for (int h = 0; h < PACK; ++h) vv[h] = v[c & 1][s][h];
                    p = fma(bi_ts_eval<NS>(vv, y), __dmul_rn(w, mu[s]), p);
                }
            }
            a.pbuf[out] = p;
        }
        __syncthreads();                                          // slots, histogram and this row stage may be reused
        if (a.stages == 2) buf ^= 1;
    }
}

template <int NS, int D, int S>
static int bi_bm_launch(const BiBmArgs& a, int smem, cudaStream_t st) {
    static int per_sm = 0, sms = 0, smem_set = 0;
    if (!per_sm || smem > smem_set) {
        int dev = 0;
        BI_CUDA_CHECK(cudaGetDevice(&dev));
        BI_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        BI_CUDA_CHECK(cudaFuncSetAttribute(k_bm_density<NS, D, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        smem_set = smem;
        BI_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bm_density<NS, D, S>, BI_BM_THREADS, smem));
        BI_REQUIRE(per_sm >= 1, "k_bm_density<%d,%d,%d> does not fit on this device (%d bytes of shared memory)", NS, D, S, smem);
    }
    int64_t blocks = (int64_t)sms * per_sm;
    if (blocks > a.n_tasks) blocks = a.n_tasks;
    k_bm_density<NS, D, S><<<(unsigned)blocks, BI_BM_THREADS, smem, st>>>(a);
    BI_LAUNCH_CHECK();
    return BI_OK;
}

template <int NS, int D>
static int bi_bm_launch_s(const BiBmArgs& a, int smem, cudaStream_t st) {
    switch (a.S) {
        case 1: return bi_bm_launch<NS, D, 1>(a, smem, st);
        case 2: return bi_bm_launch<NS, D, 2>(a, smem, st);
        case 3: return bi_bm_launch<NS, D, 3>(a, smem, st);
        case 4: return bi_bm_launch<NS, D, 4>(a, smem, st);
        case 5: return bi_bm_launch<NS, D, 5>(a, smem, st);
        case 6: return bi_bm_launch<NS, D, 6>(a, smem, st);
        case 7: return bi_bm_launch<NS, D, 7>(a, smem, st);
        default: return bi_bm_launch<NS, D, 8>(a, smem, st);
    }
}

extern "C" int bi_template_bm_supported(int32_t n_space, int32_t method, int32_t n_dims, const int32_t* n_anchors_host,
                                        int32_t n_sources, int64_t n_rows) {
    if (method != BI_LOOKUP_LINEAR || n_space < 1 || n_space > 2) return 0;
    if (n_dims < 1 || n_dims > BI_BM_MAX_DIMS || n_sources < 1 || n_sources > BI_BM_MAX_SOURCES || !n_anchors_host) return 0;
    int64_t n_cells = 1;
    for (int d = 0; d < n_dims; ++d) n_cells *= n_anchors_host[d] > 1 ? n_anchors_host[d] - 1 : 1;
    if (n_cells > BI_BM_MAX_CELLS) return 0;
    const int pack = n_space == 1 ? 2 : 4;
    if (n_rows * pack * 8 > BI_BM_MAX_ROW_BYTES) return 0;
    return 1;
}

extern "C" int32_t bi_template_bm_chunk(void) { return BI_BM_CHUNK; }

extern "C" int64_t bi_template_bm_record_doubles(int32_t n_dims, int32_t n_sources) {
    return (int64_t)(1 + n_dims + n_sources + 3) / 4 * 4;
}

extern "C" int bi_template_bm_density(const double* templates_bm_dev, int64_t n_rows, int32_t n_space,
                                      int32_t n_dims, const int32_t* n_anchors_host, int32_t n_sources, int64_t n_points,
                                      const int32_t* cell_dev, const double* frac_dev, const double* mus_dev,
                                      const int32_t* status_dev,
                                      const int32_t* task_bin_dev, const int64_t* task_start_dev,
                                      const int32_t* task_count_dev, int64_t n_tasks,
                                      const int32_t* bm_toy_dev, const int32_t* bm_src_dev, const double* bm_frac_dev,
                                      int64_t ld_bm, double* record_dev, double* density_dev, void* stream) {
    BI_REQUIRE(bi_template_bm_supported(n_space, BI_LOOKUP_LINEAR, n_dims, n_anchors_host, n_sources, n_rows),
               "bi_template_bm_density: unsupported shape (n_space=%d, n_dims=%d, n_sources=%d, n_rows=%lld)", n_space, n_dims,
               n_sources, (long long)n_rows);
    BI_REQUIRE(n_points >= 0 && n_tasks >= 0, "negative size");
    if (n_points == 0 || n_tasks == 0) return BI_OK;
    BI_REQUIRE(templates_bm_dev && cell_dev && frac_dev && mus_dev && status_dev && task_bin_dev && task_start_dev &&
                   task_count_dev && bm_toy_dev && bm_src_dev && bm_frac_dev && record_dev && density_dev,
               "bi_template_bm_density: NULL device pointer");
    BI_REQUIRE(((uintptr_t)templates_bm_dev & 15) == 0 && ((uintptr_t)record_dev & 31) == 0,
               "bi_template_bm_density: templates must be 16-byte and records 32-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    BiPlanDims dims;
    memset(&dims, 0, sizeof(dims));
    dims.n_dims = n_dims;
    BiBmStrides anchor_stride;
    memset(&anchor_stride, 0, sizeof(anchor_stride));
    int64_t n_cells = 1, n_anchor_total = 1;
    for (int d = n_dims - 1; d >= 0; --d) {
        dims.cells[d] = n_anchors_host[d] > 1 ? n_anchors_host[d] - 1 : 1;
        dims.stride[d] = (int32_t)n_cells;
        n_cells *= dims.cells[d];
        anchor_stride.v[d] = (int32_t)n_anchor_total;
        n_anchor_total *= n_anchors_host[d];
    }
    BI_REQUIRE(n_anchor_total * n_sources == n_rows, "n_rows=%lld is not anchors x sources = %lld", (long long)n_rows,
               (long long)(n_anchor_total * n_sources));
    BiBmArgs a;
    memset(&a, 0, sizeof(a));
    a.tbm = templates_bm_dev;
    a.task_bin = task_bin_dev; a.task_start = task_start_dev; a.task_count = task_count_dev;
    a.bm_toy = bm_toy_dev; a.bm_src = bm_src_dev; a.bm_frac = bm_frac_dev;
    a.rec = record_dev; a.pbuf = density_dev;
    a.n_tasks = n_tasks; a.ld_bm = ld_bm;
    a.n_rows = (int32_t)n_rows; a.pack = n_space == 1 ? 2 : 4; a.S = n_sources;
    a.R = (int32_t)bi_template_bm_record_doubles(n_dims, n_sources);
    a.n_cells = (int32_t)n_cells;
    a.row_bytes = (int32_t)(n_rows * a.pack * 8);
    a.stages = 2 * ((a.row_bytes + 127) & ~127) <= 64 * 1024 ? 2 : 1;
    for (int c = 0; c < (1 << n_dims); ++c) {
        int delta = 0;
        for (int d = 0; d < n_dims; ++d)
            if (((c >> (n_dims - 1 - d)) & 1) && n_anchors_host[d] > 1) delta += anchor_stride.v[d];
        a.corner_delta[c] = delta;
    }
    k_bm_pack<<<(unsigned)((n_points + 255) / 256), 256, 0, st>>>(dims, anchor_stride, n_sources, a.R, n_points, cell_dev, frac_dev,
                                                                 mus_dev, status_dev, (int)n_cells, record_dev);
    BI_LAUNCH_CHECK();
    const int smem = a.stages * ((a.row_bytes + 127) & ~127) + n_space * BI_BM_CHUNK * 8 + 2 * BI_BM_CHUNK * 4 +
                     ((int)n_cells + 2 + 32) * 4 + 8 + 2 * 8;
#define BI_BM_CASE(NSV, DV) if (n_space == NSV && n_dims == DV) return bi_bm_launch_s<NSV, DV>(a, smem, st);
    BI_BM_CASE(1, 1) BI_BM_CASE(1, 2) BI_BM_CASE(1, 3) BI_BM_CASE(1, 4)
    BI_BM_CASE(2, 1) BI_BM_CASE(2, 2) BI_BM_CASE(2, 3) BI_BM_CASE(2, 4)
#undef BI_BM_CASE
    bi_set_error("bi_template_bm_density: unsupported configuration");
    return BI_ERR_UNSUPPORTED;
}
