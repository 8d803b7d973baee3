// blueice_b200 -- helpers of the device-side schedules (bucketing points / toys by hypercube cell in one CTA).
#pragma once
#include "bi_common.cuh"

#define BI_PLAN_THREADS 1024

struct BiPlanDims {
    int32_t n_dims;
    int32_t cells[BI_MAX_DIMS];       // cells per dim: max(n_anchors - 1, 1)
    int32_t stride[BI_MAX_DIMS];      // flat cell-index stride per dim
};

__device__ __forceinline__ int bi_flat_cell(const BiPlanDims& dims, const int32_t* __restrict__ cell_p) {
    int flat = 0;
    for (int d = 0; d < dims.n_dims; ++d) {
        const int c = cell_p[d] < 0 ? 0 : cell_p[d];          // one-point axis: cell -1
        flat += c * dims.stride[d];
    }
    return flat;
}

// exclusive scan of v[0..n) in place over one CTA (n arbitrary); returns the total; `carry` is a 33-int scratch
static __device__ int bi_block_exclusive_scan(int* v, int n, int* carry) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + BI_PLAN_THREADS - 1) / BI_PLAN_THREADS;
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
        int w = carry[lane];
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const int o = __shfl_up_sync(BI_FULL_MASK, w, k);
            if (lane >= k) w += o;
        }
        carry[lane] = w;                                       // inclusive over warps
        if (lane == 31) carry[32] = w;
    }
    __syncthreads();
    int run = incl - sum + (warp ? carry[warp - 1] : 0);
    for (int i = lo; i < hi; ++i) { const int x = v[i]; v[i] = run; run += x; }
    __syncthreads();
    return carry[32];
}

