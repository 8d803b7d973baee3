// blueice_b200 -- ABI housekeeping and roofline micro-benchmarks.
#include <stdarg.h>
#include <string.h>

#include "bi_common.cuh"

static thread_local char g_error[512] = "";

void bi_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

extern "C" const char* bi_last_error(void) { return g_error; }
extern "C" int bi_abi_version(void) { return BI_ABI_VERSION; }
extern "C" int64_t bi_num_superblocks(int64_t n_events) {
    return n_events <= 0 ? 0 : (n_events + BI_SUPERBLOCK - 1) / BI_SUPERBLOCK;
}

// ---------------------------------------------------------------------------------------------
// FP64 FMA peak: 8 independent dependency chains per thread, nothing else in the loop.
// "FP64 FMA peak ... not recorded -- the builder must microbenchmark".
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_fma(int64_t iters, double* sink) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000000001, c = 1e-12;
    for (int64_t i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456) sink[0] = r;  // never true; keeps the chains alive
}

extern "C" int bi_bench_fp64_fma(int64_t fma_per_thread, int32_t n_blocks, double* sink_dev,
                                 float* ms_host, double* flops_host, void* stream) {
    BI_REQUIRE(fma_per_thread > 0 && n_blocks > 0 && sink_dev && ms_host && flops_host,
               "bi_bench_fp64_fma: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t iters = fma_per_thread / 8;
    cudaEvent_t e0, e1;
    BI_CUDA_CHECK(cudaEventCreate(&e0));
    BI_CUDA_CHECK(cudaEventCreate(&e1));
    k_fp64_fma<<<n_blocks, 256, 0, st>>>(iters / 8 + 1, sink_dev);  // warm-up
    BI_CUDA_CHECK(cudaEventRecord(e0, st));
    k_fp64_fma<<<n_blocks, 256, 0, st>>>(iters, sink_dev);
    BI_CUDA_CHECK(cudaEventRecord(e1, st));
    BI_LAUNCH_CHECK();
    BI_CUDA_CHECK(cudaEventSynchronize(e1));
    BI_CUDA_CHECK(cudaEventElapsedTime(ms_host, e0, e1));
    *flops_host = 2.0 * 8.0 * (double)iters * 256.0 * (double)n_blocks;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return BI_OK;
}

// FP64 tensor-pipe peak: 8 independent DMMA.8x8x4 accumulation chains per warp, nothing else in the loop.
// (On sm_100a DMMA and DFMA share one pipe; this is the denominator of the DMMA K2 kernel's roofline.)
__global__ void __launch_bounds__(128) k_fp64_mma(int64_t iters, double* sink) {
    double d[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = 0.0;
    const double a = 1e-3 * (threadIdx.x + 1), b = 1.0000000001;
    for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a), "d"(b));
    }
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += d[i][0] + d[i][1];
    if (r == 123.456) sink[0] = r;
}

extern "C" int bi_bench_fp64_mma(int64_t mma_per_warp, int32_t n_blocks, double* sink_dev,
                                 float* ms_host, double* flops_host, void* stream) {
    BI_REQUIRE(mma_per_warp > 0 && n_blocks > 0 && sink_dev && ms_host && flops_host,
               "bi_bench_fp64_mma: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t iters = mma_per_warp / 8;
    cudaEvent_t e0, e1;
    BI_CUDA_CHECK(cudaEventCreate(&e0));
    BI_CUDA_CHECK(cudaEventCreate(&e1));
    k_fp64_mma<<<n_blocks, 128, 0, st>>>(iters / 8 + 1, sink_dev);  // warm-up
    BI_CUDA_CHECK(cudaEventRecord(e0, st));
    k_fp64_mma<<<n_blocks, 128, 0, st>>>(iters, sink_dev);
    BI_CUDA_CHECK(cudaEventRecord(e1, st));
    BI_LAUNCH_CHECK();
    BI_CUDA_CHECK(cudaEventSynchronize(e1));
    BI_CUDA_CHECK(cudaEventElapsedTime(ms_host, e0, e1));
    *flops_host = 2.0 * 256.0 * 8.0 * (double)iters * 4.0 * (double)n_blocks;   // 8x8x4 FMAs per DMMA, 4 warps per CTA
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return BI_OK;
}

// ---------------------------------------------------------------------------------------------
// Streaming read (sum of all doubles, 16-byte loads, grid-stride): the read-only HBM ceiling the
// streaming likelihood kernel is compared with next to MEASURED_PEAKS.json's copy figure.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_stream_read(const double2* __restrict__ src, int64_t n2, double* sink) {
    double acc0 = 0, acc1 = 0;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        double2 a = __ldg(src + i), b = __ldg(src + i + stride);
        double2 c = __ldg(src + i + 2 * stride), d = __ldg(src + i + 3 * stride);
        acc0 += (a.x + b.x) + (c.x + d.x);
        acc1 += (a.y + b.y) + (c.y + d.y);
    }
    for (; i < n2; i += stride) { double2 a = __ldg(src + i); acc0 += a.x; acc1 += a.y; }
    double r = acc0 + acc1;
    if (r == 123.456) sink[0] = r;
}

extern "C" int bi_bench_stream_read(const double* src_dev, int64_t n_doubles, double* sink_dev,
                                    float* ms_host, void* stream) {
    BI_REQUIRE(src_dev && sink_dev && ms_host && n_doubles >= 2, "bi_bench_stream_read: bad arguments");
    BI_REQUIRE(((uintptr_t)src_dev & 15) == 0, "bi_bench_stream_read: src must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    BI_CUDA_CHECK(cudaEventCreate(&e0));
    BI_CUDA_CHECK(cudaEventCreate(&e1));
    int blocks = 148 * 8;
    BI_CUDA_CHECK(cudaEventRecord(e0, st));
    k_stream_read<<<blocks, 256, 0, st>>>((const double2*)src_dev, n_doubles / 2, sink_dev);
    BI_CUDA_CHECK(cudaEventRecord(e1, st));
    BI_LAUNCH_CHECK();
    BI_CUDA_CHECK(cudaEventSynchronize(e1));
    BI_CUDA_CHECK(cudaEventElapsedTime(ms_host, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return BI_OK;
}
