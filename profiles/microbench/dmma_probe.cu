// Microbenchmark (not product code): FP64 vector (DFMA) vs FP64 tensor (DMMA, mma.sync m8n8k4 / m16n8k8)
// throughput on sm_100a, whether the two pipes overlap, and DMMA's accumulation order.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma1684(double (&d)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}

// mode 0: DFMA only (8 chains); 1: m8n8k4 (8 chains); 2: m16n8k8 (4 chains); 3: m16n8k8 x4 + 32 DFMA interleaved;
// 4: m16n8k8, accumulator re-zeroed each time (no chain dependency on C) + 4 DMUL epilogue per MMA
template <int MODE>
__global__ void __launch_bounds__(128) k_rate(int64_t iters, double* sink) {
    const double m = 1.0000000001, c = 1e-12;
    double f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 1e-9 + i;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    double a[4] = {1e-3 * threadIdx.x, 2e-3, 3e-3, 4e-3}, b[2] = {1e-3, 2e-3};
    double prod[4] = {1.0, 1.0, 1.0, 1.0};
    int xacc = 0;
    for (int64_t it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = fma(f[i], m, c);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { dmma884(acc[i][0], acc[i][1], a[0], b[0]); dmma884(acc[i][2], acc[i][3], a[1], b[1]); }
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma1688(acc[i], a, b);
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dmma1688(acc[i], a, b);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fma(f[j], m, c);
            }
        } else if (MODE == 5 || MODE == 6) {
            // m8n8k4, fresh output each time: C = 0 (5) or a loop-invariant non-zero C (6); outputs consumed by integer XOR
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                double d0 = (MODE == 5) ? 0.0 : b[0], d1 = (MODE == 5) ? 0.0 : b[1];
                a[i & 3] += 1.0;   // keeps the operands live (1 DADD per DMMA)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(d0), "+d"(d1) : "d"(a[i & 3]), "d"(b[0]));
                xacc ^= __double2hiint(d0) ^ __double2loint(d1);
            }
        } else if (MODE == 7) {
            // m8n8k4 fresh output (C = 0), no DADD on the operands at all
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                double d0 = 0.0, d1 = 0.0;
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(d0), "+d"(d1) : "d"(a[i & 3]), "d"(b[i & 1]));
                xacc ^= __double2hiint(d0) ^ __double2loint(d1);
                a[i & 3] = __hiloint2double(__double2hiint(a[i & 3]), __double2loint(a[i & 3]) ^ (xacc & 1));
            }
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double d[4] = {0.0, 0.0, 0.0, 0.0};
                a[0] += 1e-9;
                dmma1688(d, a, b);
                prod[i] = __dmul_rn(prod[i], __dmul_rn(__dmul_rn(d[0], d[1]), __dmul_rn(d[2], d[3])));
            }
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += f[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) { r += prod[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) r += acc[i][j]; }
    if (r == 123.456 || xacc == 0x12345) sink[0] = r;
}

template <int MODE>
static void run_rate(const char* name, double fma_per_thread_iter, int ctas_per_sm, double* sink) {
    const int64_t iters = 20000;
    const int blocks = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    k_rate<MODE><<<blocks, 128>>>(iters / 10, sink);
    CHECK(cudaEventRecord(e0));
    k_rate<MODE><<<blocks, 128>>>(iters, sink);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * fma_per_thread_iter * iters * 128.0 * blocks;
    printf("%-44s ctas/sm=%d  %.3f ms  %.2f TFLOP/s\n", name, ctas_per_sm, ms, flops / ms * 1e-9);
}

// accumulation order: D = A(16x8) * B(8x8) with C = 0, compared with fma chains in k order / reversed / pairwise
__global__ void k_order(const double* A, const double* B, double* D) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    double a[4] = {A[g * 8 + t], A[(g + 8) * 8 + t], A[g * 8 + t + 4], A[(g + 8) * 8 + t + 4]};
    double b[2] = {B[t * 8 + g], B[(t + 4) * 8 + g]};
    double d[4] = {0.0, 0.0, 0.0, 0.0};
    dmma1688(d, a, b);
    D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1];
    D[(g + 8) * 8 + 2 * t] = d[2]; D[(g + 8) * 8 + 2 * t + 1] = d[3];
    // m8n8k4 twice (k 0..3 then 4..7) on rows 0..7
    double e0 = 0.0, e1 = 0.0;
    dmma884(e0, e1, A[g * 8 + t], B[t * 8 + g]);
    dmma884(e0, e1, A[g * 8 + t + 4], B[(t + 4) * 8 + g]);
    D[128 + g * 8 + 2 * t] = e0; D[128 + g * 8 + 2 * t + 1] = e1;
}

int main() {
    double* sink; CHECK(cudaMalloc(&sink, 8));
    for (int c = 1; c <= 4; c *= 2) {
        run_rate<0>("DFMA only (32/iter)", 32, c, sink);
        run_rate<1>("DMMA m8n8k4 x8 (8*256/32 fma/thr)", 8 * 256.0 / 32, c, sink);
        run_rate<2>("DMMA m16n8k8 x4 (4*1024/32)", 4 * 1024.0 / 32, c, sink);
        run_rate<3>("DMMA m16n8k8 x4 + 32 DFMA", 4 * 1024.0 / 32 + 32, c, sink);
        run_rate<4>("DMMA m16n8k8 x4 (C=0) + 4x4 DMUL [mma flops only]", 4 * 1024.0 / 32, c, sink);
        run_rate<5>("DMMA m8n8k4 x16 C=RZ fresh (+16 DADD) [mma only]", 16 * 256.0 / 32, c, sink);
        run_rate<6>("DMMA m8n8k4 x16 C=reg fresh (+16 DADD) [mma only]", 16 * 256.0 / 32, c, sink);
        run_rate<7>("DMMA m8n8k4 x16 C=RZ fresh, int-only glue", 16 * 256.0 / 32, c, sink);
    }
    // order test
    double hA[128], hB[64], hD[192];
    srand(12345);
    int n_seq = 0, n_rev = 0, n_pair = 0, n_seq4 = 0, n_tot = 0, n_884 = 0;
    for (int trial = 0; trial < 200; ++trial) {
        for (int i = 0; i < 128; ++i) hA[i] = (rand() / (double)RAND_MAX - 0.3) * pow(10.0, rand() % 6 - 3);
        for (int i = 0; i < 64; ++i) hB[i] = (rand() / (double)RAND_MAX - 0.3) * pow(10.0, rand() % 6 - 3);
        double *dA, *dB, *dD;
        CHECK(cudaMalloc(&dA, sizeof hA)); CHECK(cudaMalloc(&dB, sizeof hB)); CHECK(cudaMalloc(&dD, sizeof hD));
        CHECK(cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice));
        k_order<<<1, 32>>>(dA, dB, dD);
        CHECK(cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost));
        for (int r = 0; r < 16; ++r)
            for (int c = 0; c < 8; ++c) {
                double seq = 0, rev = 0, lo = 0, hi = 0;
                for (int k = 0; k < 8; ++k) seq = fma(hA[r * 8 + k], hB[k * 8 + c], seq);
                for (int k = 7; k >= 0; --k) rev = fma(hA[r * 8 + k], hB[k * 8 + c], rev);
                for (int k = 0; k < 4; ++k) lo = fma(hA[r * 8 + k], hB[k * 8 + c], lo);
                for (int k = 4; k < 8; ++k) hi = fma(hA[r * 8 + k], hB[k * 8 + c], hi);
                double got = hD[r * 8 + c];
                n_tot++;
                n_seq += (got == seq); n_rev += (got == rev); n_pair += (got == lo + hi);
                if (r < 8) n_884 += (hD[128 + r * 8 + c] == seq);
            }
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    printf("order: m16n8k8 == sequential-k fma chain: %d/%d; reversed: %d; (k0-3)+(k4-7): %d; 2x m8n8k4 == seq: %d/%d\n",
           n_seq, n_tot, n_rev, n_pair, n_884, n_tot / 2);
    (void)n_seq4;
    return 0;
}
