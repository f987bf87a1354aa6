// Microbenchmark: FP64 tensor-core (DMMA, mma.sync f64) vs FP64 CUDA-core (DFMA) throughput on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(int iters, double* sink) {
    double a[8];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], m, c);
    double r = 0;
    for (int i = 0; i < 8; i++) r += a[i];
    if (r == 1.2345) sink[0] = r;
}

__global__ void k_dmma884(int iters, double* sink) {
    double c[8][2];
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    double r = 0;
    for (int i = 0; i < 8; i++) r += c[i][0] + c[i][1];
    if (r == 1.2345) sink[0] = r;
}

__global__ void k_dmma16816(int iters, double* sink) {
    double c[4][4];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = 0.0;
    double a[8], b[4];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 4; i++) b[i] = 1.0 + threadIdx.x * 1e-6 * i;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    double r = 0;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r += c[i][j];
    if (r == 1.2345) sink[0] = r;
}

// DMMA and DFMA in the same warp: per iteration 8 DMMA (4096 flops / warp) + NF x 8 DFMA per thread (NF*8*64 flops / warp).
// If the two instruction classes run on separate pipes the time is max(t_dmma, t_dfma), otherwise the sum.
template <int NF>
__global__ void k_mixed(int iters, double* sink) {
    double c[8][2], f[8];
    for (int i = 0; i < 8; i++) { c[i][0] = c[i][1] = 0.0; f[i] = threadIdx.x * 1e-3 + i; }
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    const double m = 1.0000001, cc = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
            for (int q = 0; q < NF; q++) f[(i + q) & 7] = fma(f[(i + q) & 7], m, cc);
        }
    }
    double r = 0;
    for (int i = 0; i < 8; i++) r += c[i][0] + c[i][1] + f[i];
    if (r == 1.2345) sink[0] = r;
}

template <class F>
float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    double* sink; cudaMalloc(&sink, 64);
    const int blocks = 148 * 8, thr = 256, iters = 100000;
    float t = timeit([&] { k_dfma<<<blocks, thr>>>(iters, sink); });
    printf("DFMA      : %.2f TFLOP/s\n", (double)blocks * thr * iters * 8 * 2 / (t * 1e-3) / 1e12);
    t = timeit([&] { k_dmma884<<<blocks, thr>>>(iters, sink); });
    printf("DMMA 8x8x4: %.2f TFLOP/s\n", (double)blocks * (thr / 32) * iters * 8 * (8.0 * 8 * 4 * 2) / (t * 1e-3) / 1e12);
    t = timeit([&] { k_dmma16816<<<blocks, thr>>>(iters, sink); });
    printf("DMMA 16x8x16: %.2f TFLOP/s\n", (double)blocks * (thr / 32) * iters * 4 * (16.0 * 8 * 16 * 2) / (t * 1e-3) / 1e12);
    {
        float tm = timeit([&] { k_mixed<4><<<blocks, thr>>>(iters / 4, sink); });
        double fl = (double)blocks * (thr / 32) * (iters / 4) * (8 * 512.0 + 4 * 8 * 64.0);
        printf("mixed 8 DMMA + 32 DFMA per warp-iter : %.2f TFLOP/s total (%.3f ms)\n", fl / (tm * 1e-3) / 1e12, tm);
        tm = timeit([&] { k_mixed<8><<<blocks, thr>>>(iters / 4, sink); });
        fl = (double)blocks * (thr / 32) * (iters / 4) * (8 * 512.0 + 8 * 8 * 64.0);
        printf("mixed 8 DMMA + 64 DFMA per warp-iter : %.2f TFLOP/s total (%.3f ms)\n", fl / (tm * 1e-3) / 1e12, tm);
        tm = timeit([&] { k_mixed<2><<<blocks, thr>>>(iters / 4, sink); });
        fl = (double)blocks * (thr / 32) * (iters / 4) * (8 * 512.0 + 2 * 8 * 64.0);
        printf("mixed 8 DMMA + 16 DFMA per warp-iter : %.2f TFLOP/s total (%.3f ms)\n", fl / (tm * 1e-3) / 1e12, tm);
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
