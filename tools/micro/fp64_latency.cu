// Microbenchmark: latency of a dependent FP64 chain (DFMA / DADD / DSETP+select) on sm_100a, alone and while other warps
// of the same SM keep the FP64 pipe busy with DMMA (as k_als's neighbours do) or with independent DFMA streams.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

// block = 4 + 4*nload warps: warps 0..3 (one per SM sub-partition) run the dependent chain; the rest generate load.
__global__ void k_lat(int iters, int mode, long long* out, double* sink) {
    const int warp = threadIdx.x >> 5;
    if (warp < 4) {
        double a = threadIdx.x * 1e-3 + 1.0, m = 1.0000001, c = 1e-9;
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int k = 0; k < 16; k++) a = fma(a, m, c);
        }
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[warp] = t1 - t0;
        if (a == 1.2345) sink[0] = a;
    } else if (mode == 1) {   // DMMA load
        double c[8][2];
        for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
        double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
        for (int it = 0; it < iters * 4; it++)
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        double r = 0;
        for (int i = 0; i < 8; i++) r += c[i][0] + c[i][1];
        if (r == 1.2345) sink[1] = r;
    } else if (mode == 2) {   // independent DFMA load
        double f[8];
        for (int i = 0; i < 8; i++) f[i] = threadIdx.x * 1e-3 + i;
        for (int it = 0; it < iters * 16; it++)
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] = fma(f[i], 1.0000001, 1e-9);
        double r = 0;
        for (int i = 0; i < 8; i++) r += f[i];
        if (r == 1.2345) sink[2] = r;
    }
}

int main() {
    long long* out; double* sink;
    cudaMalloc(&out, 64); cudaMalloc(&sink, 64);
    const int iters = 20000;
    const char* names[3] = {"alone", "with DMMA warps", "with DFMA warps"};
    for (int nload = 0; nload <= 3; nload++)
        for (int mode = 0; mode < 3; mode++) {
            if ((nload == 0) != (mode == 0)) continue;
            const int threads = 32 * (4 + 4 * nload);
            k_lat<<<148, threads>>>(iters, mode, out, sink);
            cudaDeviceSynchronize();
            long long h[4];
            cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
            printf("dependent DFMA chain, %-16s (%d load warps per sub-partition): %.1f cycles per DFMA   [%s]\n", names[mode], nload,
                   (double)h[0] / ((double)iters * 16), cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
