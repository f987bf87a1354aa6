// Shared definitions for the libmvmc.so kernels (sm_100a).
#pragma once
#include "mvmc.h"

#ifdef MVMC_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#define MVMC_LAUNCH(kernel, grid, block, smem, stream, ...)                         \
    do {                                                                            \
        kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);   \
        mvmc_count_launch();                                                        \
    } while (0)
#define MVMC_DYN_SMEM(type, name)                                   \
    extern __shared__ __align__(16) unsigned char name##_raw_[];    \
    type* name = reinterpret_cast<type*>(name##_raw_)
#endif

#define MVMC_FULL 0xffffffffu

void mvmc_count_launch();
int mvmc_set_cuda_error(cudaError_t e, const char* where);

#define MVMC_CUDA_OK(expr)                                           \
    do {                                                             \
        cudaError_t e__ = (expr);                                    \
        if (e__ != cudaSuccess) return mvmc_set_cuda_error(e__, #expr); \
    } while (0)

#define MVMC_CHECK_LAUNCH(where)                                     \
    do {                                                             \
        cudaError_t e__ = cudaGetLastError();                        \
        if (e__ != cudaSuccess) return mvmc_set_cuda_error(e__, where); \
    } while (0)

// ---- joint tables (reference: src/pose_def.py:273-298; src/inverse_kinematics.py:366-378) ----
// joints shared by a BASIC_18 3D pose and a COCO 2D pose, in BASIC_18 order
#define MVMC_N_COMMON 15
#define MVMC_N_IKJ 16

namespace mvmc {

constexpr int kCocoLShoulder = 5, kCocoRShoulder = 6, kCocoLHip = 11, kCocoRHip = 12;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MVMC_FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(MVMC_FULL, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MVMC_FULL, v, o);
    return v;
}

// Block-wide sum with a fixed reduction order (deterministic run to run). `scratch` >= 32 doubles.
// Every thread of the block must call it; the result is returned to all threads.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < nw; i++) r += scratch[i];
    return r;
}
__device__ __forceinline__ double block_max(double v, double* scratch) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double r = scratch[0];
    for (int i = 1; i < nw; i++) r = fmax(r, scratch[i]);
    return r;
}

}  // namespace mvmc
