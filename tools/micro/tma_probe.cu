// Microbenchmark: ways of staging a 16 x 96 (k x i) FP64 operand chunk (12 KB) of a row-major matrix into shared memory
// on sm_100a, as k_als needs it: 2-D / 3-D tensor-map TMA with the 128-byte swizzle, 1-D bulk copies (one per row),
// and 16-byte cp.async by all threads. Also checks the landing pattern of the swizzled boxes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu && ./tma_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(s32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* m, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     s32(dst)),
                 "l"(m), "r"(s32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma3d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     s32(dst)),
                 "l"(m), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma4d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            s32(dst)),
        "l"(m), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void bulk1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src),
                 "r"(bytes), "r"(s32(bar))
                 : "memory");
}

constexpr int NP = 304, LD = 304, NS = 4, CHUNK = 16 * 96;   // doubles per chunk

// mode 0: i-major chunk (96 rows x 16 k) by one 3-D TMA {16, 96, 1} (dims: k, row, clip), swizzle 128B
// mode 1: k-major chunk (16 k-rows x 96 i) by one 4-D TMA {16, 16, 6, 1} (dims: i%16, k-row, i/16, clip), swizzle 128B
// mode 2: i-major by 96 one-row bulk copies of 128 B;  mode 3: k-major by 16 one-row bulk copies of 768 B
// mode 4: k-major by 16-byte cp.async from all 256 threads (3 per thread)
__global__ void __launch_bounds__(256, 2) k_stage(const __grid_constant__ CUtensorMap mi, const __grid_constant__ CUtensorMap mk,
                                                  const double* base, int mode, int iters, double* sink, int* bad) {
    extern __shared__ __align__(1024) unsigned char raw[];
    double* st = reinterpret_cast<double*>(raw);
    __shared__ unsigned long long full[NS];
    const int clip = blockIdx.x;
    const double* M = base + (size_t)clip * NP * LD;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double acc = 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto issue = [&](int c) {   // chunk c -> stage c % NS; walks the matrix: tile column (c % 3) * 96, k0 = ((c / 3) % 18) * 16
        const int s = c % NS;
        double* dst = st + s * CHUNK;
        const int i0 = (c % 3) * 96, k0 = ((c / 3) % 18) * 16;
        if (mode == 0) {
            if (threadIdx.x == 0) {
                mbar_expect(&full[s], CHUNK * 8);
                tma3d(dst, &mi, k0, i0, clip, &full[s]);
            }
        } else if (mode == 1) {
            if (threadIdx.x == 0) {
                mbar_expect(&full[s], CHUNK * 8);
                tma4d(dst, &mk, 0, k0, i0 / 16, clip, &full[s]);
            }
        } else if (mode == 2) {
            if (warp == 0) {
                if (lane == 0) mbar_expect(&full[s], CHUNK * 8);
                __syncwarp();
                for (int i = lane; i < 96; i += 32) bulk1d(dst + i * 16, M + (size_t)(i0 + i) * LD + k0, 128, &full[s]);
            }
        } else if (mode == 3) {
            if (warp == 0) {
                if (lane == 0) mbar_expect(&full[s], CHUNK * 8);
                __syncwarp();
                if (lane < 16) bulk1d(dst + lane * 96, M + (size_t)(k0 + lane) * LD + i0, 768, &full[s]);
            }
        }
    };
    if (mode == 4) {
        // classic multistage cp.async ring
        auto issue4 = [&](int c) {
            const int s = c % NS;
            double* dst = st + s * CHUNK;
            const int i0 = (c % 3) * 96, k0 = ((c / 3) % 18) * 16;
            for (int e = threadIdx.x; e < 16 * 48; e += 256) {
                const int k = e / 48, i = (e % 48) * 2;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(dst + k * 96 + i)),
                             "l"(M + (size_t)(k0 + k) * LD + i0 + i)
                             : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int c = 0; c < NS - 1; c++) issue4(c);
        for (int c = 0; c < iters; c++) {
            asm volatile("cp.async.wait_group %0;" ::"n"(NS - 2) : "memory");
            __syncthreads();
            issue4(c + NS - 1);
            acc += st[(c % NS) * CHUNK + threadIdx.x];
        }
    } else {
        for (int c = 0; c < NS; c++) issue(c);
        for (int c = 0; c < iters; c++) {
            const int s = c % NS;
            mbar_wait(&full[s], (c / NS) & 1);
            acc += st[s * CHUNK + threadIdx.x];
            __syncthreads();   // everyone is done with the stage
            issue(c + NS);
        }
    }
    if (acc == 1.2345) sink[0] = acc;
    // landing-pattern check of the swizzled boxes (chunk 0 re-loaded into stage 0)
    if (mode <= 1 && blockIdx.x == 0) {
        __syncthreads();
        // drain: wait for the NS chunks still in flight
        for (int c = iters; c < iters + NS; c++) mbar_wait(&full[c % NS], (c / NS) & 1);
        __syncthreads();
        const int c = iters + NS;
        issue(c);
        mbar_wait(&full[c % NS], (c / NS) & 1);
        const int i0 = (c % 3) * 96, k0 = ((c / 3) % 18) * 16;
        const double* S = st + (c % NS) * CHUNK;
        int nbad = 0;
        for (int e = threadIdx.x; e < 16 * 96; e += 256) {
            const int k = e % 16, i = e / 16;
            const double want = M[(size_t)(mode == 0 ? i0 + i : k0 + k) * LD + (mode == 0 ? k0 + k : i0 + i)];
            double got;
            if (mode == 0) got = S[i * 16 + (((k >> 1) ^ (i & 7)) << 1) + (k & 1)];                              // row i, 16-byte chunk k/2 ^ i%8
            else got = S[(i >> 4) * 256 + k * 16 + ((((i & 15) >> 1) ^ (k & 7)) << 1) + (i & 1)];               // block i/16, row k
            if (got != want) nbad++;
        }
        atomicAdd(bad + mode, nbad);
    }
}

int main() {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q);
    printf("entry point: %s (%d) %p\n", cudaGetErrorString(e), (int)q, (void*)encode);
    if (!encode) return 1;
    const int B = 296;
    double* d;
    cudaMalloc(&d, (size_t)B * NP * LD * 8);
    std::vector<double> h((size_t)NP * LD);
    for (size_t i = 0; i < h.size(); i++) h[i] = (double)i;
    for (int b = 0; b < B; b++) cudaMemcpy(d + (size_t)b * NP * LD, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    CUtensorMap mi, mk;
    {
        cuuint64_t dims[3] = {LD, NP, (cuuint64_t)B};
        cuuint64_t strides[2] = {LD * 8, (cuuint64_t)NP * LD * 8};
        cuuint32_t box[3] = {16, 96, 1}, es[3] = {1, 1, 1};
        CUresult r = encode(&mi, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode i-major 3D: %d\n", (int)r);
    }
    {
        cuuint64_t dims[4] = {16, NP, LD / 16, (cuuint64_t)B};
        cuuint64_t strides[3] = {LD * 8, 128, (cuuint64_t)NP * LD * 8};
        cuuint32_t box[4] = {16, 16, 6, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = encode(&mk, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode k-major 4D (stride[1] = 128 B < stride[0]): %d\n", (int)r);
    }
    double* sink;
    int* bad;
    cudaMalloc(&sink, 64);
    cudaMalloc(&bad, 64);
    cudaMemset(bad, 0, 64);
    const size_t smem = NS * CHUNK * 8 + 1024;
    cudaFuncSetAttribute(k_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char* names[5] = {"TMA 3D i-major box {16,96,1} swizzle128", "TMA 4D k-major box {16,16,6,1} swizzle128",
                            "96 x 128 B bulk copies", "16 x 768 B bulk copies", "16 B cp.async x 768"};
    const int iters = 20000;
    for (int mode = 0; mode < 5; mode++) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        k_stage<<<B, 256, smem>>>(mi, mk, d, mode, 200, sink, bad);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k_stage<<<B, 256, smem>>>(mi, mk, d, mode, iters, sink, bad);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)B * iters * CHUNK * 8;
        printf("%-44s : %8.3f ms  %7.1f GB/s  %7.0f ns/chunk/CTA  err=%s\n", names[mode], ms, bytes / (ms * 1e-3) / 1e9,
               ms * 1e6 / iters, cudaGetErrorString(cudaGetLastError()));
    }
    int hb[2];
    cudaMemcpy(hb, bad, 8, cudaMemcpyDeviceToHost);
    printf("landing-pattern mismatches: i-major %d, k-major %d (of 1536 per check, 2 checks each)\n", hb[0], hb[1]);
    return 0;
}
