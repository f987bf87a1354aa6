// TEST INFRASTRUCTURE — a tiny single-process CUDA execution-model emulator.
//
// The build container has nvcc but no GPU. To exercise the *real kernel sources* (csrc/*.cu) in the
// `-m "not gpu"` test tier, they can be compiled with g++ against this header (-DMVMC_EMU): every CUDA
// thread of a block becomes a ucontext fiber, __syncthreads()/warp shuffles become cooperative barriers,
// blocks run one after another. It is slow and only used by tests on tiny inputs; the product library
// (libmvmc.so) is always the nvcc build and has no path into this file.
#pragma once
#include <ucontext.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <functional>
#include <algorithm>
using std::min;
using std::max;

struct uint3_ { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) double2 { double x, y; };
struct alignas(8) int2 { int x, y; };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __constant__
#define __shared__ static
#define __align__(x) __attribute__((aligned(x)))

namespace emu {

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    int state = 0;  // 0 runnable, 1 wait-block, 2 wait-warp, 3 done, 4 wait-named-barrier
    int bar = 0;    // named barrier waited on (state 4)
    uint3_ tid;
    int linear = 0;
};

struct BlockCtx {
    std::vector<Fiber> fibers;
    ucontext_t sched;
    int cur = -1;
    uint3_ bid;
    dim3 bdim, gdim;
    std::function<void()> body;
    uint64_t warp_slot[64][32];
    uint64_t warp_slot2[64][32];
    int bar_expect[16];   // threads a named barrier releases at
};

extern BlockCtx* g_blk;
extern unsigned char* g_dyn_smem;
extern size_t g_dyn_smem_cap;
extern unsigned long long g_launches;

inline Fiber& me() { return g_blk->fibers[g_blk->cur]; }
inline void yield_to_sched() { Fiber& f = me(); swapcontext(&f.ctx, &g_blk->sched); }

inline void block_sync() { me().state = 1; yield_to_sched(); }
// bar.sync id, nthreads: released when `nthreads` fibers wait on barrier `id`
inline void named_sync(int id, int nthreads) {
    g_blk->bar_expect[id] = nthreads;
    me().bar = id;
    me().state = 4;
    yield_to_sched();
}
inline void warp_sync() { me().state = 2; yield_to_sched(); }

void fiber_entry();
void run_block(BlockCtx& b);

template <class F>
void launch(dim3 grid, dim3 block, size_t smem, F body) {
    g_launches++;
    if (smem > g_dyn_smem_cap) {
        free(g_dyn_smem);
        g_dyn_smem = (unsigned char*)aligned_alloc(128, (smem + 127) / 128 * 128);
        g_dyn_smem_cap = smem;
    }
    static BlockCtx blk;
    size_t nthreads = (size_t)block.x * block.y * block.z;
    if (blk.fibers.size() < nthreads) {
        size_t old = blk.fibers.size();
        blk.fibers.resize(nthreads);
        for (size_t i = old; i < nthreads; i++) blk.fibers[i].stack = (char*)malloc(1 << 18);
    }
    blk.bdim = block;
    blk.gdim = grid;
    blk.body = body;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                blk.bid = {bx, by, bz};
                run_block(blk);
            }
}

template <class T>
inline uint64_t to_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T>
inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

template <class T>
inline T shfl_idx(T v, int src) {
    Fiber& f = me();
    int w = f.linear / 32, l = f.linear % 32;
    g_blk->warp_slot[w][l] = to_bits(v);
    warp_sync();
    T r = from_bits<T>(g_blk->warp_slot[w][src & 31]);
    warp_sync();
    return r;
}
// mma.sync.aligned.m8n8k4.row.col.f64: lane holds a = A[lane/4][lane%4], b = B[lane%4][lane/4],
// c0/c1 = C[lane/4][2*(lane%4) + {0,1}]
inline void dmma_884(double& c0, double& c1, double a, double b) {
    Fiber& f = me();
    const int w = f.linear / 32, l = f.linear % 32;
    g_blk->warp_slot[w][l] = to_bits(a);
    g_blk->warp_slot2[w][l] = to_bits(b);
    warp_sync();
    const int row = l / 4, col = 2 * (l % 4);
    for (int k = 0; k < 4; k++) {
        const double av = from_bits<double>(g_blk->warp_slot[w][row * 4 + k]);
        c0 = fma(av, from_bits<double>(g_blk->warp_slot2[w][col * 4 + k]), c0);
        c1 = fma(av, from_bits<double>(g_blk->warp_slot2[w][(col + 1) * 4 + k]), c1);
    }
    warp_sync();
}
}  // namespace emu

#define threadIdx (emu::me().tid)
#define blockIdx (emu::g_blk->bid)
#define blockDim (emu::g_blk->bdim)
#define gridDim (emu::g_blk->gdim)

inline void __syncthreads() { emu::block_sync(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_sync(); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl_idx(v, src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::shfl_idx(v, (emu::me().linear % 32) ^ m); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int d) {
    int l = emu::me().linear % 32;
    return emu::shfl_idx(v, l + d < 32 ? l + d : l);
}
template <class T> inline T __shfl_up_sync(unsigned, T v, int d) {
    int l = emu::me().linear % 32;
    return emu::shfl_idx(v, l - d >= 0 ? l - d : l);
}
inline unsigned __ballot_sync(unsigned, int pred) {
    emu::Fiber& f = emu::me();
    int w = f.linear / 32, l = f.linear % 32;
    emu::g_blk->warp_slot[w][l] = pred ? 1 : 0;
    emu::warp_sync();
    unsigned r = 0;
    int nthreads = emu::g_blk->bdim.x * emu::g_blk->bdim.y * emu::g_blk->bdim.z;
    for (int i = 0; i < 32; i++)
        if (w * 32 + i < nthreads && emu::g_blk->fibers[w * 32 + i].state != 3 && emu::g_blk->warp_slot[w][i]) r |= 1u << i;
    emu::warp_sync();
    return r;
}
inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, !p) == 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline void sincos(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
inline double __longlong_as_double(long long v) { return emu::from_bits<double>((uint64_t)v); }
inline long long __double_as_longlong(double v) { return (long long)emu::to_bits(v); }
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
inline void __threadfence() {}
inline void __threadfence_block() {}
template <class T> inline T __ldg(const T* p) { return *p; }
inline double fma_rn_(double a, double b, double c) { return fma(a, b, c); }

inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaPeekAtLastError() { return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
template <class K> inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return 0; }
template <class T> inline cudaError_t cudaMemcpyToSymbolAsync(T& sym, const void* src, size_t n, size_t off, cudaMemcpyKind, cudaStream_t = 0) {
    memcpy((char*)&sym + off, src, n);
    return 0;
}
template <class T> inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* src, size_t n, size_t off = 0, cudaMemcpyKind = cudaMemcpyHostToDevice) {
    memcpy((char*)&sym + off, src, n);
    return 0;
}

#define MVMC_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define MVMC_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::g_dyn_smem)

#ifdef MVMC_EMU_IMPL
namespace emu {
BlockCtx* g_blk = nullptr;
unsigned char* g_dyn_smem = nullptr;
size_t g_dyn_smem_cap = 0;
unsigned long long g_launches = 0;

void fiber_entry() {
    g_blk->body();
    me().state = 3;
    yield_to_sched();
}

void run_block(BlockCtx& b) {
    g_blk = &b;
    int n = b.bdim.x * b.bdim.y * b.bdim.z;
    for (int i = 0; i < n; i++) {
        Fiber& f = b.fibers[i];
        f.state = 0;
        f.linear = i;
        f.tid = {(unsigned)(i % b.bdim.x), (unsigned)((i / b.bdim.x) % b.bdim.y), (unsigned)(i / (b.bdim.x * b.bdim.y))};
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = 1 << 18;
        f.ctx.uc_link = &b.sched;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
    int done = 0;
    while (done < n) {
        bool ran = false;
        for (int i = 0; i < n; i++) {
            if (b.fibers[i].state == 0) {
                b.cur = i;
                swapcontext(&b.sched, &b.fibers[i].ctx);
                ran = true;
                if (b.fibers[i].state == 3) done++;
            }
        }
        // release warp barriers
        bool released = false;
        int nw = (n + 31) / 32;
        for (int w = 0; w < nw; w++) {
            int waiting = 0, alive = 0;
            for (int l = 0; l < 32 && w * 32 + l < n; l++) {
                int s = b.fibers[w * 32 + l].state;
                if (s != 3) alive++;
                if (s == 2) waiting++;
            }
            if (alive > 0 && waiting == alive) {
                for (int l = 0; l < 32 && w * 32 + l < n; l++)
                    if (b.fibers[w * 32 + l].state == 2) b.fibers[w * 32 + l].state = 0;
                released = true;
            }
        }
        for (int id = 0; id < 16; id++) {
            int w4 = 0;
            for (int i = 0; i < n; i++)
                if (b.fibers[i].state == 4 && b.fibers[i].bar == id) w4++;
            if (w4 > 0 && w4 >= b.bar_expect[id]) {
                for (int i = 0; i < n; i++)
                    if (b.fibers[i].state == 4 && b.fibers[i].bar == id) b.fibers[i].state = 0;
                released = true;
            }
        }
        int wb = 0, alive = 0;
        for (int i = 0; i < n; i++) {
            if (b.fibers[i].state != 3) alive++;
            if (b.fibers[i].state == 1) wb++;
        }
        if (alive > 0 && wb == alive) {
            for (int i = 0; i < n; i++)
                if (b.fibers[i].state == 1) b.fibers[i].state = 0;
            released = true;
        }
        if (!ran && !released && done < n) {
            fprintf(stderr, "cuda_emu: deadlock (divergent barrier) in block (%u,%u,%u)\n", b.bid.x, b.bid.y, b.bid.z);
            abort();
        }
    }
}
}  // namespace emu
#endif
