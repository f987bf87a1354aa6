// ALS/ADMM low-rank multi-way matcher (the reference's live association solver).
//
// Reference row: A5 mv_association.py:222-318 (match_als). Per iteration
//     Xt = Z - (Y - W + beta)/mu;  B = (inv(A^T A + a/mu I) A^T Xt)^T;  A = (inv(B^T B + a/mu I) B^T Xt^T)^T;
//     X = A B^T;  Z = clip(X + Y/mu) with same-group blocks zeroed and unit diagonal;  Y += mu (X - Z)
// until ||X - Z||_F / n < tol and mu ||X - X_prev||_F / n < tol, mu doubled / halved on a 10x imbalance.
//
// k_als: ONE CTA (8 warps) PER CLIP-FRAME, 2 resident CTAs per SM. The n x n iterates (W, Z, Y, X, Xt) and the
// n x r factors live in a per-clip global workspace (zero padded to multiples of 16 so that no copy needs a
// tail case); the seven products of an iteration (three of them 2 r n^2 flops: the dominant cost of the whole
// capture path) run as tiled FP64 tensor-core GEMMs:
//   * CTA tile 64 x 96, k-chunks of 16. Operand chunks are staged by the bulk-copy engine (cp.async.bulk, one
//     instruction per operand row, completion counted on an mbarrier) into a 4-stage ring of shared memory. One warp
//     per chunk (rotating) issues the refill of the stage it has just finished with after waiting on the stage's
//     "empty" mbarrier; the ring runs ahead across tiles, so only the first chunks of a product see memory latency and
//     there is no block-wide barrier inside a product;
//   * each warp owns a 32 x 24 piece of the tile (12 accumulator fragments) and issues mma.sync.m8n8k4.f64 (DMMA) from
//     conflict-free fragment loads (row strides chosen so the 4 x 8 fragment footprint covers every bank once);
//   * the Z / Y / X / next-Xt update and both residual norms are the epilogue of the X = A B^T product (its X_prev, Y, W
//     loads run two fragments ahead of the arithmetic), so an iteration makes three passes over n x n data;
//   * mu is a power of two throughout (64 doubled / halved), so the reference's divisions by mu are exact
//     multiplications by 1/mu here - the same bits, ~40 instructions fewer per element;
//   * Xt for the next iteration is written by that epilogue assuming mu stays (it changes a handful of times per
//     solve; then one element-wise pass rebuilds Xt from Z, Y, W with the new mu) - same arithmetic, same values.
// FP64 DMMA and DFMA share one pipe on B200 (37 TFLOP/s measured alone or mixed, tools/micro/dmma_probe.cu); DMMA is
// used because it needs 8x fewer issue slots and 4x less shared-memory bandwidth per flop.
#include "mvmc_common.cuh"
#ifndef MVMC_EMU
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)
#endif

// This file is compiled more than once with different tile shapes (-DAL_FM_=3 -DAL_THREADS_=128 -DAL_VARIANT=small for
// the small problems of 8 x 16 scenes): each build lives in its own namespace and exports its launcher under its own name;
// the default build also holds the shape dispatcher and the shared C entry points.
#ifdef AL_VARIANT
#define AL_CAT2(a, b) a##b
#define AL_CAT(a, b) AL_CAT2(a, b)
#define AL_NSNAME AL_CAT(als_, AL_VARIANT)
#define AL_LAUNCHER AL_CAT(mvmc_match_als_ordered_, AL_VARIANT)
#else
#define AL_NSNAME als_default
#define AL_LAUNCHER mvmc_match_als_ordered_default
#endif
namespace mvmc {
namespace AL_NSNAME {

#ifndef AL_THREADS_
#define AL_THREADS_ 256
#endif
#ifndef AL_FM_
#define AL_FM_ 4
#endif
constexpr int AL_THREADS = AL_THREADS_;   // 8 warps: 2 along M x 4 along N, warp tile 32 x 24 (4 warps: 2 x 2, four CTAs per SM)
constexpr int AL_WARPS = AL_THREADS / 32;
constexpr int AL_KC = 16;                 // k-chunk (one 128-byte swizzle row of doubles)
constexpr int AL_FM = AL_FM_;             // 8-row fragments per warp along M (4: warp tile 32 x 24; 3: 24 x 24)
constexpr int AL_TM = 16 * AL_FM;         // CTA tile rows (two warps along M)
constexpr int AL_TN = 24 * (AL_WARPS / 2); // CTA tile columns (warp columns x 24)
constexpr int AL_STAGE_M = AL_TM * AL_KC; // doubles: 8 KB
constexpr int AL_STAGE_N = AL_TN * AL_KC; // 12 KB
constexpr int AL_STAGE = AL_STAGE_M + AL_STAGE_N;
constexpr int AL_NS = 4;                  // stages of the operand ring
constexpr int AL_PAD = 16;                // every matrix dimension of the workspace is padded (with zeros) to this

// ---- tensor maps, mbarriers, TMA loads (emulated synchronously under the CPU emulator) ----
// Two views of a row-major FP64 matrix [rows][ld] of every clip (clip stride = workspace per clip), both with the
// 128-byte shared-memory swizzle (16-byte unit u of 128-byte row r lands at unit u ^ (r % 8)):
//   i-major  (rank 3: k = column, i = row, clip):            box {16, TI, 1}        -> smem [TI][16]
//   k-major  (rank 4: i % 16, k = row, i / 16, clip):         box {16, 16, TI/16, 1} -> smem [TI/16][16][16]
// (the second one walks the matrix in 16-column blocks: its dimension-2 stride, 128 B, is smaller than its dimension-1
// stride, the row pitch - the TMA unit does not mind; tools/micro/tma_probe.cu checks the landing pattern on the GPU).
#ifdef MVMC_EMU
#define AL_EMU_SYNC() __syncthreads()
#define AL_GRID_CONSTANT
typedef uint64_t mbar_t;
struct TMap {
    const double* base;
    int rank;
    long long dim[4], stride[4];   // stride in doubles (stride[0] = 1)
    int box[4];
};
__device__ __forceinline__ void mbar_init(mbar_t*, int) {}
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void mbar_expect_tx(mbar_t*, unsigned) {}
__device__ __forceinline__ void mbar_arrive(mbar_t*) {}
__device__ __forceinline__ void mbar_wait(mbar_t*, unsigned) {}
__device__ __forceinline__ void fence_proxy_async() {}
// box copy with the 128-byte swizzle and zero fill outside the tensor, as the TMA unit does it
inline void emu_tma(double* dst, const TMap* m, const int* crd) {
    const int b0 = m->box[0], b1 = m->box[1], b2 = m->rank > 3 ? m->box[2] : 1;
    for (int z = 0; z < b2; z++)
        for (int y = 0; y < b1; y++)
            for (int x = 0; x < b0; x++) {
                long long c[4] = {crd[0] + x, crd[1] + y, 0, 0};
                if (m->rank > 3) {
                    c[2] = crd[2] + z;
                    c[3] = crd[3];
                } else {
                    c[2] = crd[2];
                }
                bool in = true;
                long long off = 0;
                for (int d = 0; d < m->rank; d++) {
                    in = in && c[d] >= 0 && c[d] < m->dim[d];
                    off += c[d] * m->stride[d];
                }
                const int row = z * b1 + y;                       // 128-byte rows of the box, in landing order
                const int unit = (x >> 1) ^ (row & 7);
                dst[row * 16 + unit * 2 + (x & 1)] = in ? m->base[off] : 0.0;
            }
}
__device__ __forceinline__ void bulk_g2s(double* dst, const double* src, unsigned bytes, mbar_t*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void tma_load3(double* dst, const TMap* m, int c0, int c1, int c2, mbar_t*) {
    const int crd[4] = {c0, c1, c2, 0};
    emu_tma(dst, m, crd);
}
__device__ __forceinline__ void tma_load4(double* dst, const TMap* m, int c0, int c1, int c2, int c3, mbar_t*) {
    const int crd[4] = {c0, c1, c2, c3};
    emu_tma(dst, m, crd);
}
#else
#define AL_EMU_SYNC() ((void)0)
#define AL_GRID_CONSTANT __grid_constant__
typedef unsigned long long mbar_t;
typedef CUtensorMap TMap;
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(mbar_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(mbar_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(mbar_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(mbar_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "AL_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra AL_DONE;\n"
        "bra AL_WAIT;\n"
        "AL_DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load3(double* dst, const TMap* m, int c0, int c1, int c2, mbar_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load4(double* dst, const TMap* m, int c0, int c1, int c2, int c3, mbar_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// contiguous global -> shared bulk copy (16-byte aligned, size a multiple of 16), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(double* dst, const double* src, unsigned bytes, mbar_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// orders this thread's earlier generic-proxy writes before later async-proxy (TMA) reads of the same data
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
#endif

// the tensor maps of one launch (kernel parameter) and the bytes one box of each brings
enum { MAP_A_KM = 0, MAP_A_KN, MAP_A_IM, MAP_B_KM, MAP_B_KN, MAP_B_IN, MAP_XT_KN, MAP_XT_IN, MAP_T_KN, MAP_G_IM, MAP_COUNT };
struct AlsMaps {
    TMap m[MAP_COUNT];
    unsigned bytes[MAP_COUNT];
};

// D(8x8) += A(8x4) * B(4x8): lane holds a = A[lane/4][lane%4], b = B[lane%4][lane/4], c0/c1 = C[lane/4][2*(lane%4) + {0,1}]
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
#ifdef MVMC_EMU
    emu::dmma_884(c0, c1, a, b);
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
#endif
}

// layout of an operand (template argument), seen as element(k, i) with k the contraction index:
// k-major: element(k, i) = p[k*ld + i];  i-major: element(k, i) = p[i*ld + k]. The storage behind an operand is zero for
// K <= k < round16(K); whatever lies at I <= i only reaches outputs that are never stored.
constexpr bool KMAJ = true, IMAJ = false;
struct Operand {
    const TMap* map;
    unsigned bytes;
    int I;
};

// The operand ring: AL_NS stages, each [M-operand chunk | N-operand chunk]; full[s] counts the bytes landed in stage s,
// empty[s] the warps (8) that are done reading it. `gc` = chunks consumed so far by this CTA (all products), from which
// every thread derives stage and phase parity of a chunk without any shared state.
struct Ring {
    double* stages;
    mbar_t* full;
    mbar_t* empty;
    unsigned gc;
    int clip;
};

// The four DMMA steps of a 16-chunk take the k-sets {0,1,4,5}, {2,3,6,7}, {8,9,12,13}, {10,11,14,15} (lane%4 -> the
// set's element): any partition works as long as both operands use it, and with this one the 32 lanes of a fragment load
// hit every 8-byte bank pair exactly twice in BOTH swizzled layouts (the minimum for 256 bytes).
//
// Fragment loads are `ld.shared.f64 [register + immediate]`: everything that depends on the lane is folded into a handful
// of per-thread byte offsets (FragTab), re-based once per chunk on the stage's shared-memory address; the fragment index
// and the step only contribute compile-time immediates. (Indexing the stage through a generic pointer cost an IMAD, an
// IMAD.WIDE and a generic LD per fragment, issued right in front of the DMMA that needs it.)
//
// With q = lane % 4, p = lane / 4, kq = (q & 1) + 4 (q >> 1), step s, kb = 8 (s >> 1) + 2 (s & 1), k = kb + kq (disjoint bits):
//   k-major stage [TI/16][16 k][16 i] (bytes): element (k, i), i = I0 + p, I0 a multiple of 8, h = (I0 >> 3) & 1:
//       (I0 >> 4) 2048 + kb 128 + KT[h][s & 1],   KT[h][par] = kq 128 + (p & 1) 8 + ((((4 h + (p >> 1)) ^ kq) << 4) ^ (par << 5))
//   i-major stage [TI][16 k] (bytes):
//       I0 128 + IT[s],                            IT[s] = p 128 + (kq & 1) 8 + ((((kq >> 1) ^ p) ^ (kb >> 1)) << 4)
#ifdef MVMC_EMU
typedef const unsigned char* saddr_t;
__device__ __forceinline__ saddr_t saddr_of(const double* p) { return reinterpret_cast<saddr_t>(p); }
__device__ __forceinline__ double lds64(saddr_t a) { return *reinterpret_cast<const double*>(a); }
__device__ __forceinline__ double2 lds128(saddr_t a) { return *reinterpret_cast<const double2*>(a); }
__device__ __forceinline__ void sts128(saddr_t a, double2 v) { *reinterpret_cast<double2*>(const_cast<unsigned char*>(a)) = v; }
#else
typedef unsigned saddr_t;
__device__ __forceinline__ saddr_t saddr_of(const double* p) { return smem_u32(p); }
__device__ __forceinline__ double lds64(saddr_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds128(saddr_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(saddr_t a, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
#endif
struct FragTab {
    int kt[2][2];   // k-major lane offsets by (h, step parity)
    int it[4];      // i-major lane offsets by step
    __device__ __forceinline__ void init(int lane) {
        const int q = lane & 3, p = lane >> 2, kq = (q & 1) + 4 * (q >> 1);
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int par = 0; par < 2; par++)
                kt[h][par] = kq * 128 + (p & 1) * 8 + ((((4 * h + (p >> 1)) ^ kq) << 4) ^ (par << 5));
#pragma unroll
        for (int s4 = 0; s4 < 4; s4++) {
            const int kb = 8 * (s4 >> 1) + 2 * (s4 & 1);
            it[s4] = p * 128 + (kq & 1) * 8 + ((((kq >> 1) ^ p) ^ (kb >> 1)) << 4);
        }
    }
};
// An operand's fragment addressing for one warp: `nreg` lane/warp dependent byte offsets (re-based per chunk) and, per
// (fragment f, step s), which of them to use plus a compile-time immediate. I0 of fragment f = w0 + 8 f.
template <bool kmajor, int NF>
struct FragAddr {
    // k-major: reg[f][par] (the 16-column block and the half h of fragment f are folded in: they depend on the warp for
    // the N operand); i-major: reg[0][s] shared by all fragments, fragment f adds the immediate f * 1024.
    int reg[kmajor ? NF : 1][kmajor ? 2 : 4];
    __device__ __forceinline__ void init(const FragTab& t, int w0 /*first column of the warp inside the tile*/) {
        if (kmajor) {
#pragma unroll
            for (int f = 0; f < NF; f++) {
                const int I0 = w0 + 8 * f;
#pragma unroll
                for (int par = 0; par < 2; par++) reg[f][par] = (I0 >> 4) * 2048 + (((I0 >> 3) & 1) ? t.kt[1][par] : t.kt[0][par]);
            }
        } else {
#pragma unroll
            for (int s4 = 0; s4 < 4; s4++) reg[0][s4] = w0 * 128 + t.it[s4];
        }
    }
    // shared-memory address of fragment f, step s4, in the stage operand region starting at `base`
    __device__ __forceinline__ saddr_t at(saddr_t base, int f, int s4) const {
        if (kmajor) return base + reg[f][s4 & 1] + (8 * (s4 >> 1) + 2 * (s4 & 1)) * 128;
        return base + reg[0][s4] + f * 1024;
    }
};

// Shape of one product as the ring sees it, and a cursor over its chunks (tile-major, N tiles inside M tiles, k-chunks
// inside a tile) that every thread advances in step with the chunk loop: the position of the chunk to issue next costs
// three compares instead of two integer divisions.
struct Tiling {
    int nk, tn, total;   // k-chunks per tile, tiles along N, chunks of the whole product
};
struct ChunkCursor {
    int c, kc, tn, m0, n0;   // chunk index; k-chunk and N tile inside the current tile row; tile origin
    __device__ __forceinline__ void reset() { c = kc = tn = m0 = n0 = 0; }
    __device__ __forceinline__ void next(const Tiling& tl) {
        c++;
        if (++kc == tl.nk) {
            kc = 0;
            n0 += AL_TN;
            if (++tn == tl.tn) {
                tn = 0;
                n0 = 0;
                m0 += AL_TM;
            }
        }
    }
};

// Loads the chunk under the cursor of both operands into its stage. Called by ONE converged warp.
template <bool MK, bool NK>
__device__ __forceinline__ void ring_issue(const Ring& rg, const Operand& Mop, const Operand& Nop, const ChunkCursor& cu) {
    const int k0 = cu.kc * AL_KC;
    const unsigned g = rg.gc + (unsigned)cu.c;
    const int st = g % AL_NS;
    const unsigned use = g / AL_NS;
    if (use > 0) mbar_wait(&rg.empty[st], (use - 1) & 1);     // all 8 warps have finished with the stage's previous chunk
    if ((threadIdx.x & 31) == 0) {
        double* Sm = rg.stages + st * AL_STAGE;
        double* Sn = Sm + AL_STAGE_M;
        mbar_expect_tx(&rg.full[st], Mop.bytes + Nop.bytes);
        if (MK) tma_load4(Sm, Mop.map, 0, k0, cu.m0 >> 4, rg.clip, &rg.full[st]);
        else tma_load3(Sm, Mop.map, k0, cu.m0, rg.clip, &rg.full[st]);
        if (NK) tma_load4(Sn, Nop.map, 0, k0, cu.n0 >> 4, rg.clip, &rg.full[st]);
        else tma_load3(Sn, Nop.map, k0, cu.n0, rg.clip, &rg.full[st]);
    }
    __syncwarp();
}

// C[m][n] = sum_k Mop(k, m) * Nop(k, n) for m < Mop.I, n < Nop.I. The epilogue object receives the result fragment by
// fragment: ep.apply(buf, m, n, v0, v1) gets C[m][n], C[m][n+1] (n even; the callee guards n+1 < N) after
// ep.load(buf, m, n) was called for the same fragment two fragments earlier (its global loads overlap the arithmetic
// of the fragments in between; the first two are issued before the k-loop of the tile).
// All threads of the CTA must call it. Everything the operands point to must have been written before the call by
// this CTA (generic stores) - the fence + barrier at the top publish it to the TMA unit.
template <bool MK, bool NK, class EP>
__device__ __forceinline__ void cta_gemm(Ring& rg, const Operand Mop, const Operand Nop, int K, EP& ep) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp & 1, wn = warp >> 1;               // warp tile: rows 32*wm.., columns 24*wn..
    const int M = Mop.I, N = Nop.I;
    Tiling tl;
    tl.nk = (K + AL_KC - 1) / AL_KC;
    tl.tn = (N + AL_TN - 1) / AL_TN;
    tl.total = ((M + AL_TM - 1) / AL_TM) * tl.tn * tl.nk;
    fence_proxy_async();
    __syncthreads();
    ChunkCursor cu;   // the next chunk to issue (uniform across the CTA)
    cu.reset();
    {
        const int pre = min(AL_NS, tl.total);
        for (int c = 0; c < pre; c++) {
            if (warp == 0) ring_issue<MK, NK>(rg, Mop, Nop, cu);
            cu.next(tl);
        }
    }
    FragTab ftab;
    ftab.init(lane);
    FragAddr<MK, AL_FM> fm;   // M operand: rows 8 AL_FM wm + 8 a
    FragAddr<NK, 3> fn;       // N operand: columns 24 wn + 8 b
    fm.init(ftab, 8 * AL_FM * wm);
    fn.init(ftab, 24 * wn);
    const saddr_t ring0 = saddr_of(rg.stages);
    int c = 0;
    for (int m0 = 0; m0 < M; m0 += AL_TM) {
        const int mw0 = m0 + 8 * AL_FM * wm;
        const int mt = mw0 >= M ? 0 : min(AL_FM, (M - mw0 + 7) >> 3);       // live row tiles of this warp
        for (int n0 = 0; n0 < N; n0 += AL_TN) {
            const int nw0 = n0 + 24 * wn;
            const int nt = (nw0 >= N || mt == 0) ? 0 : min(3, (N - nw0 + 7) >> 3);
            double acc[AL_FM][3][2];
#pragma unroll
            for (int a = 0; a < AL_FM; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
            // fragment s = 3a + b of this thread: row, first column, liveness
            auto f_m = [&](int s) { return mw0 + (s / 3) * 8 + (lane >> 2); };
            auto f_n = [&](int s) { return nw0 + (s % 3) * 8 + 2 * (lane & 3); };
            auto f_on = [&](int s) { return (s / 3) < mt && (s % 3) < nt && f_m(s) < M && f_n(s) < N; };
            typename EP::Buf buf[3];
            if (f_on(0)) ep.load(buf[0], f_m(0), f_n(0));
            if (f_on(1)) ep.load(buf[1], f_m(1), f_n(1));
            for (int kc = 0; kc < tl.nk; kc++, c++) {
                const unsigned g = rg.gc + (unsigned)c;
                const int st = g % AL_NS;
                mbar_wait(&rg.full[st], (g / AL_NS) & 1);
                AL_EMU_SYNC();
                if (nt > 0) {
                    // every fragment of the warp tile, live or not (dead ones multiply zero padding or stale finite data into
                    // accumulators that are never stored): no branch and no predicate inside the DMMA sequence. The loads of
                    // step s+1 are issued before the DMMAs of step s.
                    const saddr_t Sm = ring0 + st * (AL_STAGE * 8);
                    const saddr_t Sn = Sm + AL_STAGE_M * 8;
                    double af[2][AL_FM], bf[2][3];
#pragma unroll
                    for (int b = 0; b < 3; b++) bf[0][b] = lds64(fn.at(Sn, b, 0));
#pragma unroll
                    for (int a = 0; a < AL_FM; a++) af[0][a] = lds64(fm.at(Sm, a, 0));
#pragma unroll
                    for (int s = 0; s < 4; s++) {
                        if (s + 1 < 4) {
#pragma unroll
                            for (int b = 0; b < 3; b++) bf[(s + 1) & 1][b] = lds64(fn.at(Sn, b, s + 1));
#pragma unroll
                            for (int a = 0; a < AL_FM; a++) af[(s + 1) & 1][a] = lds64(fm.at(Sm, a, s + 1));
                        }
#pragma unroll
                        for (int a = 0; a < AL_FM; a++)
#pragma unroll
                            for (int b = 0; b < 3; b++) dmma(acc[a][b][0], acc[a][b][1], af[s & 1][a], bf[s & 1][b]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&rg.empty[st]);
                AL_EMU_SYNC();
                if (cu.c < tl.total) {
                    if (warp == (c & (AL_WARPS - 1))) ring_issue<MK, NK>(rg, Mop, Nop, cu);
                    cu.next(tl);
                }
            }
#pragma unroll
            for (int s = 0; s < 3 * AL_FM; s++) {
                if (s + 2 < 3 * AL_FM && f_on(s + 2)) ep.load(buf[(s + 2) % 3], f_m(s + 2), f_n(s + 2));
                if (f_on(s)) ep.apply(buf[s % 3], f_m(s), f_n(s), acc[s / 3][s % 3][0], acc[s / 3][s % 3][1]);
            }
        }
    }
    rg.gc += (unsigned)tl.total;
}

// epilogue without operands of its own: F(m, n, v0, v1)
template <class F>
struct StoreEp {
    struct Buf {};
    F f;
    __device__ __forceinline__ void load(Buf&, int, int) const {}
    __device__ __forceinline__ void apply(const Buf&, int m, int n, double v0, double v1) { f(m, n, v0, v1); }
};
template <class F>
__device__ __forceinline__ StoreEp<F> store_ep(F f) { return StoreEp<F>{f}; }

constexpr int GJ_LD = 96;

// Fallback for r > GJB_MAXR: in place in global memory, thread t owns columns t % 32 + 32q of rows t / 32 + nw p.
__device__ void invert_spd(double* G, int r, int ldg, double* aux) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double* rk = aux;        // pivot row * (1 / pivot)
    double* ck = aux + 128;  // pivot column
    for (int k = 0; k < r; k++) {
        const double p = 1.0 / G[k * ldg + k];
        __syncthreads();
        for (int j = threadIdx.x; j < r; j += blockDim.x) {
            rk[j] = G[k * ldg + j] * p;
            ck[j] = G[j * ldg + k];
        }
        __syncthreads();
        for (int i = w; i < r; i += nw) {
            const double f = ck[i];
            double* row = G + i * ldg;
            if (i == k) {
                for (int j = lane; j < r; j += 32) row[j] = (j == k) ? p : rk[j];
            } else {
                for (int j = lane; j < r; j += 32) row[j] = (j == k) ? -f * p : row[j] - f * rk[j];
            }
        }
        __syncthreads();
    }
}

// optional per-phase cycle counters (thread 0 of every CTA; enabled by mvmc_als_phase_profile(1)): where an iteration's time goes
enum { PH_G1 = 0, PH_INV1, PH_T1, PH_B, PH_G2, PH_INV2, PH_T2, PH_A, PH_X, PH_RED, PH_MU, PH_INIT, PH_ADMM, PH_ADMM_WAIT, PH_ADMM_FENCE, PH_GJ_LOAD, PH_GJ_INV8, PH_GJ_PANEL, PH_GJ_UPD, PH_GJ_STORE, PH_COUNT };
__device__ unsigned long long g_als_phase[PH_COUNT];
__device__ int g_als_phase_on = 0;
struct PhaseClock {
    long long t;
    bool on;
    __device__ __forceinline__ void start() {
#ifndef MVMC_EMU
        on = g_als_phase_on != 0 && threadIdx.x == 0;
        if (on) t = clock64();
#endif
    }
    __device__ __forceinline__ void lap(int ph) {
#ifndef MVMC_EMU
        if (on) {
            const long long now = clock64();
            atomicAdd(&g_als_phase[ph], (unsigned long long)(now - t));
            t = now;
        }
#endif
    }
};

// ---- blocked Gauss-Jordan inverse on the tensor cores ----
// The scalar elimination above is a chain of r dependent pivots with a block barrier each (about 1 us per pivot once the
// FP64 pipe is shared with another CTA's DMMAs: 19 % of an ADMM iteration at r = 64). Here the matrix sits in shared
// memory (the idle operand ring) and is eliminated 8 columns at a time:
//     Pinv = inv(G[kb,kb])                        every warp, redundantly, in registers: 8 shuffle pivots, no barrier
//     R    = Pinv G[kb,:]                         8 x 8 blocks, two DMMA each, spread over the warps     | barrier
//     G[i,j] -= G[i,kb] R[j],  G[i,kb] = -G[i,kb] Pinv,  G[kb,:] = R, G[kb,kb] = Pinv   (row block i = warp) | barrier
// 2 barriers per 8 pivots instead of 8, and the rank-8 updates run as DMMA. Same in-place Gauss-Jordan recurrences, block
// wise; the matrix is padded with an identity to a multiple of 8.
constexpr int GJB_MAXR = AL_NS * AL_STAGE >= 96 * 92 ? 88 : 64;     // (88 x 92 + 8 x 92) doubles = 70.6 KB of the 80 KB ring; 64 x 68 + 8 x 68 = 39.2 KB
static_assert((GJB_MAXR + 8) * (GJB_MAXR + 4) <= AL_NS * AL_STAGE, "the blocked inverse works in the operand ring");

// 8 x 8 inverse in registers: lane l holds P[l/4][2(l%4)], P[l/4][2(l%4)+1] (the DMMA accumulator layout).
// Division-free Gauss-Jordan: instead of scaling the pivot row by 1/d, every other row is multiplied by the pivot,
// row_i <- (d row_i - f_i row_k) 2^-e with e the exponent of d (an exact scaling that keeps the rows O(1): without it the
// magnitudes square at every step), so a pivot step is three dependent FP64 operations - a reciprocal is ~10, and every one
// of them queues behind the neighbouring CTA's DMMAs - and the rows are normalised by ONE reciprocal each at the end.
// g = what the identity diagonal of a row not yet pivoted has grown to, t = this row's left diagonal. Same accuracy as the
// scaled elimination (checked against LAPACK over 14 decades of scale and condition numbers up to 1e8).
__device__ __forceinline__ void inv8_regs(double& e0, double& e1, int lane) {
    const int ro = lane >> 2, q = lane & 3;
    double g = 1.0, t = 1.0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int ks = k & 1, kl = k >> 1;
        const double d = __shfl_sync(MVMC_FULL, ks ? e1 : e0, 4 * k + kl);
        const double rk0 = __shfl_sync(MVMC_FULL, e0, 4 * k + q), rk1 = __shfl_sync(MVMC_FULL, e1, 4 * k + q);
        const double f = __shfl_sync(MVMC_FULL, ks ? e1 : e0, 4 * ro + kl);
        // 2^-e for d = m 2^e, 1 <= m < 2 (d > 0: the matrix is positive definite)
        const double sc = __longlong_as_double((0x7FELL - ((__double_as_longlong(d) >> 52) & 0x7FFLL)) << 52);
        const double ds = d * sc, fs = f * sc;
        if (ro != k) {
            e0 = ds * e0 - fs * rk0;
            e1 = ds * e1 - fs * rk1;
        }
        if (q == kl) {   // column k now holds column k of the (scaled) inverse part
            const double v = (ro == k) ? g : -fs * g;
            if (ks) e1 = v;
            else e0 = v;
        }
        t = (ro == k) ? d : (ro < k ? t * ds : t);
        g *= ds;
    }
    const double p = 1.0 / t;
    e0 *= p;
    e1 *= p;
}

__device__ void invert_spd_blocked(double* Gg, int r, int ldr, double* sm) {
    PhaseClock gc;
    gc.start();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ro = lane >> 2, q = lane & 3;
    const int nb = (r + 7) >> 3, np = 8 * nb, ld = np + 4;   // pitch = 4 (mod 8): every fragment pattern is conflict free
    double* Gs = sm;               // [np][ld]
    double* R = sm + np * ld;      // [8][ld]
    const saddr_t Gs_s = saddr_of(Gs), R_s = saddr_of(R);
    {
        // thread (w, lane) brings elements (w + 8a, lane + 32b): all the global loads first (they are independent; in a
        // load-store loop each one waited for the previous store), then the stores
        double v[(GJB_MAXR + AL_WARPS - 1) / AL_WARPS][(GJB_MAXR + 31) / 32];
#pragma unroll
        for (int a = 0; a < (GJB_MAXR + AL_WARPS - 1) / AL_WARPS; a++)
#pragma unroll
            for (int b = 0; b < (GJB_MAXR + 31) / 32; b++) {
                const int i = w + AL_WARPS * a, j = lane + 32 * b;
                v[a][b] = (i < r && j < r) ? Gg[(size_t)i * ldr + j] : (i == j ? 1.0 : 0.0);
            }
#pragma unroll
        for (int a = 0; a < (GJB_MAXR + AL_WARPS - 1) / AL_WARPS; a++)
#pragma unroll
            for (int b = 0; b < (GJB_MAXR + 31) / 32; b++) {
                const int i = w + AL_WARPS * a, j = lane + 32 * b;
                if (i < np && j < np) Gs[i * ld + j] = v[a][b];
            }
    }
    __syncthreads();
    gc.lap(PH_GJ_LOAD);
    for (int kb = 0; kb < nb; kb++) {
        // Pinv, in the accumulator layout
        double2 pv = lds128(Gs_s + (unsigned)(((8 * kb + ro) * ld + 8 * kb + 2 * q) * 8));
        inv8_regs(pv.x, pv.y, lane);
        gc.lap(PH_GJ_INV8);
        // Pinv as A fragments (row ro, k = 4h + q) and as B fragments (k = 4h + q, column ro)
        double pa[2], pb[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int sa = 4 * ro + 2 * h + (q >> 1);
            const double a0 = __shfl_sync(MVMC_FULL, pv.x, sa), a1 = __shfl_sync(MVMC_FULL, pv.y, sa);
            pa[h] = (q & 1) ? a1 : a0;
            const int sb = 4 * (4 * h + q) + (ro >> 1);
            const double b0 = __shfl_sync(MVMC_FULL, pv.x, sb), b1 = __shfl_sync(MVMC_FULL, pv.y, sb);
            pb[h] = (ro & 1) ? b1 : b0;
        }
        // R[j] = Pinv G[kb, j]   (shared-space addresses, 64 bytes per 8-column block)
        for (int j = w; j < nb; j += AL_WARPS) {
            if (j == kb) continue;
            double c0 = 0.0, c1 = 0.0;
            const saddr_t gb = Gs_s + (unsigned)(((8 * kb + q) * ld + 8 * j + ro) * 8);
#pragma unroll
            for (int h = 0; h < 2; h++) dmma(c0, c1, pa[h], lds64(gb + (unsigned)(h * 4 * ld * 8)));
            double2 o;
            o.x = c0;
            o.y = c1;
            sts128(R_s + (unsigned)((ro * ld + 8 * j + 2 * q) * 8), o);
        }
        __syncthreads();
        gc.lap(PH_GJ_PANEL);
        for (int i = w; i < nb; i += AL_WARPS) {
            if (i == kb) {
                for (int j = 0; j < nb; j++) {
                    const double2 v = (j == kb) ? pv : lds128(R_s + (unsigned)((ro * ld + 8 * j + 2 * q) * 8));
                    sts128(Gs_s + (unsigned)(((8 * kb + ro) * ld + 8 * j + 2 * q) * 8), v);
                }
            } else {
                double nf[2];   // -G[i,kb] as A fragments
#pragma unroll
                for (int h = 0; h < 2; h++) nf[h] = -lds64(Gs_s + (unsigned)(((8 * i + ro) * ld + 8 * kb + 4 * h + q) * 8));
                __syncwarp();   // the block G[i,kb] is overwritten below by lanes that hold other elements of it
                saddr_t cp = Gs_s + (unsigned)(((8 * i + ro) * ld + 2 * q) * 8);       // my two elements of block (i, 0)
                saddr_t rp = R_s + (unsigned)((q * ld + ro) * 8);                       // R fragment element of block 0, h = 0
                const unsigned rh = (unsigned)(4 * ld * 8);
                for (int j = 0; j < nb; j++, cp += 64, rp += 64) {
                    double c0 = 0.0, c1 = 0.0;
                    double b0 = pb[0], b1 = pb[1];
                    if (j != kb) {
                        const double2 c = lds128(cp);
                        c0 = c.x;
                        c1 = c.y;
                        b0 = lds64(rp);
                        b1 = lds64(rp + rh);
                    }
                    dmma(c0, c1, nf[0], b0);
                    dmma(c0, c1, nf[1], b1);
                    double2 o;
                    o.x = c0;
                    o.y = c1;
                    sts128(cp, o);
                }
            }
        }
        __syncthreads();
        gc.lap(PH_GJ_UPD);
    }
#pragma unroll
    for (int a = 0; a < (GJB_MAXR + AL_WARPS - 1) / AL_WARPS; a++)
#pragma unroll
        for (int b = 0; b < (GJB_MAXR + 31) / 32; b++) {
            const int i = w + AL_WARPS * a, j = lane + 32 * b;
            if (i < r && j < r) Gg[(size_t)i * ldr + j] = Gs[i * ld + j];
        }
    gc.lap(PH_GJ_STORE);
}

// Gg (r x r, ld ldr, global) <- inverse of Gg; `ring` = the (idle) operand ring, used as scratch
__device__ __noinline__ void invert_normal_matrix(double* Gg, int r, int ldr, double* aux, double* ring) {
    __syncthreads();
    if (r <= GJB_MAXR) invert_spd_blocked(Gg, r, ldr, ring);
    else invert_spd(Gg, r, ldr, aux);
    __syncthreads();
}

// Per-clip workspace. Every dimension is padded to a multiple of 16 and the padding of everything a product reads
// (Xt, A, B, T, G) is zero for the whole solve: chunks of 16 along any contraction index need no tail case.
struct AlsLayout {
    int N, NP, ldn, ldr;
    size_t zero_span;   // doubles from Xt to the end (the part zeroed at the start of every solve)
    size_t per;
    __host__ __device__ AlsLayout(int N_, int rmax) {
        N = N_;
        NP = (N_ + AL_PAD - 1) / AL_PAD * AL_PAD;
        ldn = NP;
        ldr = (rmax + AL_PAD - 1) / AL_PAD * AL_PAD;
        zero_span = (size_t)NP * ldn + (size_t)2 * NP * ldr + (size_t)ldr * ldn + (size_t)ldr * ldr;
        per = (size_t)5 * NP * ldn + zero_span;
    }
};

// The ADMM element-wise update (mv_association.py:286-296) as a streaming pass over row strips: X = A B^T only stores X
// (two X buffers alternate between "new" and "previous"); this pass then pulls strips of X_new, X_prev, Y, W through the
// (idle) operand ring with contiguous bulk copies - up to 4 strips (~78 KB) in flight per CTA, none of it held in
// registers - and writes the new Y and the next Xt with coalesced 16-byte stores (Y alternates between two buffers like
// X; Z itself is never stored: the rare mu change recomputes it from X and the previous Y, bit for bit). As an epilogue of the product the same
// update was bound by the few loads a thread can keep in flight next to its 48 accumulator registers.
#ifndef AL_PASS_NS_
#define AL_PASS_NS_ 8
#endif
constexpr int AL_PASS_NS = AL_PASS_NS_;    // most strips in flight (what fits the ring decides)
constexpr int AL_PASS_U = 1;     // pairs per thread per round (2 measured slower: a round then spans half the ring)
struct AdmmPass {
    mbar_t* bar;        // [AL_PASS_NS] bytes landed in a stage
    unsigned* cnt;      // [AL_PASS_NS] warps that have left a stage (running count)
    int rs, ns;         // rows per strip, stages (fixed per launch: functions of ldn)
    int st0, par0;      // stage and phase parity of the next pass's first strip
};

__device__ __forceinline__ void admm_pass(AdmmPass& ps, double* ring, const double* Xn, const double* Xo, const double* Yo, double* Yn,
                                          const double* W, double* Xt, const int* grp, int n, int ldn, double mu, double inv_mu,
                                          double beta, double& pacc, double& dacc, PhaseClock& pc) {
    const int rs = ps.rs, ns = ps.ns;   // (stage indices advance by compare-and-wrap: no division in the loop)
    const saddr_t ring_s = saddr_of(ring);
    const int stage_doubles = 4 * rs * ldn;
    const int total = (n + rs - 1) / rs;          // strips
    const int hpn = (n + 1) >> 1;                 // live column pairs of a row
    const int spp = rs * hpn;                     // pairs of a (full) strip
    const int P = n * hpn;                        // pairs of the whole pass
    fence_proxy_async();   // X_new (and the Y of the previous pass) were written with generic stores
    __syncthreads();
    pc.lap(PH_ADMM_FENCE);
    // Rows are visited last to first (the rows X = A B^T wrote most recently are the ones most likely still in L2):
    // "logical" row q is row n-1-q; strip s = logical rows [s rs, (s+1) rs) = one contiguous block of rows.
    auto issue = [&](int s, int st) {   // strip s into stage st
        const int hi = min(n, (s + 1) * rs), nr = hi - s * rs, r0 = n - hi;
        const unsigned bytes = (unsigned)(nr * ldn) * 8u;
        double* dst = ring + st * stage_doubles;
        const size_t o = (size_t)r0 * ldn;
        mbar_expect_tx(&ps.bar[st], 4u * bytes);
        bulk_g2s(dst, Xn + o, bytes, &ps.bar[st]);
        bulk_g2s(dst + rs * ldn, Xo + o, bytes, &ps.bar[st]);
        bulk_g2s(dst + 2 * rs * ldn, Yo + o, bytes, &ps.bar[st]);
        bulk_g2s(dst + 3 * rs * ldn, W + o, bytes, &ps.bar[st]);
    };
    if (threadIdx.x == 0) {
        const int pre = min(ns, total);
        int st = ps.st0;
        for (int s = 0; s < pre; s++) {
            issue(s, st);
            st = st + 1 == ns ? 0 : st + 1;
        }
    }
    // The pairs of the pass are one flat index space walked AL_PASS_U x 256 at a time, independent of the strip boundaries
    // (a strip of two 262-column rows is 262 pairs: strip-synchronous rounds would run half empty). A round waits for the
    // strips it touches; the last warp out of a strip refills its stage (no block barrier: the warps drift apart by up to
    // the depth of the ring). Every index advances incrementally and every per-round quantity is a compare or an add: the
    // loop skeleton (integer divisions by run-time values, in an earlier version) cost more than the arithmetic.
    int waited = 0, wait_at = 0, wst = ps.st0, wpar = ps.par0;   // strips waited for; first pair, stage, parity of strip `waited`
    int done = 0, done_at = min(P, spp), dst = ps.st0;           // strips this warp has left; end (pair index) and stage of strip `done`
    int q = 0, c = (int)threadIdx.x;    // this thread's first pair of the round: logical row q, pair c of the row
    int sq = 0, stq = ps.st0, sq_end = rs;   // strip of row q, its stage, its first row beyond
    while (c >= hpn) {
        c -= hpn;
        q++;
    }
    for (int e0 = 0; e0 < P;) {
        const int e1 = min(min(P, e0 + AL_PASS_U * AL_THREADS), (done + ns) * spp);   // never past the strips in flight
        while (waited < total && wait_at < e1) {
            mbar_wait(&ps.bar[wst], (unsigned)wpar);
            waited++;
            wait_at += spp;
            if (++wst == ns) {
                wst = 0;
                wpar ^= 1;
            }
        }
        AL_EMU_SYNC();
        while (q >= sq_end) {
            sq++;
            sq_end += rs;
            stq = stq + 1 == ns ? 0 : stq + 1;
        }
        if (e0 + (int)threadIdx.x < e1) {
            // one pair per thread and round; shared-memory loads through 32-bit shared addresses and one 32-bit byte
            // offset for the two stores (the generic-pointer version spent a third of the round on 64-bit address arithmetic)
            const int j = 2 * c, i = n - 1 - q;
            const int hi = min(n, sq_end);
            const saddr_t sx = ring_s + (unsigned)((stq * stage_doubles + (i - (n - hi)) * ldn + j) * 8);
            const unsigned mat = (unsigned)(rs * ldn * 8);
            const double2 x = lds128(sx), x0 = lds128(sx + mat), y = lds128(sx + 2 * mat), w = lds128(sx + 3 * mat);
            const int gi = grp[i];
            const int2 gj = *reinterpret_cast<const int2*>(grp + j);   // (grp[n] = -1: an odd last column is in no group)
            const bool live1 = j + 1 < n;     // the odd column of the last pair may be padding
            double2 yn, xt;
            {
                const double dd = x.x - x0.x;
                double zz = x.x + y.x * inv_mu;                 // y / mu (mu is a power of two)
                if (gi == gj.x) zz = 0.0;
                if (i == j) zz = 1.0;
                zz = zz < 0.0 ? 0.0 : zz;
                zz = zz > 1.0 ? 1.0 : zz;
                const double pd = x.x - zz;
                yn.x = y.x + mu * pd;
                xt.x = zz - (yn.x - w.x + beta) * inv_mu;       // next iteration's Xt if mu stays
                dacc += dd * dd;
                pacc += pd * pd;
            }
            {
                const double dd = x.y - x0.y;
                double zz = x.y + y.y * inv_mu;
                if (gi == gj.y) zz = 0.0;
                if (i == j + 1) zz = 1.0;
                zz = zz < 0.0 ? 0.0 : zz;
                zz = zz > 1.0 ? 1.0 : zz;
                const double pd = x.y - zz;
                yn.y = y.y + mu * pd;
                xt.y = live1 ? zz - (yn.y - w.y + beta) * inv_mu : 0.0;   // the padding column of Xt stays zero
                if (live1) {
                    dacc += dd * dd;
                    pacc += pd * pd;
                }
            }
            const unsigned ob = (unsigned)(i * ldn + j) * 8u;
            *reinterpret_cast<double2*>(reinterpret_cast<char*>(Yn) + ob) = yn;
            *reinterpret_cast<double2*>(reinterpret_cast<char*>(Xt) + ob) = xt;
        }
        // every thread moves on by the size of the round
        c += e1 - e0;
        while (c >= hpn) {
            c -= hpn;
            q++;
        }
        // strips this warp has finished with: the last of the 8 warps out of a strip refills its stage
        __syncwarp();
        while (done < total && done_at <= e1) {
            if ((threadIdx.x & 31) == 0) {
                __threadfence_block();
                const unsigned was = atomicAdd(&ps.cnt[dst], 1u);
                if ((was & (AL_WARPS - 1)) == AL_WARPS - 1 && done + ns < total) issue(done + ns, dst);   // refill the stage just left
            }
            done++;
            done_at = min(P, done_at + spp);
            dst = dst + 1 == ns ? 0 : dst + 1;
        }
        AL_EMU_SYNC();
        e0 = e1;
    }
    {   // where the next pass starts in the ring (one division per pass)
        const int adv = ps.st0 + total;
        ps.par0 ^= (adv / ns) & 1;
        ps.st0 = adv % ns;
    }
    __syncthreads();   // (the ring is handed back to the products)
}

#ifndef AL_CTAS_
#define AL_CTAS_ (512 / AL_THREADS_)
#endif
#if defined(AL_MAXREG_) && !defined(MVMC_EMU)
// (a register cap below 65536 / (CTAs x threads): leaves room on the SM for a co-resident IK solver warp)
__global__ void __maxnreg__(AL_MAXREG_)
#else
__global__ void __launch_bounds__(AL_THREADS, AL_CTAS_)
#endif
    k_als(const AL_GRID_CONSTANT AlsMaps maps, const double* __restrict__ sim, const int* __restrict__ dim_groups, int n_groups,
          const int* __restrict__ f32_first_iter, const double* __restrict__ rand_stream, const int* __restrict__ order,
          int N, int rmax, double* __restrict__ ws, uint32_t* __restrict__ xbin, int* __restrict__ n_iter_out, double alpha,
          double beta, double tol, int max_iter) {
    MVMC_DYN_SMEM(unsigned char, smem_raw);
    // the swizzle pattern is a function of the shared-memory address: stages start on a 1024-byte boundary
#ifdef MVMC_EMU
    double* smem = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
#else
    // (an offset added to the __shared__ array, not integer arithmetic on a generic pointer: the compiler then knows that
    //  everything derived from it is shared memory - 32-bit addresses, LDS / STS instead of generic loads and stores)
    double* smem = reinterpret_cast<double*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
#endif
    const int b = order ? order[blockIdx.x] : blockIdx.x;
    const int* dg = dim_groups + b * (n_groups + 1);
    const int n = dg[n_groups];
    const int NW = (N + 31) / 32;
    uint32_t* xb = xbin + (size_t)b * N * NW;
    if (n <= 0) {
        if (threadIdx.x == 0) n_iter_out[b] = 0;
        return;
    }
    int maxsz = 0;
    for (int g = 0; g < n_groups; g++) maxsz = max(maxsz, dg[g + 1] - dg[g]);
    int r = min(n, 2 * maxsz);
    r = min(r, rmax);

    Ring rg;
    rg.stages = smem;                                             // [AL_NS * AL_STAGE]
    double* scratch = smem + AL_NS * AL_STAGE;                    // [32]
    double* inv_aux = scratch + 32;                               // [4 * 96] pivot row / column of the Gauss-Jordan inverse
    rg.full = reinterpret_cast<mbar_t*>(inv_aux + 4 * GJ_LD);     // [AL_NS]
    rg.empty = rg.full + AL_NS;                                   // [AL_NS]
    rg.gc = 0;
    rg.clip = b;
    AdmmPass ps;
    ps.bar = rg.empty + AL_NS;                                    // [AL_PASS_NS]
    ps.st0 = 0;
    ps.par0 = 0;
    {
        const int ring_doubles = AL_NS * AL_STAGE, ldn_ = (N + AL_PAD - 1) / AL_PAD * AL_PAD;
#ifndef AL_PASS_TARGET_
#define AL_PASS_TARGET_ 4
#endif
        ps.rs = max(1, ring_doubles / (AL_PASS_TARGET_ * 4 * ldn_));   // strips of about a quarter of the ring (4 matrices per strip)
        ps.ns = max(1, min(AL_PASS_NS, ring_doubles / (4 * ps.rs * ldn_)));   // stages that fit
    }
    ps.cnt = reinterpret_cast<unsigned*>(ps.bar + AL_PASS_NS);    // [AL_PASS_NS]
    int* s_grp = reinterpret_cast<int*>(ps.cnt + AL_PASS_NS + (AL_PASS_NS & 1));   // [N + 1], 8-byte aligned
    if (threadIdx.x == 0) {
        for (int s = 0; s < AL_NS; s++) {
            mbar_init(&rg.full[s], 1);
            mbar_init(&rg.empty[s], AL_WARPS);
        }
        for (int s = 0; s < AL_PASS_NS; s++) {
            mbar_init(&ps.bar[s], 1);
            ps.cnt[s] = 0;
        }
        mbar_fence_init();
    }

    const AlsLayout L(N, rmax);
    const int ldn = L.ldn, ldr = L.ldr;
    double* W = ws + (size_t)b * L.per;
    double* Yn = W + (size_t)L.NP * ldn;       // Y being formed        } swap every iteration
    double* Y = Yn + (size_t)L.NP * ldn;       // Y of the last update  }
    double* Xm = Y + (size_t)L.NP * ldn;       // X of the previous iteration  } the two buffers swap roles
    double* Xn = Xm + (size_t)L.NP * ldn;      // X being formed               } every iteration
    double* Xt = Xn + (size_t)L.NP * ldn;      // [NP][ldn]   (from here on: zeroed below, padding stays zero)
    double* A = Xt + (size_t)L.NP * ldn;       // [NP][ldr]
    double* Bm = A + (size_t)L.NP * ldr;       // [NP][ldr]
    double* Tm = Bm + (size_t)L.NP * ldr;      // [ldr][ldn]
    double* Gg = Tm + (size_t)ldr * ldn;       // [ldr][ldr]

    PhaseClock pc;
    pc.start();
    const double* S = sim + (size_t)b * N * N;
    const bool f32 = f32_first_iter != nullptr && f32_first_iter[b] != 0;

    for (int i = threadIdx.x; i <= N; i += blockDim.x) {
        int g = 0;
        for (int q = 0; q < n_groups; q++)
            if (i >= dg[q] && i < dg[q + 1]) g = q;  // an index belongs to the group whose [start, end) contains it
        s_grp[i] = i < n ? g : -1;
    }
    {
        double2 zz;
        zz.x = zz.y = 0.0;
        double2* z2 = reinterpret_cast<double2*>(Xt);
        for (size_t e = threadIdx.x; e < L.zero_span / 2; e += blockDim.x) z2[e] = zz;
    }
    __syncthreads();
    double mu = 64.0, inv_mu = 1.0 / 64.0;
    for (int i = threadIdx.x >> 5; i < n; i += AL_WARPS)
        for (int j = threadIdx.x & 31; j < n; j += 32) {
            const size_t o = (size_t)i * ldn + j;
            double w;
            if (f32) w = (double)(0.5f * ((float)S[(size_t)i * N + j] + (float)S[(size_t)j * N + i]));
            else w = 0.5 * (S[(size_t)i * N + j] + S[(size_t)j * N + i]);
            W[o] = w;
            Xm[o] = w;
            Y[o] = 0.0;
            // first Xt (Z = W, Y = 0); the float32 no-track path of the reference keeps float32 through this expression
            if (f32) Xt[o] = (double)((float)w - ((0.0f - (float)w) + (float)beta) / (float)mu);
            else Xt[o] = w - (0.0 - w + beta) / mu;
        }
    for (int e = threadIdx.x; e < n * r; e += blockDim.x) A[(size_t)(e / r) * ldr + (e % r)] = rand_stream[e];
    __syncthreads();

    auto op = [&](int which, int I) { return Operand{&maps.m[which], maps.bytes[which], I}; };
    pc.lap(PH_INIT);
    int it = 0;
    for (it = 0; it < max_iter; it++) {
        const double reg = alpha / mu;
        // ---- G = A^T A + reg I, inverted ----
        {
            auto ep = store_ep([&](int m, int nn, double v0, double v1) {
                Gg[(size_t)m * ldr + nn] = (m == nn) ? v0 + reg * 1.0 : v0 + reg * 0.0;
                if (nn + 1 < r) Gg[(size_t)m * ldr + nn + 1] = (m == nn + 1) ? v1 + reg * 1.0 : v1 + reg * 0.0;
            });
            cta_gemm<KMAJ, KMAJ>(rg, op(MAP_A_KM, r), op(MAP_A_KN, r), n, ep);
        }
        pc.lap(PH_G1);
        invert_normal_matrix(Gg, r, ldr, inv_aux, rg.stages);
        pc.lap(PH_INV1);
        // ---- T = A^T Xt ----
        {
            auto ep = store_ep([&](int m, int nn, double v0, double v1) {
                Tm[(size_t)m * ldn + nn] = v0;
                if (nn + 1 < n) Tm[(size_t)m * ldn + nn + 1] = v1;
            });
            cta_gemm<KMAJ, KMAJ>(rg, op(MAP_A_KM, r), op(MAP_XT_KN, n), n, ep);
        }
        pc.lap(PH_T1);
        // ---- B = (Ginv T)^T ----
        {
            auto ep = store_ep([&](int m, int nn, double v0, double v1) {
                Bm[(size_t)nn * ldr + m] = v0;
                if (nn + 1 < n) Bm[(size_t)(nn + 1) * ldr + m] = v1;
            });
            cta_gemm<IMAJ, KMAJ>(rg, op(MAP_G_IM, r), op(MAP_T_KN, n), r, ep);
        }
        pc.lap(PH_B);
        // ---- H = B^T B + reg I, inverted ----
        {
            auto ep = store_ep([&](int m, int nn, double v0, double v1) {
                Gg[(size_t)m * ldr + nn] = (m == nn) ? v0 + reg : v0;
                if (nn + 1 < r) Gg[(size_t)m * ldr + nn + 1] = (m == nn + 1) ? v1 + reg : v1;
            });
            cta_gemm<KMAJ, KMAJ>(rg, op(MAP_B_KM, r), op(MAP_B_KN, r), n, ep);
        }
        pc.lap(PH_G2);
        invert_normal_matrix(Gg, r, ldr, inv_aux, rg.stages);
        pc.lap(PH_INV2);
        // ---- T = B^T Xt^T : T[m][i] = sum_j B[j][m] Xt[i][j] ----
        {
            auto ep = store_ep([&](int m, int nn, double v0, double v1) {
                Tm[(size_t)m * ldn + nn] = v0;
                if (nn + 1 < n) Tm[(size_t)m * ldn + nn + 1] = v1;
            });
            cta_gemm<KMAJ, IMAJ>(rg, op(MAP_B_KM, r), op(MAP_XT_IN, n), n, ep);
        }
        pc.lap(PH_T2);
        // ---- A = (Hinv T)^T ----
        {
            auto ep = store_ep([&](int m, int nn, double v0, double v1) {
                A[(size_t)nn * ldr + m] = v0;
                if (nn + 1 < n) A[(size_t)(nn + 1) * ldr + m] = v1;
            });
            cta_gemm<IMAJ, KMAJ>(rg, op(MAP_G_IM, r), op(MAP_T_KN, n), r, ep);
        }
        pc.lap(PH_A);
        // ---- X = A B^T ----
        {
            auto ep = store_ep([&](int i, int j, double v0, double v1) {
                double2 xv;
                xv.x = v0;
                xv.y = v1;
                *reinterpret_cast<double2*>(Xn + (size_t)i * ldn + j) = xv;   // (an odd last column lands in the padding)
            });
            cta_gemm<IMAJ, IMAJ>(rg, op(MAP_A_IM, n), op(MAP_B_IN, n), r, ep);
        }
        pc.lap(PH_X);
        // ---- Z / Y / next-Xt updates and both residual norms ----
        double pacc = 0.0, dacc = 0.0;
        admm_pass(ps, rg.stages, Xn, Xm, Y, Yn, W, Xt, s_grp, n, ldn, mu, inv_mu, beta, pacc, dacc, pc);
        {
            double* t = Xm;
            Xm = Xn;
            Xn = t;
            t = Y;
            Y = Yn;
            Yn = t;
        }
        pc.lap(PH_ADMM);
        const double psum = block_sum(pacc, scratch);
        const double dsum = block_sum(dacc, scratch);
        const double p_res = sqrt(psum) / n;
        const double d_res = mu * sqrt(dsum) / n;
        pc.lap(PH_RED);
        if (p_res < tol && d_res < tol) {
            it++;
            break;
        }
        double mu_new = mu;
        if (p_res > 10.0 * d_res) mu_new = 2.0 * mu;
        else if (d_res > 10.0 * p_res) mu_new = mu / 2.0;
        if (mu_new != mu) {
            // Xt = Z - (Y - W + beta) / mu_new, with Z recomputed exactly as the pass formed it (same operands: the new X, the
            // Y before the update - now in Yn after the swap - and the old mu)
            const double inv_old = inv_mu;
            mu = mu_new;
            inv_mu = 1.0 / mu;
            __syncthreads();
            for (int i = threadIdx.x >> 5; i < n; i += AL_WARPS) {
                const int gi = s_grp[i];
                for (int j = threadIdx.x & 31; j < n; j += 32) {
                    const size_t o = (size_t)i * ldn + j;
                    double zz = Xm[o] + Yn[o] * inv_old;
                    if (gi == s_grp[j]) zz = 0.0;
                    if (i == j) zz = 1.0;
                    if (zz < 0.0) zz = 0.0;
                    if (zz > 1.0) zz = 1.0;
                    Xt[o] = zz - (Y[o] - W[o] + beta) * inv_mu;
                }
            }
            pc.lap(PH_MU);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) n_iter_out[b] = it;
    // X_bin = 0.5 (X + X^T) > 0.5, as row bitmasks
    for (int e = threadIdx.x; e < n * NW; e += blockDim.x) {
        const int i = e / NW, w = e % NW;
        uint32_t bits = 0;
        for (int q = 0; q < 32; q++) {
            const int j = w * 32 + q;
            if (j < n) {
                const double v = 0.5 * (Xm[(size_t)i * ldn + j] + Xm[(size_t)j * ldn + i]);
                if (v > 0.5) bits |= 1u << q;
            }
        }
        xb[(size_t)i * NW + w] = bits;
    }
}

// Longest-processing-time-first launch order: clips sorted by the iteration count of their previous frame (a good
// predictor of this frame's), descending, so that the long solves start first and the tail of the launch is short ones.
// One CTA; counting sort over 0..1023 iterations. The order only affects scheduling, never a result.
__global__ void __launch_bounds__(1024) k_als_order(const int* __restrict__ prev_iter, int B, int* __restrict__ order) {
    __shared__ int cnt[1024];
    const int t = threadIdx.x;
    cnt[t] = 0;
    __syncthreads();
    for (int b = t; b < B; b += 1024) atomicAdd(&cnt[1023 - min(max(prev_iter[b], 0), 1023)], 1);
    __syncthreads();
    // exclusive prefix sum of cnt (bucket 0 = the longest solves), 1024 entries, Hillis-Steele in place
    int v = cnt[t];
    for (int o = 1; o < 1024; o <<= 1) {
        const int add = t >= o ? cnt[t - o] : 0;
        __syncthreads();
        cnt[t] += add;
        __syncthreads();
    }
    const int excl = cnt[t] - v;
    __syncthreads();
    cnt[t] = excl;
    __syncthreads();
    for (int b = t; b < B; b += 1024) {
        const int pos = atomicAdd(&cnt[1023 - min(max(prev_iter[b], 0), 1023)], 1);
        order[pos] = b;
    }
}

}  // namespace AL_NSNAME
}  // namespace mvmc

using namespace mvmc;
using namespace mvmc::AL_NSNAME;

#ifndef AL_VARIANT
int mvmc_als_order(const int* prev_iter, int B, int* order, void* stream) {
    if (!prev_iter || !order || B <= 0) return MVMC_ERR_INVALID;
    MVMC_LAUNCH(k_als_order, dim3(1), dim3(1024), 0, stream, prev_iter, B, order);
    MVMC_CHECK_LAUNCH("k_als_order");
    return MVMC_OK;
}
#endif

static size_t als_smem_bytes(int N) {
    return 1024 + (size_t)(AL_NS * AL_STAGE + 32 + 4 * GJ_LD) * sizeof(double) + (2 * AL_NS + 2 * AL_PASS_NS) * sizeof(mbar_t) +
           (size_t)(N + 2) * sizeof(int);
}

#ifndef AL_VARIANT
extern "C" size_t mvmc_match_als_workspace_bytes(int B, int N, int rmax) {
    const AlsLayout L(N, rmax);
    return (size_t)B * L.per * sizeof(double);
}
#endif

// ---- tensor maps over the workspace ----
#ifndef MVMC_EMU
typedef CUresult (*mvmc_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static mvmc_encode_fn als_encoder() {
    static mvmc_encode_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (mvmc_encode_fn)p;
    }
    return fn;
}
#endif

// view of the matrix at `base` ([rows][ld] doubles per clip, clip stride `per` doubles); kmajor: box {16, 16, ti/16, 1}, else
// box {16, ti, 1}. Box extents are clamped to the tensor so that no box is larger than what it is cut from.
static int als_make_map(TMap* out, unsigned* bytes, double* base, int rows, int ld, size_t per, int B, bool kmajor, int ti) {
#ifdef MVMC_EMU
    TMap m;
    m.base = base;
    if (kmajor) {
        m.rank = 4;
        m.dim[0] = 16; m.dim[1] = rows; m.dim[2] = ld / 16; m.dim[3] = B;
        m.stride[0] = 1; m.stride[1] = ld; m.stride[2] = 16; m.stride[3] = (long long)per;
        m.box[0] = 16; m.box[1] = 16; m.box[2] = min(ti / 16, ld / 16); m.box[3] = 1;
        *bytes = 16u * 16u * (unsigned)m.box[2] * 8u;
    } else {
        m.rank = 3;
        m.dim[0] = ld; m.dim[1] = rows; m.dim[2] = B; m.dim[3] = 1;
        m.stride[0] = 1; m.stride[1] = ld; m.stride[2] = (long long)per; m.stride[3] = 0;
        m.box[0] = 16; m.box[1] = min(ti, rows); m.box[2] = 1; m.box[3] = 1;
        *bytes = 16u * (unsigned)m.box[1] * 8u;
    }
    *out = m;
    return MVMC_OK;
#else
    mvmc_encode_fn enc = als_encoder();
    if (!enc) return mvmc_set_cuda_error(cudaErrorNotSupported, "cuTensorMapEncodeTiled entry point");
    CUresult r;
    if (kmajor) {
        const cuuint64_t dims[4] = {16, (cuuint64_t)rows, (cuuint64_t)(ld / 16), (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)ld * 8, 128, (cuuint64_t)per * 8};
        const cuuint32_t blocks = (cuuint32_t)((ti / 16 < ld / 16) ? ti / 16 : ld / 16);
        const cuuint32_t box[4] = {16, 16, blocks, 1}, es[4] = {1, 1, 1, 1};
        r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        *bytes = 16u * 16u * blocks * 8u;
    } else {
        const cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)rows, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)per * 8};
        const cuuint32_t brows = (cuuint32_t)(ti < rows ? ti : rows);
        const cuuint32_t box[3] = {16, brows, 1}, es[3] = {1, 1, 1};
        r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        *bytes = 16u * brows * 8u;
    }
    if (r != CUDA_SUCCESS) return mvmc_set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled");
    return MVMC_OK;
#endif
}

static int als_build_maps(AlsMaps* M, double* ws, int B, int N, int rmax) {
    const AlsLayout L(N, rmax);
    double* Xt = ws + (size_t)5 * L.NP * L.ldn;
    double* A = Xt + (size_t)L.NP * L.ldn;
    double* Bm = A + (size_t)L.NP * L.ldr;
    double* Tm = Bm + (size_t)L.NP * L.ldr;
    double* Gg = Tm + (size_t)L.ldr * L.ldn;
    struct Spec {
        int which;
        double* base;
        int rows, ld;
        bool kmajor;
        int ti;
    } spec[MAP_COUNT] = {
        {MAP_A_KM, A, L.NP, L.ldr, true, AL_TM},   {MAP_A_KN, A, L.NP, L.ldr, true, AL_TN},    {MAP_A_IM, A, L.NP, L.ldr, false, AL_TM},
        {MAP_B_KM, Bm, L.NP, L.ldr, true, AL_TM},  {MAP_B_KN, Bm, L.NP, L.ldr, true, AL_TN},   {MAP_B_IN, Bm, L.NP, L.ldr, false, AL_TN},
        {MAP_XT_KN, Xt, L.NP, L.ldn, true, AL_TN}, {MAP_XT_IN, Xt, L.NP, L.ldn, false, AL_TN}, {MAP_T_KN, Tm, L.ldr, L.ldn, true, AL_TN},
        {MAP_G_IM, Gg, L.ldr, L.ldr, false, AL_TM},
    };
    for (int q = 0; q < MAP_COUNT; q++) {
        const Spec& sp = spec[q];
        const int rc = als_make_map(&M->m[sp.which], &M->bytes[sp.which], sp.base, sp.rows, sp.ld, L.per, B, sp.kmajor, sp.ti);
        if (rc) return rc;
    }
    return MVMC_OK;
}

// `order` (device, [B], may be null): clip index solved by CTA i - the pipeline passes the clips sorted by their previous
// frame's iteration count, longest first, so the last wave of CTAs is the short solves.
int AL_LAUNCHER(const double* sim, const int* dim_groups, int n_groups, const int* f32_first_iter,
                const double* rand_stream, const int* order, int B, int N, int rmax, void* workspace, uint32_t* xbin,
                int* n_iter, void* stream) {
    if (!sim || !dim_groups || !rand_stream || !workspace || !xbin || !n_iter) return MVMC_ERR_INVALID;
    if (B <= 0 || N <= 0 || N > 1024 || rmax <= 0 || rmax > 128 || n_groups <= 0 || n_groups > MVMC_MAX_VIEWS + 1)
        return MVMC_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return MVMC_ERR_INVALID;   // tensor maps need a 16-byte aligned base
    AlsMaps maps;
    const int rc = als_build_maps(&maps, (double*)workspace, B, N, rmax);
    if (rc) return rc;
    const size_t smem = als_smem_bytes(N);
    MVMC_CUDA_OK(cudaFuncSetAttribute(k_als, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MVMC_LAUNCH(k_als, dim3(B), dim3(AL_THREADS), smem, stream, maps, sim, dim_groups, n_groups, f32_first_iter, rand_stream,
                order, N, rmax, (double*)workspace, xbin, n_iter, 50.0, 0.1, 1e-4, 1000);
    MVMC_CHECK_LAUNCH("k_als");
    return MVMC_OK;
}

#ifndef AL_VARIANT
// Tile shape by problem size: N <= 192 (scenes of up to 16 people per view: n ~ 130, r = 32) runs the build with 48 x 48
// CTA tiles and 4-warp CTAs (als.cu compiled with -DAL_VARIANT=small -DAL_FM_=3 -DAL_THREADS_=128) - 64 x 96 tiles leave most
// of their fragments on padding there (n = 131 is two 64-row tiles plus 3 rows). Same arithmetic per element; the two
// builds differ in which rows a CTA's warps own, never in a summation order along k.
int mvmc_match_als_ordered_small(const double* sim, const int* dim_groups, int n_groups, const int* f32_first_iter,
                                 const double* rand_stream, const int* order, int B, int N, int rmax, void* workspace,
                                 uint32_t* xbin, int* n_iter, void* stream);
static int g_als_force_variant = -1;   // -1 auto, 0 default tiles, 1 small tiles (mvmc_als_force_variant: measurements)
extern "C" int mvmc_als_force_variant(int v) {
    g_als_force_variant = v;
    return MVMC_OK;
}
int mvmc_match_als_ordered(const double* sim, const int* dim_groups, int n_groups, const int* f32_first_iter,
                           const double* rand_stream, const int* order, int B, int N, int rmax, void* workspace, uint32_t* xbin,
                           int* n_iter, void* stream) {
    // By problem size (N = max_tracks + views x max_poses <= 192 covers scenes of up to 16 people per view, n ~ 130, r = 32) and
    // by batch size: the small build's 4-warp CTAs finish MORE clips per second once they queue for SMs (4 per SM, 592
    // resident) but each clip takes longer than under an 8-warp CTA, so a launch that fits in one wave (strong scaling: 170
    // clips per group at 8 GPUs, where a frame takes as long as its slowest clip) keeps the big build.
    const bool small = g_als_force_variant < 0 ? (N <= 192 && B > 592) : g_als_force_variant == 1;
    if (small)
        return mvmc_match_als_ordered_small(sim, dim_groups, n_groups, f32_first_iter, rand_stream, order, B, N, rmax, workspace,
                                            xbin, n_iter, stream);
    return mvmc_match_als_ordered_default(sim, dim_groups, n_groups, f32_first_iter, rand_stream, order, B, N, rmax, workspace,
                                          xbin, n_iter, stream);
}

// enable != 0: start counting (resets); enable == 0: stop. out (may be null) receives the PH_COUNT cycle sums so far.
extern "C" int mvmc_als_phase_profile(int enable, double* out) {
#ifndef MVMC_EMU
    unsigned long long h[PH_COUNT];
    MVMC_CUDA_OK(cudaDeviceSynchronize());
    MVMC_CUDA_OK(cudaMemcpyFromSymbol(h, g_als_phase, sizeof(h)));
    if (out) for (int q = 0; q < PH_COUNT; q++) out[q] = (double)h[q];
    const int on = enable ? 1 : 0;
    if (enable) {
        memset(h, 0, sizeof(h));
        MVMC_CUDA_OK(cudaMemcpyToSymbol(g_als_phase, h, sizeof(h)));
    }
    MVMC_CUDA_OK(cudaMemcpyToSymbol(g_als_phase_on, &on, sizeof(on)));
#else
    if (out) for (int q = 0; q < PH_COUNT; q++) out[q] = 0.0;
    (void)enable;
#endif
    return MVMC_OK;
}

extern "C" int mvmc_match_als(const double* sim, const int* dim_groups, int n_groups, const int* f32_first_iter,
                              const double* rand_stream, int B, int N, int rmax, void* workspace, uint32_t* xbin,
                              int* n_iter, void* stream) {
    return mvmc_match_als_ordered(sim, dim_groups, n_groups, f32_first_iter, rand_stream, nullptr, B, N, rmax, workspace, xbin,
                                  n_iter, stream);
}
#endif   // AL_VARIANT
