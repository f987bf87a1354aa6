// ALS/ADMM low-rank multi-way matcher (the reference's live association solver).
//
// Reference row: A5 mv_association.py:222-318 (match_als). Per iteration
//     Xt = Z - (Y - W + beta)/mu;  B = (inv(A^T A + a/mu I) A^T Xt)^T;  A = (inv(B^T B + a/mu I) B^T Xt^T)^T;
//     X = A B^T;  Z = clip(X + Y/mu) with same-group blocks zeroed and unit diagonal;  Y += mu (X - Z)
// until ||X - Z||_F / n < tol and mu ||X - X_prev||_F / n < tol, mu doubled / halved on a 10x imbalance.
//
// k_als: ONE CTA (8 warps) PER CLIP-FRAME, 2 resident CTAs per SM. The n x n iterates (W, Z, Y, X, Xt) and the
// n x r factors live in a per-clip global workspace; the seven products of an iteration (three of them 2 r n^2
// flops: the dominant cost of the whole capture path) run as tiled FP64 tensor-core GEMMs:
//   * CTA tile 64 x 96, k-chunks of 16 staged in shared memory by 16-byte cp.async (LDGSTS) with zero fill, double buffered;
//   * each warp owns a 32 x 24 piece of the tile (12 accumulator fragments) and issues mma.sync.m8n8k4.f64 (DMMA) from conflict-free
//     fragment loads (row strides chosen so the 4 x 8 fragment footprint covers every bank once);
//   * the Z / Y / X / next-Xt update and both residual norms are the epilogue of the X = A B^T product, so an
//     iteration makes three passes over n x n data (Xt twice, the epilogue's X, Y, W once) instead of seven;
//   * Xt for the next iteration is written by that epilogue assuming mu stays (it changes a handful of times per
//     solve; then one element-wise pass rebuilds Xt from Z, Y, W with the new mu) - same arithmetic, same values.
// FP64 DMMA and DFMA have the same peak on B200 (37 TFLOP/s measured, tools/micro/dmma_probe.cu); DMMA is used because it
// needs 8x fewer issue slots and 4x less shared-memory bandwidth per flop, which is what bounds a small-tile GEMM.
#include "mvmc_common.cuh"

namespace mvmc {

constexpr int AL_THREADS = 256;           // 8 warps: 2 along M x 4 along N, warp tile 32 x 24
constexpr int AL_KC = 16;                 // k-chunk
constexpr int AL_TM = 64;                 // CTA tile rows
constexpr int AL_TN = 96;                 // CTA tile columns (4 warp columns x 24)
constexpr int AL_SKM = AL_TM + 8;         // k-major tile row stride, M operand  (= 64 B mod 128 B)
constexpr int AL_SKN = AL_TN + 8;         // k-major tile row stride, N operand
constexpr int AL_SI = AL_KC + 4;          // i-major tile row stride               (= 32 B mod 128 B)
constexpr int AL_STAGE_M = (AL_KC * AL_SKM > AL_TM * AL_SI) ? AL_KC * AL_SKM : AL_TM * AL_SI;   // doubles
constexpr int AL_STAGE_N = (AL_KC * AL_SKN > AL_TN * AL_SI) ? AL_KC * AL_SKN : AL_TN * AL_SI;
constexpr int AL_STAGE = AL_STAGE_M + AL_STAGE_N;
constexpr int AL_POOL = 8192;             // doubles of shared memory shared by the two stages and the r x r inverse
constexpr int AL_RSMEM = 88;              // largest r whose normal matrix is inverted in registers (11 rows x 8 warps)
static_assert(2 * AL_STAGE <= AL_POOL, "stages must fit the pool");

// ---- asynchronous 16-byte global -> shared copies with zero fill ----
__device__ __forceinline__ void cp16(double* dst, const double* src, int n_valid /*0,1,2 doubles*/) {
#ifdef MVMC_EMU
    dst[0] = n_valid > 0 ? src[0] : 0.0;
    dst[1] = n_valid > 1 ? src[1] : 0.0;
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int bytes = n_valid * 8;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void cp_commit() {
#ifndef MVMC_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_wait() {
#ifndef MVMC_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

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

// An operand of a product, seen as element(k, i), 0 <= k < K (contraction index), 0 <= i < I.
struct Operand {
    const double* p;
    int ld;
    int I;
};
// layout of an operand (template argument): k-major: element(k, i) = p[k*ld + i];  i-major: element(k, i) = p[i*ld + k]
constexpr bool KMAJ = true, IMAJ = false;

// Staging of the [k0, k0+KC) x [i0, i0+TI) block of an operand into shared memory (zero filled outside K x I), 16 bytes
// per cp.async. The (shared offset, global pointer, validity) of the <= 4 pieces a thread copies are computed once per
// CTA tile; advancing to the next k-chunk is one pointer increment per piece.
template <int TI, bool kmajor>
struct Stager {
    static constexpr int NP = (AL_KC * (TI / 2) + AL_THREADS - 1) / AL_THREADS;   // pieces per thread (3 or 4)
    const double* base;
    int goff[NP];     // offset (doubles) of the piece in chunk 0 from `base`
    unsigned valid;   // 2 bits per piece: k-major = valid doubles along i (0..2); i-major = 2 if row i is inside I
    int step;         // offset increment per k-chunk

    // piece q of this thread -> (shared offset, k offset inside the chunk); constant divisors only
    __device__ __forceinline__ static void where(int q, int& soff, int& kk) {
        const int e = threadIdx.x + q * AL_THREADS;
        if (kmajor) {
            kk = e / (TI / 2);
            soff = kk * (TI + 8) + (e % (TI / 2)) * 2;
        } else {
            kk = (e % (AL_KC / 2)) * 2;
            soff = (e / (AL_KC / 2)) * AL_SI + kk;
        }
    }
    __device__ __forceinline__ void init(const Operand& op, int i0) {
        base = op.p;
        step = kmajor ? AL_KC * op.ld : AL_KC;
        valid = 0;
#pragma unroll
        for (int q = 0; q < NP; q++) {
            const int e = threadIdx.x + q * AL_THREADS;
            goff[q] = 0;
            if (e >= AL_KC * (TI / 2)) continue;
            if (kmajor) {
                const int k = e / (TI / 2), i = i0 + (e % (TI / 2)) * 2;
                int nv = op.I - i;
                nv = nv < 0 ? 0 : (nv > 2 ? 2 : nv);
                valid |= (unsigned)nv << (2 * q);
                goff[q] = k * op.ld + (nv > 0 ? i : 0);
            } else {
                const int i = i0 + e / (AL_KC / 2), k = (e % (AL_KC / 2)) * 2;
                const bool in = i < op.I;
                valid |= (in ? 2u : 0u) << (2 * q);
                goff[q] = (in ? i * op.ld : 0) + k;
            }
        }
    }
    // copy chunk kc (k0 = kc * KC) into S
    __device__ __forceinline__ void issue(double* S, int kc, int K) const {
        const int k0 = kc * AL_KC;
#pragma unroll
        for (int q = 0; q < NP; q++) {
            if (threadIdx.x + q * AL_THREADS >= AL_KC * (TI / 2)) continue;
            int soff, kk;
            where(q, soff, kk);
            const int nvi = (valid >> (2 * q)) & 3;
            int nv;
            if (kmajor) {
                nv = (k0 + kk < K) ? nvi : 0;
            } else {
                nv = K - (k0 + kk);
                nv = nv < 0 ? 0 : (nv > 2 ? 2 : nv);
                nv = nvi ? nv : 0;
            }
            cp16(S + soff, nv > 0 ? base + goff[q] + kc * step : base, nv);
        }
    }
};

template <int TI, bool kmajor>
__device__ __forceinline__ double frag(const double* S, int i8, int kk, int lane) {
    // element(k = kk + lane%4, i = i8 + lane/4)
    return kmajor ? S[(kk + (lane & 3)) * (TI + 8) + i8 + (lane >> 2)] : S[(i8 + (lane >> 2)) * AL_SI + kk + (lane & 3)];
}

// C[m][n] = sum_k Mop(k, m) * Nop(k, n) for m < Mop.I, n < Nop.I; ep(m, n, v0, v1) receives C[m][n], C[m][n+1]
// (n even; the caller guards n+1 < N); slot = 0/1 tells which of the two fragments of the row it is, and pre(slot, m, n)
// is called for both fragments before either ep so that an epilogue can issue its global loads together.
// All threads of the CTA must call it. `pool` holds the two stages.
#define MVMC_ALS_GEMM_ATTR __forceinline__   // (a non-inlined copy per product measured 35 % slower: operand structs through memory)
template <bool MK, bool NK, class EP, class PRE>
__device__ MVMC_ALS_GEMM_ATTR void cta_gemm(const Operand Mop, const Operand Nop, int K, double* pool, EP ep, PRE pre) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp & 1, wn = warp >> 1;               // warp tile: rows 32*wm.., columns 24*wn..
    const int M = Mop.I, N = Nop.I;
    const int nk = (K + AL_KC - 1) / AL_KC;
    for (int m0 = 0; m0 < M; m0 += AL_TM) {
        const int mw0 = m0 + 32 * wm;
        const int mt = mw0 >= M ? 0 : min(4, (M - mw0 + 7) >> 3);       // live row tiles of this warp
        for (int n0 = 0; n0 < N; n0 += AL_TN) {
            const int nw0 = n0 + 24 * wn;
            const int nt = (nw0 >= N || mt == 0) ? 0 : min(3, (N - nw0 + 7) >> 3);
            double acc[4][3][2];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
            Stager<AL_TM, MK> sm_;
            Stager<AL_TN, NK> sn_;
            sm_.init(Mop, m0);
            sn_.init(Nop, n0);
            __syncthreads();   // the pool may still be read by the previous tile / another phase
            sm_.issue(pool, 0, K);
            sn_.issue(pool + AL_STAGE_M, 0, K);
            cp_commit();
            for (int kc = 0; kc < nk; kc++) {
                double* cur = pool + (kc & 1) * AL_STAGE;
                if (kc + 1 < nk) {
                    double* nxt = pool + ((kc + 1) & 1) * AL_STAGE;
                    sm_.issue(nxt, kc + 1, K);
                    sn_.issue(nxt + AL_STAGE_M, kc + 1, K);
                    cp_commit();
                    cp_wait<1>();
                } else {
                    cp_wait<0>();
                }
                __syncthreads();
                if (nt > 0) {
                    const double* Sm = cur;
                    const double* Sn = cur + AL_STAGE_M;
#pragma unroll
                    for (int kk = 0; kk < AL_KC; kk += 4) {
                        double bf[3];
#pragma unroll
                        for (int b = 0; b < 3; b++) bf[b] = b < nt ? frag<AL_TN, NK>(Sn, 24 * wn + 8 * b, kk, lane) : 0.0;
#pragma unroll
                        for (int a = 0; a < 4; a++) {
                            if (a < mt) {
                                const double av = frag<AL_TM, MK>(Sm, 32 * wm + 8 * a, kk, lane);
#pragma unroll
                                for (int b = 0; b < 3; b++)
                                    if (b < nt) dmma(acc[a][b][0], acc[a][b][1], av, bf[b]);
                            }
                        }
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int a = 0; a < 4; a++) {
                if (a < mt) {
                    const int m = mw0 + a * 8 + (lane >> 2);
                    // loads of all fragments of the row first, then the arithmetic and the stores
#pragma unroll
                    for (int b = 0; b < 3; b++) {
                        const int nn = nw0 + 8 * b + 2 * (lane & 3);
                        if (b < nt && m < M && nn < N) pre(b, m, nn);
                    }
#pragma unroll
                    for (int b = 0; b < 3; b++) {
                        const int nn = nw0 + 8 * b + 2 * (lane & 3);
                        if (b < nt && m < M && nn < N) ep(b, m, nn, acc[a][b][0], acc[a][b][1]);
                    }
                }
            }
        }
    }
}
struct NoPre {
    __device__ __forceinline__ void operator()(int, int, int) const {}
};

// Inverse of the SPD r x r matrix Gg (global, leading dimension ldr) by Gauss-Jordan without pivoting, REGISTER resident:
// thread (warp w, lane l) owns the elements (i = w + 8a, j = l + 32b), a < NA, b < NB, for the whole elimination, so a
// pivot step is: the owners of row k / column k publish them (double-buffered in `aux`, [2][2][96]), ONE barrier, then
// every thread does NA*NB fused multiply-adds on its own registers - no shared-memory read-modify-write, no index
// arithmetic; the pivot row / column themselves are patched afterwards under (nearly) uniform branches.
constexpr int GJ_LD = 96;
template <int NA, int NB>
__device__ void invert_spd_regs(double* Gg, int r, int ldr, double* aux) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double g[NA][NB];
#pragma unroll
    for (int a = 0; a < NA; a++)
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const int i = w + 8 * a, j = lane + 32 * b;
            g[a][b] = (i < r && j < r) ? Gg[(size_t)i * ldr + j] : 0.0;
        }
    for (int k = 0; k < r; k++) {
        double* rk = aux + (k & 1) * 2 * GJ_LD;   // pivot row (raw)
        double* ck = rk + GJ_LD;                  // pivot column (raw)
        const int ka = k >> 3, kb = k >> 5;
        if (w == (k & 7)) {                       // this warp owns row k
#pragma unroll
            for (int a = 0; a < NA; a++)
                if (a == ka) {
#pragma unroll
                    for (int b = 0; b < NB; b++) rk[lane + 32 * b] = g[a][b];
                }
        }
        if (lane == (k & 31)) {                   // this lane owns column k (in every warp)
#pragma unroll
            for (int b = 0; b < NB; b++)
                if (b == kb) {
#pragma unroll
                    for (int a = 0; a < NA; a++) ck[w + 8 * a] = g[a][b];
                }
        }
        __syncthreads();
        const double p = 1.0 / rk[k];
        double rkp[NB], f[NA];
#pragma unroll
        for (int b = 0; b < NB; b++) rkp[b] = rk[lane + 32 * b] * p;
#pragma unroll
        for (int a = 0; a < NA; a++) f[a] = ck[w + 8 * a];
#pragma unroll
        for (int a = 0; a < NA; a++)
#pragma unroll
            for (int b = 0; b < NB; b++) g[a][b] = g[a][b] - f[a] * rkp[b];
        if (w == (k & 7)) {                       // row k becomes the scaled pivot row
#pragma unroll
            for (int a = 0; a < NA; a++)
                if (a == ka) {
#pragma unroll
                    for (int b = 0; b < NB; b++) g[a][b] = rkp[b];
                }
        }
        if (lane == (k & 31)) {                   // column k becomes -f p, the pivot itself p
#pragma unroll
            for (int b = 0; b < NB; b++)
                if (b == kb) {
#pragma unroll
                    for (int a = 0; a < NA; a++) g[a][b] = (w + 8 * a == k) ? p : -f[a] * p;
                }
        }
        // (the next pivot publishes into the other half of aux; the barrier after it orders the reuse of this half)
    }
#pragma unroll
    for (int a = 0; a < NA; a++)
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const int i = w + 8 * a, j = lane + 32 * b;
            if (i < r && j < r) Gg[(size_t)i * ldr + j] = g[a][b];
        }
}

// Fallback for r > 88: in place in global memory, thread t owns columns t % 32 + 32q of rows t / 32 + nw p.
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

// Gg (r x r, ld ldr, global) <- inverse of Gg
__device__ void invert_normal_matrix(double* Gg, int r, int ldr, double* aux) {
    __syncthreads();
    if (r <= 32) invert_spd_regs<4, 1>(Gg, r, ldr, aux);
    else if (r <= 64) invert_spd_regs<8, 2>(Gg, r, ldr, aux);
    else if (r <= AL_RSMEM) invert_spd_regs<11, 3>(Gg, r, ldr, aux);
    else invert_spd(Gg, r, ldr, aux);
    __syncthreads();
}

struct AlsLayout {
    int N, ldn, ldr;
    size_t per;
    __host__ __device__ AlsLayout(int N_, int rmax) {
        N = N_;
        ldn = (N_ + 7) & ~7;
        ldr = (rmax + 7) & ~7;
        per = (size_t)5 * N * ldn + (size_t)2 * N * ldr + (size_t)ldr * ldn + (size_t)ldr * ldr;
    }
};

__global__ void __launch_bounds__(AL_THREADS, 2)
    k_als(const double* __restrict__ sim, const int* __restrict__ dim_groups, int n_groups,
          const int* __restrict__ f32_first_iter, const double* __restrict__ rand_stream, int N, int rmax,
          double* __restrict__ ws, uint32_t* __restrict__ xbin, int* __restrict__ n_iter_out, double alpha, double beta,
          double tol, int max_iter) {
    MVMC_DYN_SMEM(double, smem);
    const int b = blockIdx.x;
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

    double* pool = smem;                       // [AL_POOL] stages / inverse
    double* scratch = pool + AL_POOL;          // [32]
    double* inv_aux = scratch + 32;            // [4 * 96] pivot row / column of the Gauss-Jordan inverse, double buffered
    int* s_grp = reinterpret_cast<int*>(inv_aux + 4 * GJ_LD);   // [N]

    const AlsLayout L(N, rmax);
    const int ldn = L.ldn, ldr = L.ldr;
    double* W = ws + (size_t)b * L.per;
    double* Z = W + (size_t)N * ldn;
    double* Y = Z + (size_t)N * ldn;
    double* Xm = Y + (size_t)N * ldn;
    double* Xt = Xm + (size_t)N * ldn;
    double* A = Xt + (size_t)N * ldn;          // [n][ldr]
    double* Bm = A + (size_t)N * ldr;          // [n][ldr]
    double* Tm = Bm + (size_t)N * ldr;         // [r][ldn]
    double* Gg = Tm + (size_t)ldr * ldn;       // [r][ldr]

    const double* S = sim + (size_t)b * N * N;
    const bool f32 = f32_first_iter != nullptr && f32_first_iter[b] != 0;

    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int g = 0;
        for (int q = 0; q < n_groups; q++)
            if (i >= dg[q] && i < dg[q + 1]) g = q;  // an index belongs to the group whose [start, end) contains it
        s_grp[i] = g;
    }
    double mu = 64.0;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e / n, j = e % n;
        const size_t o = (size_t)i * ldn + j;
        double w;
        if (f32) w = (double)(0.5f * ((float)S[(size_t)i * N + j] + (float)S[(size_t)j * N + i]));
        else w = 0.5 * (S[(size_t)i * N + j] + S[(size_t)j * N + i]);
        W[o] = w;
        Z[o] = w;
        Xm[o] = w;
        Y[o] = 0.0;
        // first Xt (Z = W, Y = 0); the float32 no-track path of the reference keeps float32 through this expression
        if (f32) Xt[o] = (double)((float)w - ((0.0f - (float)w) + (float)beta) / (float)mu);
        else Xt[o] = w - (0.0 - w + beta) / mu;
    }
    for (int e = threadIdx.x; e < n * r; e += blockDim.x) A[(size_t)(e / r) * ldr + (e % r)] = rand_stream[e];
    __syncthreads();

    int it = 0;
    for (it = 0; it < max_iter; it++) {
        const double reg = alpha / mu;
        // ---- G = A^T A + reg I, inverted ----
        {
            const Operand opA{A, ldr, r};
            cta_gemm<KMAJ, KMAJ>(opA, opA, n, pool, [&](int, int m, int nn, double v0, double v1) {
                Gg[(size_t)m * ldr + nn] = (m == nn) ? v0 + reg * 1.0 : v0 + reg * 0.0;
                if (nn + 1 < r) Gg[(size_t)m * ldr + nn + 1] = (m == nn + 1) ? v1 + reg * 1.0 : v1 + reg * 0.0;
            }, NoPre());
        }
        invert_normal_matrix(Gg, r, ldr, inv_aux);
        // ---- T = A^T Xt ----
        cta_gemm<KMAJ, KMAJ>(Operand{A, ldr, r}, Operand{Xt, ldn, n}, n, pool, [&](int, int m, int nn, double v0, double v1) {
            Tm[(size_t)m * ldn + nn] = v0;
            if (nn + 1 < n) Tm[(size_t)m * ldn + nn + 1] = v1;
        }, NoPre());
        __syncthreads();
        // ---- B = (Ginv T)^T ----
        cta_gemm<IMAJ, KMAJ>(Operand{Gg, ldr, r}, Operand{Tm, ldn, n}, r, pool, [&](int, int m, int nn, double v0, double v1) {
            Bm[(size_t)nn * ldr + m] = v0;
            if (nn + 1 < n) Bm[(size_t)(nn + 1) * ldr + m] = v1;
        }, NoPre());
        __syncthreads();
        // ---- H = B^T B + reg I, inverted ----
        {
            const Operand opB{Bm, ldr, r};
            cta_gemm<KMAJ, KMAJ>(opB, opB, n, pool, [&](int, int m, int nn, double v0, double v1) {
                Gg[(size_t)m * ldr + nn] = (m == nn) ? v0 + reg : v0;
                if (nn + 1 < r) Gg[(size_t)m * ldr + nn + 1] = (m == nn + 1) ? v1 + reg : v1;
            }, NoPre());
        }
        invert_normal_matrix(Gg, r, ldr, inv_aux);
        // ---- T = B^T Xt^T : T[m][i] = sum_j B[j][m] Xt[i][j] ----
        cta_gemm<KMAJ, IMAJ>(Operand{Bm, ldr, r}, Operand{Xt, ldn, n}, n, pool, [&](int, int m, int nn, double v0, double v1) {
            Tm[(size_t)m * ldn + nn] = v0;
            if (nn + 1 < n) Tm[(size_t)m * ldn + nn + 1] = v1;
        }, NoPre());
        __syncthreads();
        // ---- A = (Hinv T)^T ----
        cta_gemm<IMAJ, KMAJ>(Operand{Gg, ldr, r}, Operand{Tm, ldn, n}, r, pool, [&](int, int m, int nn, double v0, double v1) {
            A[(size_t)nn * ldr + m] = v0;
            if (nn + 1 < n) A[(size_t)(nn + 1) * ldr + m] = v1;
        }, NoPre());
        __syncthreads();
        // ---- X = A B^T, fused with the Z / Y / next-Xt updates and both residual norms ----
        double pacc = 0.0, dacc = 0.0;
        double2 ex0[3], ey[3], ew[3];   // X_prev, Y, W of the three fragments of a row (16-byte loads, issued together)
        auto one = [&](int i, int j, double x, double x0, double y, double w, bool live, double& yn, double& z, double& xt) {
            const double dd = x - x0;
            z = x + y / mu;
            if (s_grp[i] == s_grp[live ? j : i]) z = 0.0;
            if (i == j) z = 1.0;
            if (z < 0.0) z = 0.0;
            if (z > 1.0) z = 1.0;
            const double pd = x - z;
            yn = y + mu * pd;
            xt = z - (yn - w + beta) / mu;   // next iteration's Xt if mu stays
            if (live) {
                dacc += dd * dd;
                pacc += pd * pd;
            }
        };
        cta_gemm<IMAJ, IMAJ>(Operand{A, ldr, n}, Operand{Bm, ldr, n}, r, pool,
                 [&](int slot, int i, int j, double v0, double v1) {
                     const size_t o = (size_t)i * ldn + j;
                     const bool live1 = j + 1 < n;   // the odd column of the last pair may be padding (written, never read)
                     double2 yn, z, xt;
                     one(i, j, v0, ex0[slot].x, ey[slot].x, ew[slot].x, true, yn.x, z.x, xt.x);
                     one(i, j + 1, v1, ex0[slot].y, ey[slot].y, ew[slot].y, live1, yn.y, z.y, xt.y);
                     *reinterpret_cast<double2*>(Y + o) = yn;
                     *reinterpret_cast<double2*>(Z + o) = z;
                     double2 xv;
                     xv.x = v0;
                     xv.y = v1;
                     *reinterpret_cast<double2*>(Xm + o) = xv;
                     *reinterpret_cast<double2*>(Xt + o) = xt;
                 },
                 [&](int slot, int i, int j) {
                     const size_t o = (size_t)i * ldn + j;
                     ex0[slot] = *reinterpret_cast<const double2*>(Xm + o);
                     ey[slot] = *reinterpret_cast<const double2*>(Y + o);
                     ew[slot] = *reinterpret_cast<const double2*>(W + o);
                 });
        const double psum = block_sum(pacc, scratch);
        const double dsum = block_sum(dacc, scratch);
        const double p_res = sqrt(psum) / n;
        const double d_res = mu * sqrt(dsum) / n;
        if (p_res < tol && d_res < tol) {
            it++;
            break;
        }
        double mu_new = mu;
        if (p_res > 10.0 * d_res) mu_new = 2.0 * mu;
        else if (d_res > 10.0 * p_res) mu_new = mu / 2.0;
        if (mu_new != mu) {
            mu = mu_new;
            __syncthreads();
            for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
                const size_t o = (size_t)(e / n) * ldn + (e % n);
                Xt[o] = Z[o] - (Y[o] - W[o] + beta) / mu;
            }
        }
        __syncthreads();
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

}  // namespace mvmc

using namespace mvmc;

static size_t als_smem_bytes(int N) { return (size_t)(AL_POOL + 32 + 4 * GJ_LD) * sizeof(double) + (size_t)N * sizeof(int); }

extern "C" size_t mvmc_match_als_workspace_bytes(int B, int N, int rmax) {
    const AlsLayout L(N, rmax);
    return (size_t)B * L.per * sizeof(double);
}

extern "C" int mvmc_match_als(const double* sim, const int* dim_groups, int n_groups, const int* f32_first_iter,
                              const double* rand_stream, int B, int N, int rmax, void* workspace, uint32_t* xbin,
                              int* n_iter, void* stream) {
    if (!sim || !dim_groups || !rand_stream || !workspace || !xbin || !n_iter) return MVMC_ERR_INVALID;
    if (B <= 0 || N <= 0 || N > 1024 || rmax <= 0 || rmax > 128 || n_groups <= 0 || n_groups > MVMC_MAX_VIEWS + 1)
        return MVMC_ERR_INVALID;
    const size_t smem = als_smem_bytes(N);
    MVMC_CUDA_OK(cudaFuncSetAttribute(k_als, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MVMC_LAUNCH(k_als, dim3(B), dim3(AL_THREADS), smem, stream, sim, dim_groups, n_groups, f32_first_iter, rand_stream, N,
                rmax, (double*)workspace, xbin, n_iter, 50.0, 0.1, 1e-4, 1000);
    MVMC_CHECK_LAUNCH("k_als");
    return MVMC_OK;
}
