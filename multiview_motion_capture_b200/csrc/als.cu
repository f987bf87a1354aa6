// ALS/ADMM low-rank multi-way matcher + closure/parse/decode.
//
// Reference rows: A5 mv_association.py:222-318 (match_als),
// A6 mv_association.py:99-121 (transform_closure) + motion_capture.py:417-446 (parse_match_result)
// + motion_capture.py:762-808 / :618-624 (group decoding).
//
// k_als: ONE CTA PER INSTANCE (clip-frame). All iterates (W,Z,Y,Xt,X: n x n; A,B: n x r) live in a
// per-instance global workspace that stays L2-resident (<= 4.6 MB at n=320); the r x r normal matrix
// is inverted in shared memory. The three n*n*r products per iteration are FP64 CUDA-core tiled GEMMs
// (64x64 tile, 4x4 per thread) — this stage is DFMA-issue bound, not HBM bound.
#include "mvmc_common.cuh"

namespace mvmc {

constexpr int ALS_TS = 64;
constexpr int ALS_KC = 16;
constexpr int ALS_LD = ALS_TS + 2;
constexpr int ALS_THREADS = 256;

// S[kk][ii] = src[(k0+kk)*ld + (i0+ii)]  (ii contiguous in memory)
__device__ __forceinline__ void load_kmajor(double* S, const double* __restrict__ src, int ld, int i0, int k0, int I,
                                            int K) {
    const int ii = threadIdx.x & 63;
    for (int kk = threadIdx.x >> 6; kk < ALS_KC; kk += 4) {
        const int k = k0 + kk, i = i0 + ii;
        S[kk * ALS_LD + ii] = (k < K && i < I) ? src[(size_t)k * ld + i] : 0.0;
    }
}
// S[kk][ii] = src[(i0+ii)*ld + (k0+kk)]  (kk contiguous in memory)
__device__ __forceinline__ void load_imajor(double* S, const double* __restrict__ src, int ld, int i0, int k0, int I,
                                            int K) {
    const int kk = threadIdx.x & 15;
    for (int ii = threadIdx.x >> 4; ii < ALS_TS; ii += 16) {
        const int k = k0 + kk, i = i0 + ii;
        S[kk * ALS_LD + ii] = (k < K && i < I) ? src[(size_t)i * ld + k] : 0.0;
    }
}

template <class LA, class LB, class EP>
__device__ __forceinline__ void gemm_tiles(int M, int N, int K, LA la, LB lb, EP ep, double* As, double* Bs) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int m0 = 0; m0 < M; m0 += ALS_TS)
        for (int n0 = 0; n0 < N; n0 += ALS_TS) {
            double acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
            for (int k0 = 0; k0 < K; k0 += ALS_KC) {
                la(As, m0, k0);
                lb(Bs, n0, k0, m0);
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < ALS_KC; kk++) {
                    double a[4], b[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) a[i] = As[kk * ALS_LD + ty * 4 + i];
#pragma unroll
                    for (int j = 0; j < 4; j++) b[j] = Bs[kk * ALS_LD + tx * 4 + j];
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
                }
                __syncthreads();
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int m = m0 + ty * 4 + i, nn = n0 + tx * 4 + j;
                    if (m < M && nn < N) ep(m, nn, acc[i][j]);
                }
        }
}

// In-place inverse of the SPD r x r matrix G (leading dimension ldg) by Gauss-Jordan without pivoting.
__device__ void invert_spd(double* G, int r, int ldg) {
    for (int k = 0; k < r; k++) {
        const double p = 1.0 / G[k * ldg + k];
        __syncthreads();
        for (int j = threadIdx.x; j < r; j += blockDim.x)
            if (j != k) G[k * ldg + j] *= p;
        __syncthreads();
        for (int e = threadIdx.x; e < r * r; e += blockDim.x) {
            const int i = e / r, j = e % r;
            if (i != k && j != k) G[i * ldg + j] -= G[i * ldg + k] * G[k * ldg + j];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < r; i += blockDim.x) {
            if (i != k) G[i * ldg + k] = -G[i * ldg + k] * p;
            else G[k * ldg + k] = p;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(ALS_THREADS)
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
    const int ldg = r + 1;

    double* As = smem;
    double* Bs = As + ALS_KC * ALS_LD;
    double* Gs = Bs + ALS_KC * ALS_LD;
    double* scratch = Gs + rmax * (rmax + 1);
    int* s_grp = reinterpret_cast<int*>(scratch + 32);

    const size_t per = (size_t)5 * N * N + (size_t)3 * N * rmax;
    double* W = ws + (size_t)b * per;
    double* Z = W + (size_t)N * N;
    double* Y = Z + (size_t)N * N;
    double* Xt = Y + (size_t)N * N;
    double* Xm = Xt + (size_t)N * N;
    double* A = Xm + (size_t)N * N;
    double* Bm = A + (size_t)N * rmax;
    double* Tm = Bm + (size_t)N * rmax;

    const double* S = sim + (size_t)b * N * N;
    const bool f32 = f32_first_iter != nullptr && f32_first_iter[b] != 0;

    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int g = 0;
        for (int q = 0; q < n_groups; q++)
            if (dg[q] <= i) g = q;  // last group whose offset is <= i (empty groups share an offset)
        // an index belongs to the group whose [start,end) contains it
        for (int q = 0; q < n_groups; q++)
            if (i >= dg[q] && i < dg[q + 1]) g = q;
        s_grp[i] = g;
    }
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e / n, j = e % n;
        double w;
        if (f32) w = (double)(0.5f * ((float)S[(size_t)i * N + j] + (float)S[(size_t)j * N + i]));
        else w = 0.5 * (S[(size_t)i * N + j] + S[(size_t)j * N + i]);
        W[e] = w;
        Z[e] = w;
        Xm[e] = w;
        Y[e] = 0.0;
    }
    for (int e = threadIdx.x; e < n * r; e += blockDim.x) A[e] = rand_stream[e];
    __syncthreads();

    double mu = 64.0;
    int it = 0;
    for (it = 0; it < max_iter; it++) {
        const double reg = alpha / mu;
        const bool first_f32 = f32 && it == 0;
        // ---- G = A^T A + reg I ----
        gemm_tiles(r, r, n,
                   [&](double* Sm, int m0, int k0) { load_kmajor(Sm, A, r, m0, k0, r, n); },
                   [&](double* Sm, int n0, int k0, int) { load_kmajor(Sm, A, r, n0, k0, r, n); },
                   [&](int m, int nn, double v) { Gs[m * ldg + nn] = (m == nn) ? v + reg * 1.0 : v + reg * 0.0; }, As,
                   Bs);
        __syncthreads();
        invert_spd(Gs, r, ldg);
        // ---- T = A^T Xt, with Xt = Z - (Y - W + beta)/mu formed on the fly (and stored once) ----
        gemm_tiles(r, n, n,
                   [&](double* Sm, int m0, int k0) { load_kmajor(Sm, A, r, m0, k0, r, n); },
                   [&](double* Sm, int n0, int k0, int m0) {
                       const int ii = threadIdx.x & 63;
                       for (int kk = threadIdx.x >> 6; kk < ALS_KC; kk += 4) {
                           const int k = k0 + kk, i = n0 + ii;
                           double v = 0.0;
                           if (k < n && i < n) {
                               const size_t o = (size_t)k * n + i;
                               if (first_f32)
                                   v = (double)((float)Z[o] - (((float)Y[o] - (float)W[o]) + (float)beta) / (float)mu);
                               else
                                   v = Z[o] - (Y[o] - W[o] + beta) / mu;
                               if (m0 == 0) Xt[o] = v;
                           }
                           Sm[kk * ALS_LD + ii] = v;
                       }
                   },
                   [&](int m, int nn, double v) { Tm[(size_t)m * n + nn] = v; }, As, Bs);
        __syncthreads();
        // ---- B = (Ginv T)^T ----
        gemm_tiles(r, n, r,
                   [&](double* Sm, int m0, int k0) { load_kmajor(Sm, Gs, ldg, m0, k0, r, r); },
                   [&](double* Sm, int n0, int k0, int) { load_kmajor(Sm, Tm, n, n0, k0, n, r); },
                   [&](int m, int nn, double v) { Bm[(size_t)nn * r + m] = v; }, As, Bs);
        __syncthreads();
        // ---- H = B^T B + reg I ----
        gemm_tiles(r, r, n,
                   [&](double* Sm, int m0, int k0) { load_kmajor(Sm, Bm, r, m0, k0, r, n); },
                   [&](double* Sm, int n0, int k0, int) { load_kmajor(Sm, Bm, r, n0, k0, r, n); },
                   [&](int m, int nn, double v) { Gs[m * ldg + nn] = (m == nn) ? v + reg : v; }, As, Bs);
        __syncthreads();
        invert_spd(Gs, r, ldg);
        // ---- T = B^T Xt^T : T[m][i] = sum_j B[j][m] Xt[i][j] ----
        gemm_tiles(r, n, n,
                   [&](double* Sm, int m0, int k0) { load_kmajor(Sm, Bm, r, m0, k0, r, n); },
                   [&](double* Sm, int n0, int k0, int) { load_imajor(Sm, Xt, n, n0, k0, n, n); },
                   [&](int m, int nn, double v) { Tm[(size_t)m * n + nn] = v; }, As, Bs);
        __syncthreads();
        // ---- A = (Hinv T)^T ----
        gemm_tiles(r, n, r,
                   [&](double* Sm, int m0, int k0) { load_kmajor(Sm, Gs, ldg, m0, k0, r, r); },
                   [&](double* Sm, int n0, int k0, int) { load_kmajor(Sm, Tm, n, n0, k0, n, r); },
                   [&](int m, int nn, double v) { A[(size_t)nn * r + m] = v; }, As, Bs);
        __syncthreads();
        // ---- X = A B^T, fused with the Z / Y updates and both residual norms ----
        double pacc = 0.0, dacc = 0.0;
        gemm_tiles(n, n, r,
                   [&](double* Sm, int m0, int k0) { load_imajor(Sm, A, r, m0, k0, n, r); },
                   [&](double* Sm, int n0, int k0, int) { load_imajor(Sm, Bm, r, n0, k0, n, r); },
                   [&](int i, int j, double x) {
                       const size_t o = (size_t)i * n + j;
                       const double x0 = Xm[o];
                       const double dd = x - x0;
                       dacc += dd * dd;
                       const double y = Y[o];
                       double z = x + y / mu;
                       if (s_grp[i] == s_grp[j]) z = 0.0;
                       if (i == j) z = 1.0;
                       if (z < 0.0) z = 0.0;
                       if (z > 1.0) z = 1.0;
                       const double pd = x - z;
                       pacc += pd * pd;
                       Y[o] = y + mu * pd;
                       Z[o] = z;
                       Xm[o] = x;
                   },
                   As, Bs);
        const double psum = block_sum(pacc, scratch);
        const double dsum = block_sum(dacc, scratch);
        const double p_res = sqrt(psum) / n;
        const double d_res = mu * sqrt(dsum) / n;
        if (p_res < tol && d_res < tol) {
            it++;
            break;
        }
        if (p_res > 10.0 * d_res) mu = 2.0 * mu;
        else if (d_res > 10.0 * p_res) mu = mu / 2.0;
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
                const double v = 0.5 * (Xm[(size_t)i * n + j] + Xm[(size_t)j * n + i]);
                if (v > 0.5) bits |= 1u << q;
            }
        }
        xb[(size_t)i * NW + w] = bits;
    }
}

// ------------------------------------------------------------------------------------------------
// A6: closure quirk, leader assignment, first-kept-column parse, group decoding. One warp per
// instance; lane w owns bit-word w of every N-bit row (N <= 1024).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
    k_assign(const uint32_t* __restrict__ xbin, const int* __restrict__ dim_groups, const int* __restrict__ idx_view,
             const int* __restrict__ idx_pose, const int* __restrict__ n_trk, int C, int N, int Tmax, int max_new,
             int* __restrict__ trk_nsel, int* __restrict__ trk_sel, int* __restrict__ new_n, int* __restrict__ new_nsel,
             int* __restrict__ new_sel, int* __restrict__ counts, int* __restrict__ err) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const int NW = (N + 31) / 32;
    const int n = dim_groups[b * (C + 2) + C + 1];
    const int T = min(n_trk[b], Tmax);
    const uint32_t* xb = xbin + (size_t)b * N * NW;
    const int* iv = idx_view + (size_t)b * N;
    const int* ip = idx_pose + (size_t)b * N;
    int* tn = trk_nsel + (size_t)b * Tmax;
    int* ts = trk_sel + (size_t)b * Tmax * MVMC_MAX_SEL * 2;
    int* nn_ = new_nsel + (size_t)b * max_new;
    int* ns = new_sel + (size_t)b * max_new * MVMC_MAX_SEL * 2;
    // members[c] (bit rows) of the kept leader columns are re-derived on the fly; we only need, per
    // leader c, temp[c] = X[c] | (X[c][n-1] ? X[n-1] : 0).
    __shared__ int s_members[MVMC_MAX_TRACKS + MVMC_MAX_VIEWS * MVMC_MAX_POSES];
    for (int t = lane; t < Tmax; t += 32) tn[t] = -1;
    if (lane == 0) {
        new_n[b] = 0;
        for (int q = 0; q < 4; q++) counts[4 * b + q] = 0;
        err[b] = 0;
    }
    if (n <= 0) return;
    const uint32_t last_row = (lane < NW) ? xb[(size_t)(n - 1) * NW + lane] : 0u;
    uint32_t vis = 0;       // this lane's word of `vis`
    uint32_t assigned = 0;  // this lane's word of "row already attached to a kept column"
    int n_new = 0, dup = 0, error = 0, n_single = 0, n_trunc = 0;
    const bool has_trk = T > 0;
    for (int i = 0; i < n; i++) {
        const uint32_t vw = __shfl_sync(MVMC_FULL, vis, i >> 5);
        if ((vw >> (i & 31)) & 1u) continue;  // uniform across the warp
        uint32_t row = (lane < NW) ? xb[(size_t)i * NW + lane] : 0u;
        const uint32_t lw = __shfl_sync(MVMC_FULL, row, (n - 1) >> 5);
        if ((lw >> ((n - 1) & 31)) & 1u) row |= last_row;  // temp[i] = X[i] | X[i][n-1] * X[n-1]
        vis |= row;
        const int cnt = warp_sum_i(__popc(row));
        if (cnt < 2) continue;  // column kept only with >= 2 members (sum > 1.9)
        // rows join the FIRST kept column they belong to
        uint32_t mine = row & ~assigned;
        assigned |= row;
        // enumerate members in ascending order into shared memory
        int base = 0;
        for (int w = 0; w < NW; w++) {
            const uint32_t word = __shfl_sync(MVMC_FULL, mine, w);
            if (lane == 0) {
                uint32_t x = word;
                while (x) {
                    const int bit = __ffs((int)x) - 1;
                    s_members[base++] = w * 32 + bit;
                    x &= x - 1;
                }
            }
            base = __shfl_sync(MVMC_FULL, base, 0);
        }
        __syncwarp();
        if (base == 0) continue;  // empty group (`if cur_matches:`)
        if (lane == 0) {
            int t_idx = -1;
            if (has_trk)
                for (int q = 0; q < base; q++)
                    if (s_members[q] < T) {
                        t_idx = s_members[q];
                        break;
                    }
            int sel[MVMC_MAX_SEL][2];
            int nsel = 0;
            uint32_t seen_views = 0;
            bool over = false;
            for (int q = 0; q < base; q++) {
                const int g = s_members[q];
                if (has_trk && g < T) continue;
                const int v = iv[g];
                if (has_trk) {
                    if ((seen_views >> v) & 1u) {
                        dup++;
                        continue;
                    }
                    seen_views |= 1u << v;
                }
                if (nsel < MVMC_MAX_SEL) {
                    sel[nsel][0] = v;
                    sel[nsel][1] = ip[g];
                    nsel++;
                } else {
                    over = true;  // keep the first MVMC_MAX_SEL poses (include/mvmc.h: n_truncated)
                }
            }
            if (over) n_trunc++;
            if (nsel > 0) {
                if (t_idx >= 0) {
                    tn[t_idx] = nsel;
                    for (int q = 0; q < nsel; q++) {
                        ts[(t_idx * MVMC_MAX_SEL + q) * 2] = sel[q][0];
                        ts[(t_idx * MVMC_MAX_SEL + q) * 2 + 1] = sel[q][1];
                    }
                } else if (nsel < 2) {
                    n_single++;  // a 2D-only group the one-pose-per-view rule shrank to one pose: never born
                } else if (n_new < max_new) {
                    nn_[n_new] = nsel;
                    for (int q = 0; q < nsel; q++) {
                        ns[(n_new * MVMC_MAX_SEL + q) * 2] = sel[q][0];
                        ns[(n_new * MVMC_MAX_SEL + q) * 2 + 1] = sel[q][1];
                    }
                    n_new++;
                } else {
                    error = MVMC_ERR_CAPACITY;
                }
            }
        }
        __syncwarp();
    }
    if (lane == 0) {
        new_n[b] = n_new;
        counts[4 * b] = dup;
        counts[4 * b + 1] = n_single;
        counts[4 * b + 2] = n_trunc;
        err[b] = error;
    }
}

}  // namespace mvmc

using namespace mvmc;

static size_t als_smem_bytes(int N, int rmax) {
    return (size_t)(2 * ALS_KC * ALS_LD + rmax * (rmax + 1) + 32) * sizeof(double) + (size_t)N * sizeof(int);
}

extern "C" size_t mvmc_match_als_workspace_bytes(int B, int N, int rmax) {
    return (size_t)B * ((size_t)5 * N * N + (size_t)3 * N * rmax) * sizeof(double);
}

extern "C" int mvmc_match_als(const double* sim, const int* dim_groups, int n_groups, const int* f32_first_iter,
                              const double* rand_stream, int B, int N, int rmax, void* workspace, uint32_t* xbin,
                              int* n_iter, void* stream) {
    if (!sim || !dim_groups || !rand_stream || !workspace || !xbin || !n_iter) return MVMC_ERR_INVALID;
    if (B <= 0 || N <= 0 || N > 1024 || rmax <= 0 || rmax > 128 || n_groups <= 0 || n_groups > MVMC_MAX_VIEWS + 1)
        return MVMC_ERR_INVALID;
    const size_t smem = als_smem_bytes(N, rmax);
    MVMC_CUDA_OK(cudaFuncSetAttribute(k_als, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MVMC_LAUNCH(k_als, dim3(B), dim3(ALS_THREADS), smem, stream, sim, dim_groups, n_groups, f32_first_iter, rand_stream, N,
                rmax, (double*)workspace, xbin, n_iter, 50.0, 0.1, 1e-4, 1000);
    MVMC_CHECK_LAUNCH("k_als");
    return MVMC_OK;
}

extern "C" int mvmc_assign(const uint32_t* xbin, const int* dim_groups, const int* idx_view, const int* idx_pose,
                           const int* n_trk, int B, int C, int N, int Tmax, int max_new, int* trk_nsel, int* trk_sel,
                           int* new_n, int* new_nsel, int* new_sel, int* counts, int* err, void* stream) {
    if (!xbin || !dim_groups || !idx_view || !idx_pose || !n_trk || !trk_nsel || !trk_sel || !new_n || !new_nsel ||
        !new_sel || !counts || !err)
        return MVMC_ERR_INVALID;
    if (B <= 0 || N <= 0 || N > 1024 || C <= 0 || C > MVMC_MAX_VIEWS || Tmax < 0 || Tmax > MVMC_MAX_TRACKS ||
        max_new <= 0)
        return MVMC_ERR_INVALID;
    MVMC_LAUNCH(k_assign, dim3(B), dim3(32), 0, stream, xbin, dim_groups, idx_view, idx_pose, n_trk, C, N, Tmax, max_new,
                trk_nsel, trk_sel, new_n, new_nsel, new_sel, counts, err);
    MVMC_CHECK_LAUNCH("k_assign");
    return MVMC_OK;
}
