// Cross-view affinity stage: pose filtering + index layout, fundamental matrices, distance matrix,
// NaN fill + similarity. FP64 CUDA-core kernels (no dense contraction on this stage).
//
// Reference rows: A0 motion_capture.py:1023-1043, A1 mv_math_util.py:57-77,
// A2 mv_math_util.py:80-115, A3 motion_capture.py:403-414, A4 motion_capture.py:643-756,
// A7 motion_capture.py:597-631 + mv_math_util.py:267-351.
#include "mvmc_common.cuh"

namespace mvmc {

__constant__ int c_common_b18[MVMC_N_COMMON] = {1, 2, 3, 4, 5, 6, 9, 10, 11, 12, 13, 14, 15, 16, 17};
__constant__ int c_common_coco[MVMC_N_COMMON] = {11, 13, 15, 12, 14, 16, 5, 7, 9, 6, 8, 10, 0, 3, 4};

// ------------------------------------------------------------------------------------------------
// A0 + index layout. One CTA per clip.
// ------------------------------------------------------------------------------------------------
__global__ void k_prepare(const double* __restrict__ kps, const int* __restrict__ n_pose,
                          const int* __restrict__ n_trk, int C, int Pmax, int Tmax, uint8_t* __restrict__ keep,
                          int* __restrict__ dim_groups, int* __restrict__ idx_view, int* __restrict__ idx_pose) {
    __shared__ uint8_t s_keep[MVMC_MAX_VIEWS * MVMC_MAX_POSES];
    const int b = blockIdx.x;
    const int N = Tmax + C * Pmax;
    for (int q = threadIdx.x; q < C * Pmax; q += blockDim.x) {
        const int v = q / Pmax, p = q % Pmax;
        bool ok = false;
        if (p < n_pose[b * C + v]) {
            const double* k = kps + ((size_t)(b * C + v) * Pmax + p) * (MVMC_N_COCO * 3);
            int cnt = 0;
            double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
            for (int j = 0; j < MVMC_N_COCO; j++) {
                if (k[3 * j + 2] > 0.01) {
                    cnt++;
                    xmin = fmin(xmin, k[3 * j]);
                    xmax = fmax(xmax, k[3 * j]);
                    ymin = fmin(ymin, k[3 * j + 1]);
                    ymax = fmax(ymax, k[3 * j + 1]);
                }
            }
            ok = cnt >= 4 && !((xmax - xmin) < 5.0 || (ymax - ymin) < 5.0);
        }
        s_keep[q] = ok;
        keep[(size_t)b * C * Pmax + q] = ok;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int T = n_trk[b];
        if (T > Tmax) T = Tmax;
        int* dg = dim_groups + b * (C + 2);
        int* iv = idx_view + (size_t)b * N;
        int* ip = idx_pose + (size_t)b * N;
        dg[0] = 0;
        dg[1] = T;
        int pos = T;
        for (int t = 0; t < T; t++) {
            iv[t] = -1;
            ip[t] = t;
        }
        for (int v = 0; v < C; v++) {
            for (int p = 0; p < Pmax; p++)
                if (s_keep[v * Pmax + p]) {
                    iv[pos] = v;
                    ip[pos] = p;
                    pos++;
                }
            dg[v + 2] = pos;
        }
        for (; pos < N; pos++) {
            iv[pos] = -2;
            ip[pos] = -1;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// A1. One thread per ordered camera pair. F[i][j] entry (r, c) = det of the 4x4 made of two rows of
// P_i (picked by c) stacked on two rows of P_j (picked by r).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double det4_rows(const double* r0, const double* r1, const double* r2, const double* r3) {
    // Laplace expansion by complementary 2x2 minors of rows (0,1) and (2,3)
    const double s0 = r0[0] * r1[1] - r0[1] * r1[0];
    const double s1 = r0[0] * r1[2] - r0[2] * r1[0];
    const double s2 = r0[0] * r1[3] - r0[3] * r1[0];
    const double s3 = r0[1] * r1[2] - r0[2] * r1[1];
    const double s4 = r0[1] * r1[3] - r0[3] * r1[1];
    const double s5 = r0[2] * r1[3] - r0[3] * r1[2];
    const double c5 = r2[2] * r3[3] - r2[3] * r3[2];
    const double c4 = r2[1] * r3[3] - r2[3] * r3[1];
    const double c3 = r2[1] * r3[2] - r2[2] * r3[1];
    const double c2 = r2[0] * r3[3] - r2[3] * r3[0];
    const double c1 = r2[0] * r3[2] - r2[2] * r3[0];
    const double c0 = r2[0] * r3[1] - r2[1] * r3[0];
    return s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
}

__global__ void k_fundamental(const double* __restrict__ P, double* __restrict__ F, int B, int C) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * C * C) return;
    const int b = t / (C * C), i = (t / C) % C, j = t % C;
    const double* p1 = P + (size_t)(b * C + i) * 12;
    const double* p2 = P + (size_t)(b * C + j) * 12;
    double* f = F + (size_t)t * 9;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            f[r * 3 + c] = det4_rows(p1 + 4 * ((c + 1) % 3), p1 + 4 * ((c + 2) % 3), p2 + 4 * ((r + 1) % 3),
                                     p2 + 4 * ((r + 2) % 3));
}

// ------------------------------------------------------------------------------------------------
// A7 helper. F = K0^-T (R0 R1^T) K1^T [K1 R1 R0^T (T0 - R0 R1^T T1)]_x, float64 math, float32 result.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mm3(const double* a, const double* b, double* o) {
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) o[r * 3 + c] = a[r * 3] * b[c] + a[r * 3 + 1] * b[3 + c] + a[r * 3 + 2] * b[6 + c];
}
__device__ __forceinline__ void mv3(const double* a, const double* x, double* o) {
    for (int r = 0; r < 3; r++) o[r] = a[r * 3] * x[0] + a[r * 3 + 1] * x[1] + a[r * 3 + 2] * x[2];
}
__device__ __forceinline__ void tr3(const double* a, double* o) {
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) o[r * 3 + c] = a[c * 3 + r];
}
__device__ __forceinline__ void inv3(const double* m, double* o) {
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    const double id = 1.0 / det;
    o[0] = c00 * id;
    o[1] = (m[2] * m[7] - m[1] * m[8]) * id;
    o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    o[3] = c01 * id;
    o[4] = (m[0] * m[8] - m[2] * m[6]) * id;
    o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    o[6] = c02 * id;
    o[7] = (m[1] * m[6] - m[0] * m[7]) * id;
    o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

__global__ void k_fundamental_krt(const double* __restrict__ K, const double* __restrict__ Rt, float* __restrict__ F32,
                                  int B, int C) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * C * C) return;
    const int b = t / (C * C), i = (t / C) % C, j = t % C;
    const double* K0 = K + (size_t)(b * C + i) * 9;
    const double* K1 = K + (size_t)(b * C + j) * 9;
    const double* rt0 = Rt + (size_t)(b * C + i) * 12;
    const double* rt1 = Rt + (size_t)(b * C + j) * 12;
    double R0[9], R1[9], T0[3], T1[3];
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) {
            R0[r * 3 + c] = rt0[r * 4 + c];
            R1[r * 3 + c] = rt1[r * 4 + c];
        }
        T0[r] = rt0[r * 4 + 3];
        T1[r] = rt1[r * 4 + 3];
    }
    double R0t[9], R1t[9], K1t[9], K0i[9], K0it[9], RR[9], a[9], bb[9], v[3], w[3];
    tr3(R0, R0t);
    tr3(R1, R1t);
    tr3(K1, K1t);
    inv3(K0, K0i);
    tr3(K0i, K0it);
    mm3(R0, R1t, RR);              // R0 R1^T
    mv3(RR, T1, v);                // (R0 R1^T) T1
    for (int r = 0; r < 3; r++) v[r] = T0[r] - v[r];
    mm3(K1, R1, a);                // K1 R1
    mm3(a, R0t, bb);               // (K1 R1) R0^T
    mv3(bb, v, w);                 // e = K1 R1 R0^T (T0 - R0 R1^T T1)
    const double sk[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    mm3(K0it, RR, a);              // K0^-T (R0 R1^T)
    mm3(a, K1t, bb);               // ... K1^T
    mm3(bb, sk, a);                // ... [e]_x
    float f[9];
    float sum = 0.f;
    for (int q = 0; q < 9; q++) {
        f[q] = (float)a[q];
        sum += f[q];
    }
    if (sum == 0.f)
        for (int q = 0; q < 9; q++) f[q] += 1e-12f;
    for (int q = 0; q < 9; q++) F32[(size_t)t * 9 + q] = f[q];
}

// ------------------------------------------------------------------------------------------------
// A2/A3/A4/A7 distance entries. CTA = 16x16 tile of (i, j); the 32 items of the tile are staged in
// shared memory (coalesced 8-byte loads of the (x,y,score) triples), camera matrices come through L1.
// ------------------------------------------------------------------------------------------------
#define AFF_TILE 16
#define AFF_ITEM 54  // 18 joints x 3 doubles (2D poses use the first 51)

__device__ __forceinline__ void line_from(const double* f, bool transpose, double x, double y, double& a, double& b,
                                          double& c) {
    if (!transpose) {
        a = f[0] * x + f[1] * y + f[2];
        b = f[3] * x + f[4] * y + f[5];
        c = f[6] * x + f[7] * y + f[8];
    } else {
        a = f[0] * x + f[3] * y + f[6];
        b = f[1] * x + f[4] * y + f[7];
        c = f[2] * x + f[5] * y + f[8];
    }
    double nu = a * a + b * b;
    nu = nu != 0.0 ? 1.0 / sqrt(nu) : 1.0;
    a *= nu;
    b *= nu;
    c *= nu;
}

// symmetric point-to-epiline distance, mv_math_util.py:80-115 (f = F[vi][vj])
__device__ double epipolar_error(const double* f, const double* ki, const double* kj) {
    double total = 0.0;
    int cnt = 0;
    for (int q = 0; q < MVMC_N_COCO; q++) {
        const double s = ki[3 * q + 2] * kj[3 * q + 2];
        if (!(s > 0.1)) continue;
        const double x1 = ki[3 * q], y1 = ki[3 * q + 1], x2 = kj[3 * q], y2 = kj[3 * q + 1];
        double a, b, c;
        line_from(f, false, x1, y1, a, b, c);
        const double d1 = fabs(a * x2 + b * y2 + c) / sqrt(a * a + b * b);
        line_from(f, true, x2, y2, a, b, c);
        const double d2 = fabs(a * x1 + b * y1 + c) / sqrt(a * a + b * b);
        total = total + 0.5 * (d1 + d2);
        cnt++;
    }
    return cnt ? total / cnt : NAN;
}

// mean reprojection distance of a BASIC_18 track pose against a COCO 2D pose, motion_capture.py:403-414
__device__ double reprojection_error(const double* P, const double* trk, const double* k2) {
    double total = 0.0;
    int cnt = 0;
    for (int q = 0; q < MVMC_N_COMMON; q++) {
        const int jb = c_common_b18[q], jc = c_common_coco[q];
        if (!(k2[3 * jc + 2] * 1.0 > 0.1)) continue;
        const double X = trk[3 * jb], Y = trk[3 * jb + 1], Z = trk[3 * jb + 2];
        const double pu = P[0] * X + P[1] * Y + P[2] * Z + P[3];
        const double pv = P[4] * X + P[5] * Y + P[6] * Z + P[7];
        const double pw = P[8] * X + P[9] * Y + P[10] * Z + P[11];
        const double du = pu / (1e-5 + pw) - k2[3 * jc];
        const double dv = pv / (1e-5 + pw) - k2[3 * jc + 1];
        total += sqrt(du * du + dv * dv);
        cnt++;
    }
    return cnt ? total / cnt : NAN;
}

// un-normalised |l . x| averaged over all 17 joints, l = normalise(F^T x0): mv_math_util.py:288-317
__device__ double projected_distance(const float* f32, const double* k0, const double* k1) {
    double f[9];
    for (int q = 0; q < 9; q++) f[q] = (double)f32[q];
    double total = 0.0;
    for (int q = 0; q < MVMC_N_COCO; q++) {
        double a, b, c;
        line_from(f, true, k0[3 * q], k0[3 * q + 1], a, b, c);
        total += fabs(a * k1[3 * q] + b * k1[3 * q + 1] + c);
    }
    return total / MVMC_N_COCO;
}

__global__ void __launch_bounds__(AFF_TILE* AFF_TILE)
    k_affinity(const double* __restrict__ kps, const double* __restrict__ P, const double* __restrict__ F,
               const float* __restrict__ F32, const double* __restrict__ trk_joints, const int* __restrict__ n_trk,
               const int* __restrict__ dim_groups, const int* __restrict__ idx_view, const int* __restrict__ idx_pose,
               int C, int Pmax, int Tmax, int force_f64, double* __restrict__ dst) {
    const int b = blockIdx.z;
    const int N = Tmax + C * Pmax;
    const int n = dim_groups[b * (C + 2) + C + 1];
    const int i0 = blockIdx.y * AFF_TILE, j0 = blockIdx.x * AFF_TILE;
    if (i0 >= n || j0 >= n) return;
    __shared__ double s_item[2 * AFF_TILE][AFF_ITEM];
    __shared__ int s_view[2 * AFF_TILE], s_pose[2 * AFF_TILE];
    const int tid = threadIdx.y * AFF_TILE + threadIdx.x;
    const int T = min(n_trk[b], Tmax);
    // stage: items 0..15 = rows i0.., 16..31 = cols j0.. - every thread brings a few elements of several items and all of
    // its loads are issued before the first store (item by item, the 32 global-load latencies used to run back to back)
    if (tid < 2 * AFF_TILE) {
        const int g = (tid < AFF_TILE) ? i0 + tid : j0 + tid - AFF_TILE;
        s_view[tid] = g < n ? idx_view[(size_t)b * N + g] : -2;
        s_pose[tid] = g < n ? idx_pose[(size_t)b * N + g] : 0;
    }
    __syncthreads();
    {
        constexpr int PER = (2 * AFF_TILE * AFF_ITEM + AFF_TILE * AFF_TILE - 1) / (AFF_TILE * AFF_TILE);
        double val[PER];
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int e = tid + q * AFF_TILE * AFF_TILE, it = e / AFF_ITEM, k = e - it * AFF_ITEM;
            val[q] = 0.0;
            if (it < 2 * AFF_TILE) {
                const int v = s_view[it], p = s_pose[it];
                const int len = (v < 0) ? MVMC_N_B18 * 3 : MVMC_N_COCO * 3;
                if (v != -2 && k < len) {
                    const double* src = (v < 0) ? trk_joints + ((size_t)b * Tmax + p) * (MVMC_N_B18 * 3)
                                                : kps + ((size_t)(b * C + v) * Pmax + p) * (MVMC_N_COCO * 3);
                    val[q] = src[k];
                }
            }
        }
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int e = tid + q * AFF_TILE * AFF_TILE, it = e / AFF_ITEM, k = e - it * AFF_ITEM;
            if (it < 2 * AFF_TILE) s_item[it][k] = val[q];
        }
    }
    __syncthreads();
    const int i = i0 + threadIdx.y, j = j0 + threadIdx.x;
    if (i >= n || j >= n) return;
    const int vi = s_view[threadIdx.y], vj = s_view[AFF_TILE + threadIdx.x];
    const double* ki = s_item[threadIdx.y];
    const double* kj = s_item[AFF_TILE + threadIdx.x];
    double d;
    if (T > 0 || force_f64) {
        if (i == j) d = 0.0;
        else if (vi >= 0 && vi == vj) d = NAN;
        else if (vi >= 0 && vj >= 0) d = epipolar_error(F + ((size_t)(b * C + vi) * C + vj) * 9, ki, kj);
        else if (vi >= 0) d = reprojection_error(P + (size_t)(b * C + vi) * 12, kj, ki);
        else if (vj >= 0) d = reprojection_error(P + (size_t)(b * C + vj) * 12, ki, kj);
        else d = NAN;
    } else {
        // float32 path of match_spatial: every stored distance is rounded to float
        if (i == j) d = 0.0;
        else if (vi == vj) d = 50.0;
        else {
            const bool fwd = vi < vj;
            const int h = fwd ? vi : vj, k = fwd ? vj : vi;
            const double* ph = fwd ? ki : kj;
            const double* pk = fwd ? kj : ki;
            const double m = 0.5 * (projected_distance(F32 + ((size_t)(b * C + h) * C + k) * 9, ph, pk) +
                                    projected_distance(F32 + ((size_t)(b * C + k) * C + h) * 9, pk, ph));
            d = (double)(float)m;
        }
    }
    dst[(size_t)b * N * N + (size_t)i * N + j] = d;
}

// ------------------------------------------------------------------------------------------------
// NumPy's float32 arithmetic, restated (the reference's no-track affinity, mv_math_util.py:348-350, is float32 NumPy and
// its last bit decides assignments when the matcher does not converge). None of this is in /root/reference: it is NumPy
// (requirements.txt numpy==1.17.4; container 2.3.5), restated from its published algorithms and checked bit for bit against
// the container's NumPy (tests/test_oracle_golden.py::test_numpy_float32_restatement, tests/stage_checks.py):
//   * add.reduce over a contiguous float32 array = pairwise summation (numpy/_core/src/umath/loops_utils.h.src
//     pairwise_sum): blocks of <= 128 elements summed with 8 strided accumulators, combined as ((r0+r1)+(r2+r3))+((r4+r5)+
//     (r6+r7)) plus a sequential remainder; longer ranges split at n/2 rounded down to a multiple of 8;
//   * mean = sum / count, var = sum((x - mean)^2) / count, std = sqrt(var), all float32 (numpy/_core/_methods.py _mean, _var);
//   * exp (numpy/_core/src/umath/loops_exponent_log.dispatch.c.src, AVX2/AVX512F kernel): k = round(x log2 e) by the
//     1.5 2^23 trick, Cody-Waite reduction r = x - k ln2 in two FMAs, a (5,2) rational minimax in r by Horner FMAs, one
//     division, scaling by 2^k.
// Every operation is an explicitly rounded intrinsic so that nvcc cannot contract a multiply and an add into an FMA where
// NumPy has two roundings (or the other way round).
// ------------------------------------------------------------------------------------------------
namespace np32 {
constexpr int MAX_LEAVES = 2048;   // n*n <= 102400 elements, leaves hold 64..128
#ifdef MVMC_EMU
__device__ __forceinline__ float fadd(float a, float b) { volatile float r = a + b; return r; }
__device__ __forceinline__ float fsub(float a, float b) { volatile float r = a - b; return r; }
__device__ __forceinline__ float fmul(float a, float b) { volatile float r = a * b; return r; }
__device__ __forceinline__ float fdiv(float a, float b) { volatile float r = a / b; return r; }
__device__ __forceinline__ float fsqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return fmaf(a, b, c); }
#else
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#endif

__device__ __forceinline__ float exp(float x) {
    if (x != x) return x;
    if (x > 88.72283905206835f) return INFINITY;
    if (x < -103.97208121f) return 0.0f;
    const float magic = 12582912.0f;   // 1.5 * 2^23
    float k = fmul(x, 1.442695040888963407359924681001892137f);
    k = fsub(fadd(k, magic), magic);
    float r = ffma(k, -6.93145752e-1f, x);
    r = ffma(k, -1.42860677e-6f, r);
    float num = ffma(5.082762527590693718096e-04f, r, 6.757896990527504603057e-03f);
    num = ffma(num, r, 5.114512081637298353406e-02f);
    num = ffma(num, r, 2.473615434895520810817e-01f);
    num = ffma(num, r, 7.257664613233124478488e-01f);
    num = ffma(num, r, 9.999999999980870924916e-01f);
    float den = ffma(2.159509375685829852307e-02f, r, -2.742335390411667452936e-01f);
    den = ffma(den, r, 1.0f);
    return ldexpf(fdiv(num, den), (int)k);
}

// element e of the n x n matrix (row pitch N doubles, values are float32), optionally (x - mean)^2
__device__ __forceinline__ float elem(const double* D, int n, int N, int e, float mean, bool sq) {
    const float v = (float)D[(size_t)(e / n) * N + (e % n)];
    if (!sq) return v;
    const float x = fsub(v, mean);
    return fmul(x, x);
}
__device__ float leaf_sum(const double* D, int n, int N, int off, int len, float mean, bool sq) {
    if (len < 8) {
        float res = 0.0f;
        for (int i = 0; i < len; i++) res = fadd(res, elem(D, n, N, off + i, mean, sq));
        return res;
    }
    float r[8];
    for (int j = 0; j < 8; j++) r[j] = elem(D, n, N, off + j, mean, sq);
    int i = 8;
    for (; i < len - (len % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] = fadd(r[j], elem(D, n, N, off + i + j, mean, sq));
    float res = fadd(fadd(fadd(r[0], r[1]), fadd(r[2], r[3])), fadd(fadd(r[4], r[5]), fadd(r[6], r[7])));
    for (; i < len; i++) res = fadd(res, elem(D, n, N, off + i, mean, sq));
    return res;
}
// leaves of the pairwise recursion over [off, off + len), left to right
__device__ void split(int off, int len, int* leaf, int& nl) {
    if (len <= 128) {
        leaf[nl++] = off;
        return;
    }
    int n2 = len / 2;
    n2 -= n2 % 8;
    split(off, n2, leaf, nl);
    split(off + n2, len - n2, leaf, nl);
}
__device__ float combine(int len, const float* sum, int& k) {
    if (len <= 128) return sum[k++];
    int n2 = len / 2;
    n2 -= n2 % 8;
    const float a = combine(n2, sum, k);
    const float b = combine(len - n2, sum, k);
    return fadd(a, b);
}
// np.add.reduce of the flattened matrix (or of (x - mean)^2), float32; every thread of the CTA calls it and gets the result
__device__ float reduce_sum(const double* D, int n, int N, float mean, bool sq, int* s_leaf, float* s_sum, int& s_nleaf) {
    const int total = n * n;
    __syncthreads();
    if (threadIdx.x == 0) {
        int nl = 0;
        split(0, total, s_leaf, nl);
        s_leaf[nl] = total;
        s_nleaf = nl;
    }
    __syncthreads();
    const int nl = s_nleaf;
    for (int q = threadIdx.x; q < nl; q += blockDim.x) s_sum[q] = leaf_sum(D, n, N, s_leaf[q], s_leaf[q + 1] - s_leaf[q], mean, sq);
    __syncthreads();
    if (threadIdx.x == 0) {
        int k = 0;
        s_sum[0] = fadd(0.0f, combine(total, s_sum, k));   // the reduction starts from add's identity
    }
    __syncthreads();
    const float r = s_sum[0];
    __syncthreads();
    return r;
}
}  // namespace np32

// ------------------------------------------------------------------------------------------------
// NaN fill + similarity. One CTA per clip (n*n <= 102400 entries).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_simfill(const int* __restrict__ n_trk, const int* __restrict__ dim_groups, int C, int N, double* __restrict__ dst,
              double* __restrict__ sim) {
    __shared__ double scratch[32];
    __shared__ int s_leaf[np32::MAX_LEAVES + 1];
    __shared__ float s_sum[np32::MAX_LEAVES];
    __shared__ int s_nleaf;
    const int b = blockIdx.x;
    const int n = dim_groups[b * (C + 2) + C + 1];
    double* D = dst + (size_t)b * N * N;
    double* S = sim + (size_t)b * N * N;
    const int T = n_trk[b];
    if (n == 0) return;
    if (T > 0) {
        double mx = -INFINITY;
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
            const double d = D[(size_t)(e / n) * N + (e % n)];
            if (d == d) mx = fmax(mx, d);
        }
        mx = block_max(mx, scratch);
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
            const size_t o = (size_t)(e / n) * N + (e % n);
            double d = D[o];
            if (d != d) {
                d = mx + 1.0;
                D[o] = d;
            }
            double s = (d - 15.0) / 30.0;
            s = 1.0 / (1.0 + exp(5.0 * s));
            if (s < 1e-3) s = 0.0;
            if (s > 1.0) s = 1.0;
            S[o] = s;
        }
    } else {
        // affinity = sigmoid(5 * -(D - mean)/std) in float32 (mv_math_util.py:348-350), bit for bit what NumPy computes
        // (np32 below): pairwise float32 sums for mean and std, NumPy's own float32 exp.
        const int nn = n * n;
        const float cnt = (float)nn;
        const float mean = np32::fdiv(np32::reduce_sum(D, n, N, 0.0f, false, s_leaf, s_sum, s_nleaf), cnt);
        const float var = np32::fdiv(np32::reduce_sum(D, n, N, mean, true, s_leaf, s_sum, s_nleaf), cnt);
        const float sd = np32::fsqrt(var);
        for (int e = threadIdx.x; e < nn; e += blockDim.x) {
            const size_t o = (size_t)(e / n) * N + (e % n);
            const float a = np32::fdiv(-np32::fsub((float)D[o], mean), sd);
            const float s = np32::fdiv(1.0f, np32::fadd(1.0f, np32::exp(np32::fmul(-5.0f, a))));
            S[o] = (double)s;
        }
    }
}

}  // namespace mvmc

using namespace mvmc;

extern "C" int mvmc_fundamental(const double* P, double* F, int B, int C, void* stream) {
    if (!P || !F || B <= 0 || C <= 0 || C > MVMC_MAX_VIEWS) return MVMC_ERR_INVALID;
    const int n = B * C * C;
    MVMC_LAUNCH(k_fundamental, dim3((n + 127) / 128), dim3(128), 0, stream, P, F, B, C);
    MVMC_CHECK_LAUNCH("k_fundamental");
    return MVMC_OK;
}

extern "C" int mvmc_fundamental_krt(const double* K, const double* Rt, float* F32, int B, int C, void* stream) {
    if (!K || !Rt || !F32 || B <= 0 || C <= 0 || C > MVMC_MAX_VIEWS) return MVMC_ERR_INVALID;
    const int n = B * C * C;
    MVMC_LAUNCH(k_fundamental_krt, dim3((n + 127) / 128), dim3(128), 0, stream, K, Rt, F32, B, C);
    MVMC_CHECK_LAUNCH("k_fundamental_krt");
    return MVMC_OK;
}

extern "C" int mvmc_prepare(const double* kps, const int* n_pose, const int* n_trk, int B, int C, int Pmax, int Tmax,
                            uint8_t* keep, int* dim_groups, int* idx_view, int* idx_pose, void* stream) {
    if (!kps || !n_pose || !n_trk || !keep || !dim_groups || !idx_view || !idx_pose) return MVMC_ERR_INVALID;
    if (B <= 0 || C <= 0 || C > MVMC_MAX_VIEWS || Pmax <= 0 || Pmax > MVMC_MAX_POSES || Tmax < 0 ||
        Tmax > MVMC_MAX_TRACKS)
        return MVMC_ERR_INVALID;
    MVMC_LAUNCH(k_prepare, dim3(B), dim3(256), 0, stream, kps, n_pose, n_trk, C, Pmax, Tmax, keep, dim_groups, idx_view,
                idx_pose);
    MVMC_CHECK_LAUNCH("k_prepare");
    return MVMC_OK;
}

extern "C" int mvmc_affinity(const double* kps, const double* P, const double* F, const float* F32,
                             const double* trk_joints, const int* n_trk, const int* dim_groups, const int* idx_view,
                             const int* idx_pose, int B, int C, int Pmax, int Tmax, double* dst, double* sim,
                             void* stream) {
    if (!kps || !P || !F || !F32 || !trk_joints || !n_trk || !dim_groups || !idx_view || !idx_pose || !dst || !sim)
        return MVMC_ERR_INVALID;
    if (B <= 0 || C <= 0 || C > MVMC_MAX_VIEWS || Pmax <= 0 || Pmax > MVMC_MAX_POSES || Tmax < 0 ||
        Tmax > MVMC_MAX_TRACKS)
        return MVMC_ERR_INVALID;
    const int N = Tmax + C * Pmax;
    const int tiles = (N + AFF_TILE - 1) / AFF_TILE;
    MVMC_LAUNCH(k_affinity, dim3(tiles, tiles, B), dim3(AFF_TILE, AFF_TILE), 0, stream, kps, P, F, F32, trk_joints, n_trk,
                dim_groups, idx_view, idx_pose, C, Pmax, Tmax, 0, dst);
    MVMC_CHECK_LAUNCH("k_affinity");
    MVMC_LAUNCH(k_simfill, dim3(B), dim3(256), 0, stream, n_trk, dim_groups, C, N, dst, sim);
    MVMC_CHECK_LAUNCH("k_simfill");
    return MVMC_OK;
}

// The float64 distance matrix alone (A2 / A3 entries, NaN for same-view and track-track pairs, no NaN fill, no similarity),
// also when a clip has no tracks: what the Hungarian matchers of matchers.cu take their costs from.
extern "C" int mvmc_distances(const double* kps, const double* P, const double* F, const double* trk_joints, const int* n_trk,
                              const int* dim_groups, const int* idx_view, const int* idx_pose, int B, int C, int Pmax, int Tmax,
                              double* dst, void* stream) {
    if (!kps || !P || !F || !trk_joints || !n_trk || !dim_groups || !idx_view || !idx_pose || !dst) return MVMC_ERR_INVALID;
    if (B <= 0 || C <= 0 || C > MVMC_MAX_VIEWS || Pmax <= 0 || Pmax > MVMC_MAX_POSES || Tmax < 0 || Tmax > MVMC_MAX_TRACKS)
        return MVMC_ERR_INVALID;
    const int N = Tmax + C * Pmax;
    const int tiles = (N + AFF_TILE - 1) / AFF_TILE;
    MVMC_LAUNCH(k_affinity, dim3(tiles, tiles, B), dim3(AFF_TILE, AFF_TILE), 0, stream, kps, P, F, (const float*)nullptr, trk_joints,
                n_trk, dim_groups, idx_view, idx_pose, C, Pmax, Tmax, 1, dst);
    MVMC_CHECK_LAUNCH("k_affinity");
    return MVMC_OK;
}
