// Warp-synchronous trust-region-reflective least squares (unbounded), forward-difference Jacobian.
//
// Restates scipy.optimize.least_squares(method='trf', jac='2-point', tr_solver='exact', x_scale=1) as used by the
// reference (inverse_kinematics.py:236,274; mv_math_util.py:207): same step-size rule for the finite
// differences, same Moré iteration on the Levenberg-Marquardt parameter, same accept/reject/radius logic.
//
// ONE WARP PER SOLVE, no block barrier anywhere:
//   * lanes own Jacobian columns (one perturbed model evaluation each),
//   * J^T J is accumulated view by view from a 16-row chunk of J staged in shared memory, 8x8 register tiles
//     (lane = tile of the lower triangle, 64 DFMA per 16 doubles loaded),
//   * instead of SciPy's SVD of J, J^T J = Q T Q^T is reduced ONCE per Jacobian to tridiagonal form by Householder
//     reflections; every evaluation of the secular function phi(alpha) = ||(J^T J + alpha I)^-1 g|| - Delta and of
//     its derivative is then an O(n) tridiagonal LDL^T solve in the Q basis (mathematically the same phi, phi' SciPy
//     evaluates from singular values), and the step is rotated back with the stored reflectors.
// Every scalar the control flow depends on comes out of xor-butterfly reductions, which leave bitwise-identical
// values on all lanes, so the warp never diverges on a decision.
#pragma once
#include "mvmc_common.cuh"

namespace mvmc {

constexpr int WS_NC = 56;    // live Jacobian columns per solve (multiple of 8; BASIC_18: 39 / 49, triangulation refine: 54)
constexpr int WS_NT = WS_NC / 8;
constexpr int WS_LDA = 57;   // row stride of A (odd: conflict-free column walks with 64-bit accesses)
constexpr int WS_LDJ = 58;   // row stride of the J chunk (16-byte aligned rows, odd multiple of 16 B: conflict-free LDS.128)
constexpr int WS_CH = 8;     // rows of J per chunk (16 measured the same speed; 8 lets a fifth solver CTA fit on an SM)
constexpr double kSqrtEps = 1.4901161193847656e-08;   // sqrt(2^-52)
constexpr double kEps = 2.220446049250313e-16;
constexpr double kDblMax = 1.79769313486231570e308;

// pointers that cross a __noinline__ call lose their address space; the solver's workspaces are always shared memory
#ifdef MVMC_EMU
#define MVMC_ASSUME_SHARED(p) ((void)0)
#else
#define MVMC_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#endif

#ifdef MVMC_EMU
#define DMUL(a, b) ((a) * (b))
#define DSUB(a, b) ((a) - (b))
#define DDIV(a, b) ((a) / (b))
#else
#define DMUL(a, b) __dmul_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))
#define DDIV(a, b) __ddiv_rn((a), (b))
#endif

// NC = the most live columns the instance is laid out for (A and the FD scratch are NC wide; the vectors keep WS_NC slots).
// The track-update solver never has more than 49 live columns: its instance is TrfWarpT<50>, 5 KB smaller, which - with the
// trial residual vector parked in the idle J chunk - lets a sixth solver CTA fit on an SM.
template <int NC_>
struct alignas(16) TrfWarpT {
    static constexpr int NC = NC_;
    static constexpr int LDA = NC_ | 1;   // row stride of A (odd: conflict-free column walks with 64-bit accesses)
    double A[NC * LDA];             // J^T J, then the Householder vectors; aliased by the FD scratch while J is formed
    double Jc[WS_CH * WS_LDJ];      // current chunk of J, row major
    double x[MVMC_N_PARAM], xn[MVMC_N_PARAM];
    double g[WS_NC], gt[WS_NC], p[WS_NC], pt[WS_NC], d[WS_NC], e[WS_NC], tau[WS_NC], w[WS_NC], u[WS_NC], dx[WS_NC];
    double sc[8];
    int act[WS_NC];                 // parameter index behind each Jacobian column
};
typedef TrfWarpT<WS_NC> TrfWarp;
static_assert(TrfWarp::LDA == WS_LDA, "WS_LDA is the stride of the full-width instance");

struct TrfResult {
    int nfev, njev, status;
    double cost;
};

__device__ __forceinline__ double warp_max_abs(double v) { return warp_max(fabs(v)); }

// lane -> tile (ti >= tj) of the lower triangle of a WS_NT x WS_NT tile grid
__device__ __forceinline__ void lane_tile(int lane, int& ti, int& tj) {
    int t = lane;
    ti = 0;
    while (t > ti) {
        t -= ti + 1;
        ti++;
    }
    tj = t;
}

// ---- Jacobian, g = J^T f, A = J^T J, then A = Q T Q^T ---------------------------------------------------------------
// Res interface (all members are called by the whole warp):
//   int  m() const;                      residual rows
//   int  n_chunks() const;               J is produced chunk by chunk (<= WS_CH rows each)
//   int  chunk_rows(int c) const;        rows of chunk c;  row index of its first row = chunk_row0(c)
//   void eval(const double* x, double* f);                 all residuals at x (x, f in shared memory)
//   void fd_prepare(TW& s, int ncol);                      per column: perturb, evaluate the model, park the state in s.A
//   void fd_chunk(TW& s, int ncol, int c, const double* f);        fill s.Jc[r][col] = (r'(x + h e_col) - f) / dx for chunk c
template <class TW, class Res>
__device__ __noinline__ void trf_jacobian(TW& s, Res& res, int ncol, const double* f) {
    constexpr int LDA = TW::LDA;
    MVMC_ASSUME_SHARED(&s);
    if constexpr (!Res::kGlobalF) MVMC_ASSUME_SHARED(f);   // (the many-view birth solver keeps its residual vectors in global memory)
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < WS_CH * WS_LDJ; e += 32) s.Jc[e] = 0.0;
    // SciPy's 2-point rule: h = sqrt(eps) * sign(x) * max(1, |x|) with sign(0) = +1, dx = (x + h) - x
    for (int c = lane; c < ncol; c += 32) {
        const double xi = s.x[s.act[c]];
        const double h = kSqrtEps * (xi >= 0.0 ? 1.0 : -1.0) * fmax(1.0, fabs(xi));
        s.dx[c] = DSUB(xi + h, xi);
        s.tau[c] = 1.0 / s.dx[c];   // (tau is free until the tridiagonalisation: the residuals' fd_chunk multiply by it)
        s.w[c] = xi + h;    // perturbed value of the column's parameter
    }
    __syncwarp();
    res.fd_prepare(s, ncol);
    __syncwarp();
    int ti, tj;
    lane_tile(lane, ti, tj);
    const bool tile_on = ti < WS_NT && 8 * tj < ncol && 8 * ti < ((ncol + 7) & ~7);
    double acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) acc[a][b] = 0.0;
    double g0 = 0.0, g1 = 0.0;
    const int nch = res.n_chunks();
    for (int c = 0; c < nch; c++) {
        res.fd_chunk(s, ncol, c, f);
        __syncwarp();
        const int rows = res.chunk_rows(c);
        const double* fr = f + res.chunk_row0(c);
        if (tile_on) {
            const double* pa = s.Jc + 8 * ti;
            const double* pb = s.Jc + 8 * tj;
            for (int r = 0; r < rows; r++) {
                double a[8], b[8];
#pragma unroll
                for (int q = 0; q < 8; q += 2) {
                    const double2 va = *reinterpret_cast<const double2*>(pa + r * WS_LDJ + q);
                    const double2 vb = *reinterpret_cast<const double2*>(pb + r * WS_LDJ + q);
                    a[q] = va.x;
                    a[q + 1] = va.y;
                    b[q] = vb.x;
                    b[q + 1] = vb.y;
                }
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
            }
        }
        for (int r = 0; r < rows; r++) {
            const double fv = fr[r];
            g0 = fma(s.Jc[r * WS_LDJ + lane], fv, g0);
            if (lane + 32 < WS_NC) g1 = fma(s.Jc[r * WS_LDJ + lane + 32], fv, g1);
        }
        __syncwarp();
    }
    // the FD scratch (aliasing A) is dead from here on
    for (int e = lane; e < TW::NC * LDA; e += 32) s.A[e] = 0.0;
    __syncwarp();
    if (tile_on) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int r = 8 * ti + i, c = 8 * tj + j;
                if (TW::NC == WS_NC || (r < TW::NC && c < TW::NC)) {   // (the last tiles of a narrower instance hang over its edge)
                    s.A[r * LDA + c] = acc[i][j];
                    s.A[c * LDA + r] = acc[i][j];   // (on diagonal tiles both orders are written with the same products)
                }
            }
    }
    if (lane < ncol) s.g[lane] = g0;
    if (lane + 32 < ncol) s.g[lane + 32] = g1;
    __syncwarp();
    // diagonal tiles: acc[i][j] and acc[j][i] are the same sum in the same order, so A is exactly symmetric
}

// A (n x n, symmetric, full storage) -> tridiagonal T = Q^T A Q: d[0..n), e[0..n-1); reflector k is stored in
// A[k+2.., k] (v[k+1] = 1 implicit) with s.tau[k]. LAPACK dsytd2 (lower) arithmetic, one warp.
template <class TW>
__device__ __noinline__ void warp_tridiagonalise(TW& s, int n) {
    constexpr int LDA = TW::LDA;
    MVMC_ASSUME_SHARED(&s);
    const int lane = threadIdx.x & 31;
    for (int k = 0; k + 1 < n; k++) {
        const int r0 = k + 1 + lane, r1 = r0 + 32;
        const double a0 = r0 < n ? s.A[r0 * LDA + k] : 0.0;
        const double a1 = r1 < n ? s.A[r1 * LDA + k] : 0.0;
        const double alpha = __shfl_sync(MVMC_FULL, a0, 0);
        const double xn2 = warp_sum(fma(a1, a1, lane == 0 ? 0.0 : a0 * a0));
        if (lane == 0) s.d[k] = s.A[k * LDA + k];
        if (xn2 == 0.0) {
            if (lane == 0) {
                s.tau[k] = 0.0;
                s.e[k] = alpha;
            }
            __syncwarp();
            continue;
        }
        const double beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
        const double tau = (beta - alpha) / beta;
        const double scal = 1.0 / (alpha - beta);
        const double v0 = lane == 0 ? 1.0 : a0 * scal;
        const double v1 = a1 * scal;
        if (r0 < n) {
            s.u[r0] = v0;
            if (lane != 0) s.A[r0 * LDA + k] = v0;
        }
        if (r1 < n) {
            s.u[r1] = v1;
            s.A[r1 * LDA + k] = v1;
        }
        if (lane == 0) {
            s.tau[k] = tau;
            s.e[k] = beta;
        }
        __syncwarp();
        // w = tau * A22 v
        double w0 = 0.0, w1 = 0.0;
        if (r0 < n) {
            const double* row = s.A + r0 * LDA;
#pragma unroll 4
            for (int c = k + 1; c < n; c++) w0 = fma(row[c], s.u[c], w0);
        }
        if (r1 < n) {
            const double* row = s.A + r1 * LDA;
#pragma unroll 4
            for (int c = k + 1; c < n; c++) w1 = fma(row[c], s.u[c], w1);
        }
        w0 *= tau;
        w1 *= tau;
        const double wv = warp_sum(fma(r1 < n ? w1 : 0.0, v1, r0 < n ? w0 * v0 : 0.0));
        const double kk = -0.5 * tau * wv;
        w0 = fma(kk, v0, w0);
        w1 = fma(kk, v1, w1);
        if (r0 < n) s.w[r0] = w0;
        if (r1 < n) s.w[r1] = w1;
        __syncwarp();
        // A22 -= v w^T + w v^T
        if (r0 < n) {
            double* row = s.A + r0 * LDA;
#pragma unroll 4
            for (int c = k + 1; c < n; c++) row[c] = fma(-w0, s.u[c], fma(-v0, s.w[c], row[c]));
        }
        if (r1 < n) {
            double* row = s.A + r1 * LDA;
#pragma unroll 4
            for (int c = k + 1; c < n; c++) row[c] = fma(-w1, s.u[c], fma(-v1, s.w[c], row[c]));
        }
        __syncwarp();
    }
    if (lane == 0) {
        s.d[n - 1] = s.A[(n - 1) * LDA + (n - 1)];
        if (n >= 1) s.e[n - 1] = 0.0;
    }
    __syncwarp();
}

// y <- H_k y for k = 0..n-3 (forward = true: y <- Q^T y) or k = n-3..0 (y <- Q y). y in shared memory.
template <class TW>
__device__ __noinline__ void warp_apply_q(const TW& s, int n, double* y, bool transpose) {
    constexpr int LDA = TW::LDA;
    MVMC_ASSUME_SHARED(&s);
    MVMC_ASSUME_SHARED(y);
    const int lane = threadIdx.x & 31;
    for (int q = 0; q + 2 < n; q++) {
        const int k = transpose ? q : n - 3 - q;
        const double tau = s.tau[k];
        if (tau == 0.0) continue;  // uniform
        const int r0 = k + 1 + lane, r1 = r0 + 32;
        const double v0 = r0 < n ? (lane == 0 ? 1.0 : s.A[r0 * LDA + k]) : 0.0;
        const double v1 = r1 < n ? s.A[r1 * LDA + k] : 0.0;
        const double y0 = r0 < n ? y[r0] : 0.0;
        const double y1 = r1 < n ? y[r1] : 0.0;
        const double dot = tau * warp_sum(fma(v1, y1, v0 * y0));
        if (r0 < n) y[r0] = fma(-dot, v0, y0);
        if (r1 < n) y[r1] = fma(-dot, v1, y1);
        __syncwarp();
    }
}

// ---- (T + alpha I) y = b for the symmetric tridiagonal T (d, e), n <= 64, by parallel cyclic reduction -------------
// Lane l holds rows l and l + 32 (rows >= n are identity padding). Six reduction levels (strides 1..32) decouple every
// row; neighbours come through warp shuffles, nothing touches shared memory. The per-level elimination factors are kept
// so that a second right-hand side costs six fused multiply-add levels only. PCR is Gaussian elimination without
// pivoting in a different order: stable for the positive definite T + alpha I it is used on.
struct PcrFactors {
    double k1[6][2], k2[6][2];   // elimination factors per level and row slot
    double binv[2];              // 1 / final diagonal
    bool pd;                     // every intermediate diagonal stayed positive
};

__device__ __forceinline__ void pcr_neigh(double v0, double v1, int s, int lane, double& lo0, double& lo1, double& hi0,
                                          double& hi1, double fill) {
    // values of rows (i - s) -> lo, (i + s) -> hi for i = lane (slot 0) and i = lane + 32 (slot 1); out of range -> fill
    if (s == 32) {
        lo0 = fill;
        lo1 = v0;
        hi0 = v1;
        hi1 = fill;
        return;
    }
    const double d0 = __shfl_sync(MVMC_FULL, v0, (lane - s) & 31), d1 = __shfl_sync(MVMC_FULL, v1, (lane - s) & 31);
    const double u0 = __shfl_sync(MVMC_FULL, v0, (lane + s) & 31), u1 = __shfl_sync(MVMC_FULL, v1, (lane + s) & 31);
    const bool wrap_lo = lane - s < 0, wrap_hi = lane + s >= 32;
    lo0 = wrap_lo ? fill : d0;       // row lane - s
    lo1 = wrap_lo ? d0 : d1;         // row lane + 32 - s: slot 0 of lane - s + 32 when it wraps, else slot 1 of lane - s
    hi0 = wrap_hi ? u1 : u0;         // row lane + s: slot 1 of lane + s - 32 when it wraps
    hi1 = wrap_hi ? fill : u1;       // row lane + 32 + s
}

// factor + solve: b0/b1 = right-hand side rows (lane, lane+32) -> y0/y1
template <class TW>
__device__ __forceinline__ void pcr_solve(const TW& s, int n, double alpha, double floor_, double& r0, double& r1,
                                          PcrFactors& F) {
    const int lane = threadIdx.x & 31;
    const int i0 = lane, i1 = lane + 32;
    double a0 = (i0 > 0 && i0 < n) ? s.e[i0 - 1] : 0.0, a1 = (i1 < n) ? s.e[i1 - 1] : 0.0;
    double c0 = (i0 + 1 < n) ? s.e[i0] : 0.0, c1 = (i1 + 1 < n) ? s.e[i1] : 0.0;
    double b0 = (i0 < n) ? s.d[i0] + alpha : 1.0, b1 = (i1 < n) ? s.d[i1] + alpha : 1.0;
    bool pd = true;
#pragma unroll
    for (int lv = 0; lv < 6; lv++) {
        const int st = 1 << lv;
        if (!(b0 > floor_)) {
            pd = false;
            b0 = floor_;
        }
        if (!(b1 > floor_)) {
            pd = false;
            b1 = floor_;
        }
        double bl0, bl1, bh0, bh1, al0, al1, ah0, ah1, cl0, cl1, ch0, ch1, rl0, rl1, rh0, rh1;
        // every row inverts its own diagonal once and hands the reciprocal to its neighbours (two divisions per level and
        // lane instead of four; a double-precision division is ~25 instructions with a slow path)
        pcr_neigh(1.0 / b0, 1.0 / b1, st, lane, bl0, bl1, bh0, bh1, 1.0);
        pcr_neigh(a0, a1, st, lane, al0, al1, ah0, ah1, 0.0);
        pcr_neigh(c0, c1, st, lane, cl0, cl1, ch0, ch1, 0.0);
        pcr_neigh(r0, r1, st, lane, rl0, rl1, rh0, rh1, 0.0);
        const double k10 = a0 * bl0, k20 = c0 * bh0, k11 = a1 * bl1, k21 = c1 * bh1;
        F.k1[lv][0] = k10;
        F.k2[lv][0] = k20;
        F.k1[lv][1] = k11;
        F.k2[lv][1] = k21;
        b0 = fma(-ah0, k20, fma(-cl0, k10, b0));
        b1 = fma(-ah1, k21, fma(-cl1, k11, b1));
        r0 = fma(-rh0, k20, fma(-rl0, k10, r0));
        r1 = fma(-rh1, k21, fma(-rl1, k11, r1));
        a0 = -al0 * k10;
        a1 = -al1 * k11;
        c0 = -ch0 * k20;
        c1 = -ch1 * k21;
    }
    if (!(b0 > floor_)) {
        pd = false;
        b0 = floor_;
    }
    if (!(b1 > floor_)) {
        pd = false;
        b1 = floor_;
    }
    F.binv[0] = 1.0 / b0;
    F.binv[1] = 1.0 / b1;
    r0 *= F.binv[0];
    r1 *= F.binv[1];
    F.pd = __all_sync(MVMC_FULL, pd);
}
// another right-hand side for the matrix factored by pcr_solve
__device__ __forceinline__ void pcr_resolve(const PcrFactors& F, double& r0, double& r1) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int lv = 0; lv < 6; lv++) {
        double rl0, rl1, rh0, rh1;
        pcr_neigh(r0, r1, 1 << lv, lane, rl0, rl1, rh0, rh1, 0.0);
        r0 = fma(-rh0, F.k2[lv][0], fma(-rl0, F.k1[lv][0], r0));
        r1 = fma(-rh1, F.k2[lv][1], fma(-rl1, F.k1[lv][1], r1));
    }
    r0 *= F.binv[0];
    r1 *= F.binv[1];
}

// SciPy solve_lsq_trust_region in the Q basis. In: s.d, s.e, s.gt (= Q^T g), delta, alpha (warm start), full_rank.
// Out: s.pt (step in the Q basis, already rescaled), returns alpha; sc[1] = ||p||, sc[2] = predicted reduction.
// Every lane computes the same scalars (butterfly reductions), so the control flow is warp uniform.
template <class TW>
__device__ __noinline__ double trf_subproblem(TW& s, int n, double delta, double alpha, bool full_rank) {
    MVMC_ASSUME_SHARED(&s);
    const int lane = threadIdx.x & 31;
    const int i0 = lane, i1 = lane + 32;
    const double g0 = i0 < n ? s.gt[i0] : 0.0, g1 = i1 < n ? s.gt[i1] : 0.0;
    const double gn2 = warp_sum(fma(g1, g1, g0 * g0));
    const double dmax = warp_max(fmax(i0 < n ? fabs(s.d[i0]) : 0.0, i1 < n ? fabs(s.d[i1]) : 0.0));
    const double floor_ = kEps * kEps * fmax(dmax, 1e-300);  // only guards against non-positive pivots
    // SciPy works from singular values, so exactly rank-deficient directions (s = 0, s*uf = 0) drop out even when the
    // Moré iteration ends at alpha ~ 0 (e.g. the 2-view triangulation refine, where phi(alpha) < 0 for every alpha and
    // the pseudo-inverse step gets stretched to the radius). J^T J + alpha I has no such luxury: keep the linear algebra
    // numerically positive definite with a floor on alpha far below anything the data resolves (16 eps lambda_max),
    // which turns the alpha -> 0 limit into the same pseudo-inverse step. The iteration on alpha itself is unchanged.
    const double amin = 16.0 * kEps * dmax;
    PcrFactors F;
    double y0 = 0.0, y1 = 0.0;
    double a_lo = 0.0, a_hi = sqrt(gn2) / delta;
    // One loop, one call site of the solver (code size): phase 0 = Gauss-Newton probe at alpha = 0 (full rank only),
    // phase 1 = SciPy's <= 10 Newton iterations on alpha, phase 2 = the step at the final alpha.
    int phase = full_rank ? 0 : 1, it = 0;
    bool gn_step = false;
    if (phase == 1 && alpha == 0.0) alpha = fmax(0.001 * a_hi, sqrt(a_lo * a_hi));
#pragma unroll 1
    for (;;) {
        if (phase == 1 && (alpha < a_lo || alpha > a_hi)) alpha = fmax(0.001 * a_hi, sqrt(a_lo * a_hi));
        const double a_eval = phase == 0 ? amin : fmax(alpha, amin);
        y0 = g0;
        y1 = g1;
        pcr_solve(s, n, a_eval, floor_, y0, y1, F);
        if (phase == 2) break;
        const double pn = sqrt(warp_sum(fma(y1, y1, y0 * y0)));
        if (phase == 0) {
            if (F.pd && pn <= delta) {
                alpha = 0.0;
                gn_step = true;  // Gauss-Newton step
                break;
            }
            if (!F.pd) {         // numerically rank deficient after all
                if (alpha == 0.0) alpha = fmax(0.001 * a_hi, sqrt(a_lo * a_hi));
                phase = 1;
                continue;
            }
        }
        double z0 = y0, z1 = y1;
        pcr_resolve(F, z0, z1);
        const double q = warp_sum(fma(y1, z1, y0 * z0));
        const double phi = pn - delta, dphi = -q / pn;
        if (phase == 0) {
            a_lo = -phi / dphi;
            phase = 1;
            continue;
        }
        if (phi < 0.0) a_hi = alpha;
        const double ratio = phi / dphi;
        a_lo = fmax(a_lo, alpha - ratio);
        alpha -= (phi + delta) * ratio / delta;
        it++;
        if (fabs(phi) < 0.01 * delta || it == 10) phase = 2;
    }
    const double scale_to = gn_step ? 0.0 : delta;
    // p~ = -y (rescaled to the radius unless it is the Gauss-Newton step)
    double nn = warp_sum(fma(y1, y1, y0 * y0));
    double sc = -1.0;
    if (scale_to > 0.0) sc = -scale_to / sqrt(nn);
    const double p0 = y0 * sc, p1 = y1 * sc;
    nn = warp_sum(fma(p1, p1, p0 * p0));
    __syncwarp();
    if (i0 < n) s.pt[i0] = p0;
    if (i1 < n) s.pt[i1] = p1;
    __syncwarp();
    // predicted reduction -(0.5 p^T A p + g^T p) = -(0.5 p~^T T p~ + g~^T p~)
    double tp = 0.0, gp = 0.0;
    for (int i = lane; i < n; i += 32) {
        const double pi = s.pt[i];
        double t = s.d[i] * pi;
        if (i > 0) t = fma(s.e[i - 1], s.pt[i - 1], t);
        if (i + 1 < n) t = fma(s.e[i], s.pt[i + 1], t);
        tp = fma(pi, t, tp);
        gp = fma(s.gt[i], pi, gp);
    }
    tp = warp_sum(tp);
    gp = warp_sum(gp);
    if (lane == 0) {
        s.sc[0] = alpha;
        s.sc[1] = sqrt(nn);
        s.sc[2] = -(0.5 * tp + gp);
    }
    __syncwarp();
    return alpha;
}

// scipy.optimize.least_squares(fun, x0, max_nfev=...), method='trf', jac='2-point', unbounded.
// s.x holds the full parameter vector; s.act[0..ncol) the optimised parameters that can move a residual
// ("live" columns); n_opt = number of optimised parameters including structurally dead ones (they only enter
// SciPy's norms of x and its m >= n rank test); x2_dead = sum of squares of the dead optimised parameters.
template <class TW, class Res>
__device__ TrfResult trf_solve_warp(TW& s, Res& res, int ncol, int n_opt, double x2_dead, bool has_dead, int max_nfev,
                                    double* f, double* fn) {
    const int lane = threadIdx.x & 31;
    const int m = res.m();
    const double ftol = 1e-8, xtol = 1e-8, gtol = 1e-8;
    res.eval(s.x, f);
    __syncwarp();
    double part = 0.0;
    for (int r = lane; r < m; r += 32) part = fma(f[r], f[r], part);
    double cost = 0.5 * warp_sum(part);
    int nfev = 1, njev = 1;
    trf_jacobian(s, res, ncol, f);
    part = 0.0;
    for (int c = lane; c < ncol; c += 32) part = fma(s.x[s.act[c]], s.x[s.act[c]], part);
    double delta = sqrt(warp_sum(part) + x2_dead);
    if (delta == 0.0) delta = 1.0;
    double alpha = 0.0;
    int status = -1;
    while (true) {
        double gn = 0.0;
        for (int c = lane; c < ncol; c += 32) gn = fmax(gn, fabs(s.g[c]));
        gn = warp_max(gn);
        if (gn < gtol) status = 1;
        if (status != -1 || nfev == max_nfev) break;
        // zero columns of J?  (exactly zero diagonal of J^T J)
        int zc = 0;
        for (int c = lane; c < ncol; c += 32) zc |= (s.A[c * TW::LDA + c] == 0.0) ? 1 : 0;
        zc = warp_sum_i(zc);
        const bool full_rank = (m >= n_opt) && !has_dead && zc == 0;
        for (int c = lane; c < ncol; c += 32) s.gt[c] = s.g[c];
        __syncwarp();
        warp_tridiagonalise(s, ncol);
        warp_apply_q(s, ncol, s.gt, true);
        double actual = -1.0, cost_new = cost;
        while (actual <= 0.0 && nfev < max_nfev) {
            alpha = trf_subproblem(s, ncol, delta, alpha, full_rank);
            const double p_norm = s.sc[1], predicted = s.sc[2];
            for (int c = lane; c < ncol; c += 32) s.p[c] = s.pt[c];
            __syncwarp();
            warp_apply_q(s, ncol, s.p, false);
            for (int i = lane; i < MVMC_N_PARAM; i += 32) s.xn[i] = s.x[i];
            __syncwarp();
            for (int c = lane; c < ncol; c += 32) s.xn[s.act[c]] = s.x[s.act[c]] + s.p[c];
            __syncwarp();
            res.eval(s.xn, fn);
            __syncwarp();
            nfev++;
            part = 0.0;
            for (int r = lane; r < m; r += 32) part = fma(fn[r], fn[r], part);
            cost_new = 0.5 * warp_sum(part);
            if (!(cost_new <= kDblMax)) {  // a non-finite residual poisons the sum (NaN or inf)
                delta = 0.25 * p_norm;
                continue;
            }
            actual = cost - cost_new;
            double ratio;
            if (predicted > 0.0) ratio = actual / predicted;
            else if (predicted == 0.0 && actual == 0.0) ratio = 1.0;
            else ratio = 0.0;
            double delta_new = delta;
            if (ratio < 0.25) delta_new = 0.25 * p_norm;
            else if (ratio > 0.75 && p_norm > 0.95 * delta) delta_new = delta * 2.0;
            part = 0.0;
            for (int c = lane; c < ncol; c += 32) part = fma(s.x[s.act[c]], s.x[s.act[c]], part);
            const double x_norm = sqrt(warp_sum(part) + x2_dead);
            const bool f_ok = actual < ftol * cost && ratio > 0.25;
            const bool x_ok = p_norm < xtol * (xtol + x_norm);
            if (f_ok && x_ok) status = 4;
            else if (f_ok) status = 2;
            else if (x_ok) status = 3;
            if (status != -1) break;
            alpha *= delta / delta_new;
            delta = delta_new;
        }
        if (actual > 0.0) {
            for (int i = lane; i < MVMC_N_PARAM; i += 32) s.x[i] = s.xn[i];
            for (int r = lane; r < m; r += 32) f[r] = fn[r];
            __syncwarp();
            cost = cost_new;
            // SciPy re-evaluates J here even when the loop is about to stop; x, cost and the counters do not depend on it
            if (status == -1 && nfev < max_nfev) trf_jacobian(s, res, ncol, f);
            njev++;
        }
    }
    if (status == -1) status = 0;
    TrfResult out;
    out.nfev = nfev;
    out.njev = njev;
    out.status = status;
    out.cost = cost;
    return out;
}

}  // namespace mvmc
