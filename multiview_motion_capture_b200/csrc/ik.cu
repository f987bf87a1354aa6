// Forward kinematics, DLT triangulation and the reprojection IK (trust-region-reflective least squares
// with a forward-difference Jacobian) as sm_100a CUDA kernels.
//
// Reference rows: B1/B2 mv_math_util.py:152-240; I0 inverse_kinematics.py:339-348;
// I1 inverse_kinematics.py:176-199 + Quaternions.py:97-115,335-366,442-462; I2/I3 :202-277;
// I4 scipy.optimize.least_squares(method='trf', jac='2-point', tr_solver='exact') restated on device
// (SURVEY.md §3.3); I5 :380-433; I6 kinematics.py:18-31.
//
// k_ik_solve: ONE CTA PER SOLVE. Lanes own Jacobian columns (one perturbed FK + reprojection each);
// J (n x m) lives in an L2-resident global scratch row per parameter; J J^T (n x n, n <= 68) and its
// eigenvectors live in shared memory and are diagonalised by a parallel cyclic Jacobi method
// (round-robin ordering, one 2x2 block per thread). The trust-region sub-problem is solved from the
// eigen-decomposition exactly as SciPy does from the SVD (s^2 = lambda, s*uf = V^T g).
#include "mvmc_common.cuh"

namespace mvmc {

// ---- BASIC_18 skeleton (inverse_kinematics.py:120-173) ----
struct SkelConst {
    int parents[MVMC_N_B18];
    int side_to_full[MVMC_N_B18];
    double dirs[MVMC_N_B18][3];
    double ref_side_lens[11];
};
__constant__ SkelConst c_skel;
__constant__ int c_ik_skel_idx[MVMC_N_IKJ] = {1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17};
__constant__ int c_ik_obs_idx[MVMC_N_IKJ] = {11, 13, 15, 12, 14, 16, 17, 5, 7, 9, 6, 8, 10, 0, 3, 4};

#ifdef MVMC_EMU
#define DMUL(a, b) ((a) * (b))
#define DSUB(a, b) ((a) - (b))
#define DDIV(a, b) ((a) / (b))
#else
#define DMUL(a, b) __dmul_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))
#define DDIV(a, b) __ddiv_rn((a), (b))
#endif

constexpr double kSqrtEps = 1.4901161193847656e-08;   // sqrt(2^-52)
constexpr double kEps = 2.220446049250313e-16;

// local rotation R = Rx(a) Ry(b) Rz(c) via half-angle quaternions, as Quaternions.from_euler + transforms
__device__ __forceinline__ void euler_to_mat(double ea, double eb, double ec, double* m) {
    const double k = 1.0 / (1.0 + 1e-10);  // axis / (|axis| + 1e-10), Quaternions.py:444
    double sx, cx, sy, cy, sz, cz;
    sincos(ea / 2.0, &sx, &cx);
    sincos(eb / 2.0, &sy, &cy);
    sincos(ec / 2.0, &sz, &cz);
    sx *= k;
    sy *= k;
    sz *= k;
    // t = qy * qz, q = qx * t (reference operand order)
    const double t0 = cz * cy, t1 = sz * sy, t2 = cz * sy, t3 = sz * cy;
    const double qw = t0 * cx - t1 * sx;
    const double qx = t0 * sx + t1 * cx;
    const double qy = t2 * cx - t3 * sx;
    const double qz = t2 * sx + t3 * cx;
    const double x2 = qx + qx, y2 = qy + qy, z2 = qz + qz;
    const double xx = qx * x2, yy = qy * y2, wx = qw * x2, xy = qx * y2, yz = qy * z2, wy = qw * y2, xz = qx * z2,
                 zz = qz * z2, wz = qw * z2;
    m[0] = 1.0 - (yy + zz);
    m[1] = xy - wz;
    m[2] = xz + wy;
    m[3] = xy + wz;
    m[4] = 1.0 - (xx + zz);
    m[5] = yz - wx;
    m[6] = xz - wy;
    m[7] = yz + wx;
    m[8] = 1.0 - (xx + yy);
}

// x = [root(3) | euler(54)], lens = 11 side lengths -> pos[18][3]
__device__ void fk_b18(const double* x, const double* lens, double (*pos)[3]) {
    double R[MVMC_N_B18][9];
    for (int j = 0; j < MVMC_N_B18; j++) {
        double m[9];
        euler_to_mat(x[3 + 3 * j], x[3 + 3 * j + 1], x[3 + 3 * j + 2], m);
        if (j == 0) {
            for (int q = 0; q < 9; q++) R[0][q] = m[q];
            pos[0][0] = x[0];
            pos[0][1] = x[1];
            pos[0][2] = x[2];
        } else {
            const int p = c_skel.parents[j];
            const double len = lens[c_skel.side_to_full[j]];
            const double o0 = c_skel.dirs[j][0] * len, o1 = c_skel.dirs[j][1] * len, o2 = c_skel.dirs[j][2] * len;
            const double* Rp = R[p];
            for (int r = 0; r < 3; r++) {
                for (int c = 0; c < 3; c++)
                    R[j][r * 3 + c] = Rp[r * 3] * m[c] + Rp[r * 3 + 1] * m[3 + c] + Rp[r * 3 + 2] * m[6 + c];
                pos[j][r] = Rp[r * 3] * o0 + Rp[r * 3 + 1] * o1 + Rp[r * 3 + 2] * o2 + pos[p][r];
            }
        }
    }
}

// ---- residual functors. eval(x, emit) calls emit(row, value) for every residual row. ----
struct IkResidual {
    const double* obs;   // [V][16][3] gathered at c_ik_obs_idx (shared)
    const double* P;     // [V][12] (shared)
    const double* lens;  // fixed side lengths when with_lens == 0 (shared)
    int V;
    int with_lens;       // 1: x[57..67] are the lengths
    __device__ int m() const { return V * MVMC_N_IKJ * 2; }
    template <class Emit>
    __device__ void eval(const double* x, Emit emit) const {
        double pos[MVMC_N_B18][3];
        fk_b18(x, with_lens ? x + 57 : lens, pos);
        for (int v = 0; v < V; v++) {
            const double* Pv = P + v * 12;
            for (int q = 0; q < MVMC_N_IKJ; q++) {
                const double* X = pos[c_ik_skel_idx[q]];
                const double* o = obs + (v * MVMC_N_IKJ + q) * 3;
                const double pu = Pv[0] * X[0] + Pv[1] * X[1] + Pv[2] * X[2] + Pv[3];
                const double pv = Pv[4] * X[0] + Pv[5] * X[1] + Pv[6] * X[2] + Pv[7];
                const double pw = Pv[8] * X[0] + Pv[9] * X[1] + Pv[10] * X[2] + Pv[11];
                const double den = 1e-5 + pw;
                emit((v * MVMC_N_IKJ + q) * 2, DMUL(DSUB(DDIV(pu, den), o[0]), o[2]));
                emit((v * MVMC_N_IKJ + q) * 2 + 1, DMUL(DSUB(DDIV(pv, den), o[1]), o[2]));
            }
        }
    }
};

struct TriResidual {  // mv_math_util.py:190-202
    const double* obs;  // [V][K][3] (shared)
    const double* P;    // [V][12]
    int V, K;
    __device__ int m() const { return V * K; }
    template <class Emit>
    __device__ void eval(const double* x, Emit emit) const {
        for (int v = 0; v < V; v++) {
            const double* Pv = P + v * 12;
            for (int k = 0; k < K; k++) {
                const double* X = x + 3 * k;
                const double* o = obs + (v * K + k) * 3;
                const double pu = Pv[0] * X[0] + Pv[1] * X[1] + Pv[2] * X[2] + Pv[3];
                const double pv = Pv[4] * X[0] + Pv[5] * X[1] + Pv[6] * X[2] + Pv[7];
                const double pw = Pv[8] * X[0] + Pv[9] * X[1] + Pv[10] * X[2] + Pv[11];
                const double den = pw + 1e-6;
                const double du = DSUB(DDIV(pu, den), o[0]), dv = DSUB(DDIV(pv, den), o[1]);
                emit(v * K + k, DMUL(sqrt(DMUL(du, du) + DMUL(dv, dv)), o[2]));
            }
        }
    }
};

// ---- shared-memory plan of one solver CTA ----
constexpr int TRF_NMAX = 68;
constexpr int TRF_LD = 69;
constexpr int TRF_MMAX = 512;  // 32 residuals per 2D pose x MVMC_MAX_SEL
constexpr int TRF_CH = 32;  // rows of J per chunk when forming J J^T

struct TrfShared {
    double A[TRF_NMAX * TRF_LD];
    double Vm[TRF_NMAX * TRF_LD];   // first used as the J chunk buffer, then as eigenvectors
    double x[TRF_NMAX], xn[TRF_NMAX], p[TRF_NMAX], g[TRF_NMAX], lam[TRF_NMAX], suf[TRF_NMAX];
    double f[TRF_MMAX], fn[TRF_MMAX];
    double cs[TRF_NMAX / 2 + 1], sn[TRF_NMAX / 2 + 1];
    double scratch[32];
    double sc[8];  // scalars broadcast from warp 0: alpha, delta, pnorm, ...
    int pr[TRF_NMAX / 2 + 1], qr[TRF_NMAX / 2 + 1];
    int pos[TRF_NMAX + 2], pos2[TRF_NMAX + 2];
    int act[TRF_NMAX];
    int flag;
};

// Eigen-decomposition of the symmetric n x n matrix s.A: on exit lam[k] = eigenvalue, column k of Vm
// the eigenvector. Parallel cyclic Jacobi, round-robin pair ordering.
__device__ void jacobi_eig(TrfShared& s, int n) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int np = n + (n & 1);
    const int half = np / 2;
    for (int e = tid; e < n * n; e += nt) s.Vm[(e / n) * TRF_LD + (e % n)] = ((e / n) == (e % n)) ? 1.0 : 0.0;
    for (int k = tid; k < np; k += nt) s.pos[k] = k;
    __syncthreads();
    for (int sweep = 0; sweep < 16; sweep++) {
        if (tid == 0) s.flag = 0;
        __syncthreads();
        for (int round = 0; round < np - 1; round++) {
            for (int k = tid; k < half; k += nt) {
                int p = s.pos[k], q = s.pos[np - 1 - k];
                if (p > q) {
                    const int t = p;
                    p = q;
                    q = t;
                }
                double c = 1.0, sn = 0.0;
                if (q < n) {
                    const double app = s.A[p * TRF_LD + p], aqq = s.A[q * TRF_LD + q], apq = s.A[p * TRF_LD + q];
                    if (apq != 0.0 && fabs(apq) > kEps * sqrt(fabs(app) * fabs(aqq))) {
                        const double tau = (aqq - app) / (2.0 * apq);
                        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + t * t);
                        sn = t * c;
                        s.flag = 1;
                    }
                }
                s.pr[k] = p;
                s.qr[k] = q;
                s.cs[k] = c;
                s.sn[k] = sn;
            }
            __syncthreads();
            // A <- J^T A J, one 2x2 block per work item
            for (int e = tid; e < half * half; e += nt) {
                const int kk = e / half, ll = e % half;
                const int p1 = s.pr[kk], q1 = s.qr[kk], p2 = s.pr[ll], q2 = s.qr[ll];
                const double c1 = s.cs[kk], s1 = s.sn[kk], c2 = s.cs[ll], s2 = s.sn[ll];
                if (c1 == 1.0 && c2 == 1.0 && s1 == 0.0 && s2 == 0.0) continue;
                const bool v1 = q1 < n, v2 = q2 < n;  // a pair with the dummy index only has its p member
                const double m00 = s.A[p1 * TRF_LD + p2];
                const double m01 = v2 ? s.A[p1 * TRF_LD + q2] : 0.0;
                const double m10 = v1 ? s.A[q1 * TRF_LD + p2] : 0.0;
                const double m11 = (v1 && v2) ? s.A[q1 * TRF_LD + q2] : 0.0;
                const double t00 = c1 * m00 - s1 * m10, t01 = c1 * m01 - s1 * m11;
                const double t10 = s1 * m00 + c1 * m10, t11 = s1 * m01 + c1 * m11;
                double r00 = c2 * t00 - s2 * t01, r01 = s2 * t00 + c2 * t01;
                double r10 = c2 * t10 - s2 * t11, r11 = s2 * t10 + c2 * t11;
                if (kk == ll) {
                    r01 = 0.0;
                    r10 = 0.0;
                }
                s.A[p1 * TRF_LD + p2] = r00;
                if (v2) s.A[p1 * TRF_LD + q2] = r01;
                if (v1) s.A[q1 * TRF_LD + p2] = r10;
                if (v1 && v2) s.A[q1 * TRF_LD + q2] = r11;
            }
            // V <- V J
            for (int e = tid; e < n * half; e += nt) {
                const int i = e / half, kk = e % half;
                const int p = s.pr[kk], q = s.qr[kk];
                const double c = s.cs[kk], sn = s.sn[kk];
                if (q >= n || (c == 1.0 && sn == 0.0)) continue;
                const double vp = s.Vm[i * TRF_LD + p], vq = s.Vm[i * TRF_LD + q];
                s.Vm[i * TRF_LD + p] = c * vp - sn * vq;
                s.Vm[i * TRF_LD + q] = sn * vp + c * vq;
            }
            // next round: position 0 stays, the others rotate by one
            for (int k = tid; k < np; k += nt) s.pos2[k] = (k == 0) ? s.pos[0] : (k == 1 ? s.pos[np - 1] : s.pos[k - 1]);
            __syncthreads();
            for (int k = tid; k < np; k += nt) s.pos[k] = s.pos2[k];
            __syncthreads();
        }
        if (s.flag == 0) break;
        __syncthreads();
    }
    for (int k = tid; k < n; k += nt) s.lam[k] = s.A[k * TRF_LD + k];
    __syncthreads();
}

struct TrfResult {
    int nfev, njev, status;
    double cost;
};

// Jacobian by forward differences (SciPy's default 2-point rule), then g = J^T f and A = J^T J.
// Jg: global scratch, row k (parameter act[k]) holds the m residual derivatives.
template <class Res>
__device__ void trf_jacobian(TrfShared& s, const Res& res, int n, int m, int nfull, double* __restrict__ Jg) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid < n) {
        double xl[TRF_NMAX];
        for (int i = 0; i < nfull; i++) xl[i] = s.x[i];
        const int i = s.act[tid];
        const double xi = xl[i];
        const double h = kSqrtEps * (xi >= 0.0 ? 1.0 : -1.0) * fmax(1.0, fabs(xi));
        const double xp = xi + h;
        const double dx = DSUB(xp, xi);
        xl[i] = xp;
        double* row = Jg + (size_t)tid * TRF_MMAX;
        const double* f0 = s.f;
        res.eval(xl, [&](int r, double v) { row[r] = DDIV(DSUB(v, f0[r]), dx); });
    }
    __syncthreads();
    // g[k] = sum_r J[k][r] f[r]
    for (int k = tid; k < n; k += nt) {
        const double* row = Jg + (size_t)k * TRF_MMAX;
        double acc = 0.0;
        for (int r = 0; r < m; r++) acc += row[r] * s.f[r];
        s.g[k] = acc;
    }
    // A = J J^T over chunks of residual rows staged in shared memory (buffer aliases Vm)
    double* chunk = s.Vm;  // [n][TRF_CH+1]
    const int nn = n * n;
    double acc[(TRF_NMAX * TRF_NMAX + 127) / 128];  // outputs per thread for >= 128 threads
    const int per = (nn + nt - 1) / nt;
    for (int q = 0; q < per; q++) acc[q] = 0.0;
    for (int r0 = 0; r0 < m; r0 += TRF_CH) {
        __syncthreads();
        for (int e = tid; e < n * TRF_CH; e += nt) {
            const int k = e / TRF_CH, rr = e % TRF_CH;
            chunk[k * (TRF_CH + 1) + rr] = (r0 + rr < m) ? Jg[(size_t)k * TRF_MMAX + r0 + rr] : 0.0;
        }
        __syncthreads();
        for (int q = 0; q < per; q++) {
            const int e = tid + q * nt;
            if (e < nn) {
                const int k = e / n, l = e % n;
                double a = acc[q];
                for (int rr = 0; rr < TRF_CH; rr++) a += chunk[k * (TRF_CH + 1) + rr] * chunk[l * (TRF_CH + 1) + rr];
                acc[q] = a;
            }
        }
    }
    __syncthreads();
    for (int q = 0; q < per; q++) {
        const int e = tid + q * nt;
        if (e < nn) s.A[(e / n) * TRF_LD + (e % n)] = acc[q];
    }
    __syncthreads();
}

// scipy.optimize.least_squares(fun, x0, max_nfev=...) with method='trf', jac='2-point', no bounds.
// s.x holds the full parameter vector (nfull entries); s.act[0..n) the optimised entries.
template <class Res>
__device__ TrfResult trf_solve(TrfShared& s, const Res& res, int n, int nfull, int max_nfev, double* __restrict__ Jg) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int m = res.m();
    const double ftol = 1e-8, xtol = 1e-8, gtol = 1e-8;
    TrfResult out;
    if (tid == 0) res.eval(s.x, [&](int r, double v) { s.f[r] = v; });
    __syncthreads();
    double part = 0.0;
    for (int r = tid; r < m; r += nt) part += s.f[r] * s.f[r];
    double cost = 0.5 * block_sum(part, s.scratch);
    int nfev = 1, njev = 1;
    trf_jacobian(s, res, n, m, nfull, Jg);
    // Delta = ||x0|| over the optimised entries (1.0 if zero)
    part = 0.0;
    for (int k = tid; k < n; k += nt) part += s.x[s.act[k]] * s.x[s.act[k]];
    double delta = sqrt(block_sum(part, s.scratch));
    if (delta == 0.0) delta = 1.0;
    double alpha = 0.0;
    int status = -1;
    while (true) {
        double gn = 0.0;
        for (int k = 0; k < n; k++) gn = fmax(gn, fabs(s.g[k]));
        if (gn < gtol) status = 1;
        if (status != -1 || nfev == max_nfev) break;
        jacobi_eig(s, n);
        // suf = V^T g  (= s * U^T f)
        for (int k = tid; k < n; k += nt) {
            double acc = 0.0;
            for (int i = 0; i < n; i++) acc += s.Vm[i * TRF_LD + k] * s.g[i];
            s.suf[k] = acc;
            if (s.lam[k] < 0.0) s.lam[k] = 0.0;
        }
        __syncthreads();
        double lmax = 0.0, lmin = INFINITY;
        for (int k = 0; k < n; k++) {
            lmax = fmax(lmax, s.lam[k]);
            lmin = fmin(lmin, s.lam[k]);
        }
        const bool full_rank = (m >= n) && (sqrt(lmin) > kEps * m * sqrt(lmax));
        double actual = -1.0, cost_new = cost;
        while (actual <= 0.0 && nfev < max_nfev) {
            // ---- trust-region sub-problem (warp 0), result: s.p (free entries), alpha ----
            if (tid < 32) {
                double a_new = alpha;
                bool gn_step = false;
                double scale_to = 0.0;  // 0: no rescale
                if (full_rank) {
                    double pn2 = 0.0;
                    for (int k = lane; k < n; k += 32) {
                        const double w = s.suf[k] / s.lam[k];
                        pn2 += w * w;
                    }
                    pn2 = warp_sum(pn2);
                    if (sqrt(pn2) <= delta) {
                        gn_step = true;
                        a_new = 0.0;
                    }
                }
                if (!gn_step) {
                    double sn2 = 0.0;
                    for (int k = lane; k < n; k += 32) sn2 += s.suf[k] * s.suf[k];
                    sn2 = warp_sum(sn2);
                    double a_hi = sqrt(sn2) / delta, a_lo = 0.0;
                    if (full_rank) {
                        double q1 = 0.0, q3 = 0.0;
                        for (int k = lane; k < n; k += 32) {
                            const double den = s.lam[k];
                            const double w = s.suf[k] / den;
                            q1 += w * w;
                            q3 += s.suf[k] * s.suf[k] / (den * den * den);
                        }
                        q1 = warp_sum(q1);
                        q3 = warp_sum(q3);
                        const double pn = sqrt(q1);
                        a_lo = -(pn - delta) / (-q3 / pn);
                    }
                    double al = alpha;
                    if (!full_rank && al == 0.0) al = fmax(0.001 * a_hi, sqrt(a_lo * a_hi));
                    for (int itn = 0; itn < 10; itn++) {
                        if (al < a_lo || al > a_hi) al = fmax(0.001 * a_hi, sqrt(a_lo * a_hi));
                        double q1 = 0.0, q3 = 0.0;
                        for (int k = lane; k < n; k += 32) {
                            const double den = s.lam[k] + al;
                            const double w = s.suf[k] / den;
                            q1 += w * w;
                            q3 += s.suf[k] * s.suf[k] / (den * den * den);
                        }
                        q1 = warp_sum(q1);
                        q3 = warp_sum(q3);
                        const double pn = sqrt(q1);
                        const double phi = pn - delta, dphi = -q3 / pn;
                        if (phi < 0.0) a_hi = al;
                        const double ratio = phi / dphi;
                        a_lo = fmax(a_lo, al - ratio);
                        al -= (phi + delta) * ratio / delta;
                        if (fabs(phi) < 0.01 * delta) break;
                    }
                    a_new = al;
                    scale_to = delta;
                }
                // p = -V w, w = suf / (lam + alpha)   (w parked in s.fn, which is rewritten by the next eval)
                for (int k = lane; k < n; k += 32) s.fn[k] = s.suf[k] / (s.lam[k] + a_new);
                __syncwarp();
                double pn2 = 0.0;
                for (int i = lane; i < n; i += 32) {
                    double acc = 0.0;
                    for (int k = 0; k < n; k++) acc += s.Vm[i * TRF_LD + k] * s.fn[k];
                    s.p[i] = -acc;
                    pn2 += acc * acc;
                }
                pn2 = warp_sum(pn2);
                double sc = 1.0;
                if (scale_to > 0.0) {
                    sc = scale_to / sqrt(pn2);
                    double pn2b = 0.0;
                    for (int i = lane; i < n; i += 32) {
                        s.p[i] *= sc;
                        pn2b += s.p[i] * s.p[i];
                    }
                    pn2 = warp_sum(pn2b);
                }
                __syncwarp();
                // predicted reduction: -(0.5 |J p|^2 + g.p), |J p|^2 = sum lam_k (V^T p)_k^2, V^T p = -sc w
                double jp2 = 0.0, gp = 0.0;
                for (int k = lane; k < n; k += 32) {
                    const double w = s.fn[k] * sc;
                    jp2 += s.lam[k] * w * w;
                    gp += s.g[k] * s.p[k];
                }
                jp2 = warp_sum(jp2);
                gp = warp_sum(gp);
                if (lane == 0) {
                    s.sc[0] = a_new;
                    s.sc[1] = sqrt(pn2);
                    s.sc[2] = -(0.5 * jp2 + gp);
                }
            }
            __syncthreads();
            alpha = s.sc[0];
            const double p_norm = s.sc[1], predicted = s.sc[2];
            for (int i = tid; i < nfull; i += nt) s.xn[i] = s.x[i];
            __syncthreads();
            for (int k = tid; k < n; k += nt) s.xn[s.act[k]] = s.x[s.act[k]] + s.p[k];
            __syncthreads();
            if (tid == 0) res.eval(s.xn, [&](int r, double v) { s.fn[r] = v; });
            __syncthreads();
            nfev++;
            part = 0.0;
            int bad = 0;
            for (int r = tid; r < m; r += nt) {
                const double v = s.fn[r];
                part += v * v;
                if (!(fabs(v) <= 1.79769313486231570e308)) bad = 1;
            }
            cost_new = 0.5 * block_sum(part, s.scratch);
            if (!(cost_new <= 1.79769313486231570e308)) bad = 1;  // any non-finite residual poisons the sum
            (void)bad;
            if (!(cost_new == cost_new) || cost_new > 1.79769313486231570e308) {
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
            for (int k = tid; k < n; k += nt) part += s.x[s.act[k]] * s.x[s.act[k]];
            const double x_norm = sqrt(block_sum(part, s.scratch));
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
            __syncthreads();
            for (int i = tid; i < nfull; i += nt) s.x[i] = s.xn[i];
            for (int r = tid; r < m; r += nt) s.f[r] = s.fn[r];
            __syncthreads();
            cost = cost_new;
            trf_jacobian(s, res, n, m, nfull, Jg);
            njev++;
        }
    }
    if (status == -1) status = 0;
    out.nfev = nfev;
    out.njev = njev;
    out.status = status;
    out.cost = cost;
    return out;
}

// ---- DLT: null vector of the (2V x 4) system by one-sided Jacobi (single thread, tiny) ----
__device__ void dlt_point(const double* P /*[V][12]*/, const double* xy /*[V][2]*/, const int* sel, int nsel, double* out3) {
    double a[2 * MVMC_MAX_SEL][4];
    double v[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    const int rows = 2 * nsel;
    for (int q = 0; q < nsel; q++) {
        const double* Pv = P + sel[q] * 12;
        const double px = xy[sel[q] * 2], py = xy[sel[q] * 2 + 1];
        for (int c = 0; c < 4; c++) {
            a[2 * q][c] = px * Pv[8 + c] - Pv[c];
            a[2 * q + 1][c] = py * Pv[8 + c] - Pv[4 + c];
        }
    }
    for (int sweep = 0; sweep < 30; sweep++) {
        bool rotated = false;
        for (int p = 0; p < 3; p++)
            for (int q = p + 1; q < 4; q++) {
                double app = 0, aqq = 0, apq = 0;
                for (int r = 0; r < rows; r++) {
                    app += a[r][p] * a[r][p];
                    aqq += a[r][q] * a[r][q];
                    apq += a[r][p] * a[r][q];
                }
                if (apq == 0.0 || fabs(apq) <= kEps * sqrt(app * aqq)) continue;
                rotated = true;
                const double tau = (aqq - app) / (2.0 * apq);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = t * c;
                for (int r = 0; r < rows; r++) {
                    const double ap = a[r][p], aq = a[r][q];
                    a[r][p] = c * ap - sn * aq;
                    a[r][q] = sn * ap + c * aq;
                }
                for (int r = 0; r < 4; r++) {
                    const double vp = v[r][p], vq = v[r][q];
                    v[r][p] = c * vp - sn * vq;
                    v[r][q] = sn * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    int best = 0;
    double bn = INFINITY;
    for (int c = 0; c < 4; c++) {
        double nn = 0;
        for (int r = 0; r < rows; r++) nn += a[r][c] * a[r][c];
        if (nn < bn) {
            bn = nn;
            best = c;
        }
    }
    out3[0] = v[0][best] / v[3][best];
    out3[1] = v[1][best] / v[3][best];
    out3[2] = v[2][best] / v[3][best];
}

// obs [V][K][3] -> out [K][4]; one thread per joint. mv_math_util.py:152-186
__device__ void triangulate_joint(const double* obs, const double* P, int V, int K, int k, double min_score, double* out4) {
    int sel[MVMC_MAX_SEL];
    double xy[MVMC_MAX_SEL * 2];
    int nsel = 0;
    for (int v = 0; v < V; v++) {
        xy[2 * v] = obs[(v * K + k) * 3];
        xy[2 * v + 1] = obs[(v * K + k) * 3 + 1];
        if (obs[(v * K + k) * 3 + 2] >= min_score) sel[nsel++] = v;
    }
    if (nsel < 2) {
        nsel = V;
        for (int v = 0; v < V; v++) sel[v] = v;
    }
    double sc = 0.0;
    for (int q = 0; q < nsel; q++) sc += obs[(sel[q] * K + k) * 3 + 2];
    dlt_point(P, xy, sel, nsel, out4);
    out4[3] = sc / nsel;
}

// COCO pose (17,3) -> row 17 = synthetic mid-spine, inverse_kinematics.py:339-348
__device__ __forceinline__ void mid_spine(const double* k, double* o) {
    const double* ls = k + 3 * kCocoLShoulder;
    const double* rs = k + 3 * kCocoRShoulder;
    const double* lh = k + 3 * kCocoLHip;
    const double* rh = k + 3 * kCocoRHip;
    for (int c = 0; c < 2; c++) o[c] = 0.5 * (0.5 * (ls[c] + rs[c]) + 0.5 * (lh[c] + rh[c]));
    double sc = ls[2] * rs[2];
    sc *= lh[2] * rh[2];
    o[2] = sc;
}

constexpr int IK_THREADS = 256;

struct IkShared {
    TrfShared t;
    double obs18[MVMC_MAX_SEL * 18 * 3];
    double obs16[MVMC_MAX_SEL * MVMC_N_IKJ * 3];
    double P[MVMC_MAX_SEL * 12];
    double lens[11];
    double p3[18 * 4];
};

__global__ void __launch_bounds__(IK_THREADS)
    k_ik_solve(const double* __restrict__ kps2d, const double* __restrict__ Psel, const int* __restrict__ n_views,
               const double* __restrict__ x0, const uint8_t* __restrict__ birth, const int* __restrict__ max_nfev,
               const uint8_t* __restrict__ free_mask, int M, int V, double* __restrict__ ws, double* __restrict__ x_out,
               double* __restrict__ joints, int* __restrict__ info, double* __restrict__ cost_out) {
    MVMC_DYN_SMEM(IkShared, shp);
    IkShared& sh = *shp;
    const int tid = threadIdx.x, nt = blockDim.x;
    double* Jg = ws + (size_t)blockIdx.x * TRF_NMAX * TRF_MMAX;  // one Jacobian scratch per resident CTA
  for (int mI = blockIdx.x; mI < M; mI += gridDim.x) {  // persistent CTAs stride over the work slots
    const int nv = n_views[mI];
    if (nv < 2) continue;  // the reference never solves from fewer than two views (motion_capture.py:926,942)
    __syncthreads();
    const bool is_birth = birth != nullptr && birth[mI] != 0;
    const int nfev_cap = max_nfev[mI];
    // stage observations (+ mid spine) and projection matrices
    for (int e = tid; e < nv * MVMC_N_COCO * 3; e += nt) {
        const int v = e / (MVMC_N_COCO * 3), q = e % (MVMC_N_COCO * 3);
        sh.obs18[v * 54 + q] = kps2d[((size_t)mI * V + v) * (MVMC_N_COCO * 3) + q];
    }
    for (int e = tid; e < nv * 12; e += nt) sh.P[e] = Psel[(size_t)mI * V * 12 + e];
    __syncthreads();
    if (tid < nv) mid_spine(sh.obs18 + tid * 54, sh.obs18 + tid * 54 + 51);
    __syncthreads();
    for (int e = tid; e < nv * MVMC_N_IKJ * 3; e += nt) {
        const int v = e / (MVMC_N_IKJ * 3), q = (e / 3) % MVMC_N_IKJ, c = e % 3;
        sh.obs16[e] = sh.obs18[v * 54 + c_ik_obs_idx[q] * 3 + c];
    }
    __syncthreads();
    if (is_birth) {
        // triangulate 18 joints (min score 0.01), refine with a 2-nfev TRF, inverse_kinematics.py:389-396
        if (tid < 18) triangulate_joint(sh.obs18, sh.P, nv, 18, tid, 0.01, sh.p3 + tid * 4);
        __syncthreads();
        for (int e = tid; e < 54; e += nt) {
            sh.t.x[e] = sh.p3[(e / 3) * 4 + (e % 3)];
            sh.t.act[e] = e;
        }
        __syncthreads();
        TriResidual tr{sh.obs18, sh.P, nv, 18};
        trf_solve(sh.t, tr, 54, 54, 2, Jg);
        __syncthreads();
        if (tid == 0) {
            double root[3];
            for (int c = 0; c < 3; c++) root[c] = 0.5 * (sh.t.x[kCocoLHip * 3 + c] + sh.t.x[kCocoRHip * 3 + c]);
            for (int c = 0; c < 3; c++) sh.t.x[c] = root[c];
            for (int e = 3; e < 57; e++) sh.t.x[e] = 0.0;
            for (int e = 0; e < 11; e++) sh.t.x[57 + e] = c_skel.ref_side_lens[e];
        }
    } else {
        for (int e = tid; e < MVMC_N_PARAM; e += nt) sh.t.x[e] = x0[(size_t)mI * MVMC_N_PARAM + e];
    }
    __syncthreads();
    // ---- solve 1: root + angles, lengths fixed (solve_pose_reproj) ----
    if (tid < 11) sh.lens[tid] = sh.t.x[57 + tid];
    if (tid == 0) {
        int n = 0;
        for (int e = 0; e < 57; e++)
            if (!free_mask || free_mask[e]) sh.t.act[n++] = e;
        sh.t.flag = n;
    }
    __syncthreads();
    int n1 = sh.t.flag;
    __syncthreads();
    IkResidual r1{sh.obs16, sh.P, sh.lens, nv, 0};
    TrfResult a = {1, 1, 0, 0.0};
    if (n1 > 0) a = trf_solve(sh.t, r1, n1, 57, nfev_cap, Jg);
    __syncthreads();
    // ---- solve 2: root + angles + lengths (solve_pose_bone_lens_reproj) ----
    if (tid == 0) {
        int n = 0;
        for (int e = 0; e < MVMC_N_PARAM; e++)
            if (!free_mask || free_mask[e]) sh.t.act[n++] = e;
        sh.t.flag = n;
    }
    __syncthreads();
    int n2 = sh.t.flag;
    __syncthreads();
    IkResidual r2{sh.obs16, sh.P, sh.lens, nv, 1};
    TrfResult bres = {1, 1, 0, 0.0};
    if (n2 > 0) bres = trf_solve(sh.t, r2, n2, MVMC_N_PARAM, nfev_cap, Jg);
    __syncthreads();
    for (int e = tid; e < MVMC_N_PARAM; e += nt) x_out[(size_t)mI * MVMC_N_PARAM + e] = sh.t.x[e];
    if (tid == 0) {
        double pos[MVMC_N_B18][3];
        fk_b18(sh.t.x, sh.t.x + 57, pos);
        for (int j = 0; j < MVMC_N_B18; j++)
            for (int c = 0; c < 3; c++) joints[((size_t)mI * MVMC_N_B18 + j) * 3 + c] = pos[j][c];
        int* inf = info + (size_t)mI * 8;
        inf[0] = a.nfev;
        inf[1] = a.njev;
        inf[2] = a.status;
        inf[3] = n1;
        inf[4] = bres.nfev;
        inf[5] = bres.njev;
        inf[6] = bres.status;
        inf[7] = n2;
        cost_out[(size_t)mI * 2] = a.cost;
        cost_out[(size_t)mI * 2 + 1] = bres.cost;
    }
  }
}

struct TriShared {
    TrfShared t;
    double obs[MVMC_MAX_SEL * 18 * 3];
    double P[MVMC_MAX_SEL * 12];
    double p3[18 * 4];
};

__global__ void __launch_bounds__(IK_THREADS)
    k_triangulate(const double* __restrict__ obs, const double* __restrict__ Psel, const int* __restrict__ n_views, int M,
                  int V, int K, double min_score, int refine_nfev, double* __restrict__ ws, double* __restrict__ out) {
    MVMC_DYN_SMEM(TriShared, shp);
    TriShared& sh = *shp;
    const int tid = threadIdx.x, nt = blockDim.x;
  for (int mI = blockIdx.x; mI < M; mI += gridDim.x) {
    const int nv = n_views[mI];
    if (nv < 1) continue;
    __syncthreads();
    for (int e = tid; e < nv * K * 3; e += nt) sh.obs[e] = obs[(size_t)mI * V * K * 3 + e];
    for (int e = tid; e < nv * 12; e += nt) sh.P[e] = Psel[(size_t)mI * V * 12 + e];
    __syncthreads();
    if (tid < K) triangulate_joint(sh.obs, sh.P, nv, K, tid, min_score, sh.p3 + tid * 4);
    __syncthreads();
    if (refine_nfev > 0) {
        for (int e = tid; e < 3 * K; e += nt) {
            sh.t.x[e] = sh.p3[(e / 3) * 4 + (e % 3)];
            sh.t.act[e] = e;
        }
        __syncthreads();
        TriResidual tr{sh.obs, sh.P, nv, K};
        trf_solve(sh.t, tr, 3 * K, 3 * K, refine_nfev, ws + (size_t)blockIdx.x * TRF_NMAX * TRF_MMAX);
        __syncthreads();
        for (int e = tid; e < 3 * K; e += nt) sh.p3[(e / 3) * 4 + (e % 3)] = sh.t.x[e];
        __syncthreads();
    }
    for (int e = tid; e < K * 4; e += nt) out[(size_t)mI * K * 4 + e] = sh.p3[e];
  }
}

__global__ void k_fk(const double* __restrict__ params, int M, double* __restrict__ joints) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    double x[MVMC_N_PARAM];
    for (int e = 0; e < MVMC_N_PARAM; e++) x[e] = params[(size_t)i * MVMC_N_PARAM + e];
    double pos[MVMC_N_B18][3];
    fk_b18(x, x + 57, pos);
    for (int j = 0; j < MVMC_N_B18; j++)
        for (int c = 0; c < 3; c++) joints[((size_t)i * MVMC_N_B18 + j) * 3 + c] = pos[j][c];
}

// generic chain: one warp per instance, lane = joint (two joints per lane when J > 32); the chain is
// resolved level by level through shared memory.
__global__ void __launch_bounds__(64)
    k_fk_chain(const double* __restrict__ rot, const double* __restrict__ offsets, const int* __restrict__ parents,
               const double* __restrict__ root, int M, int J, double* __restrict__ joints) {
    __shared__ double sR[64][9];
    __shared__ double sT[64][3];
    __shared__ int sDone[64];
    const int i = blockIdx.x, j = threadIdx.x;
    if (j < J) sDone[j] = 0;
    __syncthreads();
    for (int level = 0; level < J; level++) {
        bool work = false;
        if (j < J && !sDone[j]) {
            const int p = parents[j];
            if (p < 0 || sDone[p]) work = true;
        }
        __syncthreads();
        if (work) {
            const int p = parents[j];
            const double* m = rot + ((size_t)i * J + j) * 9;
            if (p < 0) {
                for (int q = 0; q < 9; q++) sR[j][q] = m[q];
                for (int c = 0; c < 3; c++) sT[j][c] = root ? root[(size_t)i * 3 + c] : offsets[j * 3 + c];
            } else {
                for (int r = 0; r < 3; r++) {
                    for (int c = 0; c < 3; c++)
                        sR[j][r * 3 + c] = sR[p][r * 3] * m[c] + sR[p][r * 3 + 1] * m[3 + c] + sR[p][r * 3 + 2] * m[6 + c];
                    sT[j][r] = sR[p][r * 3] * offsets[j * 3] + sR[p][r * 3 + 1] * offsets[j * 3 + 1] +
                               sR[p][r * 3 + 2] * offsets[j * 3 + 2] + sT[p][r];
                }
            }
        }
        __syncthreads();
        if (work) sDone[j] = 1;
        __syncthreads();
        int all = 1;
        for (int q = 0; q < J; q++) all &= sDone[q];
        if (all) break;
    }
    if (j < J)
        for (int c = 0; c < 3; c++) joints[((size_t)i * J + j) * 3 + c] = sT[j][c];
}

}  // namespace mvmc

using namespace mvmc;

static bool g_skel_ready = false;

static int ensure_skeleton() {
    if (g_skel_ready) return MVMC_OK;
    static const double off[MVMC_N_B18][3] = {
        {0, 0, 0},     {0.15, 0, 0}, {0, 0, -0.5}, {0, 0, -0.5},  {-0.15, 0, 0}, {0, 0, -0.5},
        {0, 0, -0.5},  {0, 0, 0.3},  {0, 0, 0.3},  {0.2, 0, 0},   {0.3, 0, 0},   {0.3, 0, 0},
        {-0.2, 0, 0},  {-0.3, 0, 0}, {-0.3, 0, 0}, {0, -0.02, 0.15}, {0.07, 0.02, 0.1}, {-0.07, 0.02, 0.1}};
    static const int parents[MVMC_N_B18] = {-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 10, 8, 12, 13, 8, 15, 15};
    static const int s2f[MVMC_N_B18] = {7, 0, 1, 2, 0, 1, 2, 8, 9, 3, 4, 5, 3, 4, 5, 10, 6, 6};
    static const int side_src[11] = {1, 2, 3, 9, 10, 11, 16, 0, 7, 8, 15};
    SkelConst h;
    double lens[MVMC_N_B18];
    for (int j = 0; j < MVMC_N_B18; j++) {
        h.parents[j] = parents[j];
        h.side_to_full[j] = s2f[j];
        volatile double a = off[j][0] * off[j][0];
        volatile double b = off[j][1] * off[j][1];
        volatile double c = off[j][2] * off[j][2];
        volatile double s = a + b;
        s = s + c;
        lens[j] = sqrt(s);
        for (int q = 0; q < 3; q++) h.dirs[j][q] = (j == 0) ? off[j][q] : off[j][q] / lens[j];
    }
    for (int e = 0; e < 11; e++) h.ref_side_lens[e] = lens[side_src[e]];
    MVMC_CUDA_OK(cudaMemcpyToSymbol(c_skel, &h, sizeof(h)));
    g_skel_ready = true;
    return MVMC_OK;
}

// persistent grid: at most this many CTAs stride over the work slots (148 SMs x 2 resident CTAs x 4)
constexpr int IK_MAX_GRID = 1184;
static int ik_grid(int M) { return M < IK_MAX_GRID ? M : IK_MAX_GRID; }

extern "C" size_t mvmc_ik_workspace_bytes(int M, int V) {
    (void)V;
    return (size_t)ik_grid(M) * TRF_NMAX * TRF_MMAX * sizeof(double);
}

extern "C" int mvmc_ik_solve(const double* kps2d, const double* Psel, const int* n_views, const double* x0,
                             const uint8_t* birth, const int* max_nfev, const uint8_t* free_mask, int M, int V,
                             void* workspace, double* x_out, double* joints, int* info, double* cost, void* stream) {
    if (!kps2d || !Psel || !n_views || !x0 || !max_nfev || !workspace || !x_out || !joints || !info || !cost)
        return MVMC_ERR_INVALID;
    if (M <= 0 || V < 2 || V > MVMC_MAX_SEL) return MVMC_ERR_INVALID;
    int rc = ensure_skeleton();
    if (rc) return rc;
    MVMC_CUDA_OK(cudaFuncSetAttribute(k_ik_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IkShared)));
    MVMC_LAUNCH(k_ik_solve, dim3(ik_grid(M)), dim3(IK_THREADS), sizeof(IkShared), stream, kps2d, Psel, n_views, x0, birth,
                max_nfev, free_mask, M, V, (double*)workspace, x_out, joints, info, cost);
    MVMC_CHECK_LAUNCH("k_ik_solve");
    return MVMC_OK;
}

extern "C" int mvmc_triangulate(const double* obs, const double* Psel, const int* n_views, int M, int V, int K,
                                double min_score, int refine_nfev, double* out, void* stream) {
    if (!obs || !Psel || !n_views || !out) return MVMC_ERR_INVALID;
    if (M <= 0 || V < 1 || V > MVMC_MAX_SEL || K < 1 || K > 18 || refine_nfev < 0) return MVMC_ERR_INVALID;
    double* ws = nullptr;
    if (refine_nfev > 0) {
        // the refine needs a Jacobian scratch; allocated per call (stage API only; the clip pipeline owns its own)
        MVMC_CUDA_OK(cudaMalloc((void**)&ws, mvmc_ik_workspace_bytes(M, V)));
    }
    MVMC_CUDA_OK(cudaFuncSetAttribute(k_triangulate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TriShared)));
    MVMC_LAUNCH(k_triangulate, dim3(ik_grid(M)), dim3(IK_THREADS), sizeof(TriShared), stream, obs, Psel, n_views, M, V, K,
                min_score, refine_nfev, ws, out);
    cudaError_t e = cudaGetLastError();
    if (ws) {
        cudaStreamSynchronize((cudaStream_t)stream);
        cudaFree(ws);
    }
    if (e != cudaSuccess) return mvmc_set_cuda_error(e, "k_triangulate");
    return MVMC_OK;
}

extern "C" int mvmc_fk(const double* params, int M, double* joints, void* stream) {
    if (!params || !joints || M <= 0) return MVMC_ERR_INVALID;
    int rc = ensure_skeleton();
    if (rc) return rc;
    MVMC_LAUNCH(k_fk, dim3((M + 63) / 64), dim3(64), 0, stream, params, M, joints);
    MVMC_CHECK_LAUNCH("k_fk");
    return MVMC_OK;
}

extern "C" int mvmc_fk_chain(const double* rot, const double* offsets, const int* parents, const double* root, int M,
                             int J, double* joints, void* stream) {
    if (!rot || !offsets || !parents || !joints || M <= 0 || J <= 0 || J > 64) return MVMC_ERR_INVALID;
    MVMC_LAUNCH(k_fk_chain, dim3(M), dim3(64), 0, stream, rot, offsets, parents, root, M, J, joints);
    MVMC_CHECK_LAUNCH("k_fk_chain");
    return MVMC_OK;
}

int mvmc_ensure_skeleton() { return ensure_skeleton(); }
