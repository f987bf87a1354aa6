// Forward kinematics, DLT triangulation and the reprojection IK (trust-region-reflective least squares with a
// forward-difference Jacobian) as sm_100a CUDA kernels.
//
// Reference rows: B1/B2 mv_math_util.py:152-240; I0 inverse_kinematics.py:339-348;
// I1 inverse_kinematics.py:176-199 + Quaternions.py:97-115,335-366,442-462; I2/I3 :202-277;
// I4 scipy.optimize.least_squares(method='trf', jac='2-point', tr_solver='exact') restated on device
// (trf_warp.cuh); I5 :380-433; I6 kinematics.py:18-31.
//
// k_ik_solve: ONE WARP (one 32-thread CTA) PER SOLVE, persistent CTAs pulling work slots from an atomic counter.
// Lanes own Jacobian columns: each runs the fully unrolled BASIC_18 chain in registers on its own perturbed
// parameter vector and parks the 16 observed joint positions in shared memory; the Jacobian is then produced
// half a camera view at a time and folded into J^T J by 8x8 register tiles. 4 resident CTAs per SM (53 KB of
// shared memory each); FP64 CUDA-core work throughout (no dense contraction large enough for tensor cores).
#include "trf_warp.cuh"
#include "det_math.cuh"

namespace mvmc {

// ---- BASIC_18 skeleton (inverse_kinematics.py:120-173) ----
struct SkelConst {
    int parents[MVMC_N_B18];
    int side_to_full[MVMC_N_B18];
    double dirs[MVMC_N_B18][3];
    double ref_side_lens[11];
    unsigned char param_dead[MVMC_N_PARAM];   // 1: the parameter cannot move any joint (leaf rotations, root bone length)
    unsigned char leaf[MVMC_N_B18];
    signed char ik_slot[MVMC_N_B18];           // slot among the 16 observed joints, -1 for Mid_Hip / Neck
};
__constant__ SkelConst c_skel;
__constant__ int c_ik_obs_idx[MVMC_N_IKJ] = {11, 13, 15, 12, 14, 16, 17, 5, 7, 9, 6, 8, 10, 0, 3, 4};

// compile-time copies of the topology for the unrolled chain (kept equal to ensure_skeleton()'s tables)
struct Topo {
    static __host__ __device__ constexpr int parent(int j) {
        constexpr int p[MVMC_N_B18] = {-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 10, 8, 12, 13, 8, 15, 15};
        return p[j];
    }
    static __host__ __device__ constexpr int s2f(int j) {
        constexpr int t[MVMC_N_B18] = {7, 0, 1, 2, 0, 1, 2, 8, 9, 3, 4, 5, 3, 4, 5, 10, 6, 6};
        return t[j];
    }
    static __host__ __device__ constexpr bool leaf(int j) { return j == 3 || j == 6 || j == 11 || j == 14 || j == 16 || j == 17; }
    // slot of skeleton joint j among the 16 joints the IK observes (skeleton idx 1..7, 9..17), -1 otherwise
    static __host__ __device__ constexpr int ik_slot(int j) { return (j == 0 || j == 8) ? -1 : (j < 8 ? j - 1 : j - 2); }
};

// local rotation R = Rx(a) Ry(b) Rz(c) via half-angle quaternions, as Quaternions.from_euler + transforms
__device__ __forceinline__ void euler_to_mat(double ea, double eb, double ec, double* m) {
    const double k = 1.0 / (1.0 + 1e-10);  // axis / (|axis| + 1e-10), Quaternions.py:444
    double sx, cx, sy, cy, sz, cz;
    // (own sincos and no implicit FMA contraction in this file - it is compiled with -fmad=false, every fused operation
    //  is an explicit fma(): the solver's results are then bit-identical on the GPU and in the CPU build of these sources)
    det_sincos(ea / 2.0, &sx, &cx);
    det_sincos(eb / 2.0, &sy, &cy);
    det_sincos(ec / 2.0, &sz, &cz);
    sx *= k;
    sy *= k;
    sz *= k;
    // t = qy * qz, q = qx * t (reference operand order)
    const double t0 = cz * cy, t1 = sz * sy, t2 = cz * sy, t3 = sz * cy;
    const double qw = fma(t0, cx, -(t1 * sx));
    const double qx = fma(t0, sx, t1 * cx);
    const double qy = fma(t2, cx, -(t3 * sx));
    const double qz = fma(t2, sx, t3 * cx);
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

// x = [root(3) | euler(54)], lens = 11 side lengths -> pos[18][3]   (generic, table driven; used by k_fk)
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

// The chain inside the solver. A Jacobian column perturbs ONE parameter, so only one local rotation differs from the
// unperturbed pose: the 12 non-leaf local rotations of the current x are formed once (one lane each, `local_rots`) and
// parked in shared memory; every lane then walks the chain with its own override of a single joint (and of the root
// translation / a bone length, through getx). Same arithmetic as evaluating the whole pose per column, 12x fewer
// sincos. The walk is a compact rolled loop with everything in registers: only three global rotations are ever live -
// the root's, the neck's (joint 8, parent of both arms and the head) and the current chain's - so the parent of a
// joint is a 3-way select instead of an indexed (local-memory) array. Leaf rotations are never formed (no joint
// position depends on them). Kept rolled on purpose: unrolled it is ~4.5k instructions per call site and the solver
// kernel outgrew the instruction cache.
__constant__ unsigned char c_fk_psel[MVMC_N_B18] = {0, 1, 0, 0, 1, 0, 0, 1, 0, 2, 0, 0, 2, 0, 0, 2, 0, 0};   // 0 cur, 1 root, 2 neck

// Rloc[j][9] <- local rotation of joint j at pose x, non-leaf joints only (lane j computes joint j)
__device__ __noinline__ void local_rots(const double* x, double* Rloc) {
    MVMC_ASSUME_SHARED(x);
    MVMC_ASSUME_SHARED(Rloc);
    const int j = threadIdx.x & 31;
    if (j < MVMC_N_B18 && !c_skel.leaf[j]) {
        double m[9];
        euler_to_mat(x[3 + 3 * j], x[4 + 3 * j], x[5 + 3 * j], m);
#pragma unroll
        for (int q = 0; q < 9; q++) Rloc[j * 9 + q] = m[q];
    }
    __syncwarp();
}

// Walks the chain of pose x with parameter `prm` replaced by `xp` (prm < 0: no override) and stores joint j at
// out[j*3*stride + k*stride] - the 16 observed joints only (slot order) when `ik_slots`, else all 18. One copy for the
// whole kernel. `Rloc` must hold local_rots(x).
__device__ __noinline__ void fk_store(const double* x, const double* Rloc, int prm, double xp, double* out, int stride,
                                      bool ik_slots, bool on) {
    MVMC_ASSUME_SHARED(x);
    MVMC_ASSUME_SHARED(Rloc);
    auto getx = [&](int i) { return i == prm ? xp : x[i]; };
    // the one local rotation this call overrides (computed by every lane to stay convergent; unused when jo < 0)
    const int jo = (prm >= 3 && prm < 57) ? (prm - 3) / 3 : -1;
    double mo[9];
    {
        const int jj = jo < 0 ? 0 : jo;
        euler_to_mat(getx(3 + 3 * jj), getx(4 + 3 * jj), getx(5 + 3 * jj), mo);
    }
    double Gc[9], Gr[9], Gn[9], pc[3], pr[3], pn[3];
#pragma unroll
    for (int q = 0; q < 9; q++) Gr[q] = jo == 0 ? mo[q] : Rloc[q];
    pr[0] = getx(0);
    pr[1] = getx(1);
    pr[2] = getx(2);
#pragma unroll
    for (int q = 0; q < 9; q++) Gc[q] = Gn[q] = Gr[q];
#pragma unroll
    for (int q = 0; q < 3; q++) pc[q] = pn[q] = pr[q];
    auto emit = [&](int j, double px, double py, double pz) {
        const int sl = ik_slots ? c_skel.ik_slot[j] : j;
        if (sl >= 0 && on) {
            out[(sl * 3) * stride] = px;
            out[(sl * 3 + 1) * stride] = py;
            out[(sl * 3 + 2) * stride] = pz;
        }
    };
    emit(0, pr[0], pr[1], pr[2]);
#pragma unroll 1
    for (int j = 1; j < MVMC_N_B18; j++) {
        const int sel = c_fk_psel[j];
        double Rp[9], pp[3];
#pragma unroll
        for (int q = 0; q < 9; q++) Rp[q] = sel == 0 ? Gc[q] : (sel == 1 ? Gr[q] : Gn[q]);
#pragma unroll
        for (int q = 0; q < 3; q++) pp[q] = sel == 0 ? pc[q] : (sel == 1 ? pr[q] : pn[q]);
        const double len = getx(57 + c_skel.side_to_full[j]);
        const double o0 = c_skel.dirs[j][0] * len, o1 = c_skel.dirs[j][1] * len, o2 = c_skel.dirs[j][2] * len;
        double pj[3];
#pragma unroll
        for (int r = 0; r < 3; r++) pj[r] = fma(Rp[r * 3 + 2], o2, fma(Rp[r * 3 + 1], o1, fma(Rp[r * 3], o0, pp[r])));
        emit(j, pj[0], pj[1], pj[2]);
        if (!c_skel.leaf[j]) {   // uniform branch
            double m[9];
#pragma unroll
            for (int q = 0; q < 9; q++) m[q] = j == jo ? mo[q] : Rloc[j * 9 + q];
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) Gc[r * 3 + c] = fma(Rp[r * 3 + 2], m[6 + c], fma(Rp[r * 3 + 1], m[3 + c], Rp[r * 3] * m[c]));
#pragma unroll
            for (int q = 0; q < 3; q++) pc[q] = pj[q];
            if (j == 8) {
#pragma unroll
                for (int q = 0; q < 9; q++) Gn[q] = Gc[q];
#pragma unroll
                for (int q = 0; q < 3; q++) pn[q] = pj[q];
            }
        }
    }
}

__device__ __forceinline__ void project3(const double* Pv, double X, double Y, double Z, double& pu, double& pv, double& pw) {
    pu = fma(Pv[2], Z, fma(Pv[1], Y, fma(Pv[0], X, Pv[3])));
    pv = fma(Pv[6], Z, fma(Pv[5], Y, fma(Pv[4], X, Pv[7])));
    pw = fma(Pv[10], Z, fma(Pv[9], Y, fma(Pv[8], X, Pv[11])));
}

// ---- reprojection residual of the BASIC_18 pose (inverse_kinematics.py:202-277) ----
// rows: (view v, observed joint q, {u, v}) -> (v*16 + q)*2 + {0,1};  a chunk = 4 joints of one view (8 rows)
struct IkRes {
    static constexpr bool kGlobalF = false;
    const double* obs;   // [V][16][3] gathered at c_ik_obs_idx (shared)
    const double* P;     // [V][12] (shared)
    double* posb;        // [16][3] scratch (shared)
    double* Rloc;        // [18][9] local rotations of the pose being differentiated / evaluated (shared)
    int V;
    __device__ int m() const { return V * MVMC_N_IKJ * 2; }
    __device__ int n_chunks() const { return 4 * V; }
    __device__ int chunk_rows(int) const { return 8; }
    __device__ int chunk_row0(int c) const { return 8 * c; }

    __device__ __noinline__ void eval(const double* x, double* f) {
        MVMC_ASSUME_SHARED(x);
        MVMC_ASSUME_SHARED(f);
        MVMC_ASSUME_SHARED(obs);
        MVMC_ASSUME_SHARED(P);
        MVMC_ASSUME_SHARED(posb);
        MVMC_ASSUME_SHARED(Rloc);
        const int lane = threadIdx.x & 31;
        double* pb = posb;
        // the chain is evaluated redundantly by every lane (uniform control flow); lane 0 parks the positions
        local_rots(x, Rloc);
        fk_store(x, Rloc, -1, 0.0, pb, 1, true, lane == 0);
        __syncwarp();
        for (int it = lane; it < V * MVMC_N_IKJ; it += 32) {
            const int v = it >> 4, q = it & 15;
            const double* o = obs + it * 3;
            double pu, pv, pw;
            project3(P + v * 12, pb[q * 3], pb[q * 3 + 1], pb[q * 3 + 2], pu, pv, pw);
            // one reciprocal instead of two divisions - the same expression in eval and in the Jacobian columns, so the
            // forward differences see a consistent rounding (the quotient differs from pu/den by <= 1 ulp)
            const double inv = 1.0 / (1e-5 + pw);
            f[2 * it] = DMUL(DSUB(DMUL(pu, inv), o[0]), o[2]);
            f[2 * it + 1] = DMUL(DSUB(DMUL(pv, inv), o[1]), o[2]);
        }
        __syncwarp();
    }
    // per column: perturbed chain; the 16 observed joint positions go to the scratch S[(slot*3+c)*TW::NC + col] (aliases s.A)
    template <class TW>
    __device__ void fd_prepare(TW& s, int ncol) {
        const int lane = threadIdx.x & 31;
        double* S = s.A;
        local_rots(s.x, Rloc);
        for (int c0 = 0; c0 < ncol; c0 += 32) {
            const int c = c0 + lane;
            const bool on = c < ncol;
            const int prm = on ? s.act[c] : -1;
            const double xp = on ? s.w[c] : 0.0;
            fk_store(s.x, Rloc, prm, xp, S + c, TW::NC, true, on);
        }
    }
    template <class TW>
    __device__ __forceinline__ void fd_chunk(TW& s, int ncol, int ch, const double* f) {
        MVMC_ASSUME_SHARED(&s);
        MVMC_ASSUME_SHARED(f);
        MVMC_ASSUME_SHARED(obs);
        MVMC_ASSUME_SHARED(P);
        const int lane = threadIdx.x & 31;
        const int v = ch >> 2, q0 = (ch & 3) * 4;
        const double* S = s.A;
        double Pv[12];
#pragma unroll
        for (int e = 0; e < 12; e++) Pv[e] = P[v * 12 + e];
        for (int c = lane; c < ncol; c += 32) {
            const double rdx = s.tau[c];   // 1 / dx, formed once per Jacobian
#pragma unroll
            for (int qq = 0; qq < 4; qq++) {
                const int q = q0 + qq;
                const double* o = obs + (v * MVMC_N_IKJ + q) * 3;
                double pu, pv, pw;
                project3(Pv, S[(q * 3) * TW::NC + c], S[(q * 3 + 1) * TW::NC + c], S[(q * 3 + 2) * TW::NC + c], pu, pv, pw);
                const double inv = 1.0 / (1e-5 + pw);
                const double ru = DMUL(DSUB(DMUL(pu, inv), o[0]), o[2]);
                const double rv = DMUL(DSUB(DMUL(pv, inv), o[1]), o[2]);
                const int row = (v * MVMC_N_IKJ + q) * 2;
                s.Jc[(2 * qq) * WS_LDJ + c] = DMUL(DSUB(ru, f[row]), rdx);
                s.Jc[(2 * qq + 1) * WS_LDJ + c] = DMUL(DSUB(rv, f[row + 1]), rdx);
            }
        }
    }
};

// ---- 3D-target residual (inverse_kinematics.py:280-336 solve_pose / solve_pose_bone_lens): FK joints against triangulated
// points, rows (observed joint q, coordinate c) -> 3 q + c, weighted by the point's score; 48 rows = 6 chunks of 8 ----
struct Ik3dRes {
    static constexpr bool kGlobalF = false;
    const double* tgt;   // [16][4] (x, y, z, score), gathered at the IK joints (shared)
    double* posb;        // [16][3] scratch (shared)
    double* Rloc;        // [18][9] (shared)
    __device__ int m() const { return MVMC_N_IKJ * 3; }
    __device__ int n_chunks() const { return MVMC_N_IKJ * 3 / 8; }
    __device__ int chunk_rows(int) const { return 8; }
    __device__ int chunk_row0(int c) const { return 8 * c; }
    __device__ __noinline__ void eval(const double* x, double* f) {
        MVMC_ASSUME_SHARED(x);
        MVMC_ASSUME_SHARED(f);
        MVMC_ASSUME_SHARED(tgt);
        MVMC_ASSUME_SHARED(posb);
        MVMC_ASSUME_SHARED(Rloc);
        const int lane = threadIdx.x & 31;
        local_rots(x, Rloc);
        fk_store(x, Rloc, -1, 0.0, posb, 1, true, lane == 0);
        __syncwarp();
        for (int it = lane; it < MVMC_N_IKJ * 3; it += 32) {
            const int q = it / 3, c = it % 3;
            f[it] = DMUL(DSUB(posb[q * 3 + c], tgt[q * 4 + c]), tgt[q * 4 + 3]);
        }
        __syncwarp();
    }
    __device__ void fd_prepare(TrfWarp& s, int ncol) {
        const int lane = threadIdx.x & 31;
        double* S = s.A;
        local_rots(s.x, Rloc);
        for (int c0 = 0; c0 < ncol; c0 += 32) {
            const int c = c0 + lane;
            const bool on = c < ncol;
            fk_store(s.x, Rloc, on ? s.act[c] : -1, on ? s.w[c] : 0.0, S + c, WS_NC, true, on);
        }
    }
    __device__ __forceinline__ void fd_chunk(TrfWarp& s, int ncol, int ch, const double* f) {
        MVMC_ASSUME_SHARED(&s);
        MVMC_ASSUME_SHARED(f);
        MVMC_ASSUME_SHARED(tgt);
        const int lane = threadIdx.x & 31;
        const double* S = s.A;
        for (int c = lane; c < ncol; c += 32) {
            const double rdx = s.tau[c];   // 1 / dx, formed once per Jacobian
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int row = 8 * ch + r, q = row / 3, cc = row % 3;
                const double v = DMUL(DSUB(S[(q * 3 + cc) * WS_NC + c], tgt[q * 4 + cc]), tgt[q * 4 + 3]);
                s.Jc[r * WS_LDJ + c] = DMUL(DSUB(v, f[row]), rdx);
            }
        }
    }
};

// ---- triangulation refine residual (mv_math_util.py:190-202): rows (view v, point k) -> v*K + k ----
// a chunk = a third of the points of one view (K <= 18: at most 6 rows)
struct TriRes {
    static constexpr bool kGlobalF = false;
    const double* obs;  // [V][K][3] (shared)
    const double* P;    // [V][12]
    int V, K;
    __device__ int third() const { return (K + 2) / 3; }
    __device__ int part0(int t) const { return min(K, t * third()); }           // first point of part t (t = 0..3)
    __device__ int m() const { return V * K; }
    __device__ int n_chunks() const { return 3 * V; }
    __device__ int chunk_rows(int c) const { return part0(c % 3 + 1) - part0(c % 3); }
    __device__ int chunk_row0(int c) const { return (c / 3) * K + part0(c % 3); }
    __device__ double one(const double* Pv, const double* o, double X, double Y, double Z) const {
        double pu, pv, pw;
        project3(Pv, X, Y, Z, pu, pv, pw);
        const double den = pw + 1e-6;
        const double du = DSUB(DDIV(pu, den), o[0]), dv = DSUB(DDIV(pv, den), o[1]);
        return DMUL(sqrt(DMUL(du, du) + DMUL(dv, dv)), o[2]);
    }
    __device__ void eval(const double* x, double* f) {
        const int lane = threadIdx.x & 31;
        for (int it = lane; it < V * K; it += 32) {
            const int v = it / K, k = it % K;
            f[it] = one(P + v * 12, obs + it * 3, x[3 * k], x[3 * k + 1], x[3 * k + 2]);
        }
        __syncwarp();
    }
    __device__ void fd_prepare(TrfWarp&, int) {}
    __device__ void fd_chunk(TrfWarp& s, int ncol, int ch, const double* f) {
        const int lane = threadIdx.x & 31;
        const int v = ch / 3, k0 = part0(ch % 3), k1 = k0 + chunk_rows(ch);
        for (int c = lane; c < ncol; c += 32) {
            const int prm = s.act[c], k = prm / 3, comp = prm % 3;
            for (int r = 0; r < k1 - k0; r++) s.Jc[r * WS_LDJ + c] = 0.0;   // (the other parts' entries are stale)
            if (k < k0 || k >= k1) continue;   // a point only moves its own residual rows
            double X[3] = {s.x[3 * k], s.x[3 * k + 1], s.x[3 * k + 2]};
            X[comp] = s.w[c];
            const double r = one(P + v * 12, obs + (v * K + k) * 3, X[0], X[1], X[2]);
            s.Jc[(k - k0) * WS_LDJ + c] = DDIV(DSUB(r, f[v * K + k]), s.dx[c]);
        }
    }
};

// ---- DLT: null vector of the (2V x 4) system by one-sided Jacobi, one thread per joint ----
// mv_math_util.py:215-240 (rows x P[2] - P[0], y P[2] - P[1] of every used view, smallest right singular vector, divide by
// w). View v owns rows 2v, 2v+1 of a fixed-size system (zero rows for views that are not used: they add exact zeros to every
// sum, so the result is that of the compacted system) and every loop is unrolled: the whole system lives in registers
// (the first version indexed local arrays by run-time row numbers and waited on local-memory loads 80 % of the time).
template <int VM>
__device__ __forceinline__ void dlt_point(const double* P /*[V][12]*/, const double* obs /*[V][K][3]*/, int V, int K, int k, unsigned use,
                                          double* out3) {
    double a[2 * VM][4];
    double v[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
#pragma unroll
    for (int q = 0; q < VM; q++) {
        const bool on = q < V && ((use >> q) & 1u);
        const double* Pv = P + (on ? q : 0) * 12;
        const double px = on ? obs[(q * K + k) * 3] : 0.0, py = on ? obs[(q * K + k) * 3 + 1] : 0.0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            a[2 * q][c] = on ? px * Pv[8 + c] - Pv[c] : 0.0;
            a[2 * q + 1][c] = on ? py * Pv[8 + c] - Pv[4 + c] : 0.0;
        }
    }
    for (int sweep = 0; sweep < 30; sweep++) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
            for (int q = p + 1; q < 4; q++) {
                double app = 0, aqq = 0, apq = 0;
#pragma unroll
                for (int r = 0; r < 2 * VM; r++) {
                    app += a[r][p] * a[r][p];
                    aqq += a[r][q] * a[r][q];
                    apq += a[r][p] * a[r][q];
                }
                if (apq == 0.0 || fabs(apq) <= kEps * sqrt(app * aqq)) continue;
                rotated = true;
                const double tau = (aqq - app) / (2.0 * apq);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = t * c;
#pragma unroll
                for (int r = 0; r < 2 * VM; r++) {
                    const double ap = a[r][p], aq = a[r][q];
                    a[r][p] = c * ap - sn * aq;
                    a[r][q] = sn * ap + c * aq;
                }
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const double vp = v[r][p], vq = v[r][q];
                    v[r][p] = c * vp - sn * vq;
                    v[r][q] = sn * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    double bn = INFINITY, b0 = 0, b1 = 0, b2 = 0, b3 = 1;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        double nn = 0;
#pragma unroll
        for (int r = 0; r < 2 * VM; r++) nn += a[r][c] * a[r][c];
        if (nn < bn) {
            bn = nn;
            b0 = v[0][c];
            b1 = v[1][c];
            b2 = v[2][c];
            b3 = v[3][c];
        }
    }
    out3[0] = b0 / b3;
    out3[1] = b1 / b3;
    out3[2] = b2 / b3;
}

// obs [V][K][3] -> out [K][4]; one thread per joint. mv_math_util.py:152-186: views with score >= min_score, all views if
// fewer than two qualify; the fourth output is the mean score of the views used.
template <int VM>
__device__ __forceinline__ void triangulate_joint(const double* obs, const double* P, int V, int K, int k, double min_score, double* out4) {
    unsigned use = 0;
    int nsel = 0;
    for (int v = 0; v < V; v++)
        if (obs[(v * K + k) * 3 + 2] >= min_score) {
            use |= 1u << v;
            nsel++;
        }
    if (nsel < 2) {
        nsel = V;
        use = V >= 32 ? 0xffffffffu : ((1u << V) - 1u);
    }
    double sc = 0.0;
    for (int v = 0; v < V; v++)
        if ((use >> v) & 1u) sc += obs[(v * K + k) * 3 + 2];
    dlt_point<VM>(P, obs, V, K, k, use, out4);
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

// ---- per-warp shared memory of the solver kernels ----
// (the staging copy of the 18-point observations and the triangulated points exist only in the kernels that give birth)
// The update-only instance is trimmed so that six one-warp CTAs fit an SM: a 50-column solver workspace (the pose model has
// at most 49 live columns; the 54-column triangulation refine exists only where births do) and no buffer of its own for the
// trial residuals - a trial is evaluated while the J chunk is idle, so they live there (trf_solve_warp copies an accepted
// trial into f before the next Jacobian is formed).
constexpr int IK_UPDATE_NC = 50;
template <int VMAX, bool BIRTH>
struct IkBirthSh {
    double obs18[VMAX * 18 * 3];
    double p3[18 * 4];
    double fn_[32 * VMAX];
    TrfWarp t;
    static constexpr int NC = TrfWarp::NC;
    __device__ __forceinline__ double* fn() { return fn_; }
};
template <int VMAX>
struct IkBirthSh<VMAX, false> {
    TrfWarpT<IK_UPDATE_NC> t;
    static constexpr int NC = IK_UPDATE_NC;
    __device__ __forceinline__ double* fn() { return t.Jc; }
    static_assert(32 * VMAX <= WS_CH * WS_LDJ, "the trial residuals must fit the J chunk");
};
template <int VMAX, bool BIRTH = true>
struct alignas(16) IkWarpSh : IkBirthSh<VMAX, BIRTH> {
    double f[32 * VMAX];
    double obs16[VMAX * MVMC_N_IKJ * 3];
    double P[VMAX * 12];
    double posb[MVMC_N_IKJ * 3];
    double Rloc[MVMC_N_B18 * 9];
};
constexpr int IK_UPDATE_CTAS = 6;   // resident update-solver CTAs per SM
static_assert(IK_UPDATE_CTAS * (sizeof(IkWarpSh<8, false>) + 1024) <= 228 * 1024, "six update-solver CTAs must fit the 228 KB of an SM");

// Triangulate K (<= 18) joints from nv views (+ optional refine_nfev-evaluation TRF refine); result in sh.p3 [K][4].
template <int VMAX>
__device__ void warp_triangulate(IkWarpSh<VMAX>& sh, const double* obs /*[nv][K][3] shared*/, int nv, int K, double min_score,
                                 int refine_nfev) {
    const int lane = threadIdx.x & 31;
    if (lane < K) triangulate_joint<VMAX>(obs, sh.P, nv, K, lane, min_score, sh.p3 + lane * 4);
    __syncwarp();
    if (refine_nfev <= 0) return;
    for (int e = lane; e < MVMC_N_PARAM; e += 32) sh.t.x[e] = 0.0;
    __syncwarp();
    for (int e = lane; e < 3 * K; e += 32) {
        sh.t.x[e] = sh.p3[(e / 3) * 4 + (e % 3)];
        sh.t.act[e] = e;
    }
    __syncwarp();
    TriRes tr{obs, sh.P, nv, K};
    trf_solve_warp(sh.t, tr, 3 * K, 3 * K, 0.0, false, refine_nfev, sh.f, sh.fn());
    __syncwarp();
    for (int e = lane; e < 3 * K; e += 32) sh.p3[(e / 3) * 4 + (e % 3)] = sh.t.x[e];
    __syncwarp();
}

// active-column list of one IK stage: optimised (free_mask) parameters in [0, npar) that can move a joint.
// returns ncol; n_opt / x2_dead / has_dead describe the optimised-but-dead parameters (they enter SciPy's norms).
template <class TW>
__device__ int ik_columns(TW& t, const uint8_t* free_mask, int npar, int& n_opt, double& x2_dead, bool& has_dead) {
    const int lane = threadIdx.x & 31;
    int ncol = 0;
    n_opt = 0;
    x2_dead = 0.0;
    has_dead = false;
    for (int e0 = 0; e0 < npar; e0 += 32) {
        const int e = e0 + lane;
        const bool opt = e < npar && (!free_mask || free_mask[e]);
        const bool live = opt && !c_skel.param_dead[e];
        const unsigned mo = __ballot_sync(MVMC_FULL, opt), ml = __ballot_sync(MVMC_FULL, live);
        if (live) {
            const int c = ncol + __popc(ml & ((1u << lane) - 1u));
            if (c < WS_NC) t.act[c] = e;
        }
        const double xv = (opt && !live) ? t.x[e] : 0.0;
        x2_dead += warp_sum(xv * xv);
        n_opt += __popc(mo);
        ncol += __popc(ml);
        has_dead = has_dead || (mo != ml);
    }
    __syncwarp();
    return ncol;
}

// ------------------------------------------------------------------------------------------------
// Births from MANY poses. A no-track frame of a crowded scene makes the reference's float32 affinity merge dozens of 2D
// poses (several per view, of several people) into one group, and it then builds the new track from ALL of them
// (motion_capture.py:618-624, 942-958: cam_poses_2d = every grouped pose, one "view" each). Groups of up to MVMC_MAX_SEL
// poses go through k_ik_solve's shared-memory staging; larger ones (up to MVMC_MAX_GROUP) are solved here, one warp per
// group, with the observations, the residual vectors and the DLT systems in a global scratch slot owned by the CTA: same
// triangulation (+ 2-evaluation refine), same two 50-evaluation TRF stages, no cap on the number of "views".
// ------------------------------------------------------------------------------------------------
struct BigScratch {
    double f[32 * MVMC_MAX_GROUP], fn[32 * MVMC_MAX_GROUP];
    double obs18[MVMC_MAX_GROUP * 18 * 3];
    double dlt[18][2 * MVMC_MAX_GROUP][4];
};
struct IkResBig {
    static constexpr bool kGlobalF = true;
    const double* obs18;  // [V][18][3] (global scratch): COCO + mid spine
    const int* vof;       // [V] camera of every pose (shared)
    const double* P;      // [C][12] (shared)
    double* posb;
    double* Rloc;
    int V;
    __device__ int m() const { return V * MVMC_N_IKJ * 2; }
    __device__ int n_chunks() const { return 4 * V; }
    __device__ int chunk_rows(int) const { return 8; }
    __device__ int chunk_row0(int c) const { return 8 * c; }
    __device__ __noinline__ void eval(const double* x, double* f) {
        const int lane = threadIdx.x & 31;
        local_rots(x, Rloc);
        fk_store(x, Rloc, -1, 0.0, posb, 1, true, lane == 0);
        __syncwarp();
        for (int it = lane; it < V * MVMC_N_IKJ; it += 32) {
            const int v = it >> 4, q = it & 15;
            const double* o = obs18 + (v * 18 + c_ik_obs_idx[q]) * 3;
            double pu, pv, pw;
            project3(P + vof[v] * 12, posb[q * 3], posb[q * 3 + 1], posb[q * 3 + 2], pu, pv, pw);
            const double inv = 1.0 / (1e-5 + pw);
            f[2 * it] = DMUL(DSUB(DMUL(pu, inv), o[0]), o[2]);
            f[2 * it + 1] = DMUL(DSUB(DMUL(pv, inv), o[1]), o[2]);
        }
        __syncwarp();
    }
    __device__ void fd_prepare(TrfWarp& s, int ncol) {
        const int lane = threadIdx.x & 31;
        double* S = s.A;
        local_rots(s.x, Rloc);
        for (int c0 = 0; c0 < ncol; c0 += 32) {
            const int c = c0 + lane;
            const bool on = c < ncol;
            fk_store(s.x, Rloc, on ? s.act[c] : -1, on ? s.w[c] : 0.0, S + c, WS_NC, true, on);
        }
    }
    __device__ __forceinline__ void fd_chunk(TrfWarp& s, int ncol, int ch, const double* f) {
        const int lane = threadIdx.x & 31;
        const int v = ch >> 2, q0 = (ch & 3) * 4;
        const double* S = s.A;
        double Pv[12];
#pragma unroll
        for (int e = 0; e < 12; e++) Pv[e] = P[vof[v] * 12 + e];
        for (int c = lane; c < ncol; c += 32) {
            const double rdx = s.tau[c];
#pragma unroll
            for (int qq = 0; qq < 4; qq++) {
                const int q = q0 + qq;
                const double* o = obs18 + (v * 18 + c_ik_obs_idx[q]) * 3;
                double pu, pv, pw;
                project3(Pv, S[(q * 3) * WS_NC + c], S[(q * 3 + 1) * WS_NC + c], S[(q * 3 + 2) * WS_NC + c], pu, pv, pw);
                const double inv = 1.0 / (1e-5 + pw);
                const double ru = DMUL(DSUB(DMUL(pu, inv), o[0]), o[2]);
                const double rv = DMUL(DSUB(DMUL(pv, inv), o[1]), o[2]);
                const int row = (v * MVMC_N_IKJ + q) * 2;
                s.Jc[(2 * qq) * WS_LDJ + c] = DMUL(DSUB(ru, f[row]), rdx);
                s.Jc[(2 * qq + 1) * WS_LDJ + c] = DMUL(DSUB(rv, f[row + 1]), rdx);
            }
        }
    }
};
struct TriResBig {
    static constexpr bool kGlobalF = true;
    const double* obs;  // [V][18][3] (global)
    const int* vof;
    const double* P;
    int V;
    static constexpr int K = 18;
    __device__ int m() const { return V * K; }
    __device__ int n_chunks() const { return 3 * V; }
    __device__ int chunk_rows(int) const { return 6; }
    __device__ int chunk_row0(int c) const { return (c / 3) * K + 6 * (c % 3); }
    __device__ double one(const double* Pv, const double* o, double X, double Y, double Z) const {
        double pu, pv, pw;
        project3(Pv, X, Y, Z, pu, pv, pw);
        const double den = pw + 1e-6;
        const double du = DSUB(DDIV(pu, den), o[0]), dv = DSUB(DDIV(pv, den), o[1]);
        return DMUL(sqrt(DMUL(du, du) + DMUL(dv, dv)), o[2]);
    }
    __device__ void eval(const double* x, double* f) {
        const int lane = threadIdx.x & 31;
        for (int it = lane; it < V * K; it += 32) {
            const int v = it / K, k = it % K;
            f[it] = one(P + vof[v] * 12, obs + it * 3, x[3 * k], x[3 * k + 1], x[3 * k + 2]);
        }
        __syncwarp();
    }
    __device__ void fd_prepare(TrfWarp&, int) {}
    __device__ void fd_chunk(TrfWarp& s, int ncol, int ch, const double* f) {
        const int lane = threadIdx.x & 31;
        const int v = ch / 3, k0 = 6 * (ch % 3), k1 = k0 + 6;
        for (int c = lane; c < ncol; c += 32) {
            const int prm = s.act[c], k = prm / 3, comp = prm % 3;
            for (int r = 0; r < 6; r++) s.Jc[r * WS_LDJ + c] = 0.0;
            if (k < k0 || k >= k1) continue;
            double X[3] = {s.x[3 * k], s.x[3 * k + 1], s.x[3 * k + 2]};
            X[comp] = s.w[c];
            const double r = one(P + vof[v] * 12, obs + (v * K + k) * 3, X[0], X[1], X[2]);
            s.Jc[(k - k0) * WS_LDJ + c] = DDIV(DSUB(r, f[v * K + k]), s.dx[c]);
        }
    }
};

// DLT of one joint from any number of views: the (2 nsel x 4) system in a global scratch block (one-sided Jacobi, the
// arithmetic of dlt_point on the compacted rows)
__device__ void dlt_point_dyn(double (*a)[4], int rows, double* out3) {
    double v[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
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

struct BigSh {
    TrfWarp t;
    double P[MVMC_MAX_VIEWS * 12];
    double p3[18 * 4];
    double posb[MVMC_N_IKJ * 3];
    double Rloc[MVMC_N_B18 * 9];
    int vof[MVMC_MAX_GROUP];
};

// items: (clip b, big group g) for g < big_n[b]; the group's (view, pose id) list is big_sel[b][g][0..nsel); the result goes
// to work slot b * S + slot0 + big_slot[b][g]. Persistent CTAs (one warp, one scratch slot each) scan the items.
__global__ void __launch_bounds__(32)
    k_ik_birth_big(const double* __restrict__ kps, const double* __restrict__ Pm, const int* __restrict__ big_n,
                   const int* __restrict__ big_nsel, const int* __restrict__ big_sel, const int* __restrict__ big_slot, int B, int C,
                   int Pmax, int G, int S, int slot0, int nfev_cap, BigScratch* __restrict__ scratch, int* __restrict__ counter,
                   double* __restrict__ x_out, double* __restrict__ joints, int* __restrict__ info, double* __restrict__ cost_out) {
    MVMC_DYN_SMEM(BigSh, shp);
    BigSh& sh = *shp;
    BigScratch& sc = scratch[blockIdx.x];
    const int lane = threadIdx.x & 31;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(counter, 1);
        item = __shfl_sync(MVMC_FULL, item, 0);
        if (item >= B * G) break;
        const int b = item / G, g = item % G;
        if (g >= big_n[b]) continue;
        const int nv = min(big_nsel[b * G + g], MVMC_MAX_GROUP);
        const int* sel = big_sel + ((size_t)b * G + g) * MVMC_MAX_GROUP * 2;
        const size_t mI = (size_t)b * S + slot0 + big_slot[b * G + g];
        __syncwarp();
        for (int e = lane; e < C * 12; e += 32) sh.P[e] = Pm[(size_t)b * C * 12 + e];
        for (int v = lane; v < nv; v += 32) sh.vof[v] = sel[2 * v];
        for (int e = lane; e < nv * MVMC_N_COCO * 3; e += 32) {
            const int v = e / (MVMC_N_COCO * 3), q = e % (MVMC_N_COCO * 3);
            sc.obs18[v * 54 + q] = kps[((size_t)(b * C + sel[2 * v]) * Pmax + sel[2 * v + 1]) * (MVMC_N_COCO * 3) + q];
        }
        __syncwarp();
        for (int v = lane; v < nv; v += 32) mid_spine(sc.obs18 + v * 54, sc.obs18 + v * 54 + 51);
        __syncwarp();
        // triangulate the 18 points (score >= 0.01, all views if fewer than two qualify), mv_math_util.py:152-186
        if (lane < 18) {
            const int k = lane;
            int nsel = 0;
            for (int v = 0; v < nv; v++) nsel += sc.obs18[(v * 18 + k) * 3 + 2] >= 0.01 ? 1 : 0;
            const bool all = nsel < 2;
            double (*a)[4] = sc.dlt[k];
            int rows = 0;
            double ssum = 0.0;
            for (int v = 0; v < nv; v++) {
                const double* o = sc.obs18 + (v * 18 + k) * 3;
                if (!(all || o[2] >= 0.01)) continue;
                const double* Pv = sh.P + sh.vof[v] * 12;
                for (int c = 0; c < 4; c++) {
                    a[rows][c] = o[0] * Pv[8 + c] - Pv[c];
                    a[rows + 1][c] = o[1] * Pv[8 + c] - Pv[4 + c];
                }
                rows += 2;
                ssum += o[2];
            }
            dlt_point_dyn(a, rows, sh.p3 + k * 4);
            sh.p3[k * 4 + 3] = ssum / (rows / 2);
        }
        __syncwarp();
        {   // 2-evaluation refine of the 54 coordinates (mv_math_util.py:188-210)
            for (int e = lane; e < MVMC_N_PARAM; e += 32) sh.t.x[e] = 0.0;
            __syncwarp();
            for (int e = lane; e < 54; e += 32) {
                sh.t.x[e] = sh.p3[(e / 3) * 4 + (e % 3)];
                sh.t.act[e] = e;
            }
            __syncwarp();
            TriResBig tr{sc.obs18, sh.vof, sh.P, nv};
            trf_solve_warp(sh.t, tr, 54, 54, 0.0, false, 2, sc.f, sc.fn);
            __syncwarp();
            for (int e = lane; e < 54; e += 32) sh.p3[(e / 3) * 4 + (e % 3)] = sh.t.x[e];
            __syncwarp();
        }
        if (lane == 0) {
            for (int c = 0; c < 3; c++) sh.t.x[c] = 0.5 * (sh.p3[kCocoLHip * 4 + c] + sh.p3[kCocoRHip * 4 + c]);
            for (int e = 3; e < 57; e++) sh.t.x[e] = 0.0;
            for (int e = 0; e < 11; e++) sh.t.x[57 + e] = c_skel.ref_side_lens[e];
        }
        __syncwarp();
        IkResBig res{sc.obs18, sh.vof, sh.P, sh.posb, sh.Rloc, nv};
        TrfResult r[2];
        int ncols[2];
#pragma unroll 1
        for (int stage = 0; stage < 2; stage++) {
            int n_opt;
            double x2_dead;
            bool has_dead;
            const int ncol = ik_columns(sh.t, nullptr, stage == 0 ? 57 : MVMC_N_PARAM, n_opt, x2_dead, has_dead);
            ncols[stage] = n_opt;
            r[stage] = trf_solve_warp(sh.t, res, ncol, n_opt, x2_dead, has_dead, nfev_cap, sc.f, sc.fn);
            __syncwarp();
        }
        for (int e = lane; e < MVMC_N_PARAM; e += 32) x_out[mI * MVMC_N_PARAM + e] = sh.t.x[e];
        local_rots(sh.t.x, sh.Rloc);
        fk_store(sh.t.x, sh.Rloc, -1, 0.0, joints + mI * MVMC_N_B18 * 3, 1, false, lane == 0);
        if (lane == 0) {
            for (int q = 0; q < 2; q++) {
                info[mI * 8 + 4 * q] = r[q].nfev;
                info[mI * 8 + 4 * q + 1] = r[q].njev;
                info[mI * 8 + 4 * q + 2] = r[q].status;
                info[mI * 8 + 4 * q + 3] = ncols[q];
                cost_out[mI * 2 + q] = r[q].cost;
            }
        }
        __syncwarp();
    }
}

template <int VMAX, bool WITH_BIRTH>
__global__ void __launch_bounds__(32)
    k_ik_solve(const double* __restrict__ kps2d, const double* __restrict__ Psel, const int* __restrict__ n_views,
               const double* __restrict__ x0, const uint8_t* __restrict__ birth, const int* __restrict__ max_nfev,
               const uint8_t* __restrict__ free_mask, int n_items, int cnt, int S, int s0, int V, int* __restrict__ counter,
               double* __restrict__ x_out, double* __restrict__ joints, int* __restrict__ info, double* __restrict__ cost_out) {
    typedef IkWarpSh<VMAX, WITH_BIRTH> Sh;
    MVMC_DYN_SMEM(Sh, shp);
    Sh& sh = *shp;
    const int lane = threadIdx.x & 31;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(counter, 1);
        item = __shfl_sync(MVMC_FULL, item, 0);
        if (item >= n_items) break;
        const int mI = (item / cnt) * S + s0 + item % cnt;
        const int nv = n_views[mI];
        if (nv < 2) continue;  // the reference never solves from fewer than two views (motion_capture.py:926,942)
        int* inf = info + (size_t)mI * 8;
        if (nv > VMAX) {       // cannot happen through the pipeline (it sizes VMAX from the slot kind)
            if (lane < 8) inf[lane] = -1;
            continue;
        }
        const bool is_birth = birth != nullptr && birth[mI] != 0;
        if (is_birth && !WITH_BIRTH) {   // births go through the WITH_BIRTH instantiation (separate launch in the pipeline)
            if (lane < 8) inf[lane] = -1;
            continue;
        }
        const int nfev_cap = max_nfev[mI];
        __syncwarp();
        // stage observations (+ mid spine) and projection matrices
        for (int e = lane; e < nv * 12; e += 32) sh.P[e] = Psel[(size_t)mI * V * 12 + e];
        if constexpr (WITH_BIRTH) {
            for (int e = lane; e < nv * MVMC_N_COCO * 3; e += 32) {
                const int v = e / (MVMC_N_COCO * 3), q = e % (MVMC_N_COCO * 3);
                sh.obs18[v * 54 + q] = kps2d[((size_t)mI * V + v) * (MVMC_N_COCO * 3) + q];
            }
            __syncwarp();
            if (lane < nv) mid_spine(sh.obs18 + lane * 54, sh.obs18 + lane * 54 + 51);
            __syncwarp();
            for (int e = lane; e < nv * MVMC_N_IKJ * 3; e += 32) {
                const int v = e / (MVMC_N_IKJ * 3), q = (e / 3) % MVMC_N_IKJ, c = e % 3;
                sh.obs16[e] = sh.obs18[v * 54 + c_ik_obs_idx[q] * 3 + c];
            }
        } else {
            // straight from global memory into the 16-joint layout; the mid spine (index 17) with mid_spine()'s arithmetic
            for (int e = lane; e < nv * MVMC_N_IKJ * 3; e += 32) {
                const int v = e / (MVMC_N_IKJ * 3), q = (e / 3) % MVMC_N_IKJ, c = e % 3;
                const double* k = kps2d + ((size_t)mI * V + v) * (MVMC_N_COCO * 3);
                const int src = c_ik_obs_idx[q];
                double val;
                if (src < MVMC_N_COCO) {
                    val = k[src * 3 + c];
                } else {
                    const double ls = k[3 * kCocoLShoulder + c], rs = k[3 * kCocoRShoulder + c], lh = k[3 * kCocoLHip + c],
                                 rh = k[3 * kCocoRHip + c];
                    if (c < 2) {
                        val = 0.5 * (0.5 * (ls + rs) + 0.5 * (lh + rh));
                    } else {
                        double sc = ls * rs;
                        sc *= lh * rh;
                        val = sc;
                    }
                }
                sh.obs16[e] = val;
            }
        }
        __syncwarp();
        bool born = false;
        if constexpr (WITH_BIRTH) {
            born = is_birth;
            if (born) {
                // triangulate 18 joints (min score 0.01), refine with a 2-nfev TRF, inverse_kinematics.py:389-396
                warp_triangulate(sh, sh.obs18, nv, 18, 0.01, 2);
                if (lane == 0) {
                    double root[3];
                    for (int c = 0; c < 3; c++) root[c] = 0.5 * (sh.p3[kCocoLHip * 4 + c] + sh.p3[kCocoRHip * 4 + c]);
                    for (int c = 0; c < 3; c++) sh.t.x[c] = root[c];
                    for (int e = 3; e < 57; e++) sh.t.x[e] = 0.0;
                    for (int e = 0; e < 11; e++) sh.t.x[57 + e] = c_skel.ref_side_lens[e];
                }
            }
        }
        if (!born) {
            for (int e = lane; e < MVMC_N_PARAM; e += 32) sh.t.x[e] = x0[(size_t)mI * MVMC_N_PARAM + e];
        }
        __syncwarp();
        IkRes res{sh.obs16, sh.P, sh.posb, sh.Rloc, nv};
        TrfResult r[2];
        int ncols[2];
#pragma unroll 1
        for (int stage = 0; stage < 2; stage++) {
            // stage 0: root + angles, lengths fixed (solve_pose_reproj); stage 1: + the 11 side lengths
            int n_opt;
            double x2_dead;
            bool has_dead;
            const int ncol = ik_columns(sh.t, free_mask, stage == 0 ? 57 : MVMC_N_PARAM, n_opt, x2_dead, has_dead);
            ncols[stage] = n_opt;
            r[stage].nfev = 1;
            r[stage].njev = 1;
            r[stage].status = 0;
            r[stage].cost = 0.0;
            if (ncol > Sh::NC) {
                r[stage].status = -2;
            } else if (ncol > 0) {
                r[stage] = trf_solve_warp(sh.t, res, ncol, n_opt, x2_dead, has_dead, nfev_cap, sh.f, sh.fn());
            }
            __syncwarp();
        }
        for (int e = lane; e < MVMC_N_PARAM; e += 32) x_out[(size_t)mI * MVMC_N_PARAM + e] = sh.t.x[e];
        local_rots(sh.t.x, sh.Rloc);
        fk_store(sh.t.x, sh.Rloc, -1, 0.0, joints + (size_t)mI * MVMC_N_B18 * 3, 1, false, lane == 0);
        if (lane == 0) {
            for (int q = 0; q < 2; q++) {
                inf[4 * q] = r[q].nfev;
                inf[4 * q + 1] = r[q].njev;
                inf[4 * q + 2] = r[q].status;
                inf[4 * q + 3] = ncols[q];
                cost_out[(size_t)mI * 2 + q] = r[q].cost;
            }
        }
        __syncwarp();
    }
}

template <int VMAX>
__global__ void __launch_bounds__(32)
    k_triangulate(const double* __restrict__ obs, const double* __restrict__ Psel, const int* __restrict__ n_views, int M,
                  int V, int K, double min_score, int refine_nfev, double* __restrict__ out) {
    MVMC_DYN_SMEM(IkWarpSh<VMAX>, shp);
    IkWarpSh<VMAX>& sh = *shp;
    const int lane = threadIdx.x & 31;
    for (int mI = blockIdx.x; mI < M; mI += gridDim.x) {
        const int nv = n_views[mI];
        if (nv < 1 || nv > VMAX) continue;
        __syncwarp();
        for (int e = lane; e < nv * K * 3; e += 32) sh.obs18[e] = obs[(size_t)mI * V * K * 3 + e];
        for (int e = lane; e < nv * 12; e += 32) sh.P[e] = Psel[(size_t)mI * V * 12 + e];
        __syncwarp();
        warp_triangulate(sh, sh.obs18, nv, K, min_score, refine_nfev);
        for (int e = lane; e < K * 4; e += 32) out[(size_t)mI * K * 4 + e] = sh.p3[e];
        __syncwarp();
    }
}

// 3D-target solves (the reference's `use_only_reproj = False` branch, inverse_kinematics.py:409-415): one warp per solve.
// target [M,16,4]; stages bit 0 = solve_pose (root + angles), bit 1 = solve_pose_bone_lens (+ the 11 side lengths).
struct Ik3dSh {
    TrfWarp t;
    double f[64], fn[64];
    double tgt[MVMC_N_IKJ * 4];
    double posb[MVMC_N_IKJ * 3];
    double Rloc[MVMC_N_B18 * 9];
};
__global__ void __launch_bounds__(32)
    k_ik_targets(const double* __restrict__ target, const double* __restrict__ x0, const int* __restrict__ max_nfev, int stages,
                 int M, double* __restrict__ x_out, double* __restrict__ joints, int* __restrict__ info, double* __restrict__ cost_out) {
    MVMC_DYN_SMEM(Ik3dSh, shp);
    Ik3dSh& sh = *shp;
    const int lane = threadIdx.x & 31;
    for (int mI = blockIdx.x; mI < M; mI += gridDim.x) {
        __syncwarp();
        for (int e = lane; e < MVMC_N_IKJ * 4; e += 32) sh.tgt[e] = target[(size_t)mI * MVMC_N_IKJ * 4 + e];
        for (int e = lane; e < MVMC_N_PARAM; e += 32) sh.t.x[e] = x0[(size_t)mI * MVMC_N_PARAM + e];
        __syncwarp();
        Ik3dRes res{sh.tgt, sh.posb, sh.Rloc};
        TrfResult r[2];
        int ncols[2] = {0, 0};
#pragma unroll 1
        for (int stage = 0; stage < 2; stage++) {
            r[stage].nfev = 0;
            r[stage].njev = 0;
            r[stage].status = 0;
            r[stage].cost = 0.0;
            if (!((stages >> stage) & 1)) continue;
            int n_opt;
            double x2_dead;
            bool has_dead;
            const int ncol = ik_columns(sh.t, nullptr, stage == 0 ? 57 : MVMC_N_PARAM, n_opt, x2_dead, has_dead);
            ncols[stage] = n_opt;
            if (ncol > WS_NC) r[stage].status = -2;
            else if (ncol > 0) r[stage] = trf_solve_warp(sh.t, res, ncol, n_opt, x2_dead, has_dead, max_nfev[mI], sh.f, sh.fn);
            __syncwarp();
        }
        for (int e = lane; e < MVMC_N_PARAM; e += 32) x_out[(size_t)mI * MVMC_N_PARAM + e] = sh.t.x[e];
        local_rots(sh.t.x, sh.Rloc);
        fk_store(sh.t.x, sh.Rloc, -1, 0.0, joints + (size_t)mI * MVMC_N_B18 * 3, 1, false, lane == 0);
        if (lane == 0) {
            for (int q = 0; q < 2; q++) {
                info[(size_t)mI * 8 + 4 * q] = r[q].nfev;
                info[(size_t)mI * 8 + 4 * q + 1] = r[q].njev;
                info[(size_t)mI * 8 + 4 * q + 2] = r[q].status;
                info[(size_t)mI * 8 + 4 * q + 3] = ncols[q];
                cost_out[(size_t)mI * 2 + q] = r[q].cost;
            }
        }
        __syncwarp();
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
    static const int side_src[11] = {1, 2, 3, 9, 10, 11, 16, 0, 7, 8, 15};
    SkelConst h;
    double lens[MVMC_N_B18];
    bool has_child[MVMC_N_B18] = {};
    for (int j = 0; j < MVMC_N_B18; j++) {
        h.parents[j] = Topo::parent(j);
        h.side_to_full[j] = Topo::s2f(j);
        if (h.parents[j] >= 0) has_child[h.parents[j]] = true;
        volatile double a = off[j][0] * off[j][0];
        volatile double b = off[j][1] * off[j][1];
        volatile double c = off[j][2] * off[j][2];
        volatile double s = a + b;
        s = s + c;
        lens[j] = sqrt(s);
        for (int q = 0; q < 3; q++) h.dirs[j][q] = (j == 0) ? off[j][q] : off[j][q] / lens[j];
    }
    for (int e = 0; e < 11; e++) h.ref_side_lens[e] = lens[side_src[e]];
    // a parameter is dead when no joint position depends on it: rotations of childless joints, and a side length
    // that only the root uses (the root's own offset is replaced by the root translation)
    for (int e = 0; e < MVMC_N_PARAM; e++) h.param_dead[e] = 0;
    for (int j = 0; j < MVMC_N_B18; j++) {
        h.leaf[j] = Topo::leaf(j) ? 1 : 0;
        h.ik_slot[j] = (signed char)Topo::ik_slot(j);
    }
    for (int j = 0; j < MVMC_N_B18; j++) {
        if (has_child[j] != !Topo::leaf(j)) return MVMC_ERR_INVALID;
        if (!has_child[j])
            for (int c = 0; c < 3; c++) h.param_dead[3 + 3 * j + c] = 1;
    }
    for (int e = 0; e < 11; e++) {
        bool used = false;
        for (int j = 1; j < MVMC_N_B18; j++) used = used || h.side_to_full[j] == e;
        if (!used) h.param_dead[57 + e] = 1;
    }
    MVMC_CUDA_OK(cudaMemcpyToSymbol(c_skel, &h, sizeof(h)));
    g_skel_ready = true;
    return MVMC_OK;
}

// persistent one-warp CTAs: as many as are resident at once (six update solvers per SM; the kernels with the birth staging
// or the 3D targets hold fewer - their surplus CTAs find the work counter exhausted)
constexpr int IK_MAX_GRID = 148 * IK_UPDATE_CTAS;
static int ik_grid(int M) { return M < IK_MAX_GRID ? M : IK_MAX_GRID; }

extern "C" size_t mvmc_ik_workspace_bytes(int M, int V) {
    (void)M;
    (void)V;
    return 256;  // work counters of the persistent CTAs
}

// Internal launcher shared with the clip pipeline: items i in [0, n_items) map to slot (i / cnt) * S + s0 + i % cnt.
int mvmc_ik_launch(const double* kps2d, const double* Psel, const int* n_views, const double* x0, const uint8_t* birth,
                   const int* max_nfev, const uint8_t* free_mask, int n_items, int cnt, int S, int s0, int V, int vmax,
                   int* counter, double* x_out, double* joints, int* info, double* cost, void* stream) {
    int rc = ensure_skeleton();
    if (rc) return rc;
    MVMC_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(int), (cudaStream_t)stream));
    if (vmax <= 8 && birth == nullptr) {
        auto kern = k_ik_solve<8, false>;
        MVMC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IkWarpSh<8, false>)));
        MVMC_LAUNCH(kern, dim3(ik_grid(n_items)), dim3(32), sizeof(IkWarpSh<8, false>), stream, kps2d, Psel, n_views, x0, birth,
                    max_nfev, free_mask, n_items, cnt, S, s0, V, counter, x_out, joints, info, cost);
    } else {
        auto kern = k_ik_solve<MVMC_MAX_SEL, true>;
        MVMC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IkWarpSh<MVMC_MAX_SEL>)));
        MVMC_LAUNCH(kern, dim3(ik_grid(n_items)), dim3(32), sizeof(IkWarpSh<MVMC_MAX_SEL>), stream, kps2d, Psel, n_views, x0,
                    birth, max_nfev, free_mask, n_items, cnt, S, s0, V, counter, x_out, joints, info, cost);
    }
    MVMC_CHECK_LAUNCH("k_ik_solve");
    return MVMC_OK;
}

extern "C" int mvmc_ik_solve(const double* kps2d, const double* Psel, const int* n_views, const double* x0,
                             const uint8_t* birth, const int* max_nfev, const uint8_t* free_mask, int M, int V,
                             void* workspace, double* x_out, double* joints, int* info, double* cost, void* stream) {
    if (!kps2d || !Psel || !n_views || !x0 || !max_nfev || !workspace || !x_out || !joints || !info || !cost)
        return MVMC_ERR_INVALID;
    if (M <= 0 || V < 2 || V > MVMC_MAX_SEL) return MVMC_ERR_INVALID;
    return mvmc_ik_launch(kps2d, Psel, n_views, x0, birth, max_nfev, free_mask, M, M, M, 0, V, V, (int*)workspace, x_out,
                          joints, info, cost, stream);
}

constexpr int IK_BIG_CTAS = 148;
extern "C" size_t mvmc_ik_birth_big_workspace_bytes(void) { return 256 + (size_t)IK_BIG_CTAS * sizeof(BigScratch); }

// Births from groups of more than MVMC_MAX_SEL poses (see k_ik_birth_big). kps [B,C,Pmax,17,3], P [B,C,3,4]; big_n [B],
// big_nsel [B,G], big_sel [B,G,MVMC_MAX_GROUP,2] (view, pose id), big_slot [B,G]; outputs at work slot b*S + slot0 +
// big_slot[b][g] of x_out [.,68], joints [.,18,3], info [.,2,4], cost [.,2]. workspace: mvmc_ik_birth_big_workspace_bytes().
extern "C" int mvmc_ik_birth_big(const double* kps, const double* P, const int* big_n, const int* big_nsel, const int* big_sel,
                                 const int* big_slot, int B, int C, int Pmax, int G, int S, int slot0, int max_nfev, void* workspace,
                                 double* x_out, double* joints, int* info, double* cost, void* stream) {
    if (!kps || !P || !big_n || !big_nsel || !big_sel || !big_slot || !workspace || !x_out || !joints || !info || !cost)
        return MVMC_ERR_INVALID;
    if (B <= 0 || C <= 0 || C > MVMC_MAX_VIEWS || Pmax <= 0 || G <= 0 || S <= 0 || slot0 < 0 || max_nfev < 1) return MVMC_ERR_INVALID;
    int rc = ensure_skeleton();
    if (rc) return rc;
    int* counter = (int*)workspace;
    BigScratch* scratch = (BigScratch*)((char*)workspace + 256);
    MVMC_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(int), (cudaStream_t)stream));
    MVMC_CUDA_OK(cudaFuncSetAttribute(k_ik_birth_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BigSh)));
    const int grid = B * G < IK_BIG_CTAS ? B * G : IK_BIG_CTAS;
    MVMC_LAUNCH(k_ik_birth_big, dim3(grid), dim3(32), sizeof(BigSh), stream, kps, P, big_n, big_nsel, big_sel, big_slot, B, C, Pmax, G,
                S, slot0, max_nfev, scratch, counter, x_out, joints, info, cost);
    MVMC_CHECK_LAUNCH("k_ik_birth_big");
    return MVMC_OK;
}

extern "C" int mvmc_ik_solve_targets(const double* target, const double* x0, const int* max_nfev, int stages, int M,
                                     double* x_out, double* joints, int* info, double* cost, void* stream) {
    if (!target || !x0 || !max_nfev || !x_out || !joints || !info || !cost || M <= 0 || stages < 1 || stages > 3)
        return MVMC_ERR_INVALID;
    int rc = ensure_skeleton();
    if (rc) return rc;
    MVMC_CUDA_OK(cudaFuncSetAttribute(k_ik_targets, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Ik3dSh)));
    MVMC_LAUNCH(k_ik_targets, dim3(ik_grid(M)), dim3(32), sizeof(Ik3dSh), stream, target, x0, max_nfev, stages, M, x_out, joints,
                info, cost);
    MVMC_CHECK_LAUNCH("k_ik_targets");
    return MVMC_OK;
}

extern "C" int mvmc_triangulate(const double* obs, const double* Psel, const int* n_views, int M, int V, int K,
                                double min_score, int refine_nfev, double* out, void* stream) {
    if (!obs || !Psel || !n_views || !out) return MVMC_ERR_INVALID;
    if (M <= 0 || V < 1 || V > MVMC_MAX_SEL || K < 1 || K > 18 || refine_nfev < 0) return MVMC_ERR_INVALID;
    if (V <= 8) {   // (up to 8 views: the 16 x 4 system of a joint fits the register file without spills, 3 CTAs more per SM)
        MVMC_CUDA_OK(cudaFuncSetAttribute(k_triangulate<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IkWarpSh<8>)));
        MVMC_LAUNCH(k_triangulate<8>, dim3(M < 148 * 6 ? M : 148 * 6), dim3(32), sizeof(IkWarpSh<8>), stream, obs, Psel, n_views, M, V, K,
                    min_score, refine_nfev, out);
        MVMC_CHECK_LAUNCH("k_triangulate");
        return MVMC_OK;
    }
    MVMC_CUDA_OK(cudaFuncSetAttribute(k_triangulate<MVMC_MAX_SEL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(IkWarpSh<MVMC_MAX_SEL>)));
    MVMC_LAUNCH(k_triangulate<MVMC_MAX_SEL>, dim3(ik_grid(M)), dim3(32), sizeof(IkWarpSh<MVMC_MAX_SEL>), stream, obs, Psel,
                n_views, M, V, K, min_score, refine_nfev, out);
    MVMC_CHECK_LAUNCH("k_triangulate");
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
