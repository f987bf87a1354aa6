// Alternative matchers of the reference, all built on one device solver of the linear sum assignment
// problem:
//   * mvmc_linear_sum_assignment     scipy.optimize.linear_sum_assignment (third party: the rectangular shortest augmenting
//                                    path algorithm of Crouse 2016 as SciPy implements it, same scan order and tie rule)
//   * mvmc_match_views_hungarian     motion_capture.py:166-241 match_objects_across_views + :44-101 PoseAssociation: greedy
//                                    view-by-view Hungarian grouping on the epipolar distances (the README's "greedy" matcher)
//   * mvmc_tracklet_pose_association motion_capture.py:844-871 tracklet_to_pose_2d_cost / tracklet_to_poses_association with
//                                    mv_math_util.py:11-32 (3D joints against the rays of a view's 2D joints) + Hungarian
// One warp per problem; the column scan of an augmenting-path step runs across the lanes.
#include "mvmc_common.cuh"

namespace mvmc {

constexpr int LS_MAXC = MVMC_MAX_VIEWS * MVMC_MAX_POSES;   // longest side of a problem (columns after the transpose)
constexpr int LS_MAXR = 64;                                 // shortest side

struct LsapWs {
    double u[LS_MAXR], v[LS_MAXC], spc[LS_MAXC];
    int path[LS_MAXC], col4row[LS_MAXR], row4col[LS_MAXC], remaining[LS_MAXC];
    unsigned char SR[LS_MAXR], SC[LS_MAXC];
};

// cost(i, j) for i < nr <= nc: element [i*ld + j], or [j*ld + i] when `tr` (the caller's matrix had more rows than columns).
// Returns 0, or -1 when the problem is infeasible / has a NaN (SciPy raises ValueError). col4row[i] = column of row i.
__device__ int lsap_warp(const double* cost, int ld, bool tr, int nr, int nc, LsapWs& w) {
    const int lane = threadIdx.x & 31;
    auto C = [&](int i, int j) { return tr ? cost[(size_t)j * ld + i] : cost[(size_t)i * ld + j]; };
    int bad = 0;
    for (int e = lane; e < nr * nc; e += 32) {
        const double c = C(e / nc, e % nc);
        if (c != c || c == -INFINITY) bad = 1;
    }
    if (__any_sync(MVMC_FULL, bad)) return -1;
    for (int i = lane; i < nr; i += 32) {
        w.u[i] = 0.0;
        w.col4row[i] = -1;
    }
    for (int j = lane; j < nc; j += 32) {
        w.v[j] = 0.0;
        w.row4col[j] = -1;
        w.path[j] = -1;
    }
    __syncwarp();
    for (int cur = 0; cur < nr; cur++) {
        // ---- shortest augmenting path from row `cur` ----
        double minVal = 0.0;
        int num_remaining = nc;
        for (int it = lane; it < nc; it += 32) {
            w.remaining[it] = nc - it - 1;
            w.SC[it] = 0;
            w.spc[it] = INFINITY;
        }
        for (int i = lane; i < nr; i += 32) w.SR[i] = 0;
        __syncwarp();
        int sink = -1, i = cur;
        while (sink == -1) {
            if (lane == 0) w.SR[i] = 1;
            // scan the remaining columns: relax, then pick the lowest (ties: SciPy takes the first minimum it meets, replaced
            // by any LATER equal one whose column is still unassigned)
            double best = INFINITY;
            int best_it = -1, best_free_it = -1;
            const double ui = w.u[i];
            for (int it = lane; it < num_remaining; it += 32) {
                const int j = w.remaining[it];
                const double r = minVal + C(i, j) - ui - w.v[j];
                if (r < w.spc[j]) {
                    w.path[j] = i;
                    w.spc[j] = r;
                }
                const double s = w.spc[j];
                if (s < best) {
                    best = s;
                    best_it = it;
                    best_free_it = (w.row4col[j] == -1) ? it : -1;
                } else if (s == best && w.row4col[j] == -1) {
                    best_free_it = it;
                }
            }
            // warp reduction: lowest value; among equal values the largest free `it` if any is free, else the smallest `it`
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(MVMC_FULL, best, o);
                const int oi = __shfl_xor_sync(MVMC_FULL, best_it, o), of = __shfl_xor_sync(MVMC_FULL, best_free_it, o);
                if (ob < best) {
                    best = ob;
                    best_it = oi;
                    best_free_it = of;
                } else if (ob == best) {
                    best_it = (best_it < 0) ? oi : (oi < 0 ? best_it : min(best_it, oi));
                    best_free_it = max(best_free_it, of);
                }
            }
            if (best == INFINITY) return -1;
            minVal = best;
            const int index = best_free_it >= 0 ? best_free_it : best_it;
            const int j = w.remaining[index];
            if (w.row4col[j] == -1) sink = j;
            else i = w.row4col[j];
            __syncwarp();
            if (lane == 0) {
                w.SC[j] = 1;
                w.remaining[index] = w.remaining[num_remaining - 1];
            }
            num_remaining--;
            __syncwarp();
        }
        // ---- dual update and augmentation ----
        if (lane == 0) w.u[cur] += minVal;
        __syncwarp();
        for (int r = lane; r < nr; r += 32)
            if (w.SR[r] && r != cur) w.u[r] += minVal - w.spc[w.col4row[r]];
        for (int j = lane; j < nc; j += 32)
            if (w.SC[j]) w.v[j] -= minVal - w.spc[j];
        __syncwarp();
        if (lane == 0) {
            int j = sink;
            for (;;) {
                const int r = w.path[j];
                w.row4col[j] = r;
                const int t = w.col4row[r];
                w.col4row[r] = j;
                j = t;
                if (r == cur) break;
            }
        }
        __syncwarp();
    }
    return 0;
}

// SciPy's front end: transpose when there are more rows than columns; result per ORIGINAL row (-1 = unassigned)
__device__ int lsap_rows(const double* cost, int ld, int nr, int nc, LsapWs& w, int* col_of_row) {
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < nr; i += 32) col_of_row[i] = -1;
    __syncwarp();
    if (nr == 0 || nc == 0) return 0;
    const bool tr = nc < nr;
    const int a = tr ? nc : nr, b = tr ? nr : nc;
    if (a > LS_MAXR || b > LS_MAXC) return -2;
    const int rc = lsap_warp(cost, ld, tr, a, b, w);
    if (rc) return rc;
    for (int i = lane; i < a; i += 32) {
        if (tr) col_of_row[w.col4row[i]] = i;
        else col_of_row[i] = w.col4row[i];
    }
    __syncwarp();
    return 0;
}

__global__ void __launch_bounds__(32)
    k_lsap(const double* __restrict__ cost, const int* __restrict__ n_rows, const int* __restrict__ n_cols, int R, int Cc,
           int* __restrict__ col_of_row, int* __restrict__ status) {
    MVMC_DYN_SMEM(LsapWs, wsp);
    const int b = blockIdx.x;
    const int rc = lsap_rows(cost + (size_t)b * R * Cc, Cc, min(n_rows[b], R), min(n_cols[b], Cc), *wsp, col_of_row + (size_t)b * R);
    if (threadIdx.x == 0) status[b] = rc;
}

// ---- greedy Hungarian grouping across views (motion_capture.py:166-241) ----
// dst: the float64 distance matrix of mvmc_distances (2D-2D entries = calc_epipolar_error), index layout of mvmc_prepare with
// T = dim_groups[1]. Groups come out as lists of global pose indices in merge order; group_of [N] = group of every pose.
struct ViewsSh {
    LsapWs ws;
    int members[LS_MAXC];        // group member lists, concatenated in creation order ... (linked through next[])
    int head[LS_MAXC], tail[LS_MAXC], len[LS_MAXC], next[LS_MAXC];
    int match[LS_MAXC];          // LSAP result of the current view
};
__global__ void __launch_bounds__(32)
    k_match_views(const double* __restrict__ dst, const int* __restrict__ dim_groups, int C, int N, double threshold,
                  double* __restrict__ cost_ws, int* __restrict__ group_of, int* __restrict__ n_groups, int* __restrict__ status) {
    MVMC_DYN_SMEM(ViewsSh, shp);
    ViewsSh& sh = *shp;
    const int b = blockIdx.x, lane = threadIdx.x;
    const int* dg = dim_groups + (size_t)b * (C + 2);
    const double* D = dst + (size_t)b * N * N;
    double* cost = cost_ws + (size_t)b * LS_MAXC * MVMC_MAX_POSES;   // [n_hyp][n_poses of the view], ld = MVMC_MAX_POSES
    int* gof = group_of + (size_t)b * N;
    for (int e = lane; e < N; e += 32) gof[e] = -1;
    // view with the most poses (first maximum) seeds one group per pose
    int init = 0, best = -1;
    for (int v = 0; v < C; v++) {
        const int c = dg[v + 2] - dg[v + 1];
        if (c > best) {
            best = c;
            init = v;
        }
    }
    int ng = 0;
    __syncwarp();
    if (lane == 0) {
        for (int g = dg[init + 1]; g < dg[init + 2]; g++) {
            sh.head[ng] = sh.tail[ng] = g;
            sh.len[ng] = 1;
            sh.next[g] = -1;
            gof[g] = ng;
            ng++;
        }
    }
    ng = __shfl_sync(MVMC_FULL, ng, 0);
    int rc_all = 0;
    for (int v = 0; v < C; v++) {
        if (v == init) continue;
        const int p0 = dg[v + 1], np = dg[v + 2] - dg[v + 1];
        if (np < 1) continue;
        __syncwarp();
        // cost[h][p] = mean over the members q of group h (merge order) of dst[q][p]; too_bad once a running total > threshold
        // (the mask is kept in the sign bit-free side array `match` later; here: bit array in registers per (h, p) is too big,
        //  so the mask is recomputed for the matched pairs only)
        for (int e = lane; e < ng * np; e += 32) {
            const int h = e / np, p = e % np;
            double total = 0.0;
            for (int q = sh.head[h]; q >= 0; q = sh.next[q]) total += D[(size_t)q * N + p0 + p];
            cost[(size_t)h * MVMC_MAX_POSES + p] = total / sh.len[h];
        }
        __syncwarp();
        const int rc = lsap_rows(cost, MVMC_MAX_POSES, ng, np, sh.ws, sh.match);
        if (rc) {
            rc_all = rc;
            break;
        }
        if (lane == 0) {
            unsigned matched = 0;   // np <= 32
            const int ng0 = ng;
            for (int h = 0; h < ng0; h++) {
                const int p = sh.match[h];
                if (p < 0) continue;
                matched |= 1u << p;
                const int g = p0 + p;
                bool too_bad = false;
                double total = 0.0;
                for (int q = sh.head[h]; q >= 0; q = sh.next[q]) {
                    total += D[(size_t)q * N + g];
                    if (total > threshold) too_bad = true;
                }
                sh.next[g] = -1;
                if (too_bad) {            // even the closest hypothesis is too far: the pose starts its own group
                    sh.head[ng] = sh.tail[ng] = g;
                    sh.len[ng] = 1;
                    gof[g] = ng++;
                } else {
                    sh.next[sh.tail[h]] = g;
                    sh.tail[h] = g;
                    sh.len[h]++;
                    gof[g] = h;
                }
            }
            for (int p = 0; p < np; p++)
                if (!((matched >> p) & 1u)) {
                    const int g = p0 + p;
                    sh.head[ng] = sh.tail[ng] = g;
                    sh.len[ng] = 1;
                    sh.next[g] = -1;
                    gof[g] = ng++;
                }
        }
        ng = __shfl_sync(MVMC_FULL, ng, 0);
    }
    if (lane == 0) {
        n_groups[b] = ng;
        status[b] = rc_all;
    }
}

// ---- tracklet <-> 2D pose association by 3D ray distance (motion_capture.py:844-871, mv_math_util.py:11-32) ----
// COCO joint order of the 15 joints shared with BASIC_18 (map_to_common_keypoints(pose_2d, pose_3d): source = COCO)
__constant__ int c_ray_coco[15] = {0, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
__constant__ int c_ray_b18[15] = {15, 16, 17, 9, 12, 10, 13, 11, 14, 1, 4, 2, 5, 3, 6};
struct RaySh {
    LsapWs ws;
    double cost[MVMC_MAX_TRACKS * MVMC_MAX_POSES];
    int pid[MVMC_MAX_POSES], col[MVMC_MAX_TRACKS];
};
__global__ void __launch_bounds__(32)
    k_tracklet_pose_assoc(const double* __restrict__ trk_joints, const int* __restrict__ n_trk, const double* __restrict__ kps,
                          const uint8_t* __restrict__ keep, const double* __restrict__ Kr_inv, const double* __restrict__ cam_loc,
                          int C, int Pmax, int Tmax, double max_dst, int* __restrict__ match, double* __restrict__ cost_out,
                          int* __restrict__ status) {
    MVMC_DYN_SMEM(RaySh, shp);
    RaySh& sh = *shp;
    const int b = blockIdx.x / C, v = blockIdx.x % C, lane = threadIdx.x;
    const int T = min(n_trk[b], Tmax);
    int* mt = match + ((size_t)b * C + v) * Tmax;
    for (int t = lane; t < Tmax; t += 32) mt[t] = -1;
    int np = 0;
    if (lane == 0) {
        for (int p = 0; p < Pmax; p++)
            if (keep[((size_t)b * C + v) * Pmax + p]) sh.pid[np++] = p;
    }
    np = __shfl_sync(MVMC_FULL, np, 0);
    __syncwarp();
    if (T == 0 || np == 0) {
        if (lane == 0) status[blockIdx.x] = 0;
        return;
    }
    const double* Ki = Kr_inv + ((size_t)b * C + v) * 9;
    const double* cl = cam_loc + ((size_t)b * C + v) * 3;
    for (int e = lane; e < T * np; e += 32) {
        const int t = e / np, p = sh.pid[e % np];
        const double* k2 = kps + (((size_t)b * C + v) * Pmax + p) * (MVMC_N_COCO * 3);
        const double* j3 = trk_joints + ((size_t)b * Tmax + t) * (MVMC_N_B18 * 3);
        // np.mean of 15 distances: NumPy's pairwise block (8 accumulators, then the remainder)
        double r8[8], rest = 0.0, dd[15];
        for (int q = 0; q < 15; q++) {
            const double x = k2[c_ray_coco[q] * 3], y = k2[c_ray_coco[q] * 3 + 1];
            double ray[3];
            for (int a = 0; a < 3; a++) ray[a] = Ki[a * 3] * x + Ki[a * 3 + 1] * y + Ki[a * 3 + 2];
            const double nrm = sqrt(ray[0] * ray[0] + ray[1] * ray[1] + ray[2] * ray[2]);
            for (int a = 0; a < 3; a++) ray[a] /= nrm;
            const double* p3 = j3 + c_ray_b18[q] * 3;
            const double d0 = p3[0] - cl[0], d1 = p3[1] - cl[1], d2 = p3[2] - cl[2];
            const double c0 = d1 * ray[2] - d2 * ray[1], c1 = d2 * ray[0] - d0 * ray[2], c2 = d0 * ray[1] - d1 * ray[0];
            dd[q] = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
        }
        for (int q = 0; q < 8; q++) r8[q] = dd[q];
        double res = ((r8[0] + r8[1]) + (r8[2] + r8[3])) + ((r8[4] + r8[5]) + (r8[6] + r8[7]));
        for (int q = 8; q < 15; q++) res += dd[q];
        (void)rest;
        const double c = res / 15.0;
        sh.cost[t * MVMC_MAX_POSES + (e % np)] = c;
        if (cost_out) cost_out[(((size_t)b * C + v) * Tmax + t) * Pmax + p] = c;
    }
    __syncwarp();
    const int rc = lsap_rows(sh.cost, MVMC_MAX_POSES, T, np, sh.ws, sh.col);
    if (rc == 0)
        for (int t = lane; t < T; t += 32) {
            const int c = sh.col[t];
            if (c >= 0 && !(sh.cost[t * MVMC_MAX_POSES + c] > max_dst)) mt[t] = sh.pid[c];
        }
    if (lane == 0) status[blockIdx.x] = rc;
}

}  // namespace mvmc

using namespace mvmc;

extern "C" int mvmc_linear_sum_assignment(const double* cost, const int* n_rows, const int* n_cols, int B, int R, int Cc,
                                          int* col_of_row, int* status, void* stream) {
    if (!cost || !n_rows || !n_cols || !col_of_row || !status || B <= 0 || R <= 0 || Cc <= 0) return MVMC_ERR_INVALID;
    if ((R < Cc ? R : Cc) > LS_MAXR || (R > Cc ? R : Cc) > LS_MAXC) return MVMC_ERR_INVALID;
    MVMC_CUDA_OK(cudaFuncSetAttribute(k_lsap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LsapWs)));
    MVMC_LAUNCH(k_lsap, dim3(B), dim3(32), sizeof(LsapWs), stream, cost, n_rows, n_cols, R, Cc, col_of_row, status);
    MVMC_CHECK_LAUNCH("k_lsap");
    return MVMC_OK;
}

extern "C" size_t mvmc_match_views_workspace_bytes(int B) { return (size_t)B * LS_MAXC * MVMC_MAX_POSES * sizeof(double); }

extern "C" int mvmc_match_views_hungarian(const double* dst, const int* dim_groups, int B, int C, int N, double threshold,
                                          void* workspace, int* group_of, int* n_groups, int* status, void* stream) {
    if (!dst || !dim_groups || !workspace || !group_of || !n_groups || !status) return MVMC_ERR_INVALID;
    if (B <= 0 || C <= 0 || C > MVMC_MAX_VIEWS || N <= 0 || N > LS_MAXC) return MVMC_ERR_INVALID;
    MVMC_CUDA_OK(cudaFuncSetAttribute(k_match_views, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ViewsSh)));
    MVMC_LAUNCH(k_match_views, dim3(B), dim3(32), sizeof(ViewsSh), stream, dst, dim_groups, C, N, threshold, (double*)workspace,
                group_of, n_groups, status);
    MVMC_CHECK_LAUNCH("k_match_views");
    return MVMC_OK;
}

extern "C" int mvmc_tracklet_pose_association(const double* trk_joints, const int* n_trk, const double* kps, const uint8_t* keep,
                                              const double* Kr_inv, const double* cam_loc, int B, int C, int Pmax, int Tmax,
                                              double max_dst, int* match, double* cost, int* status, void* stream) {
    if (!trk_joints || !n_trk || !kps || !keep || !Kr_inv || !cam_loc || !match || !status) return MVMC_ERR_INVALID;
    if (B <= 0 || C <= 0 || C > MVMC_MAX_VIEWS || Pmax <= 0 || Pmax > MVMC_MAX_POSES || Tmax <= 0 || Tmax > MVMC_MAX_TRACKS)
        return MVMC_ERR_INVALID;
    MVMC_CUDA_OK(cudaFuncSetAttribute(k_tracklet_pose_assoc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RaySh)));
    MVMC_LAUNCH(k_tracklet_pose_assoc, dim3(B * C), dim3(32), sizeof(RaySh), stream, trk_joints, n_trk, kps, keep, Kr_inv, cam_loc,
                C, Pmax, Tmax, max_dst, match, cost, status);
    MVMC_CHECK_LAUNCH("k_tracklet_pose_assoc");
    return MVMC_OK;
}
