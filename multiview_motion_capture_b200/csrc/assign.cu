// Closure quirk + leader assignment + first-kept-column parse + group decoding of the matcher's X_bin.
//
// Reference rows: A6 mv_association.py:99-121 (transform_closure) + motion_capture.py:417-446
// (parse_match_result) + motion_capture.py:762-808 / :618-624 (group decoding). Integer/bit work, one warp per clip.
#include "mvmc_common.cuh"

namespace mvmc {

// ------------------------------------------------------------------------------------------------
// A6: closure quirk, leader assignment, first-kept-column parse, group decoding. One warp per
// instance; lane w owns bit-word w of every N-bit row (N <= 1024).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
    k_assign(const uint32_t* __restrict__ xbin, const int* __restrict__ dim_groups, const int* __restrict__ idx_view,
             const int* __restrict__ idx_pose, const int* __restrict__ n_trk, int C, int N, int Tmax, int max_new,
             int* __restrict__ trk_nsel, int* __restrict__ trk_sel, int* __restrict__ new_n, int* __restrict__ new_nsel,
             int* __restrict__ new_sel, int* __restrict__ counts, int* __restrict__ err, int* __restrict__ new_seq,
             int* __restrict__ singles, int* __restrict__ big_n, int* __restrict__ big_nsel, int* __restrict__ big_sel,
             int* __restrict__ big_slot) {
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
    int n_new = 0, dup = 0, error = 0, n_single = 0, n_trunc = 0, seq = 0;   // seq: position among the 2D-only groups
    int n_big = 0;                                                          // groups of more than MVMC_MAX_SEL poses so far
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
            int nsel = 0, ntot = 0;   // poses stored in `sel` / poses of the group
            uint32_t seen_views = 0;
            bool over = false;
            // a no-track group may hold many poses per view: all of them go to the overflow list as well (kept if > MVMC_MAX_SEL)
            int* bs = (!has_trk && big_sel && n_big < MVMC_MAX_BIG) ? big_sel + ((size_t)b * MVMC_MAX_BIG + n_big) * MVMC_MAX_GROUP * 2
                                                                     : nullptr;
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
                }
                if (bs && ntot < MVMC_MAX_GROUP) {
                    bs[2 * ntot] = v;
                    bs[2 * ntot + 1] = ip[g];
                }
                ntot++;
            }
            // more poses than the fast path stages: solved from all of them by the many-pose birth solver if an overflow slot
            // is free (and the group fits MVMC_MAX_GROUP), else cut to the first MVMC_MAX_SEL and counted (n_truncated)
            const bool big = ntot > MVMC_MAX_SEL && bs != nullptr && ntot <= MVMC_MAX_GROUP && t_idx < 0 && n_new < max_new;
            over = ntot > MVMC_MAX_SEL && !big;
            if (over) n_trunc++;
            if (nsel > 0) {
                if (t_idx >= 0) {
                    tn[t_idx] = nsel;
                    for (int q = 0; q < nsel; q++) {
                        ts[(t_idx * MVMC_MAX_SEL + q) * 2] = sel[q][0];
                        ts[(t_idx * MVMC_MAX_SEL + q) * 2 + 1] = sel[q][1];
                    }
                } else if (nsel < 2) {
                    // a 2D-only group the one-pose-per-view rule shrank to one pose: listed by the reference, never born
                    if (singles && n_single < max_new) {
                        int* sg = singles + ((size_t)b * max_new + n_single) * 3;
                        sg[0] = seq;
                        sg[1] = sel[0][0];
                        sg[2] = sel[0][1];
                    }
                    n_single++;
                    seq++;
                } else if (n_new < max_new) {
                    if (new_seq) new_seq[(size_t)b * max_new + n_new] = seq;
                    seq++;
                    nn_[n_new] = big ? ntot : nsel;   // (> MVMC_MAX_SEL: the list is in big_sel, new_sel holds its first poses)
                    if (big) {
                        big_nsel[b * MVMC_MAX_BIG + n_big] = ntot;
                        big_slot[b * MVMC_MAX_BIG + n_big] = n_new;
                        n_big++;
                    }
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
        if (big_n) big_n[b] = n_big;
    }
}

// A6, first half on its own (the drop-in `mv_association.transform_closure` / `match_als` seams return this matrix):
// match[j][i] = 1 for every unvisited leader i and every j with temp[i][j], temp = X | X[:, n-1] (x) X[n-1, :].
// One warp per instance, lane w owns bit-word w of a row; output as bytes [N][N] (leading n x n block written).
__global__ void __launch_bounds__(32)
    k_closure(const uint32_t* __restrict__ xbin, const int* __restrict__ n_of, int N, uint8_t* __restrict__ match) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const int NW = (N + 31) / 32;
    const int n = min(n_of[b], N);
    const uint32_t* xb = xbin + (size_t)b * N * NW;
    uint8_t* M = match + (size_t)b * N * N;
    for (int e = lane; e < n * n; e += 32) M[(size_t)(e / n) * N + (e % n)] = 0;
    __syncwarp();
    if (n <= 0) return;
    const uint32_t last_row = (lane < NW) ? xb[(size_t)(n - 1) * NW + lane] : 0u;
    uint32_t vis = 0;
    for (int i = 0; i < n; i++) {
        const uint32_t vw = __shfl_sync(MVMC_FULL, vis, i >> 5);
        if ((vw >> (i & 31)) & 1u) continue;
        uint32_t row = (lane < NW) ? xb[(size_t)i * NW + lane] : 0u;
        const uint32_t lw = __shfl_sync(MVMC_FULL, row, (n - 1) >> 5);
        if ((lw >> ((n - 1) & 31)) & 1u) row |= last_row;
        vis |= row;
        uint32_t x = row;
        while (x) {
            const int j = lane * 32 + __ffs((int)x) - 1;
            if (j < n) M[(size_t)j * N + i] = 1;
            x &= x - 1;
        }
    }
}

}  // namespace mvmc

using namespace mvmc;

extern "C" int mvmc_transform_closure(const uint32_t* xbin, const int* n, int B, int N, uint8_t* match_mat, void* stream) {
    if (!xbin || !n || !match_mat || B <= 0 || N <= 0 || N > 1024) return MVMC_ERR_INVALID;
    MVMC_LAUNCH(k_closure, dim3(B), dim3(32), 0, stream, xbin, n, N, match_mat);
    MVMC_CHECK_LAUNCH("k_closure");
    return MVMC_OK;
}

extern "C" int mvmc_assign_groups(const uint32_t* xbin, const int* dim_groups, const int* idx_view, const int* idx_pose,
                                  const int* n_trk, int B, int C, int N, int Tmax, int max_new, int* trk_nsel, int* trk_sel,
                                  int* new_n, int* new_nsel, int* new_sel, int* counts, int* err, int* new_seq, int* singles,
                                  int* big_n, int* big_nsel, int* big_sel, int* big_slot, void* stream) {
    if ((big_n != nullptr) != (big_sel != nullptr) || (big_n != nullptr) != (big_nsel != nullptr) ||
        (big_n != nullptr) != (big_slot != nullptr))
        return MVMC_ERR_INVALID;
    if (!xbin || !dim_groups || !idx_view || !idx_pose || !n_trk || !trk_nsel || !trk_sel || !new_n || !new_nsel ||
        !new_sel || !counts || !err)
        return MVMC_ERR_INVALID;
    if (B <= 0 || N <= 0 || N > MVMC_MAX_TRACKS + MVMC_MAX_VIEWS * MVMC_MAX_POSES || C <= 0 || C > MVMC_MAX_VIEWS || Tmax < 0 ||
        Tmax > MVMC_MAX_TRACKS || max_new <= 0)
        return MVMC_ERR_INVALID;   // (k_assign lists a group's members in shared memory sized for the largest layout)
    MVMC_LAUNCH(k_assign, dim3(B), dim3(32), 0, stream, xbin, dim_groups, idx_view, idx_pose, n_trk, C, N, Tmax, max_new,
                trk_nsel, trk_sel, new_n, new_nsel, new_sel, counts, err, new_seq, singles, big_n, big_nsel, big_sel, big_slot);
    MVMC_CHECK_LAUNCH("k_assign");
    return MVMC_OK;
}

extern "C" int mvmc_assign_listed(const uint32_t* xbin, const int* dim_groups, const int* idx_view, const int* idx_pose,
                                  const int* n_trk, int B, int C, int N, int Tmax, int max_new, int* trk_nsel, int* trk_sel,
                                  int* new_n, int* new_nsel, int* new_sel, int* counts, int* err, int* new_seq, int* singles,
                                  void* stream) {
    return mvmc_assign_groups(xbin, dim_groups, idx_view, idx_pose, n_trk, B, C, N, Tmax, max_new, trk_nsel, trk_sel, new_n,
                              new_nsel, new_sel, counts, err, new_seq, singles, nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int mvmc_assign(const uint32_t* xbin, const int* dim_groups, const int* idx_view, const int* idx_pose,
                           const int* n_trk, int B, int C, int N, int Tmax, int max_new, int* trk_nsel, int* trk_sel,
                           int* new_n, int* new_nsel, int* new_sel, int* counts, int* err, void* stream) {
    return mvmc_assign_listed(xbin, dim_groups, idx_view, idx_pose, n_trk, B, C, N, Tmax, max_new, trk_nsel, trk_sel, new_n,
                              new_nsel, new_sel, counts, err, nullptr, nullptr, stream);
}
