// Clip-batch pipeline: B independent clips advance one frame per step with a fixed sequence of kernel
// launches and no host synchronisation (CUDA-graph capturable). Track lifecycle lives on the device.
//
// Reference rows: L1 motion_capture.py:288-400 (MvTracklet/TrackState) and
// :873-963 (MvTracker.update_4d); the run loop :1046-1129 is the caller (one step per frame).
#include "mvmc_common.cuh"

#include <atomic>
#include <mutex>
#include <new>
#include <string.h>
#include <string>
#include <vector>

int mvmc_ensure_skeleton();
int mvmc_als_order(const int* prev_iter, int B, int* order, void* stream);
int mvmc_match_als_ordered(const double* sim, const int* dim_groups, int n_groups, const int* f32_first_iter,
                           const double* rand_stream, const int* order, int B, int N, int rmax, void* workspace, uint32_t* xbin,
                           int* n_iter, void* stream);
int mvmc_assign_groups(const uint32_t* xbin, const int* dim_groups, const int* idx_view, const int* idx_pose, const int* n_trk, int B,
                       int C, int N, int Tmax, int max_new, int* trk_nsel, int* trk_sel, int* new_n, int* new_nsel, int* new_sel,
                       int* counts, int* err, int* new_seq, int* singles, int* big_n, int* big_nsel, int* big_sel, int* big_slot,
                       void* stream);
size_t mvmc_ik_birth_big_workspace_bytes(void);
int mvmc_ik_birth_big(const double* kps, const double* P, const int* big_n, const int* big_nsel, const int* big_sel,
                      const int* big_slot, int B, int C, int Pmax, int G, int S, int slot0, int max_nfev, void* workspace,
                      double* x_out, double* joints, int* info, double* cost, void* stream);
int mvmc_ik_launch(const double* kps2d, const double* Psel, const int* n_views, const double* x0, const uint8_t* birth,
                   const int* max_nfev, const uint8_t* free_mask, int n_items, int cnt, int S, int s0, int V, int vmax,
                   int* counter, double* x_out, double* joints, int* info, double* cost, void* stream);

#define MVMC_N_STATS 8
#define MVMC_N_STAGES 5
#define MVMC_PROF_STEPS 512

// ------------------------------------------------------------------------------------------------
// library-level state
// ------------------------------------------------------------------------------------------------
static std::atomic<unsigned long long> g_launches{0};
static std::mutex g_err_mu;
static std::string g_last_cuda_error = "";

void mvmc_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int mvmc_set_cuda_error(cudaError_t e, const char* where) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_last_cuda_error = std::string(where) + ": " + cudaGetErrorString(e);
    return MVMC_ERR_CUDA;
}

extern "C" unsigned long long mvmc_launch_count(void) {
#ifdef MVMC_EMU
    return emu::g_launches;
#else
    return g_launches.load();
#endif
}
extern "C" int mvmc_version(void) { return 100; }
extern "C" const char* mvmc_last_cuda_error(void) { return g_last_cuda_error.c_str(); }
extern "C" const char* mvmc_error_string(int code) {
    switch (code) {
        case MVMC_OK: return "ok";
        case MVMC_ERR_INVALID: return "invalid argument";
        case MVMC_ERR_CUDA: return "CUDA error";
        case MVMC_ERR_CAPACITY: return "per-clip capacity exceeded";
        case MVMC_ERR_NO_DEVICE: return "no CUDA device";
        default: return "unknown error";
    }
}

// numpy.random.RandomState(0).rand(): MT19937 seeded by init_genrand(0), 53-bit doubles
extern "C" int mvmc_rand_stream_host(double* out, int n) {
    if (!out || n < 0) return MVMC_ERR_INVALID;
    uint32_t mt[624];
    mt[0] = 0u;
    for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    int idx = 624;
    auto next = [&]() -> uint32_t {
        if (idx >= 624) {
            for (int k = 0; k < 624; k++) {
                const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    };
    for (int i = 0; i < n; i++) {
        const uint32_t a = next() >> 5, b = next() >> 6;
        out[i] = (a * 67108864.0 + b) / 9007199254740992.0;
    }
    return MVMC_OK;
}

// ------------------------------------------------------------------------------------------------
// device-side lifecycle kernels
// ------------------------------------------------------------------------------------------------
namespace mvmc {

struct ClipState {
    int* n_trk;      // [B]
    int* next_id;    // [B]
    int* id;         // [B,Tmax]
    int* state;
    int* hits;
    int* tsu;
    int* length;
    double* param;   // [B,Tmax,68]
    double* joints;  // [B,Tmax,54]
};

// predict(): time_since_update += 1 for every alive track (motion_capture.py:874-875); also raises the
// float32-first-iteration flag of the matcher when the clip has no alive track (A7 path).
__global__ void k_predict(ClipState st, int B, int Tmax, int* __restrict__ f32_flag) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int T = st.n_trk[b];
    for (int t = 0; t < T; t++) st.tsu[b * Tmax + t] += 1;
    f32_flag[b] = (T == 0) ? 1 : 0;
}

// Build the IK work slots of one clip: slot t < Tmax = update of alive track t, slot Tmax+k = birth k.
__global__ void __launch_bounds__(128)
    k_gather(const double* __restrict__ kps, const double* __restrict__ P, ClipState st, const int* __restrict__ trk_nsel,
             const int* __restrict__ trk_sel, const int* __restrict__ new_n, const int* __restrict__ new_nsel,
             const int* __restrict__ new_sel, int C, int Pmax, int Tmax, int max_new, int nfev_update, int nfev_birth,
             double* __restrict__ w_kps, double* __restrict__ w_P, int* __restrict__ w_nv, double* __restrict__ w_x0,
             uint8_t* __restrict__ w_birth, int* __restrict__ w_nfev) {
    const int b = blockIdx.x;
    const int S = Tmax + max_new;
    const int T = st.n_trk[b];
    const int nb = new_n[b];
    for (int slot = 0; slot < S; slot++) {
        const size_t m = (size_t)b * S + slot;
        int nsel = 0;
        const int* sel = nullptr;
        bool birth = false;
        if (slot < Tmax) {
            if (slot < T && trk_nsel[b * Tmax + slot] >= 2) {
                nsel = trk_nsel[b * Tmax + slot];
                sel = trk_sel + ((size_t)b * Tmax + slot) * MVMC_MAX_SEL * 2;
            }
        } else {
            const int k = slot - Tmax;
            if (k < nb && new_nsel[b * max_new + k] >= 2) {
                nsel = min(new_nsel[b * max_new + k], MVMC_MAX_SEL);   // (a many-pose group: its first poses; solved again from all)
                sel = new_sel + ((size_t)b * max_new + k) * MVMC_MAX_SEL * 2;
                birth = true;
            }
        }
        if (threadIdx.x == 0) {
            w_nv[m] = nsel;
            w_birth[m] = birth ? 1 : 0;
            w_nfev[m] = birth ? nfev_birth : nfev_update;
        }
        if (nsel == 0) continue;
        for (int e = threadIdx.x; e < nsel * MVMC_N_COCO * 3; e += blockDim.x) {
            const int q = e / (MVMC_N_COCO * 3), r = e % (MVMC_N_COCO * 3);
            const int v = sel[q * 2], p = sel[q * 2 + 1];
            w_kps[(m * MVMC_MAX_SEL + q) * (MVMC_N_COCO * 3) + r] = kps[((size_t)(b * C + v) * Pmax + p) * (MVMC_N_COCO * 3) + r];
        }
        for (int e = threadIdx.x; e < nsel * 12; e += blockDim.x) {
            const int q = e / 12, r = e % 12;
            w_P[(m * MVMC_MAX_SEL + q) * 12 + r] = P[(size_t)(b * C + sel[q * 2]) * 12 + r];
        }
        if (!birth)
            for (int e = threadIdx.x; e < MVMC_N_PARAM; e += blockDim.x)
                w_x0[m * MVMC_N_PARAM + e] = st.param[((size_t)b * Tmax + slot) * MVMC_N_PARAM + e];
    }
}

// Lifecycle commit (motion_capture.py:920-963): matched tracks take their new pose, unmatched ones are
// marked missed, births are appended in group order, dead tracks are dropped (stable), records written.
__global__ void __launch_bounds__(128)
    k_commit(ClipState st, const int* __restrict__ trk_nsel, const int* __restrict__ trk_sel, const int* __restrict__ new_n,
             const int* __restrict__ new_nsel, const int* __restrict__ new_sel, const int* __restrict__ n_dup,
             const int* __restrict__ assign_err, const int* __restrict__ dim_groups, const int* __restrict__ als_iter,
             const double* __restrict__ x_out, const double* __restrict__ j_out, const int* __restrict__ info,
             const double* __restrict__ cost, int C, int Tmax, int max_new, int n_inits, int max_age, int frame_idx,
             mvmc_step_out* __restrict__ out, double* __restrict__ stats) {
    __shared__ int s_src[MVMC_MAX_TRACKS];   // for output slot o: source (old slot t, or Tmax+k for a birth)
    __shared__ int s_upd[MVMC_MAX_TRACKS];
    __shared__ int s_n;
    const int b = blockIdx.x;
    const int S = Tmax + max_new;
    mvmc_step_out& rec = out[b];
    if (threadIdx.x == 0) {
        const int T = st.n_trk[b];
        int n_out = 0, n_died = 0, error = assign_err[b];
        for (int t = 0; t < T; t++) {
            const int o = b * Tmax + t;
            const int nsel = trk_nsel[o];
            bool dead = false;
            int upd = 0;
            if (nsel >= 0) {
                if (nsel >= 2) {
                    upd = 1;
                    st.tsu[o] = 0;
                    st.hits[o] += 1;
                    st.length[o] += 1;
                    if (st.state[o] == 1 && st.hits[o] >= n_inits) st.state[o] = 2;
                }
            } else {
                if (st.state[o] == 1) dead = true;
                else if (st.tsu[o] > max_age) dead = true;
            }
            if (dead) rec.died_ids[n_died++] = st.id[o];
            else {
                s_src[n_out] = t;
                s_upd[n_out] = upd;
                n_out++;
            }
        }
        const int nb = new_n[b];
        for (int k = 0; k < nb; k++) {
            if (new_nsel[b * max_new + k] < 2) continue;
            if (n_out >= Tmax) {
                error = MVMC_ERR_CAPACITY;
                break;
            }
            s_src[n_out] = Tmax + k;
            s_upd[n_out] = 2;
            n_out++;
        }
        s_n = n_out;
        rec.frame_idx = frame_idx;
        rec.n_alive = n_out;
        rec.n_died = n_died;
        rec.n_total = dim_groups[b * (C + 2) + C + 1];
        rec.als_iters = als_iter[b];
        rec.n_dup_view = n_dup[4 * b];
        rec.n_truncated = n_dup[4 * b + 2];
        rec.error = error;
        // algorithmic work counters: ALS flops = I (6 r n^2 + 8 r^2 n + 4 r^3)
        const int* dg = dim_groups + b * (C + 2);
        const double n = dg[C + 1];
        int maxsz = 0;
        for (int q = 0; q < C + 1; q++) maxsz = max(maxsz, dg[q + 1] - dg[q]);
        const double r = min((double)n, 2.0 * maxsz);
        const double it = als_iter[b];
        atomicAdd(&stats[0], it * (6.0 * r * n * n + 8.0 * r * r * n + 4.0 * r * r * r));
        atomicAdd(&stats[1], it);
        atomicAdd(&stats[2], 1.0);
        double solves = 0, nfev = 0, njev = 0, ikflops = 0;
        for (int o = 0; o < n_out; o++) {
            if (!s_upd[o]) continue;
            const size_t slot = (size_t)b * S + s_src[o];
            const int* sel_n = (s_upd[o] == 1) ? trk_nsel + b * Tmax + s_src[o] : new_nsel + b * max_new + (s_src[o] - Tmax);
            const double V = *sel_n, m = 32.0 * V;
            for (int q = 0; q < 2; q++) {
                const double np_ = info[slot * 8 + q * 4 + 3], fe = info[slot * 8 + q * 4], je = info[slot * 8 + q * 4 + 1];
                const double fres = 18 * (3 * 40 + 2 * 16 + 30) + 17 * 45 + V * 16 * 30;
                solves += 1;
                nfev += fe;
                njev += je;
                ikflops += (fe + je * np_) * fres + je * (2.0 * m * np_ * np_ + 10.0 * np_ * np_ * np_);
            }
        }
        atomicAdd(&stats[3], solves);
        atomicAdd(&stats[4], nfev);
        atomicAdd(&stats[5], njev);
        atomicAdd(&stats[6], ikflops);
        atomicAdd(&stats[7], n * n);
    }
    __syncthreads();
    const int n_out = s_n;
    // phase 1: fill the records from the old state / the solver outputs
    for (int o = 0; o < n_out; o++) {
        mvmc_track_out& tr = rec.tracks[o];
        const int src = s_src[o], upd = s_upd[o];
        const size_t slot = (size_t)b * S + src;
        const double* psrc = upd ? x_out + slot * MVMC_N_PARAM : st.param + ((size_t)b * Tmax + src) * MVMC_N_PARAM;
        const double* jsrc = upd ? j_out + slot * 54 : st.joints + ((size_t)b * Tmax + src) * 54;
        for (int e = threadIdx.x; e < MVMC_N_PARAM; e += blockDim.x) tr.param[e] = psrc[e];
        for (int e = threadIdx.x; e < 54; e += blockDim.x) tr.joints[e] = jsrc[e];
        if (threadIdx.x == 0) {
            if (upd == 2) {
                tr.track_id = st.next_id[b]++;
                tr.state = 1;
                tr.hits = 1;
                tr.time_since_update = 0;
                tr.length = 1;
            } else {
                const int q = b * Tmax + src;
                tr.track_id = st.id[q];
                tr.state = st.state[q];
                tr.hits = st.hits[q];
                tr.time_since_update = st.tsu[q];
                tr.length = st.length[q];
            }
            tr.updated = upd;
            const int* sel = nullptr;
            int nsel = 0;
            if (upd == 1) {
                nsel = trk_nsel[b * Tmax + src];
                sel = trk_sel + ((size_t)b * Tmax + src) * MVMC_MAX_SEL * 2;
            } else if (upd == 2) {
                nsel = new_nsel[b * max_new + (src - Tmax)];
                sel = new_sel + ((size_t)b * max_new + (src - Tmax)) * MVMC_MAX_SEL * 2;
            } else if (trk_nsel[b * Tmax + src] == 1) {
                // single-view match: reported, not solved (motion_capture.py:929-932)
                nsel = 1;
                sel = trk_sel + ((size_t)b * Tmax + src) * MVMC_MAX_SEL * 2;
            }
            tr.n_sel = nsel;
            for (int q = 0; q < MVMC_MAX_SEL; q++) {
                tr.sel[q][0] = q < nsel ? sel[q * 2] : -1;
                tr.sel[q][1] = q < nsel ? sel[q * 2 + 1] : -1;
            }
            for (int q = 0; q < 2; q++) {
                tr.nfev[q] = upd ? info[slot * 8 + q * 4] : 0;
                tr.njev[q] = upd ? info[slot * 8 + q * 4 + 1] : 0;
                tr.status[q] = upd ? info[slot * 8 + q * 4 + 2] : 0;
                tr.cost[q] = upd ? cost[slot * 2 + q] : 0.0;
            }
            tr.pad_ = 0;
        }
    }
    __syncthreads();
    // phase 2: the records become the new alive table
    for (int o = 0; o < n_out; o++) {
        const mvmc_track_out& tr = rec.tracks[o];
        const size_t q = (size_t)b * Tmax + o;
        for (int e = threadIdx.x; e < MVMC_N_PARAM; e += blockDim.x) st.param[q * MVMC_N_PARAM + e] = tr.param[e];
        for (int e = threadIdx.x; e < 54; e += blockDim.x) st.joints[q * 54 + e] = tr.joints[e];
        if (threadIdx.x == 0) {
            st.id[q] = tr.track_id;
            st.state[q] = tr.state;
            st.hits[q] = tr.hits;
            st.tsu[q] = tr.time_since_update;
            st.length[q] = tr.length;
        }
    }
    if (threadIdx.x == 0) st.n_trk[b] = n_out;
}

// FP64 DFMA peak probe: 8 independent FMA chains per thread, `iters` x 8 x 2 flops per thread.
__global__ void __launch_bounds__(256) k_fp64_probe(int iters, double* __restrict__ sink) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c);
        a1 = fma(a1, m, c);
        a2 = fma(a2, m, c);
        a3 = fma(a3, m, c);
        a4 = fma(a4, m, c);
        a5 = fma(a5, m, c);
        a6 = fma(a6, m, c);
        a7 = fma(a7, m, c);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 1.2345) sink[0] = r;  // never true; keeps the chains alive
}

// FP64 tensor-core (DMMA m8n8k4) peak probe: 8 independent accumulator fragments per warp, `iters` x 8 x 512 flops per warp.
__global__ void __launch_bounds__(256) k_fp64_tensor_probe(int iters, double* __restrict__ sink) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
    const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
#ifdef MVMC_EMU
            emu::dmma_884(c[i][0], c[i][1], a, b);
#else
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
#endif
        }
    }
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += c[i][0] + c[i][1];
    if (r == 1.2345) sink[0] = r;  // never true; keeps the chains alive
}

}  // namespace mvmc

using namespace mvmc;

extern "C" int mvmc_fp64_tensor_probe(int blocks, int iters, double* sink, void* stream) {
    if (blocks <= 0 || iters <= 0 || !sink) return MVMC_ERR_INVALID;
    MVMC_LAUNCH(k_fp64_tensor_probe, dim3(blocks), dim3(256), 0, stream, iters, sink);
    MVMC_CHECK_LAUNCH("k_fp64_tensor_probe");
    return MVMC_OK;
}

extern "C" int mvmc_fp64_probe(int blocks, int iters, double* sink, void* stream) {
    if (blocks <= 0 || iters <= 0 || !sink) return MVMC_ERR_INVALID;
    MVMC_LAUNCH(k_fp64_probe, dim3(blocks), dim3(256), 0, stream, iters, sink);
    MVMC_CHECK_LAUNCH("k_fp64_probe");
    return MVMC_OK;
}

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
struct mvmc_clips {
    mvmc_config cfg;
    int B, C, Pmax, Tmax, N, NW, rmax, S;
    size_t bytes = 0;
    std::vector<void*> allocs;
    // calibration
    double *K = nullptr, *Rt = nullptr, *P = nullptr, *F = nullptr;
    float* F32 = nullptr;
    // state
    ClipState st;
    // per-step work
    double *kps_in = nullptr, *kps25_in = nullptr;   // (kps25_in: BODY_25 staging of mvmc_clips_step_body25_host, allocated on first use)
    int *npose_in = nullptr, *npeople_in = nullptr;
    uint8_t* keep = nullptr;
    int *dim_groups = nullptr, *idx_view = nullptr, *idx_pose = nullptr, *f32_flag = nullptr;
    double *dst = nullptr, *sim = nullptr, *rand_stream = nullptr;
    void* als_ws = nullptr;
    uint32_t* xbin = nullptr;
    int *als_iter = nullptr, *als_order = nullptr;
    int *trk_nsel = nullptr, *trk_sel = nullptr, *new_n = nullptr, *new_nsel = nullptr, *new_sel = nullptr, *n_dup = nullptr,
        *assign_err = nullptr;
    double *w_kps = nullptr, *w_P = nullptr, *w_x0 = nullptr, *w_xout = nullptr, *w_joints = nullptr, *w_cost = nullptr;
    int *w_nv = nullptr, *w_nfev = nullptr, *w_info = nullptr;
    uint8_t* w_birth = nullptr;
    void* ik_ws = nullptr;
    int *big_n = nullptr, *big_nsel = nullptr, *big_sel = nullptr, *big_slot = nullptr;   // many-pose birth groups of the step
    void* big_ws = nullptr;
    mvmc_step_out* out = nullptr;
    double* stats = nullptr;  // [MVMC_N_STATS] device counters
    // side stream for the birth solves (a handful per step, each ~10x an update: alone they are a ~5 ms latency tail)
#ifndef MVMC_EMU
    cudaStream_t birth_stream = nullptr, ik_stream = nullptr;   // both at the highest stream priority (see mvmc_clips_step)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join_ik = nullptr;
#endif
    // optional per-stage event timing
    int profiling = 0;
    int ev_step = 0;
    std::vector<cudaEvent_t> events;  // [MVMC_PROF_STEPS][MVMC_N_STAGES+1]
    double stage_ms[MVMC_N_STAGES] = {0, 0, 0, 0, 0};

    template <class T>
    int alloc(T** p, size_t count) {
        void* q = nullptr;
        const size_t nbytes = count * sizeof(T);
        cudaError_t e = cudaMalloc(&q, nbytes ? nbytes : 16);
        if (e != cudaSuccess) return mvmc_set_cuda_error(e, "cudaMalloc");
        e = cudaMemset(q, 0, nbytes ? nbytes : 16);
        if (e != cudaSuccess) return mvmc_set_cuda_error(e, "cudaMemset");
        allocs.push_back(q);
        bytes += nbytes;
        *p = (T*)q;
        return MVMC_OK;
    }
};

extern "C" void mvmc_default_config(mvmc_config* cfg) {
    if (!cfg) return;
    cfg->n_clips = 1;
    cfg->n_views = 5;
    cfg->max_poses = 8;
    cfg->max_tracks = 24;
    cfg->max_new = 8;
    cfg->n_inits = 3;
    cfg->max_age = 0;
    cfg->nfev_update = 5;
    cfg->nfev_birth = 50;
    cfg->keep_matrices = 1;
}

extern "C" void mvmc_clips_destroy(mvmc_clips* h) {
    if (!h) return;
#ifndef MVMC_EMU
    for (cudaEvent_t e : h->events) cudaEventDestroy(e);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_join_ik) cudaEventDestroy(h->ev_join_ik);
    if (h->birth_stream) cudaStreamDestroy(h->birth_stream);
    if (h->ik_stream) cudaStreamDestroy(h->ik_stream);
#endif
    for (void* p : h->allocs) cudaFree(p);
    delete h;
}

extern "C" size_t mvmc_clips_device_bytes(const mvmc_clips* h) { return h ? h->bytes : 0; }

#define TRY(x)                    \
    do {                          \
        int rc__ = (x);           \
        if (rc__ != MVMC_OK) {    \
            mvmc_clips_destroy(h); \
            return rc__;          \
        }                         \
    } while (0)

extern "C" int mvmc_clips_create(const mvmc_config* cfg, mvmc_clips** out) {
    if (!cfg || !out) return MVMC_ERR_INVALID;
    if (cfg->n_clips <= 0 || cfg->n_views <= 0 || cfg->n_views > MVMC_MAX_VIEWS || cfg->max_poses <= 0 ||
        cfg->max_poses > MVMC_MAX_POSES || cfg->max_tracks <= 0 || cfg->max_tracks > MVMC_MAX_TRACKS || cfg->max_new <= 0 ||
        cfg->max_new > 64 || cfg->nfev_update < 1 || cfg->nfev_birth < 1)
        return MVMC_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return MVMC_ERR_NO_DEVICE;
    mvmc_clips* h = new (std::nothrow) mvmc_clips();
    if (!h) return MVMC_ERR_INVALID;
    h->cfg = *cfg;
    const int B = h->B = cfg->n_clips, C = h->C = cfg->n_views, Pmax = h->Pmax = cfg->max_poses,
              Tmax = h->Tmax = cfg->max_tracks;
    const int N = h->N = Tmax + C * Pmax;
    h->NW = (N + 31) / 32;
    int rmax = 2 * (Tmax > Pmax ? Tmax : Pmax);
    if (rmax > N) rmax = N;
    if (rmax > 128) rmax = 128;
    h->rmax = rmax;
    const int S = h->S = Tmax + cfg->max_new;
    const size_t M = (size_t)B * S;
    TRY(mvmc_ensure_skeleton());
    TRY(h->alloc(&h->K, (size_t)B * C * 9));
    TRY(h->alloc(&h->Rt, (size_t)B * C * 12));
    TRY(h->alloc(&h->P, (size_t)B * C * 12));
    TRY(h->alloc(&h->F, (size_t)B * C * C * 9));
    TRY(h->alloc(&h->F32, (size_t)B * C * C * 9));
    TRY(h->alloc(&h->st.n_trk, (size_t)B));
    TRY(h->alloc(&h->st.next_id, (size_t)B));
    TRY(h->alloc(&h->st.id, (size_t)B * Tmax));
    TRY(h->alloc(&h->st.state, (size_t)B * Tmax));
    TRY(h->alloc(&h->st.hits, (size_t)B * Tmax));
    TRY(h->alloc(&h->st.tsu, (size_t)B * Tmax));
    TRY(h->alloc(&h->st.length, (size_t)B * Tmax));
    TRY(h->alloc(&h->st.param, (size_t)B * Tmax * MVMC_N_PARAM));
    TRY(h->alloc(&h->st.joints, (size_t)B * Tmax * 54));
    TRY(h->alloc(&h->kps_in, (size_t)B * C * Pmax * MVMC_N_COCO * 3));
    TRY(h->alloc(&h->npose_in, (size_t)B * C));
    TRY(h->alloc(&h->keep, (size_t)B * C * Pmax));
    TRY(h->alloc(&h->dim_groups, (size_t)B * (C + 2)));
    TRY(h->alloc(&h->idx_view, (size_t)B * N));
    TRY(h->alloc(&h->idx_pose, (size_t)B * N));
    TRY(h->alloc(&h->f32_flag, (size_t)B));
    TRY(h->alloc(&h->dst, (size_t)B * N * N));
    TRY(h->alloc(&h->sim, (size_t)B * N * N));
    TRY(h->alloc(&h->rand_stream, (size_t)N * rmax));
    {
        void* p = nullptr;
        const size_t nb = mvmc_match_als_workspace_bytes(B, N, rmax);
        cudaError_t e = cudaMalloc(&p, nb);
        if (e != cudaSuccess) {
            int rc = mvmc_set_cuda_error(e, "cudaMalloc(als workspace)");
            mvmc_clips_destroy(h);
            return rc;
        }
        h->allocs.push_back(p);
        h->bytes += nb;
        h->als_ws = p;
    }
    TRY(h->alloc(&h->xbin, (size_t)B * N * h->NW));
    TRY(h->alloc(&h->als_iter, (size_t)B));
    TRY(h->alloc(&h->als_order, (size_t)B));
    TRY(h->alloc(&h->trk_nsel, (size_t)B * Tmax));
    TRY(h->alloc(&h->trk_sel, (size_t)B * Tmax * MVMC_MAX_SEL * 2));
    TRY(h->alloc(&h->new_n, (size_t)B));
    TRY(h->alloc(&h->new_nsel, (size_t)B * cfg->max_new));
    TRY(h->alloc(&h->new_sel, (size_t)B * cfg->max_new * MVMC_MAX_SEL * 2));
    TRY(h->alloc(&h->n_dup, (size_t)B * 4));
    TRY(h->alloc(&h->assign_err, (size_t)B));
    TRY(h->alloc(&h->w_kps, M * MVMC_MAX_SEL * MVMC_N_COCO * 3));
    TRY(h->alloc(&h->w_P, M * MVMC_MAX_SEL * 12));
    TRY(h->alloc(&h->w_x0, M * MVMC_N_PARAM));
    TRY(h->alloc(&h->w_xout, M * MVMC_N_PARAM));
    TRY(h->alloc(&h->w_joints, M * 54));
    TRY(h->alloc(&h->w_cost, M * 2));
    TRY(h->alloc(&h->w_nv, M));
    TRY(h->alloc(&h->w_nfev, M));
    TRY(h->alloc(&h->w_info, M * 8));
    TRY(h->alloc(&h->w_birth, M));
    {
        void* p = nullptr;
        const size_t nb = mvmc_ik_workspace_bytes((int)M, MVMC_MAX_SEL);
        cudaError_t e = cudaMalloc(&p, nb);
        if (e != cudaSuccess) {
            int rc = mvmc_set_cuda_error(e, "cudaMalloc(ik workspace)");
            mvmc_clips_destroy(h);
            return rc;
        }
        h->allocs.push_back(p);
        h->bytes += nb;
        h->ik_ws = p;
    }
    TRY(h->alloc(&h->big_n, (size_t)B));
    TRY(h->alloc(&h->big_nsel, (size_t)B * MVMC_MAX_BIG));
    TRY(h->alloc(&h->big_slot, (size_t)B * MVMC_MAX_BIG));
    TRY(h->alloc(&h->big_sel, (size_t)B * MVMC_MAX_BIG * MVMC_MAX_GROUP * 2));
    {
        unsigned char* p = nullptr;
        TRY(h->alloc(&p, mvmc_ik_birth_big_workspace_bytes()));
        h->big_ws = p;
    }
    TRY(h->alloc(&h->out, (size_t)B));
    TRY(h->alloc(&h->stats, (size_t)MVMC_N_STATS));
    {
        std::vector<double> rs((size_t)N * rmax);
        mvmc_rand_stream_host(rs.data(), (int)rs.size());
        cudaError_t e = cudaMemcpy(h->rand_stream, rs.data(), rs.size() * sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            int rc = mvmc_set_cuda_error(e, "cudaMemcpy(rand_stream)");
            mvmc_clips_destroy(h);
            return rc;
        }
    }
#ifndef MVMC_EMU
    {
        // The IK solvers run on side streams of the HIGHEST priority: when several handles (clip groups) share the GPU, a
        // group's short, latency-bound solver CTAs are then placed ahead of the other groups' pending ALS CTAs as SM
        // resources free up, instead of waiting behind a whole ~50 ms wave of them (measured: without priorities the
        // overlapped throughput of three groups fell from 4.5 k to 3.4 k frames/s in a third of the runs).
        int prio_lo = 0, prio_hi = 0;
        cudaError_t e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->birth_stream, cudaStreamNonBlocking, prio_hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->ik_stream, cudaStreamNonBlocking, prio_hi);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join_ik, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            int rc = mvmc_set_cuda_error(e, "birth stream");
            mvmc_clips_destroy(h);
            return rc;
        }
    }
#endif
    *out = h;
    return MVMC_OK;
}

extern "C" int mvmc_clips_set_calib(mvmc_clips* h, const double* K, const double* Rt, const double* P, void* stream) {
    if (!h || !K || !Rt || !P) return MVMC_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t BC = (size_t)h->B * h->C;
    MVMC_CUDA_OK(cudaMemcpyAsync(h->K, K, BC * 9 * sizeof(double), cudaMemcpyDefault, s));
    MVMC_CUDA_OK(cudaMemcpyAsync(h->Rt, Rt, BC * 12 * sizeof(double), cudaMemcpyDefault, s));
    MVMC_CUDA_OK(cudaMemcpyAsync(h->P, P, BC * 12 * sizeof(double), cudaMemcpyDefault, s));
    int rc = mvmc_fundamental(h->P, h->F, h->B, h->C, stream);
    if (rc) return rc;
    return mvmc_fundamental_krt(h->K, h->Rt, h->F32, h->B, h->C, stream);
}

extern "C" int mvmc_clips_reset(mvmc_clips* h, void* stream) {
    if (!h) return MVMC_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    MVMC_CUDA_OK(cudaMemsetAsync(h->st.n_trk, 0, (size_t)h->B * sizeof(int), s));
    MVMC_CUDA_OK(cudaMemsetAsync(h->st.next_id, 0, (size_t)h->B * sizeof(int), s));
    return MVMC_OK;
}

extern "C" int mvmc_clips_step(mvmc_clips* h, const double* kps, const int* n_pose, int frame_idx, void* stream) {
    if (!h || !kps || !n_pose) return MVMC_ERR_INVALID;
    const int B = h->B, C = h->C, Pmax = h->Pmax, Tmax = h->Tmax, N = h->N;
    int ev = -1;
#ifndef MVMC_EMU
    if (h->profiling && h->ev_step < MVMC_PROF_STEPS) ev = (h->ev_step++) * (MVMC_N_STAGES + 1);
#define MVMC_EV(i) do { if (ev >= 0) cudaEventRecord(h->events[ev + (i)], (cudaStream_t)stream); } while (0)
#else
#define MVMC_EV(i) do { (void)ev; } while (0)
#endif
    MVMC_EV(0);
    MVMC_LAUNCH(k_predict, dim3((B + 127) / 128), dim3(128), 0, stream, h->st, B, Tmax, h->f32_flag);
    MVMC_CHECK_LAUNCH("k_predict");
    int rc = mvmc_prepare(kps, n_pose, h->st.n_trk, B, C, Pmax, Tmax, h->keep, h->dim_groups, h->idx_view, h->idx_pose,
                          stream);
    if (rc) return rc;
    rc = mvmc_affinity(kps, h->P, h->F, h->F32, h->st.joints, h->st.n_trk, h->dim_groups, h->idx_view, h->idx_pose, B, C,
                       Pmax, Tmax, h->dst, h->sim, stream);
    if (rc) return rc;
    MVMC_EV(1);
    rc = mvmc_als_order(h->als_iter, B, h->als_order, stream);   // previous frame's iteration counts: longest solves first
    if (rc) return rc;
    rc = mvmc_match_als_ordered(h->sim, h->dim_groups, C + 1, h->f32_flag, h->rand_stream, h->als_order, B, N, h->rmax, h->als_ws,
                                h->xbin, h->als_iter, stream);
    if (rc) return rc;
    MVMC_EV(2);
    rc = mvmc_assign_groups(h->xbin, h->dim_groups, h->idx_view, h->idx_pose, h->st.n_trk, B, C, N, Tmax, h->cfg.max_new,
                            h->trk_nsel, h->trk_sel, h->new_n, h->new_nsel, h->new_sel, h->n_dup, h->assign_err, nullptr, nullptr,
                            h->big_n, h->big_nsel, h->big_sel, h->big_slot, stream);
    if (rc) return rc;
    MVMC_LAUNCH(k_gather, dim3(B), dim3(128), 0, stream, kps, h->P, h->st, h->trk_nsel, h->trk_sel, h->new_n, h->new_nsel,
                h->new_sel, C, Pmax, Tmax, h->cfg.max_new, h->cfg.nfev_update, h->cfg.nfev_birth, h->w_kps, h->w_P, h->w_nv,
                h->w_x0, h->w_birth, h->w_nfev);
    MVMC_CHECK_LAUNCH("k_gather");
    MVMC_EV(3);
    // track updates use one pose per view (<= C observations); births of no-track frames may group more (<= MVMC_MAX_SEL).
    // The two launches touch disjoint work slots; the births run on a side stream next to the updates.
    void* bstream = stream;
    void* ustream = stream;
#ifndef MVMC_EMU
    MVMC_CUDA_OK(cudaEventRecord(h->ev_fork, (cudaStream_t)stream));
    MVMC_CUDA_OK(cudaStreamWaitEvent(h->birth_stream, h->ev_fork, 0));
    MVMC_CUDA_OK(cudaStreamWaitEvent(h->ik_stream, h->ev_fork, 0));
    bstream = h->birth_stream;
    ustream = h->ik_stream;
#endif
    rc = mvmc_ik_launch(h->w_kps, h->w_P, h->w_nv, h->w_x0, h->w_birth, h->w_nfev, nullptr, B * h->cfg.max_new, h->cfg.max_new,
                        h->S, Tmax, MVMC_MAX_SEL, MVMC_MAX_SEL, (int*)h->ik_ws + 16, h->w_xout, h->w_joints, h->w_info,
                        h->w_cost, bstream);
    if (rc) return rc;
    // births from groups of more than MVMC_MAX_SEL poses (crowded no-track frames): solved again, from all their poses, by the
    // many-pose solver, which overwrites the slot the fast path filled from the first MVMC_MAX_SEL (148 idle warps otherwise)
    rc = mvmc_ik_birth_big(kps, h->P, h->big_n, h->big_nsel, h->big_sel, h->big_slot, B, C, Pmax, MVMC_MAX_BIG, h->S, Tmax,
                           h->cfg.nfev_birth, h->big_ws, h->w_xout, h->w_joints, h->w_info, h->w_cost, bstream);
    if (rc) return rc;
    rc = mvmc_ik_launch(h->w_kps, h->w_P, h->w_nv, h->w_x0, nullptr, h->w_nfev, nullptr, B * Tmax, Tmax, h->S, 0,
                        MVMC_MAX_SEL, C, (int*)h->ik_ws, h->w_xout, h->w_joints, h->w_info, h->w_cost, ustream);
    if (rc) return rc;
#ifndef MVMC_EMU
    MVMC_CUDA_OK(cudaEventRecord(h->ev_join, h->birth_stream));
    MVMC_CUDA_OK(cudaEventRecord(h->ev_join_ik, h->ik_stream));
    MVMC_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)stream, h->ev_join, 0));
    MVMC_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)stream, h->ev_join_ik, 0));
#endif
    MVMC_EV(4);
    MVMC_LAUNCH(k_commit, dim3(B), dim3(128), 0, stream, h->st, h->trk_nsel, h->trk_sel, h->new_n, h->new_nsel, h->new_sel,
                h->n_dup, h->assign_err, h->dim_groups, h->als_iter, h->w_xout, h->w_joints, h->w_info, h->w_cost, C, Tmax,
                h->cfg.max_new, h->cfg.n_inits, h->cfg.max_age, frame_idx, h->out, h->stats);
    MVMC_CHECK_LAUNCH("k_commit");
    MVMC_EV(5);
    return MVMC_OK;
}

// stats: [0] ALS flops, [1] ALS iterations, [2] clip-frames, [3] IK solves, [4] nfev, [5] njev, [6] IK flops (est.),
// [7] sum n^2 (affinity entries)
extern "C" int mvmc_clips_stats_host(mvmc_clips* h, double* out, int reset, void* stream) {
    if (!h || !out) return MVMC_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    MVMC_CUDA_OK(cudaStreamSynchronize(s));
    MVMC_CUDA_OK(cudaMemcpy(out, h->stats, MVMC_N_STATS * sizeof(double), cudaMemcpyDeviceToHost));
    if (reset) MVMC_CUDA_OK(cudaMemset(h->stats, 0, MVMC_N_STATS * sizeof(double)));
    return MVMC_OK;
}

// Per-stage device time from CUDA events recorded on the step's stream:
// out[0] prepare+affinity, [1] ALS matcher, [2] assign+gather, [3] IK (updates + births), [4] commit  (ms, summed
// over the steps since the last reset). enable: 1 start recording (resets), 0 stop, -1 just read.
extern "C" int mvmc_clips_profile(mvmc_clips* h, int enable, double* out_ms, int* n_steps, void* stream) {
    if (!h) return MVMC_ERR_INVALID;
#ifndef MVMC_EMU
    cudaStream_t s = (cudaStream_t)stream;
    if (enable == 1) {
        if (h->events.empty()) {
            h->events.resize((size_t)MVMC_PROF_STEPS * (MVMC_N_STAGES + 1));
            for (auto& e : h->events) MVMC_CUDA_OK(cudaEventCreate(&e));
        }
        h->profiling = 1;
        h->ev_step = 0;
        return MVMC_OK;
    }
    MVMC_CUDA_OK(cudaStreamSynchronize(s));
    if (out_ms) {
        for (int q = 0; q < MVMC_N_STAGES; q++) out_ms[q] = 0.0;
        for (int st = 0; st < h->ev_step; st++)
            for (int q = 0; q < MVMC_N_STAGES; q++) {
                float ms = 0.f;
                MVMC_CUDA_OK(cudaEventElapsedTime(&ms, h->events[(size_t)st * (MVMC_N_STAGES + 1) + q],
                                                  h->events[(size_t)st * (MVMC_N_STAGES + 1) + q + 1]));
                out_ms[q] += ms;
            }
    }
    if (n_steps) *n_steps = h->ev_step;
    if (enable == 0) h->profiling = 0;
#else
    (void)stream;
    if (out_ms) for (int q = 0; q < MVMC_N_STAGES; q++) out_ms[q] = 0.0;
    if (n_steps) *n_steps = 0;
    (void)enable;
#endif
    return MVMC_OK;
}

// Compact result records for gathering (per track-frame (clip, track id, frame, 68 params, 54 joint
// coordinates) = 1 KB): the tracks SOLVED in the last step of every clip, in track-list order.
//   rec [B,cap,128] doubles: [0] clip id (clip0 + b), [1] track id, [2] frame, [3] state, [4] hits, [5] views used,
//   [6..73] parameters, [74..127] joints;   count [B] (tracks solved; rows beyond min(count, cap) are zero)
__global__ void __launch_bounds__(128)
    k_pack_records(const mvmc_step_out* __restrict__ out, int cap, int clip0, double* __restrict__ rec, int* __restrict__ count) {
    __shared__ int s_row[MVMC_MAX_TRACKS];
    const int b = blockIdx.x;
    const mvmc_step_out& o = out[b];
    if (threadIdx.x == 0) {
        int n = 0;
        for (int t = 0; t < o.n_alive; t++) s_row[t] = o.tracks[t].updated > 0 ? n++ : -1;
        count[b] = n;
    }
    __syncthreads();
    double* R = rec + (size_t)b * cap * 128;
    for (int e = threadIdx.x; e < cap * 128; e += blockDim.x) R[e] = 0.0;
    __syncthreads();
    for (int t = 0; t < o.n_alive; t++) {
        const int r = s_row[t];
        if (r < 0 || r >= cap) continue;
        const mvmc_track_out& tr = o.tracks[t];
        double* q = R + (size_t)r * 128;
        for (int e = threadIdx.x; e < 128; e += blockDim.x) {
            double v;
            if (e == 0) v = (double)(clip0 + b);
            else if (e == 1) v = (double)tr.track_id;
            else if (e == 2) v = (double)o.frame_idx;
            else if (e == 3) v = (double)tr.state;
            else if (e == 4) v = (double)tr.hits;
            else if (e == 5) v = (double)tr.n_sel;
            else if (e < 6 + MVMC_N_PARAM) v = tr.param[e - 6];
            else v = tr.joints[e - 6 - MVMC_N_PARAM];
            q[e] = v;
        }
    }
}

extern "C" int mvmc_clips_pack_records(mvmc_clips* h, int cap, int clip0, double* rec, int* count, void* stream) {
    if (!h || !rec || !count || cap <= 0 || cap > MVMC_MAX_TRACKS) return MVMC_ERR_INVALID;
    MVMC_LAUNCH(k_pack_records, dim3(h->B), dim3(128), 0, stream, h->out, cap, clip0, rec, count);
    MVMC_CHECK_LAUNCH("k_pack_records");
    return MVMC_OK;
}

extern "C" size_t mvmc_sizeof_step_out(void) { return sizeof(mvmc_step_out); }

extern "C" const mvmc_step_out* mvmc_clips_last_out(const mvmc_clips* h) { return h ? h->out : nullptr; }

extern "C" int mvmc_clips_step_host_async(mvmc_clips* h, const double* kps_host, const int* n_pose_host, int frame_idx,
                                          mvmc_step_out* out_host, void* stream) {
    if (!h || !kps_host || !n_pose_host) return MVMC_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t nk = (size_t)h->B * h->C * h->Pmax * MVMC_N_COCO * 3;
    MVMC_CUDA_OK(cudaMemcpyAsync(h->kps_in, kps_host, nk * sizeof(double), cudaMemcpyHostToDevice, s));
    MVMC_CUDA_OK(cudaMemcpyAsync(h->npose_in, n_pose_host, (size_t)h->B * h->C * sizeof(int), cudaMemcpyHostToDevice, s));
    int rc = mvmc_clips_step(h, h->kps_in, h->npose_in, frame_idx, stream);
    if (rc) return rc;
    if (out_host)
        MVMC_CUDA_OK(cudaMemcpyAsync(out_host, h->out, (size_t)h->B * sizeof(mvmc_step_out), cudaMemcpyDeviceToHost, s));
    return MVMC_OK;
}

int mvmc_ingest_body25(const double* kps25, const int* n_people, int B, int C, int Pin, int Pmax, double* kps, int* n_pose,
                       void* stream);

// Ingest + step: raw OpenPose BODY_25 detections [B,C,Pmax,25,3] + people counts [B,C] (HOST, e.g. straight from
// mvmc_parse_openpose_files_host) -> device, BODY_25 -> COCO gather on the device, then the step. Asynchronous like
// mvmc_clips_step_host_async.
extern "C" int mvmc_clips_step_body25_host(mvmc_clips* h, const double* kps25_host, const int* n_people_host, int frame_idx,
                                           mvmc_step_out* out_host, void* stream) {
    if (!h || !kps25_host || !n_people_host) return MVMC_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n25 = (size_t)h->B * h->C * h->Pmax * 75;
    if (!h->kps25_in) {
        int rc = h->alloc(&h->kps25_in, n25);
        if (rc) return rc;
        rc = h->alloc(&h->npeople_in, (size_t)h->B * h->C);
        if (rc) return rc;
    }
    MVMC_CUDA_OK(cudaMemcpyAsync(h->kps25_in, kps25_host, n25 * sizeof(double), cudaMemcpyHostToDevice, s));
    MVMC_CUDA_OK(cudaMemcpyAsync(h->npeople_in, n_people_host, (size_t)h->B * h->C * sizeof(int), cudaMemcpyHostToDevice, s));
    int rc = mvmc_ingest_body25(h->kps25_in, h->npeople_in, h->B, h->C, h->Pmax, h->Pmax, h->kps_in, h->npose_in, stream);
    if (rc) return rc;
    rc = mvmc_clips_step(h, h->kps_in, h->npose_in, frame_idx, stream);
    if (rc) return rc;
    if (out_host)
        MVMC_CUDA_OK(cudaMemcpyAsync(out_host, h->out, (size_t)h->B * sizeof(mvmc_step_out), cudaMemcpyDeviceToHost, s));
    return MVMC_OK;
}

extern "C" int mvmc_clips_step_host(mvmc_clips* h, const double* kps_host, const int* n_pose_host, int frame_idx,
                                    mvmc_step_out* out_host, void* stream) {
    const int rc = mvmc_clips_step_host_async(h, kps_host, n_pose_host, frame_idx, out_host, stream);
    if (rc) return rc;
    MVMC_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    return MVMC_OK;
}

extern "C" int mvmc_clips_set_tracks_host(mvmc_clips* h, const int* n_trk, const int* ids, const int* state,
                                          const int* hits, const int* tsu, const int* length, const double* param,
                                          const double* joints, const int* next_id, void* stream) {
    if (!h || !n_trk || !ids || !state || !hits || !tsu || !length || !param || !joints || !next_id) return MVMC_ERR_INVALID;
    for (int b = 0; b < h->B; b++)
        if (n_trk[b] < 0 || n_trk[b] > h->Tmax) return MVMC_ERR_CAPACITY;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t BT = (size_t)h->B * h->Tmax;
    MVMC_CUDA_OK(cudaMemcpyAsync(h->st.n_trk, n_trk, (size_t)h->B * sizeof(int), cudaMemcpyHostToDevice, s));
    MVMC_CUDA_OK(cudaMemcpyAsync(h->st.next_id, next_id, (size_t)h->B * sizeof(int), cudaMemcpyHostToDevice, s));
    MVMC_CUDA_OK(cudaMemcpyAsync(h->st.id, ids, BT * sizeof(int), cudaMemcpyHostToDevice, s));
    MVMC_CUDA_OK(cudaMemcpyAsync(h->st.state, state, BT * sizeof(int), cudaMemcpyHostToDevice, s));
    MVMC_CUDA_OK(cudaMemcpyAsync(h->st.hits, hits, BT * sizeof(int), cudaMemcpyHostToDevice, s));
    MVMC_CUDA_OK(cudaMemcpyAsync(h->st.tsu, tsu, BT * sizeof(int), cudaMemcpyHostToDevice, s));
    MVMC_CUDA_OK(cudaMemcpyAsync(h->st.length, length, BT * sizeof(int), cudaMemcpyHostToDevice, s));
    MVMC_CUDA_OK(cudaMemcpyAsync(h->st.param, param, BT * MVMC_N_PARAM * sizeof(double), cudaMemcpyHostToDevice, s));
    MVMC_CUDA_OK(cudaMemcpyAsync(h->st.joints, joints, BT * 54 * sizeof(double), cudaMemcpyHostToDevice, s));
    MVMC_CUDA_OK(cudaStreamSynchronize(s));
    return MVMC_OK;
}

extern "C" int mvmc_clips_read_big_groups_host(mvmc_clips* h, int b, int* n, int* nsel, int* slot, int* sel, void* stream) {
    if (!h || b < 0 || b >= h->B || !n || !nsel || !slot || !sel) return MVMC_ERR_INVALID;
    MVMC_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    MVMC_CUDA_OK(cudaMemcpy(n, h->big_n + b, sizeof(int), cudaMemcpyDeviceToHost));
    MVMC_CUDA_OK(cudaMemcpy(nsel, h->big_nsel + (size_t)b * MVMC_MAX_BIG, MVMC_MAX_BIG * sizeof(int), cudaMemcpyDeviceToHost));
    MVMC_CUDA_OK(cudaMemcpy(slot, h->big_slot + (size_t)b * MVMC_MAX_BIG, MVMC_MAX_BIG * sizeof(int), cudaMemcpyDeviceToHost));
    MVMC_CUDA_OK(cudaMemcpy(sel, h->big_sel + (size_t)b * MVMC_MAX_BIG * MVMC_MAX_GROUP * 2,
                            (size_t)MVMC_MAX_BIG * MVMC_MAX_GROUP * 2 * sizeof(int), cudaMemcpyDeviceToHost));
    return MVMC_OK;
}

extern "C" int mvmc_clips_read_matrices_host(mvmc_clips* h, int b, double* dst, double* sim, uint8_t* xbin, int* n_out,
                                             int* dim_groups, void* stream) {
    if (!h || b < 0 || b >= h->B || !n_out) return MVMC_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    MVMC_CUDA_OK(cudaStreamSynchronize(s));
    const int N = h->N, NW = h->NW, C = h->C;
    std::vector<int> dg(C + 2);
    MVMC_CUDA_OK(cudaMemcpy(dg.data(), h->dim_groups + (size_t)b * (C + 2), (C + 2) * sizeof(int), cudaMemcpyDeviceToHost));
    const int n = dg[C + 1];
    *n_out = n;
    if (dim_groups) memcpy(dim_groups, dg.data(), (C + 2) * sizeof(int));
    if (n <= 0) return MVMC_OK;
    std::vector<double> tmp((size_t)N * N);
    if (dst) {
        MVMC_CUDA_OK(cudaMemcpy(tmp.data(), h->dst + (size_t)b * N * N, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++) memcpy(dst + (size_t)i * n, tmp.data() + (size_t)i * N, n * sizeof(double));
    }
    if (sim) {
        MVMC_CUDA_OK(cudaMemcpy(tmp.data(), h->sim + (size_t)b * N * N, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++) memcpy(sim + (size_t)i * n, tmp.data() + (size_t)i * N, n * sizeof(double));
    }
    if (xbin) {
        std::vector<uint32_t> xb((size_t)N * NW);
        MVMC_CUDA_OK(cudaMemcpy(xb.data(), h->xbin + (size_t)b * N * NW, xb.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) xbin[(size_t)i * n + j] = (xb[(size_t)i * NW + (j >> 5)] >> (j & 31)) & 1u;
    }
    return MVMC_OK;
}
