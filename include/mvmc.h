/*
 * mvmc.h — C-ABI of the B200-native capture hot path (libmvmc.so).
 *
 * The reference (khanhha/multiview_motion_capture) is pure Python and has no FFI; the seams this
 * library replaces are plain Python calls, cited per entry point below as /root/reference paths.
 * INTEGRATION.md shows the ctypes stubs a maintainer of the reference would add.
 *
 * Conventions
 *   - All array arguments of the *stage* and *clip-batch* entry points are DEVICE pointers owned by the
 *     caller, contiguous, row-major, float64 unless stated. `stream` is a cudaStream_t passed as void*.
 *     Calls only enqueue work on `stream`; the caller keeps the buffers alive until it synchronises.
 *   - `*_host` entry points take HOST pointers and include the host<->device copies.
 *   - Return value: 0 on success, a negative MVMC_ERR_* otherwise. Nothing throws. No CPU fallback.
 *   - Index layout of the association problem ("global index"): [ T alive tracks | kept poses of view 0
 *     (ascending pose id) | view 1 | ... ], n = T + sum P_v   (reference: motion_capture.py:658-667).
 *   - Joint layouts: 2D poses are COCO-17 (x, y, score); 3D track poses are BASIC_18 (x, y, z).
 *   - A pose parameter vector is 68 doubles: root(3) | euler XYZ (18*3) | side bone lengths (11)
 *     (reference: inverse_kinematics.py:86-91 PoseShapeParam).
 */
#ifndef MVMC_H_
#define MVMC_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVMC_OK 0
#define MVMC_ERR_INVALID (-1)   /* bad argument (null pointer, size out of range) */
#define MVMC_ERR_CUDA (-2)      /* a CUDA call failed; see mvmc_last_cuda_error() */
#define MVMC_ERR_CAPACITY (-3)  /* a per-clip capacity (tracks, births, views per match) overflowed */
#define MVMC_ERR_NO_DEVICE (-4)

#define MVMC_MAX_VIEWS 8        /* cameras per clip */
#define MVMC_MAX_POSES 32       /* 2D poses per view */
#define MVMC_MAX_TRACKS 64      /* alive tracks per clip */
#define MVMC_N_COCO 17
#define MVMC_N_B18 18
#define MVMC_N_PARAM 68
#define MVMC_MAX_SEL 16         /* 2D poses per IK solve on the fast path (tracked frames use at most one per view) */
#define MVMC_MAX_GROUP 256      /* 2D poses a no-track group can hold: births from more than MVMC_MAX_SEL poses take the slow path */
#define MVMC_MAX_BIG 8          /* such groups per clip and frame */

int mvmc_version(void);
const char* mvmc_error_string(int code);
const char* mvmc_last_cuda_error(void);

/* First n doubles of numpy.random.RandomState(0).rand(...) (MT19937, 53-bit) — the ALS initial factor
 * A0[i][j] = stream[i*r + j] (reference: mv_association.py:271). HOST pointer. */
int mvmc_rand_stream_host(double* out, int n);

/* ------------------------------------------------------------------------------------------------
 * Stage entry points (one per reference seam). B = number of independent instances ("clips").
 * ---------------------------------------------------------------------------------------------- */

/* A1 — mv_math_util.py:57-77 get_fundamental_matrix for every ordered camera pair.
 * P [B,C,3,4] -> F [B,C,C,3,3] with x_j^T F[i][j] x_i = 0. */
int mvmc_fundamental(const double* P, double* F, int B, int C, void* stream);

/* A7 helper — mv_math_util.py:267-285 calc_pairwise_f_mats: F from (K, R|t), float64 math stored as
 * float32. K [B,C,3,3], Rt [B,C,3,4] -> F32 [B,C,C,3,3] (float). */
int mvmc_fundamental_krt(const double* K, const double* Rt, float* F32, int B, int C, void* stream);

/* A0 + A4 index layout — motion_capture.py:1023-1043 filter_bad_pose and :643-667.
 * kps [B,C,Pmax,17,3], n_pose [B,C], n_trk [B] ->
 * keep [B,C,Pmax] (u8), dim_groups [B,C+2], idx_view [B,N] (-1 for tracks), idx_pose [B,N] (pose id /
 * track slot), N = Tmax + C*Pmax. */
int mvmc_prepare(const double* kps, const int* n_pose, const int* n_trk, int B, int C, int Pmax, int Tmax,
                 uint8_t* keep, int* dim_groups, int* idx_view, int* idx_pose, void* stream);

/* A2 + A3 + A4 (+ A7 when n_trk[b] == 0) — motion_capture.py:634-756, :597-631;
 * mv_math_util.py:80-115, :288-351. Needs mvmc_prepare's outputs.
 * trk_joints [B,Tmax,18,3]; F [B,C,C,3,3]; F32 [B,C,C,3,3] float; dst, sim [B,N,N] (only the leading
 * n x n block of each instance, leading dimension N, is written). */
int mvmc_affinity(const double* kps, const double* P, const double* F, const float* F32,
                  const double* trk_joints, const int* n_trk, const int* dim_groups, const int* idx_view,
                  const int* idx_pose, int B, int C, int Pmax, int Tmax, double* dst, double* sim,
                  void* stream);

/* A5 — mv_association.py:222-318 match_als. sim [B,N,N] (ld N), dim_groups [B,G+1] with G = n_groups.
 * f32_first_iter [B] (may be NULL): 1 where `sim` came from the float32 no-track path (A7).
 * rand_stream: device copy of mvmc_rand_stream_host with at least N*rmax entries.
 * workspace: mvmc_match_als_workspace_bytes(B, N, rmax) bytes of device memory.
 * xbin [B,N,(N+31)/32 words] row bitmasks of X_bin (bit j of row i = X_bin[i][j]); n_iter [B]. */
size_t mvmc_match_als_workspace_bytes(int B, int N, int rmax);
int mvmc_match_als(const double* sim, const int* dim_groups, int n_groups, const int* f32_first_iter,
                   const double* rand_stream, int B, int N, int rmax, void* workspace, uint32_t* xbin,
                   int* n_iter, void* stream);

/* A6 — mv_association.py:99-121 transform_closure, motion_capture.py:417-446 parse_match_result and
 * :762-808 / :618-624 group decoding. Outputs, per instance:
 *   trk_nsel [B,Tmax]: -1 = track absent from spatial_time_matches, else number of (view,pose) pairs;
 *   trk_sel  [B,Tmax,MVMC_MAX_SEL,2]: (view, pose id);
 *   new_n [B]; new_nsel [B,Nb]; new_sel [B,Nb,MVMC_MAX_SEL,2] for the 2D-only groups with >= 2 poses (the ones
 *   the reference turns into new tracks, motion_capture.py:942), in reference order, Nb = max_new;
 *   counts [B,4]: [0] how often the "more than one pose per view" rule fired (motion_capture.py:779,798), [1] 2D-only
 *   groups that rule shrank to a single pose (listed by the reference, never born), [2] groups with more than
 *   MVMC_MAX_SEL poses, of which only the first MVMC_MAX_SEL are kept (no-track frames of crowded scenes, where the
 *   reference's float32 affinity merges dozens of poses of different people into one group), [3] reserved;
 *   err [B]: 0 or MVMC_ERR_CAPACITY. N <= MVMC_MAX_TRACKS + MVMC_MAX_VIEWS * MVMC_MAX_POSES. */
int mvmc_assign(const uint32_t* xbin, const int* dim_groups, const int* idx_view, const int* idx_pose,
                const int* n_trk, int B, int C, int N, int Tmax, int max_new, int* trk_nsel, int* trk_sel,
                int* new_n, int* new_nsel, int* new_sel, int* counts, int* err, void* stream);

/* The same with the reference's full `spatial_matches` list recoverable (the `associate_tracking` seam returns it):
 * new_seq [B,Nb] (may be NULL) = position of each stored group among ALL 2D-only groups, singles [B,Nb,3] (may be NULL) =
 * (position, view, pose id) of the counts[1] single-pose groups. */
int mvmc_assign_listed(const uint32_t* xbin, const int* dim_groups, const int* idx_view, const int* idx_pose,
                       const int* n_trk, int B, int C, int N, int Tmax, int max_new, int* trk_nsel, int* trk_sel,
                       int* new_n, int* new_nsel, int* new_sel, int* counts, int* err, int* new_seq, int* singles,
                       void* stream);

/* The same with the overflow lists of the many-pose groups (no-track frames of crowded scenes): a 2D-only group of more than
 * MVMC_MAX_SEL poses keeps ALL its poses (up to MVMC_MAX_GROUP) in big_sel [B,MVMC_MAX_BIG,MVMC_MAX_GROUP,2], its size in
 * big_nsel [B,MVMC_MAX_BIG] AND in new_nsel of its birth slot big_slot [B,MVMC_MAX_BIG]; big_n [B] such groups. Only a
 * group beyond those capacities is still cut to MVMC_MAX_SEL poses and counted in counts[2]. All four or none may be NULL. */
int mvmc_assign_groups(const uint32_t* xbin, const int* dim_groups, const int* idx_view, const int* idx_pose,
                       const int* n_trk, int B, int C, int N, int Tmax, int max_new, int* trk_nsel, int* trk_sel,
                       int* new_n, int* new_nsel, int* new_sel, int* counts, int* err, int* new_seq, int* singles,
                       int* big_n, int* big_nsel, int* big_sel, int* big_slot, void* stream);

/* A6, first half alone — mv_association.py:99-121 transform_closure (what match_als returns as `match_mat`).
 * xbin [B,N,(N+31)/32] as written by mvmc_match_als, n [B] live size of each instance ->
 * match_mat [B,N,N] bytes (leading n x n block written): match[j][i] = 1 iff i is a leader and j belongs to it. */
int mvmc_transform_closure(const uint32_t* xbin, const int* n, int B, int N, uint8_t* match_mat, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Alternative matchers: kept for A/B against the ALS matcher.
 * ---------------------------------------------------------------------------------------------- */
/* The float64 distance matrix alone: A2 (calc_epipolar_error) for 2D-2D pairs, A3 for 2D-3D pairs, NaN for same-view and
 * 3D-3D pairs, 0 on the diagonal - with or without tracks (mvmc_affinity switches to the float32 path without tracks). */
int mvmc_distances(const double* kps, const double* P, const double* F, const double* trk_joints, const int* n_trk,
                   const int* dim_groups, const int* idx_view, const int* idx_pose, int B, int C, int Pmax, int Tmax,
                   double* dst, void* stream);
/* scipy.optimize.linear_sum_assignment (third party; rectangular shortest-augmenting-path algorithm, Crouse 2016, with
 * SciPy's scan order and tie rule) for B problems: cost [B,R,Cc] row-major, n_rows/n_cols [B] live sizes ->
 * col_of_row [B,R] (-1 = unassigned), status [B] (0, -1 = NaN / infeasible: SciPy raises). min(R,Cc) <= 64, max <= 256. */
int mvmc_linear_sum_assignment(const double* cost, const int* n_rows, const int* n_cols, int B, int R, int Cc,
                               int* col_of_row, int* status, void* stream);
/* motion_capture.py:166-241 match_objects_across_views with PoseAssociation.calc_epipolar_error (:81-93): the view with the
 * most poses seeds one group per pose; every other view (ascending) is assigned to the groups by a Hungarian step on the
 * mean epipolar distance to the group's members; a matched pose whose running distance sum exceeds `threshold` starts its
 * own group, as does every unmatched pose. dst = mvmc_distances' matrix [B,N,N], dim_groups as mvmc_prepare writes them.
 * group_of [B,N]: group index of every 2D pose (-1 for track slots / padding), groups numbered in creation order;
 * n_groups [B]; status [B]. workspace: mvmc_match_views_workspace_bytes(B). */
size_t mvmc_match_views_workspace_bytes(int B);
int mvmc_match_views_hungarian(const double* dst, const int* dim_groups, int B, int C, int N, double threshold,
                               void* workspace, int* group_of, int* n_groups, int* status, void* stream);
/* motion_capture.py:844-871 tracklet_to_poses_association with :845-850 tracklet_to_pose_2d_cost and
 * mv_math_util.py:11-32 (unproject_uv_to_rays, points_to_lines_distances): per (clip, view) the cost of (track t, pose p) is
 * the mean distance of the track's 3D joints to the camera rays through the pose's 2D joints (15 common joints, no score
 * mask), assigned by a Hungarian step, matches dearer than max_dst dropped.
 * Kr_inv [B,C,3,3] = R^T K^-1, cam_loc [B,C,3] (common.py:7-17), keep [B,C,Pmax] (mvmc_prepare) ->
 * match [B,C,Tmax] pose id or -1, cost [B,C,Tmax,Pmax] (may be NULL), status [B*C]. */
int mvmc_tracklet_pose_association(const double* trk_joints, const int* n_trk, const double* kps, const uint8_t* keep,
                                   const double* Kr_inv, const double* cam_loc, int B, int C, int Pmax, int Tmax,
                                   double max_dst, int* match, double* cost, int* status, void* stream);

/* B1 + B2 — mv_math_util.py:152-240 (DLT per joint + optional 2-nfev TRF refine).
 * obs [M,V,K,3] (x,y,score), Psel [M,V,3,4], n_views [M] (<= V <= MVMC_MAX_SEL), K <= 18 joints.
 * out [M,K,4] = (x,y,z,mean score). refine_nfev = 0 disables the refine. */
int mvmc_triangulate(const double* obs, const double* Psel, const int* n_views, int M, int V, int K,
                     double min_score, int refine_nfev, double* out, void* stream);

/* I1 — inverse_kinematics.py:176-199 foward_kinematics on the BASIC_18 skeleton.
 * params [M,68] -> joints [M,18,3]. */
int mvmc_fk(const double* params, int M, double* joints, void* stream);

/* I1/I6 generic chain — kinematics.py:18-31 without its two quirks.
 * rot [M,J,3,3] local rotations, offsets [J,3], parents [J] (parent index < child index, -1 = root),
 * root [M,3] or NULL -> joints [M,J,3]. J <= 64. */
int mvmc_fk_chain(const double* rot, const double* offsets, const int* parents, const double* root, int M,
                  int J, double* joints, void* stream);

/* I0 + I2 + I3 + I4 + I5 — inverse_kinematics.py:202-277,339-433 PoseSolver.solve, with
 * scipy.optimize.least_squares(method='trf', jac='2-point') restated on the device.
 * kps2d [M,V,17,3] COCO poses feeding each solve, Psel [M,V,3,4], n_views [M];
 * x0 [M,68] warm start (ignored where birth[m] != 0: triangulate -> root = mid hip, zero angles,
 * reference bone lengths); max_nfev [M] (reference: 5 for updates, 50 for births);
 * free_mask [68] u8 or NULL: 1 = parameter is optimised (NULL = all; the reference optimises all).
 * Outputs: x_out [M,68], joints [M,18,3], info [M,2,4] = per solve (nfev, njev, status, n_free),
 * cost [M,2]. workspace: mvmc_ik_workspace_bytes(M, V). */
size_t mvmc_ik_workspace_bytes(int M, int V);
int mvmc_ik_solve(const double* kps2d, const double* Psel, const int* n_views, const double* x0,
                  const uint8_t* birth, const int* max_nfev, const uint8_t* free_mask, int M, int V,
                  void* workspace, double* x_out, double* joints, int* info, double* cost, void* stream);

/* Births from many poses — motion_capture.py:618-624, 942-958: a no-track frame of a crowded scene groups dozens of 2D poses
 * (several per view) and the reference builds the new track from ALL of them. kps [B,C,Pmax,17,3], P [B,C,3,4]; per clip
 * big_n [B] groups, big_nsel [B,G] their sizes (<= MVMC_MAX_GROUP), big_sel [B,G,MVMC_MAX_GROUP,2] (view, pose id),
 * big_slot [B,G]; the solve of group (b, g) is written to row b*S + slot0 + big_slot[b][g] of x_out [.,68], joints [.,18,3],
 * info [.,2,4], cost [.,2]. Triangulation (+ 2-evaluation refine) and the two max_nfev-evaluation TRF stages of
 * PoseSolver.solve, with no cap on the number of observations. workspace: mvmc_ik_birth_big_workspace_bytes(). */
size_t mvmc_ik_birth_big_workspace_bytes(void);
int mvmc_ik_birth_big(const double* kps, const double* P, const int* big_n, const int* big_nsel, const int* big_sel,
                      const int* big_slot, int B, int C, int Pmax, int G, int S, int slot0, int max_nfev, void* workspace,
                      double* x_out, double* joints, int* info, double* cost, void* stream);

/* SURVEY.md 8f-4 — inverse_kinematics.py:280-336 solve_pose / solve_pose_bone_lens, the 3D-target variants behind
 * PoseSolver's `use_only_reproj = False` (:402-415): residual = (FK joint - triangulated point) * point score over the 16
 * IK joints. target [M,16,4] = (x, y, z, score) gathered at the IK joints (BASIC_18 idx 1..7, 9..17 <- 18-point observation
 * idx 11,13,15,12,14,16,17,5,7,9,6,8,10,0,3,4); x0 [M,68]; max_nfev [M]; stages: bit 0 = solve_pose (root + angles), bit 1 =
 * solve_pose_bone_lens (+ 11 side lengths), run in that order. Outputs as mvmc_ik_solve (a stage not run reports nfev 0). */
int mvmc_ik_solve_targets(const double* target, const double* x0, const int* max_nfev, int stages, int M, double* x_out,
                          double* joints, int* info, double* cost, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Clip-batch pipeline: B independent clips advance one frame per step
 * (reference: MvTracker.update_4d, motion_capture.py:873-963, one call per clip and frame).
 * ---------------------------------------------------------------------------------------------- */
typedef struct mvmc_clips mvmc_clips;

typedef struct mvmc_config {
    int n_clips;      /* B */
    int n_views;      /* C <= MVMC_MAX_VIEWS */
    int max_poses;    /* Pmax <= MVMC_MAX_POSES */
    int max_tracks;   /* Tmax <= MVMC_MAX_TRACKS */
    int max_new;      /* births per clip per frame that can be solved (<= 64) */
    int n_inits;      /* hits to confirm a track (reference: 3) */
    int max_age;      /* misses tolerated by a confirmed track (reference: 0) */
    int nfev_update;  /* reference: 5 */
    int nfev_birth;   /* reference: 50 */
    int keep_matrices;/* 1: keep dst/sim/xbin readable after a step (debug/parity) */
} mvmc_config;

void mvmc_default_config(mvmc_config* cfg);
int mvmc_clips_create(const mvmc_config* cfg, mvmc_clips** out);
void mvmc_clips_destroy(mvmc_clips* h);
/* bytes of device memory held by the handle */
size_t mvmc_clips_device_bytes(const mvmc_clips* h);

/* K [B,C,3,3], Rt [B,C,3,4], P [B,C,3,4] (= K·Rt as the host computed it; reference:
 * motion_capture.py:250-272 load_calib). Host or device pointers (copied with cudaMemcpyDefault);
 * both fundamental-matrix tables are derived on the device. */
int mvmc_clips_set_calib(mvmc_clips* h, const double* K, const double* Rt, const double* P, void* stream);
/* forget all tracks of all clips */
int mvmc_clips_reset(mvmc_clips* h, void* stream);

/* Advance every clip by one frame. kps [B,C,Pmax,17,3], n_pose [B,C] DEVICE pointers.
 * frame_idx is recorded in the outputs only. */
int mvmc_clips_step(mvmc_clips* h, const double* kps, const int* n_pose, int frame_idx, void* stream);

/* Per-step result record, one per clip (device layout == host layout). */
typedef struct mvmc_track_out {
    int32_t track_id;                 /* creation order within the clip (reference: append order) */
    int32_t state;                    /* 1 tentative, 2 confirmed (dead tracks are not listed) */
    int32_t hits;
    int32_t time_since_update;
    int32_t length;                   /* frames in which the track was solved */
    int32_t updated;                  /* 1 = solved this frame (update or birth), 2 = born this frame */
    int32_t n_sel;                    /* (view, pose) pairs used this frame; a birth from a many-pose group may report more than
                                         MVMC_MAX_SEL: sel then holds the first ones, mvmc_clips_read_big_groups_host all */
    int32_t sel[MVMC_MAX_SEL][2];
    int32_t nfev[2], njev[2], status[2];
    int32_t pad_;
    double cost[2];
    double param[MVMC_N_PARAM];
    double joints[MVMC_N_B18 * 3];
} mvmc_track_out;

typedef struct mvmc_step_out {
    int32_t frame_idx;
    int32_t n_alive;                  /* tracks alive after the step */
    int32_t n_died;                   /* tracks that died in this step */
    int32_t died_ids[MVMC_MAX_TRACKS];
    int32_t n_total;                  /* n of the association problem */
    int32_t als_iters;
    int32_t n_dup_view;
    int32_t error;                    /* 0 or MVMC_ERR_CAPACITY */
    int32_t n_truncated;              /* groups cut to their first MVMC_MAX_SEL poses: only groups beyond the overflow capacities
                                         (MVMC_MAX_BIG groups of MVMC_MAX_GROUP poses per clip and frame), see mvmc_assign_groups */
    mvmc_track_out tracks[MVMC_MAX_TRACKS];
} mvmc_step_out;

/* sizeof(mvmc_step_out) as compiled into the library (binding sanity check) */
size_t mvmc_sizeof_step_out(void);

/* DEVICE pointer to the B records written by the last step (valid until the next step). */
const mvmc_step_out* mvmc_clips_last_out(const mvmc_clips* h);

/* Compact fixed-stride records of the tracks solved in the last step, for gathering results across GPUs (the path's only
 * collective): rec [B,cap,128] doubles (DEVICE) = (clip0 + b, track id, frame, state, hits, views used,
 * 68 parameters, 54 joint coordinates) = 1 KB per track-frame; count [B] (DEVICE) = tracks solved in that clip. */
int mvmc_clips_pack_records(mvmc_clips* h, int cap, int clip0, double* rec, int* count, void* stream);

/* Same step with HOST buffers: copies kps/n_pose host->device, steps, copies the B records back and
 * synchronises the stream. out_host may be NULL (then only `n_alive`-sized summaries stay on device). */
int mvmc_clips_step_host(mvmc_clips* h, const double* kps_host, const int* n_pose_host, int frame_idx,
                         mvmc_step_out* out_host, void* stream);
/* The same without the final synchronisation (pinned host buffers; the caller synchronises `stream` before it reads
 * out_host or reuses the input buffers): several handles on several streams then overlap each other's copies and kernels
 * (the host side's ClipStreams steps groups of clips that way). */
int mvmc_clips_step_host_async(mvmc_clips* h, const double* kps_host, const int* n_pose_host, int frame_idx,
                               mvmc_step_out* out_host, void* stream);

/* Ingest — motion_capture.py:974-1005 parse_openpose_kps / extract_frame_data_from_openpose and
 * pose_def.py:262-270 conversion_openpose_25_to_coco.
 * mvmc_parse_openpose_host: the text of one OpenPose `*_keypoints.json` -> out [max_people,25,3] (x, y, score; zero padded),
 *   *n_people = len(people) (people beyond max_people are dropped). Native scanner, no Python.
 * mvmc_parse_openpose_files_host: n_files paths -> out [n_files,max_people,25,3], n_people [n_files], n_threads worker threads.
 * mvmc_ingest_body25 (DEVICE pointers): kps25 [B,C,Pin,25,3], n_people [B,C] -> kps [B,C,Pmax,17,3] COCO, n_pose [B,C].
 * mvmc_clips_step_body25_host: HOST BODY_25 buffers [B,C,Pmax,25,3] + [B,C] -> copy, gather on the device, step, records
 *   back (asynchronous like mvmc_clips_step_host_async). */
int mvmc_parse_openpose_host(const char* text, size_t len, int max_people, double* out, int* n_people);
int mvmc_parse_openpose_files_host(const char* const* paths, int n_files, int max_people, double* out, int* n_people,
                                   int n_threads);
int mvmc_ingest_body25(const double* kps25, const int* n_people, int B, int C, int Pin, int Pmax, double* kps, int* n_pose,
                       void* stream);
int mvmc_clips_step_body25_host(mvmc_clips* h, const double* kps25_host, const int* n_people_host, int frame_idx,
                                mvmc_step_out* out_host, void* stream);

/* Teacher forcing / checkpoint-resume: overwrite the alive-track table of every clip.
 * n_trk [B]; ids, state, hits, tsu, length [B,Tmax]; param [B,Tmax,68]; joints [B,Tmax,18,3];
 * next_id [B]. HOST pointers. */
int mvmc_clips_set_tracks_host(mvmc_clips* h, const int* n_trk, const int* ids, const int* state,
                               const int* hits, const int* tsu, const int* length, const double* param,
                               const double* joints, const int* next_id, void* stream);

/* The many-pose birth groups of clip b in the last step (HOST pointers): *n groups, nsel [MVMC_MAX_BIG] sizes,
 * slot [MVMC_MAX_BIG] = index among the clip's births of that step (the k-th track with updated == 2 of its record),
 * sel [MVMC_MAX_BIG, MVMC_MAX_GROUP, 2] (view, pose id). */
int mvmc_clips_read_big_groups_host(mvmc_clips* h, int b, int* n, int* nsel, int* slot, int* sel, void* stream);

/* Debug/parity read-back of the last step's association matrices of clip b (HOST pointers, sized n*n
 * with n = n_total of that clip; xbin as bytes). Requires keep_matrices=1. */
int mvmc_clips_read_matrices_host(mvmc_clips* h, int b, double* dst, double* sim, uint8_t* xbin, int* n,
                                  int* dim_groups, void* stream);

/* Device-side work counters accumulated by every step (HOST out[8]): [0] ALS algorithmic flops
 * sum I(6rn^2+8r^2n+4r^3), [1] ALS iterations, [2] clip-frames, [3] IK solves, [4] nfev, [5] njev,
 * [6] IK algorithmic flops (SURVEY.md §8d formula), [7] sum n^2. Synchronises the stream. */
int mvmc_clips_stats_host(mvmc_clips* h, double* out, int reset, void* stream);

/* Per-stage device time via CUDA events on the step's stream. enable: 1 = start (resets), 0 = stop and read,
 * -1 = read. out_ms[5] = {prepare+affinity, ALS, assign+gather, IK, commit} summed over *n_steps steps. */
int mvmc_clips_profile(mvmc_clips* h, int enable, double* out_ms, int* n_steps, void* stream);

/* FP64 DFMA peak probe (roofline denominator for the FP64-issue-bound kernels): blocks x 256 threads x
 * iters x 8 FMAs. Time it with events on `stream`; flops = blocks*256*iters*16. */
int mvmc_fp64_probe(int blocks, int iters, double* sink, void* stream);
/* Same for the FP64 tensor cores (mma.sync.m8n8k4.f64, the instruction k_als runs on): blocks x 8 warps x iters x 8 DMMA;
 * flops = blocks*8*iters*8*512. */
int mvmc_fp64_tensor_probe(int blocks, int iters, double* sink, void* stream);

/* Diagnostics of mvmc_match_als: per-phase SM-cycle sums of thread 0 of every CTA since the counters were enabled
 * (enable != 0 resets and starts, 0 stops; out may be NULL, else 20 doubles: G = A^T A, its inverse, T = A^T Xt, B,
 * H = B^T B, its inverse, T = B^T Xt^T, A, X = A B^T, residual reduction, mu-change pass, init, ADMM update pass: arithmetic, waiting for its strips, its opening fence; then, inside the two inverses: load, 8 x 8 pivot-block inverses, row panels, updates, store).
 * Synchronises the device. */
#define MVMC_ALS_N_PHASES 20
int mvmc_als_phase_profile(int enable, double* out);

/* Measurement aid: which tile build of the ALS kernel mvmc_match_als runs: -1 = by shape (default), 0 = 64 x 96 CTA tiles /
 * 8 warps, 1 = 48 x 48 tiles / 4 warps. Results do not depend on it. */
int mvmc_als_force_variant(int v);

/* number of kernel launches enqueued by this library since load (for bench.py's gpu_launches) */
unsigned long long mvmc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MVMC_H_ */
