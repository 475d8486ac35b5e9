/*
 * lvba_b200.h — C ABI of the B200-native LiDAR-visual bundle-adjustment hot path.
 *
 * This is the drop-in boundary for the two solver call sites of xuankuzcr/Global-LVBA
 * (SURVEY.md §8b).  The reference has no FFI of its own: both call sites are plain C++
 * inside one translation unit, so every entry point below cites the reference code it
 * replaces.  All buffers are HOST pointers owned by the caller; the library copies in
 * at call time, keeps its own device memory, writes results back before returning and
 * retains no caller pointer.  Nothing throws across this boundary: every function
 * returns LVBA_OK (0) or a negative lvba_status, and on error in/out buffers are left
 * untouched.  There is no CPU fallback: without a CUDA device every compute entry
 * point returns LVBA_ERR_NO_DEVICE.
 *
 * Conventions shared with the reference:
 *   pose           12 doubles: R (3x3 row-major, body->world) then p      IMUST.R/.p   include/BALM/tools.hpp:147-153
 *   cluster        10 doubles: Pxx Pxy Pxz Pyy Pyz Pzz vx vy vz N          PointCluster include/BALM/tools.hpp:407-424
 *   tangent order  (dphi, dp) per pose, R <- R*Exp(dphi), p <- p + dp      include/BALM/bavoxel.hpp:722-727
 *   quaternion     {w,x,y,z} (memory order of qs[k])                      src/lvba_system.cpp:1513-1516
 *   intr[8]        fx fy cx cy k1 k2 p1 p2 (Brown-Conrady)                 include/utils.hpp:53-58
 */
#ifndef LVBA_B200_H
#define LVBA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVBA_B200_VERSION 100   /* 0.1.0 */

typedef enum lvba_status {
  LVBA_OK = 0,
  LVBA_ERR_INVALID_ARG = -1,   /* null pointer, negative size, index out of range, non-monotone CSR */
  LVBA_ERR_NO_DEVICE = -2,     /* no CUDA device / driver: the product path has no CPU fallback */
  LVBA_ERR_CUDA = -3,          /* a CUDA runtime call failed; see lvba_last_error() */
  LVBA_ERR_UNSUPPORTED = -4,   /* problem shape outside what this build handles */
  LVBA_ERR_NUMERIC = -5,       /* non-finite step / zero pivot and no recovery possible */
  LVBA_ERR_COMM = -6,          /* NCCL not loadable or a collective failed */
  LVBA_ERR_NOMEM = -7
} lvba_status;

/* ---- path A options: BALM2::damping_iter constants, include/BALM/bavoxel.hpp:664,686,760 */
typedef struct lvba_lidar_opts {
  double u0;           /* 0.01  initial damping            (bavoxel.hpp:664) */
  double v0;           /* 2.0   initial damping growth     (bavoxel.hpp:664) */
  int32_t max_iter;    /* 10                               (bavoxel.hpp:686) */
  double rel_tol;      /* 1e-6  |r1-r2|/r1 stop            (bavoxel.hpp:760); <0 disables the test */
  int32_t device;      /* CUDA device ordinal; -1 = current */
  int32_t verbose;     /* 1: print one line per LM iteration to stderr */
} lvba_lidar_opts;

/* ---- path B options: the ceres::Solver::Options in force at src/lvba_system.cpp:1572-1576
 *      (ceres-solver 2.1.0 defaults everywhere else) */
typedef struct lvba_visual_opts {
  int32_t max_iter;              /* 50      options.max_num_iterations (:1573) */
  double initial_radius;         /* 1e4  */
  double max_radius;             /* 1e16 */
  double min_radius;             /* 1e-32 */
  double min_lm_diagonal;        /* 1e-6 */
  double max_lm_diagonal;        /* 1e32 */
  double min_relative_decrease;  /* 1e-3 */
  double function_tolerance;     /* 1e-6;  <0 disables */
  double gradient_tolerance;     /* 1e-10; <0 disables */
  double parameter_tolerance;    /* 1e-8;  <0 disables */
  int32_t jacobi_scaling;        /* 1 */
  int32_t device;
  int32_t verbose;
} lvba_visual_opts;

typedef enum lvba_termination {
  LVBA_TERM_MAX_ITER = 0,
  LVBA_TERM_FUNCTION_TOL = 1,   /* A: bavoxel.hpp:760 ; B: Ceres function tolerance */
  LVBA_TERM_PARAMETER_TOL = 2,
  LVBA_TERM_GRADIENT_TOL = 3,
  LVBA_TERM_RADIUS = 4,
  LVBA_TERM_INVALID_STEPS = 5,
  LVBA_TERM_SKIPPED = 6         /* window BA: fewer than min_voxels_per_pose * W voxels, poses untouched */
} lvba_termination;

typedef struct lvba_summary {
  int32_t iterations;        /* LM loop passes executed (accepted + rejected) */
  int32_t accepted;          /* accepted steps */
  int32_t hessian_builds;    /* passes that rebuilt H/g (A) or J/S (B) */
  int32_t termination;       /* lvba_termination */
  double cost_first;         /* A: sum(lambda0)/V at entry ; B: 1/2 sum r^2 at entry */
  double cost_last;          /* same quantity at the returned state */
  double damping_last;       /* A: u ; B: trust-region radius */
  double ms_total;           /* wall time inside the call (host clock) */
  double ms_setup;           /* host symbolic analysis + H2D upload */
  double ms_build;           /* CUDA-event time in Hessian / Jacobian+Schur build kernels */
  double ms_solve;           /* CUDA-event time in factorisation + substitution */
  double ms_residual;        /* CUDA-event time in residual-only passes + retraction */
  int64_t kernel_launches;   /* launches of this library's own kernels during the call */
  int64_t h2d_bytes;
  int64_t d2h_bytes;
} lvba_summary;

/* ---- misc ------------------------------------------------------------------------------ */
int         lvba_version(void);
int         lvba_device_count(void);          /* 0 when no usable GPU; never fails */
const char* lvba_status_string(int status);
const char* lvba_last_error(void);            /* thread-local detail of the last failure */
/* Device buffers of destroyed problems are cached for reuse by later calls (cudaMalloc/cudaFree cost
 * milliseconds each); this returns the cached memory to the driver. */
int         lvba_release_cached_memory(void);
void        lvba_lidar_default_opts(lvba_lidar_opts* o);
void        lvba_visual_default_opts(lvba_visual_opts* o);

/* ======================================================================================
 * B1  LiDAR LM — replaces  void BALM2::damping_iter(vector<IMUST>& x_stats, VOX_HESS& voxhess)
 *     include/BALM/bavoxel.hpp:662-767, called at src/lvba_system.cpp:264 and :386.
 *
 *   W          number of poses (BALM2::win_size)
 *   V          number of plane voxels (VOX_HESS::plvec_voxels.size())
 *   vox_ptr    [V+1] CSR offsets; voxel a owns slots vox_ptr[a]..vox_ptr[a+1]
 *   pose_idx   [nnz] pose index i of every slot with (*plvec_voxels[a])[i].N != 0, ascending inside a voxel
 *   clusters   [nnz*10] body-frame PointCluster of that slot
 *   poses      [W*12] in/out
 * ====================================================================================== */
int lvba_lidar_lm(int32_t W, int64_t V, const int64_t* vox_ptr, const int32_t* pose_idx,
                  const double* clusters, double* poses, const lvba_lidar_opts* opts,
                  lvba_summary* summary);

/* B1 batched — every window of LvbaSystem::runWindowBA (src/lvba_system.cpp:232-302: consecutive windows of
 * `window_size` poses, one BALM2::damping_iter each, :264) in ONE call: one block-diagonal system, one CTA per
 * window in the factorisation, per-window u / v / accept-reject / stop test.  Results equal n_windows separate
 * lvba_lidar_lm calls.
 *   win_ptr    [n_windows+1] window w owns poses win_ptr[w] .. win_ptr[w+1]-1 of the concatenated `poses`
 *   pose_idx   indices into the concatenated pose array; every voxel lies inside ONE window
 *   min_voxels_per_pose   windows with fewer than this many voxels per pose are skipped and their poses left
 *              untouched (3 in the reference, :262-266); their summary carries LVBA_TERM_SKIPPED
 *   summaries  [n_windows] per-window LM summary (iterations, accepted, costs, damping, termination); may be NULL
 *   total      timing / traffic of the whole call; may be NULL
 * Windows of more than 31 poses return LVBA_ERR_UNSUPPORTED (solve those with lvba_lidar_lm). */
int lvba_lidar_lm_batch(int32_t n_windows, const int32_t* win_ptr, int64_t V, const int64_t* vox_ptr,
                        const int32_t* pose_idx, const double* clusters, double* poses,
                        int32_t min_voxels_per_pose, const lvba_lidar_opts* opts,
                        lvba_summary* summaries, lvba_summary* total);
/* Device-resident handle for the same problem: lets a caller (bench, parity tests, a ROS
 * node that re-solves after outlier removal) run single phases with inputs already in HBM. */
typedef struct lvba_lidar_problem lvba_lidar_problem;

int lvba_lidar_create(int32_t W, int64_t V, const int64_t* vox_ptr, const int32_t* pose_idx,
                      const double* clusters, const double* poses, int32_t device,
                      lvba_lidar_problem** out);
int lvba_lidar_destroy(lvba_lidar_problem* p);
int lvba_lidar_set_poses(lvba_lidar_problem* p, const double* poses);     /* H2D, W*12 */
int lvba_lidar_get_poses(lvba_lidar_problem* p, double* poses);           /* D2H, W*12 */
/* VOX_HESS::acc_evaluate2 + divide_thread (bavoxel.hpp:68-174, 597-639) at the current poses:
 * builds H and g on the device, returns sum_v lambda0 (NOT divided by V). */
int lvba_lidar_build(lvba_lidar_problem* p, double* residual_sum);
/* VOX_HESS::evaluate_only_residual (bavoxel.hpp:176-203) at `poses` (host, W*12) or at the
 * current device poses when poses == NULL. */
int lvba_lidar_residual(lvba_lidar_problem* p, const double* poses, double* residual_sum);
/* (H + u*diag(H)) dx = -g  (bavoxel.hpp:692-710); dx [W*6] to host. */
int lvba_lidar_solve(lvba_lidar_problem* p, double u, double* dx);
/* Block structure of the lower-triangular envelope that stores H: nblocks and per block (row, col). */
int lvba_lidar_structure(lvba_lidar_problem* p, int64_t* nblocks, int32_t* brow, int32_t* bcol);
/* Copy out g [W*6] and the envelope blocks [nblocks*36, row-major 6x6, block (r,c) = H[6r.., 6c..]]. */
int lvba_lidar_get_system(lvba_lidar_problem* p, double* g, double* blocks);
/* n LM passes of damping_iter starting from the handle's state (u, v, poses carried in the
 * handle; call lvba_lidar_reset_lm to restart).  Device-resident: no problem data crosses PCIe. */
int lvba_lidar_reset_lm(lvba_lidar_problem* p, const lvba_lidar_opts* opts);
/* restore the poses passed to lvba_lidar_create (device-to-device; nothing crosses PCIe) */
int lvba_lidar_reset_state(lvba_lidar_problem* p);
int lvba_lidar_iterate(lvba_lidar_problem* p, int32_t n_iter, lvba_summary* summary);
/* exact algorithmic byte / flop counters of SURVEY.md §8(d) for this problem instance */
int lvba_lidar_counts(lvba_lidar_problem* p, int64_t* nnz, int64_t* n_blocks_env,
                      int64_t* n_blocks_nonzero, int64_t* n_pairs);

/* ======================================================================================
 * B2  visual LM — replaces the Ceres block of LvbaSystem::optimizeCameraPoses(),
 *     src/lvba_system.cpp:1571-1656 (problem build :1578-1640, ceres::Solve :1643,
 *     write-back :1651-1665).  Residual functors: include/utils.hpp:51-147.
 *
 *   M, T        cameras, landmarks
 *   q_wxyz      [M*4] in/out   qs[k]   (:1513-1516)
 *   t           [M*3] in/out   ts[k]
 *   X           [T*3] in/out   Xs[pi]  (:1521-1525); landmarks without a valid plane are left untouched
 *   plane_nd    [T*4] (n, d); n == 0 marks "no valid plane" => landmark and its observations are skipped (:1598-1603)
 *   obs_ptr     [T+1] CSR; obs_cam [nnz] camera of each inlier observation; obs_uv [nnz*2] float pixel (:1624-1625)
 *   fixed_cam   camera held constant (0 in the reference, :1582-1583); -1 = none
 * ====================================================================================== */
int lvba_visual_lm(int32_t M, int64_t T, double* q_wxyz, double* t, double* X,
                   const double* plane_nd, const int64_t* obs_ptr, const int32_t* obs_cam,
                   const float* obs_uv, const double intr[8], double sigma_px, double sigma_plane,
                   int32_t fixed_cam, const lvba_visual_opts* opts, lvba_summary* summary);

typedef struct lvba_visual_problem lvba_visual_problem;

int lvba_visual_create(int32_t M, int64_t T, const double* q_wxyz, const double* t, const double* X,
                       const double* plane_nd, const int64_t* obs_ptr, const int32_t* obs_cam,
                       const float* obs_uv, const double intr[8], double sigma_px, double sigma_plane,
                       int32_t fixed_cam, int32_t device, lvba_visual_problem** out);
int lvba_visual_destroy(lvba_visual_problem* p);
int lvba_visual_set_state(lvba_visual_problem* p, const double* q_wxyz, const double* t, const double* X);
int lvba_visual_get_state(lvba_visual_problem* p, double* q_wxyz, double* t, double* X);
/* 1/2 sum r^2 over all residual blocks at the current device state */
int lvba_visual_cost(lvba_visual_problem* p, double* cost);
/* One linearisation + Schur elimination + reduced solve + back-substitution at the current state
 * with the given trust-region radius; writes the (unscaled) tangent step: cam_step [M*6] (zeros for
 * inactive cameras), pt_step [T*3] (zeros for skipped landmarks), and the model cost change. */
int lvba_visual_step(lvba_visual_problem* p, double radius, int32_t jacobi_scaling, int32_t recompute_scale,
                     double* cam_step, double* pt_step, double* model_cost_change, double* cost);
/* Reduced camera system of the last lvba_visual_step: block structure + values + rhs (for parity tests). */
int lvba_visual_structure(lvba_visual_problem* p, int32_t* n_active, int32_t* cam_of_row /* [n_active] */,
                          int64_t* nblocks, int32_t* brow, int32_t* bcol);
int lvba_visual_get_system(lvba_visual_problem* p, double* rhs /* [n_active*6] */, double* blocks /* [nblocks*36] */);
int lvba_visual_reset_lm(lvba_visual_problem* p, const lvba_visual_opts* opts);
/* restore q, t, X passed to lvba_visual_create (device-to-device) */
int lvba_visual_reset_state(lvba_visual_problem* p);
int lvba_visual_iterate(lvba_visual_problem* p, int32_t n_iter, lvba_summary* summary);
int lvba_visual_counts(lvba_visual_problem* p, int64_t* nnz_valid, int64_t* n_valid_tracks,
                       int64_t* n_blocks_env, int64_t* n_pairs);

/* ======================================================================================
 * B3  adaptive voxel map (set-up stage) — replaces the cut_voxel / recut / tras_opt sequence in front of every
 *     LiDAR solve and the plane lookup in front of the visual solve:
 *       cut_voxel per scan       include/BALM/bavoxel.hpp:799-836   src/lvba_system.cpp:248-251, 366-369, 1499-1502
 *       recut + tras_opt         include/BALM/bavoxel.hpp:420-474   src/lvba_system.cpp:255-258, 374-377, 1504-1506
 *       recompute_local_planes   src/lvba_system.cpp:1529-1566 with findCorrespondPoint, bavoxel.hpp:320-333
 *     The map is built on the device from the raw scans (sort-based, no hash table, no per-point allocation) and
 *     stays there; the plane voxels come back in exactly the layout lvba_lidar_lm takes.
 *
 *   W                  number of scans = poses of the window (win_size)
 *   scan_ptr           [W+1] CSR offsets into the point array; scan j owns points scan_ptr[j] .. scan_ptr[j+1]-1
 *   xyz                body-frame points, x y z as float at the start of every record
 *   xyz_stride_floats  record size in floats: 3 for packed xyz, 12 for an array of pcl::PointXYZINormal (48 B)
 *   poses              [W*12] pose of every scan (x_buf / anchor_poses)
 *   Voxel order: ascending (root key x, y, z), then octant path — the reference's unordered_map order is unspecified.
 *   A non-finite point, or one more than 2^30 root voxels from the origin, is LVBA_ERR_INVALID_ARG (the reference's
 *   float -> int64 cast is undefined there).
 * ====================================================================================== */
typedef struct lvba_voxel_opts {
  double voxel_size;       /* root voxel edge: stage1_root_voxel_size_ / stage2_root_voxel_size_ (lvba_system.cpp:344-345) */
  float eigen_ratio[4];    /* eigen_ratio_array per layer, bavoxel.hpp:17-22 (set_eigen_ratio_array, lvba_system.cpp:360) */
  int32_t layer_limit;     /* 2   bavoxel.hpp:13; 0..2 supported */
  int32_t min_points;      /* 15  min_ps, bavoxel.hpp:24 */
  int32_t device;          /* CUDA device ordinal; -1 = current */
} lvba_voxel_opts;

typedef struct lvba_voxel_summary {
  int64_t n_points;
  int64_t n_voxels;        /* plane voxels seen from >= 2 poses (VOX_HESS::plvec_voxels.size()) */
  int64_t nnz;             /* (voxel, pose) clusters */
  int64_t n_nodes[3];      /* octree nodes per layer that hold points (layers never reached are 0) */
  double ms_total;         /* wall time inside lvba_voxel_map_create (host clock) */
  double ms_upload;        /* validation + packing + H2D enqueue */
  double ms_device;        /* CUDA-event time of the build passes */
  int64_t kernel_launches; /* launches of this library's own kernels (the cub sorts / scans are not counted) */
  int64_t h2d_bytes;
} lvba_voxel_summary;

typedef struct lvba_voxel_map lvba_voxel_map;

void lvba_voxel_default_opts(lvba_voxel_opts* o);
int lvba_voxel_map_create(int32_t W, const int64_t* scan_ptr, const float* xyz, int32_t xyz_stride_floats,
                          const double* poses, const lvba_voxel_opts* opts, lvba_voxel_map** out,
                          lvba_voxel_summary* summary /* may be NULL */);
/* One INDEPENDENT map per window of consecutive scans, all built together — the surf_map that runWindowBA creates, recuts
 * and deletes once per window (src/lvba_system.cpp:232-258).  win_ptr [n_windows+1]: window w owns scans
 * win_ptr[w] .. win_ptr[w+1]-1; W = win_ptr[n_windows].  Voxels never merge across windows, pose indices are those of the
 * concatenated scans, voxels are ordered by (window, key, path): the export is the input of lvba_lidar_lm_batch as it
 * stands.  lvba_voxel_map_lookup is not defined on a windowed map (LVBA_ERR_UNSUPPORTED). */
int lvba_voxel_map_create_windows(int32_t n_windows, const int32_t* win_ptr, const int64_t* scan_ptr, const float* xyz,
                                  int32_t xyz_stride_floats, const double* poses, const lvba_voxel_opts* opts,
                                  lvba_voxel_map** out, lvba_voxel_summary* summary /* may be NULL */);
/* n_windows (0 for a single map) and, when vox_window != NULL, the window of every voxel [V]. */
int lvba_voxel_map_windows(lvba_voxel_map* m, int32_t* n_windows, int32_t* vox_window);
int lvba_voxel_map_summary(const lvba_voxel_map* m, lvba_voxel_summary* summary);
/* Copy out the plane voxels (sizes from the summary).  Any pointer may be NULL.
 *   vox_ptr [V+1], pose_idx [nnz], clusters [nnz*10]   the arguments of lvba_lidar_lm / lvba_lidar_create
 *   root_key [V*3] int64 voxel key; path [V*3] int8 (layer, octant1 or -1, octant2 or -1)
 *   centre, normal, eigenvalues [V*3]   judge_eigen's center / direct / value_vector (bavoxel.hpp:346-349); the sign
 *   of `normal` is not defined (nor is it by Eigen's solver) */
int lvba_voxel_map_export(lvba_voxel_map* m, int64_t* vox_ptr, int32_t* pose_idx, double* clusters, int64_t* root_key,
                          int8_t* path, double* centre, double* normal, double* eigenvalues);
/* recompute_local_planes: for n world points X [n*3] the plane (n, d) [n*4] of the PLANE node each one falls in,
 * zeros when there is none — the plane_nd argument of lvba_visual_lm. */
int lvba_voxel_map_lookup(lvba_voxel_map* m, int64_t n, const double* X, double* plane_nd);
/* tras_opt straight into B1: the map's plane voxels as a device-resident LiDAR problem (handle API above).  The cluster
 * records never leave HBM; only the CSR index arrays (12 B per cluster) visit the host for the symbolic analysis.
 * poses [W*12]: the linearisation point (normally the poses the map was built with). */
int lvba_voxel_map_lidar_create(lvba_voxel_map* m, const double* poses, lvba_lidar_problem** out);
/* ... and solved: cut_voxel + recut (the map) -> tras_opt + BALM2::damping_iter (this call), poses [W*12] in/out.
 * Fewer than min_voxels_per_pose * W voxels (the caller-side rule of src/lvba_system.cpp:262-266; pass 0 for
 * runLidarBA, which has none) or an empty map: LVBA_OK, LVBA_TERM_SKIPPED, poses untouched. */
int lvba_voxel_map_lidar_lm(lvba_voxel_map* m, double* poses, int32_t min_voxels_per_pose, const lvba_lidar_opts* opts,
                            lvba_summary* summary);
/* The window stage of runWindowBA (src/lvba_system.cpp:232-266) from a windowed map: tras_opt + damping_iter of EVERY window
 * in one batched solve (see lvba_lidar_lm_batch for summaries / total / the skip rule), clusters never leaving the device. */
int lvba_voxel_map_lidar_lm_batch(lvba_voxel_map* m, double* poses, int32_t min_voxels_per_pose, const lvba_lidar_opts* opts,
                                  lvba_summary* summaries, lvba_summary* total);
int lvba_voxel_map_destroy(lvba_voxel_map* m);

/* ======================================================================================
 * B6  anchor clouds — replaces the tail of the window loop of runWindowBA, src/lvba_system.cpp:284-301 (and its twin at
 *     :1474-1491): every scan of a window transformed into the window's anchor frame (pl_transform, include/BALM/tools.hpp:385-395:
 *     the result is stored back as float), merged, and down-sampled with down_sampling_voxel2 (tools.hpp:301-359): per voxel of
 *     edge `leaf` the ORIGINAL point closest to the voxel centre, the first in cloud order among equally close ones.
 *   win_ptr [n_windows+1] over scans; scan_ptr / xyz / xyz_stride_floats as for B3
 *   rel_poses [S*12]   pose of every scan in its anchor frame: rel.R = anchor.R^T x.R, rel.p = anchor.R^T (x.p - anchor.p) (:286-289)
 *   leaf               anchor_leaf_size_ (window_ba/anchor_leaf_size, 0.1); < 0.001 returns the transformed points unsampled (:303)
 *   export: cloud_ptr [n_windows+1], xyz [n_points*3] float — the anchor_clouds; points of a window are ordered by voxel key
 *   (the reference's unordered_map order is unspecified)
 * ====================================================================================== */
typedef struct lvba_anchor_clouds lvba_anchor_clouds;
int lvba_anchor_clouds_create(int32_t n_windows, const int32_t* win_ptr, const int64_t* scan_ptr, const float* xyz,
                              int32_t xyz_stride_floats, const double* rel_poses, double leaf, int32_t device,
                              lvba_anchor_clouds** out, int64_t* n_points_out);
int lvba_anchor_clouds_export(lvba_anchor_clouds* a, int64_t* cloud_ptr, float* xyz, double* ms_device /* may be NULL */);
int lvba_anchor_clouds_destroy(lvba_anchor_clouds* a);

/* ======================================================================================
 * B4  depth rendering — replaces the world point grid and the per-image z-buffer in front of the track fusion:
 *       buildGridMapFromOptimized   src/lvba_system.cpp:1266-1338   (0.5 m voxels of ALL world points, per-frame voxel sets,
 *                                                                    per image the voxels of the frames within +-0.5 s)
 *       generateDepthWithVoxel      src/lvba_system.cpp:835-919     (project every point of those voxels, (int) pixel,
 *                                                                    `if (d == 0 || Z < d) d = (float)Z`)
 *     The z-buffer is float(min Z) whatever the visiting order, so the images are reproducible bit for bit.
 *
 *   n_frames, scan_ptr, xyz, xyz_stride_floats, poses   as for B3: pl_fulls_ / x_buf_ (dataset_io_)
 *   frame_ts     [n_frames] LiDAR frame timestamps x_buf_[i].t, ascending (the reference binary-searches them, :1318-1319)
 *   voxel_size   0.5 in the reference (:1277)
 *   cams         [n_images*12] Rcw (row-major) then tcw per image: Rcw_all_optimized_ / tcw_all_optimized_ (:861-864)
 *   image_ts     [n_images] image timestamps; NaN = an image name that does not parse (:1309-1314) -> empty image
 *   half_window  0.5 s in the reference (:1299)
 *   intr         fx fy cx cy k1 k2 p1 p2
 *   depth        [n_images * height * width] float, row-major per image (cv::Mat CV_32FC1), 0 = no point
 * ====================================================================================== */
typedef struct lvba_depth_summary {
  int64_t n_points, n_voxels;
  int64_t n_pairs;          /* distinct (frame, voxel) pairs = sum of the per_frame_voxels set sizes */
  double ms_total, ms_upload, ms_device;
  int64_t kernel_launches, h2d_bytes, d2h_bytes;
  int64_t work_pairs;       /* render: (image, frame-in-window, voxel) triples examined */
  int64_t work_chunks;      /* render: 64-point chunks projected (every voxel of a window once) */
} lvba_depth_summary;

typedef struct lvba_depth_grid lvba_depth_grid;

int lvba_depth_grid_create(int32_t n_frames, const int64_t* scan_ptr, const float* xyz, int32_t xyz_stride_floats,
                           const double* poses, const double* frame_ts, double voxel_size, int32_t device,
                           lvba_depth_grid** out, lvba_depth_summary* summary /* may be NULL */);
int lvba_depth_render(lvba_depth_grid* g, int32_t n_images, const double* cams, const double* image_ts, double half_window,
                      const double intr[8], int32_t width, int32_t height, float* depth,
                      lvba_depth_summary* summary /* may be NULL */);
/* The depth-fused 3-D candidates of the track fusion — the loop at src/lvba_system.cpp:1020-1038 of BuildTracksAndFuse3D
 * (fetchDepthBilinear include/utils.hpp:246-275 -> backProjectPixelDepthDistorted :235-243 -> camToWorld :277-283) for
 * EVERY keypoint of every image; a pure function of (image, keypoint), so the caller indexes the result by
 * (component[t].first, component[t].second).  The depth images are rendered and sampled on the device and never copied out.
 *   kp_ptr  [n_images+1] CSR over images; kp_uv [n_kp*2] float pixel (all_keypoints_[im][kp].x / .y)
 *   Xw      [n_kp*3] out: points3d (zeros where invalid); valid [n_kp] out: valid_mask */
int lvba_depth_backproject(lvba_depth_grid* g, int32_t n_images, const double* cams, const double* image_ts, double half_window,
                           const double intr[8], int32_t width, int32_t height, const int64_t* kp_ptr, const float* kp_uv,
                           double* Xw, uint8_t* valid, lvba_depth_summary* summary /* may be NULL */);
int lvba_depth_grid_destroy(lvba_depth_grid* g);

/* ======================================================================================
 * B5  per-track numerics of the track fusion (BuildTracksAndFuse3D, src/lvba_system.cpp:921-1263), many tracks at once:
 *       TriangulateTrackDLT   src/lvba_system.cpp:52-111    lvba_tracks_triangulate
 *       ComputeMeanReproj     src/lvba_system.cpp:8-50      lvba_tracks_mean_reproj
 *     The caller keeps what is inherently sequential there — connected components over the match graph, one observation
 * per image, the greedy view-angle filter (its result depends on the iteration order of std::unordered_map) — and passes the
 * SELECTED observations of every track (`selected_ids`) as a CSR list.
 *   obs_ptr [n_tracks+1]; obs_cam [n_obs] image id of each selected observation (ids outside [0, n_cams) are skipped, :74-78);
 *   obs_uv [n_obs*2] float keypoint; cams [n_cams*12] Rcw row-major + tcw (Rcw_all_optimized_ / tcw_all_optimized_)
 *   triangulate: Xw [n_tracks*3], mean_reproj, count, ok (the function's bool) out; fewer than 4 observations / 8 rows -> ok = 0
 *   mean_reproj: Xw in (e.g. the depth-fused candidate), min_count = obser_thr_ (:1098) or 4 (:108)
 * ====================================================================================== */
int lvba_tracks_triangulate(int64_t n_tracks, const int64_t* obs_ptr, const int32_t* obs_cam, const float* obs_uv, int32_t n_cams,
                            const double* cams, const double intr[8], int32_t device, double* Xw, double* mean_reproj,
                            int32_t* count, uint8_t* ok);
int lvba_tracks_mean_reproj(int64_t n_tracks, const int64_t* obs_ptr, const int32_t* obs_cam, const float* obs_uv, int32_t n_cams,
                            const double* cams, const double intr[8], int32_t device, const double* Xw, int32_t min_count,
                            double* mean_reproj, int32_t* count, uint8_t* ok);

/* ======================================================================================
 * Boundary B7: track fusion — LvbaSystem::BuildTracksAndFuse3D
 * (reference src/lvba_system.cpp:921-1263) as one call.  In: the keypoints of all images
 * (CSR), the pairwise matches as four parallel arrays in the order the reference visits them
 * (image pairs (i < j) by ascending i then j, the matches of a pair in stored order), camera
 * poses [n_images][12] = Rcw row-major then tcw, intrinsics fx fy cx cy k1 k2 p1 p2, and the
 * depth-fused 3-D candidate of every keypoint with its validity flag (what
 * lvba_depth_backproject returns: the loop at :1020-1038).  Out: the tracks in the order the
 * reference appends them; Track::observations = the whole connected component in BFS order,
 * Track::inlier_indices as one flag per observation, Xw_fused, which candidate was chosen
 * (1 depth, 2 triangulation) and its mean reprojection error.  The (image, keypoint, inlier)
 * lists are the observation CSR of lvba_visual_lm once the inliers are kept (:1610-1617).
 * Where the reference iterates std::unordered_map<int,int> (unspecified order: the greedy
 * view-angle filter depends on it, and with it which tracks survive) the images of a component are
 * visited in the order GNU libstdc++'s container has after the reference's reserve() / insert calls
 * (map_order = LVBA_FUSE_ORDER_LIBSTDCXX, the default: what a g++ build of the reference does; with
 * it the stage reproduces the reference's own source track for track, tests/test_ref_system_pin.py,
 * tests/test_zzz_ref_gpu.py) or in ascending id (LVBA_FUSE_ORDER_ASCENDING: independent of any C++
 * library).
 * ====================================================================================== */
typedef struct lvba_fuse_opts {
  int32_t obser_thr;              /* minimum members / images / survivors (lvba_system.h:139: 3) */
  double min_view_angle_deg;      /* track_fusion/min_view_angle (8) */
  double reproj_mean_thr_px;      /* track_fusion/reproj_mean_thr (3) */
  double depth_gate_m;            /* distance to the anchor's depth point (0.12, :1050) */
  int32_t device;                 /* -1: current */
  int32_t map_order;              /* LVBA_FUSE_ORDER_*: visiting order of the three unordered_map loops (:1057, :1069, :1124) */
} lvba_fuse_opts;
#define LVBA_FUSE_ORDER_ASCENDING 0
#define LVBA_FUSE_ORDER_LIBSTDCXX 1
typedef struct lvba_fuse_summary {
  int64_t n_keypoints, n_components, n_candidates, n_tracks, n_depth_selected, n_tri_selected;
  int64_t n_rounds, n_attempts;   /* retries: a failed component is tried again from its next keypoint as BFS seed (:1199) */
  int64_t n_obs, n_inliers;       /* totals over the tracks: sizes of the export arrays */
  int64_t kernel_launches;
  double ms_total;
} lvba_fuse_summary;
typedef struct lvba_track_set lvba_track_set;
void lvba_fuse_default_opts(lvba_fuse_opts* o);
int lvba_tracks_fuse_create(int32_t n_images, const int64_t* kp_ptr /* [n_images+1] */, const float* kp_uv /* [n_kp][2] */,
                            int64_t n_matches, const int32_t* match_img_a, const int32_t* match_kp_a, const int32_t* match_img_b,
                            const int32_t* match_kp_b, const double* cams, const double intr[8], const double* kp_Xw /* [n_kp][3] */,
                            const uint8_t* kp_valid /* [n_kp] */, const lvba_fuse_opts* opts /* NULL: defaults */,
                            lvba_track_set** out, lvba_fuse_summary* summary /* may be NULL */);
int lvba_tracks_fuse_summary(const lvba_track_set* s, lvba_fuse_summary* summary);
/* arrays sized from the summary: obs_ptr [n_tracks+1], obs_* [n_obs], Xw [n_tracks][3], source / mean_reproj [n_tracks]; any but obs_ptr may be NULL */
int lvba_tracks_fuse_export(lvba_track_set* s, int64_t* obs_ptr, int32_t* obs_img, int32_t* obs_kp, uint8_t* obs_inlier, double* Xw,
                            uint8_t* source, double* mean_reproj);
int lvba_tracks_fuse_destroy(lvba_track_set* s);

/* ======================================================================================
 * The block LDL^T of the pose / camera system on its own (diagnostics, solver tests and the
 * solver line of bench.py).  Solves (A + diag(dadd)) x = rhs for a symmetric matrix of 6x6
 * blocks stored as a block envelope: row r keeps the blocks of columns first[r]..r
 * contiguously, row after row, each block row-major; first[] must be non-decreasing (what the
 * library builds from the voxel / track structure); only the lower triangle of the diagonal
 * blocks is read.  LDL^T without pivoting (A may be indefinite), as Eigen::SimplicialLDLT in
 * BALM2::damping_iter (reference include/BALM/bavoxel.hpp:695-710) and the DENSE_SCHUR
 * Cholesky of ceres::Solve (src/lvba_system.cpp:1573-1575).
 *   path  LVBA_SOLVE_AUTO: what lvba_lidar_lm / lvba_visual_lm pick for this structure;
 *         the other values pin one path and fail with LVBA_ERR_UNSUPPORTED if the structure
 *         does not allow it.  chunks: for LVBA_SOLVE_CHUNKED (0 = library default).
 *   reps  >= 1 solves; ms (may be NULL) = fastest of them, device time by CUDA events.
 *   info  (may be NULL) int32[4]: path taken, chunks, tree levels, kernel launches per solve.
 * ====================================================================================== */
#define LVBA_SOLVE_AUTO 0
#define LVBA_SOLVE_ONE_CTA 1        /* one register-window factorisation (columns of <= 30 blocks) */
#define LVBA_SOLVE_TWISTED 2        /* two-ended elimination on two SMs, joined at one separator */
#define LVBA_SOLVE_CHUNKED 3        /* substructured: chunk interiors + tree of separators, one CTA per node */
#define LVBA_SOLVE_SHARED_WINDOW 4  /* shared-memory window (columns of <= 320 blocks) */
#define LVBA_SOLVE_ANY_WIDTH 5      /* device-wide passes, any envelope */
int lvba_env_solve(int32_t n, const int32_t* first, const double* blocks, const double* dadd, const double* rhs,
                   double* x, int32_t path, int32_t chunks, int32_t reps, int32_t device, double* ms, int32_t* info);

/* ======================================================================================
 * Multi-GPU (one process per GPU).  The path shards by contiguous pose-block rows
 * (SURVEY.md §8e): voxel / track -> owner of its lowest pose / camera index.  Every rank
 * passes the FULL problem to *_create; after lvba_comm_init each rank keeps only its shard
 * of voxels / tracks on its GPU.  The pose / camera system is ROW-OWNED: the substructured
 * solver's chunks are the multi-GPU unit, a rank builds and factorises the rows of its own
 * chunks, only the <= band-width block rows a rank's voxels reach into its right neighbour's
 * range travel (ncclSend/ncclRecv), the ranks' separator complements meet in one
 * ncclAllGather (~0.8 MB per rank), the small top tree is solved redundantly and the update
 * is assembled by an all-reduce of 48 bytes per pose; g, the right-hand sides, diagonals
 * and the scalar costs (O(poses) data) use ncclAllReduce.  Structures the solver cannot
 * cut per rank fall back to an all-reduce of the full matrix.  NCCL is dlopen()ed
 * (libnccl.so.2) on first use.
 * ====================================================================================== */
#define LVBA_NCCL_ID_BYTES 128
int lvba_comm_unique_id(void* id_out /* LVBA_NCCL_ID_BYTES */);
/* Row ownership of a problem created after lvba_comm_init: this rank holds the block rows [row_begin, row_end) of H (pose rows)
 * / of the reduced camera system; *sharded = 1 when the system is row-owned and solved by the substructured solver with its
 * chunks spread over the ranks (only <= band-width boundary rows travel between neighbours), 0 when every rank
 * holds the all-reduced full system (one GPU, or a structure that cannot be cut per rank).  lvba_*_get_system returns the
 * owned rows only when sharded. */
int lvba_lidar_owned_rows(lvba_lidar_problem* p, int32_t* row_begin, int32_t* row_end, int32_t* sharded);
int lvba_visual_owned_rows(lvba_visual_problem* p, int32_t* row_begin, int32_t* row_end, int32_t* sharded);
/* NCCL payload (bytes handed to send-type calls by this rank) since the previous call of this function */
int64_t lvba_comm_bytes_sent(void);
int lvba_comm_init(int32_t n_ranks, int32_t rank, const void* id /* LVBA_NCCL_ID_BYTES */, int32_t device);
int lvba_comm_destroy(void);
int lvba_comm_info(int32_t* n_ranks, int32_t* rank);
/* Host-only shard rule (no GPU needed; used by the gloo CPU tests): owner rank of a unit whose
 * lowest pose index is `min_pose` when `n_rows` pose-block rows are split over `n_ranks`. */
int32_t lvba_shard_owner(int32_t min_pose, int32_t n_rows, int32_t n_ranks);

#ifdef __cplusplus
}
#endif
#endif /* LVBA_B200_H */
