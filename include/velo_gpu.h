/*
 * velo_gpu.h — C ABI of the B200-native VELO front end (libvelo_gpu.so).
 *
 * The reference (lichunshang/vision-enhanced-lidar-odometry) has no FFI/plugin
 * layer: its per-frame front end is a set of free functions textually included
 * into main.cpp (main.cpp:53-57).  This header is the C-ABI a maintainer binds
 * instead; include/velo_dropin.hpp keeps the reference's C++ signatures on top
 * of it.  Every entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - plain C, POD only, no exceptions; every call returns a velo_status
 *     (0 = OK) and records a message readable with velo_gpu_last_error().
 *   - one context per GPU; contexts are independent (one host thread per GPU).
 *   - calls are asynchronous on the context's stream unless they return data
 *     into host memory, in which case they synchronise before returning.
 *   - "slot" = a device-resident scan (the GPU analogue of lru.h's ScanData):
 *     ring-segmented points + neighbour index + per-camera projections.
 *   - there is NO CPU fallback: without a CUDA device velo_gpu_create fails.
 */
#ifndef VELO_GPU_H
#define VELO_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VELO_GPU_ABI_VERSION 2
#define VELO_MAX_CAMS 4          /* kitti.h:4  num_cams_actual */
#define VELO_NUM_KP_SETS 2       /* main.cpp:261 (after tracking) and main.cpp:600 (after detection) */
#define VELO_NEQ 28              /* 21 upper-triangular JtJ + 6 Jtr + 1 cost */
#define VELO_NEQ_STRIDE 64       /* doubles per normal-equation record: [0,28) robustified, [28,56) raw, [56] n_blocks, [57] n_residuals, [58] n_queries, ICP only: [59] seed candidates, [60] exhaustive candidates, [61] rings scanned, [62] rings considered */

typedef enum velo_status {
    VELO_OK = 0,
    VELO_ERR_NO_DEVICE = 1,      /* no CUDA device / wrong architecture: there is no CPU fallback */
    VELO_ERR_CUDA = 2,
    VELO_ERR_INVALID_ARG = 3,
    VELO_ERR_CAPACITY = 4,       /* more points / rings / features than the context was created for */
    VELO_ERR_STATE = 5           /* e.g. associate before project */
} velo_status;

/* residual type tags, same order as velo.h:3-8 (+3DPD) */
enum { VELO_RES_3D3D = 0, VELO_RES_3D2D = 1, VELO_RES_2D3D = 2, VELO_RES_2D2D = 3, VELO_RES_3DPD = 4 };

/* Tunables: defaults equal kitti.h:3-35 and the feature macros of main.cpp:42-49. */
typedef struct velo_gpu_params {
    int num_cams;                 /* kitti.h:3   (2) */
    int icp_skip;                 /* kitti.h:8   (200); BASELINE configs use 1 */
    int f2f_iterations;           /* kitti.h:9   (2) */
    int icp_iterations;           /* kitti.h:10  (3) */
    int enable_2d2d;              /* main.cpp:44 (1) */
    int enable_3d2d;              /* main.cpp:45 (1) */
    int abs_truncates;            /* SURVEY hazard H1: 1 = unqualified abs() bound to int abs(int) (velo.h:416,419,709) */
    int reserved0;
    double weight_3D2D;           /* kitti.h:20-35 */
    double weight_2D2D;
    double weight_3DPD;
    double loss_thresh_3D2D;
    double loss_thresh_2D2D;
    double loss_thresh_3DPD;
    double loss_thresh_3D3D;
    double depth_assoc_thresh;
    double outlier_reject;
    double correspondence_thresh_icp;
    double icp_norm_condition;
    /* capacities of the context (device memory is allocated once at create) */
    int max_slots;                /* device-resident scans (lru.h:33 keeps 50) */
    int max_points;               /* per scan; lru.h:5 says ~130 000 */
    int max_rings;                /* ring count is data dependent (kitti.h:166-173) */
    int max_features;             /* per image per set; kitti.h:7 corner_count = 3000 */
    int max_matches;              /* per camera per frame pair */
    int max_icp_passes;           /* f2f_iterations * icp_iterations per batched call */
    int ctas_per_icp_unit;        /* CTAs sharing the queries of one frame pair (all its passes); 0 = auto.  Results do not depend on it:
                                     sums are kept per run of 64 queries and added in run order */
    int reserved1;
} velo_gpu_params;

/* Calibration, the kitti.h:40-51 globals packed for the device. */
typedef struct velo_gpu_calib {
    float velo_to_cam[16];            /* row-major 4x4, kitti.h:100-105 */
    float cam_trans[VELO_MAX_CAMS][4];/* K^-1 * P[:,3], kitti.h:74-78 (w unused) */
    float cam_K[VELO_MAX_CAMS][9];    /* cam_intrinsic, kitti.h:80 */
    float cam_Kinv[VELO_MAX_CAMS][9]; /* cam_intrinsic_inv, kitti.h:81 */
    double min_x[VELO_MAX_CAMS], max_x[VELO_MAX_CAMS];  /* kitti.h:87-97 (stored as double, kitti.h:51) */
    double min_y[VELO_MAX_CAMS], max_y[VELO_MAX_CAMS];
    int img_width, img_height;        /* kitti.h:37-38, overwritten by loadImage kitti.h:196-197 */
} velo_gpu_calib;

/* One ICP correspondence (velo.h:806-874 locals), for parity checks. */
typedef struct velo_icp_corr {
    int32_t src_ring, src_idx;        /* sm, smi */
    int32_t np_s_i, np_i;             /* best ring / point        velo.h:836-842 */
    int32_t np_s_j, np_j;             /* runner-up ring / point   velo.h:843-847 */
    int32_t np_k;                     /* third point on ring np_s_i, velo.h:852-863 */
    int32_t kept;                     /* 0: <2 rings (velo.h:849), 2: degenerate normal (velo.h:873), 1: residual block added */
    float   normal[3];                /* velo.h:868-874 */
    float   v0[3];                    /* plane offset, velo.h:865 */
    double  residual;                 /* cost3DPD at the supplied pose, costfunctions.h:40-53 */
    double  jacobian[6];              /* d residual / d (angle-axis, translation) */
} velo_icp_corr;

/* One visual residual block (velo.h:662-789), in residualStats order (velo.h:934-976). */
typedef struct velo_vis_block {
    int32_t cam, match;               /* index into the match list of that camera */
    int32_t type, n_res;              /* VELO_RES_*, 3/2/2/1 */
    double  residual[3];
    double  jacobian[18];             /* row-major n_res x 6 */
} velo_vis_block;

typedef struct velo_gpu_ctx velo_gpu_ctx;

/* ---------------------------------------------------------------- lifecycle */
int  velo_gpu_abi_version(void);
/* kitti.h:3-35 defaults + capacities for one KITTI frame pair */
int  velo_gpu_default_params(velo_gpu_params *p);
/* replaces loadCalibration (kitti.h:59-108): P = 4 x (3x4 row-major), Tr = 3x4 row-major */
int  velo_gpu_calib_from_kitti(const float P[48], const float Tr[12], int img_width, int img_height,
                               velo_gpu_calib *out);
/* pixel2canonical / canonical2pixel (velo.h:10-26): n points, interleaved (x,y) */
int  velo_pixel2canonical(const velo_gpu_calib *calib, int cam, const float *pix, int n, float *canon);
int  velo_canonical2pixel(const velo_gpu_calib *calib, int cam, const float *canon, int n, float *pix);

/* ---------------------------------------------------------------- KITTI wire formats (host only) */
/* calib.txt as parsed by kitti.h:66-105: a label token then 12 floats, for P0..P3, then "Tr:" + 12 floats. */
int  velo_kitti_load_calib(const char *path, float P[48], float Tr[12]);
/* velodyne/NNNNNN.bin as read by loadPoints (kitti.h:121-152): float {x,y,z,reflectance} records.  Returns the number of
 * points through *n (at most max_points are stored; *n is the count in the file). */
int  velo_kitti_load_scan(const char *path, float *xyzr, int max_points, int *n);
/* one line of results/<seq>.txt as written by output_line (kitti.h:202-216): the first 3 rows of a row-major 4x4 pose,
 * 12 numbers separated (and followed) by a space, default ostream formatting (%g, 6 significant digits). */
int  velo_kitti_format_pose(const double T[16], char *buf, int buflen);

int  velo_gpu_create(int device, const velo_gpu_params *params, const velo_gpu_calib *calib, velo_gpu_ctx **out);
int  velo_gpu_destroy(velo_gpu_ctx *ctx);
const char *velo_gpu_last_error(const velo_gpu_ctx *ctx); /* ctx may be NULL: message of a failed create */
int  velo_gpu_sync(velo_gpu_ctx *ctx);
int  velo_gpu_device_name(velo_gpu_ctx *ctx, char *buf, int buflen);

/* pinned host memory for the batched path */
int  velo_gpu_host_alloc(void **ptr, uint64_t bytes);
int  velo_gpu_host_free(void *ptr);

/* device timing on the context's stream (CUDA events) */
int  velo_gpu_timer_begin(velo_gpu_ctx *ctx);
int  velo_gpu_timer_end(velo_gpu_ctx *ctx, float *ms);   /* synchronises */
/* per-kernel accumulated device time since the last reset; names via velo_gpu_kernel_name */
#define VELO_NUM_KERNELS 12
int  velo_gpu_profile_enable(velo_gpu_ctx *ctx, int on);
int  velo_gpu_profile_reset(velo_gpu_ctx *ctx);
int  velo_gpu_profile_read(velo_gpu_ctx *ctx, float ms[VELO_NUM_KERNELS], int launches[VELO_NUM_KERNELS]); /* synchronises */
const char *velo_gpu_kernel_name(int k);
/* search statistics of the correspondence kernel (diagnostics, off by default: 1.6 % of the kernel): slots 60..62 of every
 * normal-equation record = candidate distance evaluations, target rings scanned, ring-mask bits taken, summed over the pass's queries */
int  velo_gpu_search_stats_enable(velo_gpu_ctx *ctx, int on);

/* ---------------------------------------------------------------- single-frame path (drop-in) */
/* ScanData ctor (lru.h:12-28): loadPoints layout (kitti.h:121-152) -> segmentPoints (kitti.h:154-185)
 * -> neighbour index (replaces 64x KdTreeFLANN::setInputCloud, lru.h:17-20).  xyzr: n x float4. */
int  velo_gpu_scan_upload(velo_gpu_ctx *ctx, int slot, const float *xyzr, int n);
/* same, for a scan that is ALREADY ring-segmented in the cam-0 frame (the `scans` vector of kitti.h:156 flattened):
 * xyz1 = n x {x,y,z,*}, ring_start[n_rings+1].  Skips segmentPoints, builds the neighbour index. */
int  velo_gpu_scan_upload_rings(velo_gpu_ctx *ctx, int slot, const float *xyz1, const int *ring_start, int n_rings);
int  velo_gpu_scan_info(velo_gpu_ctx *ctx, int slot, int *n_points, int *n_rings);
/* ring-ordered cam-0-frame points (n x {x,y,z,1}) and ring_start[n_rings+1] — the `scans` vector of kitti.h:156 */
int  velo_gpu_scan_download(velo_gpu_ctx *ctx, int slot, float *xyz1, int *ring_start);

/* projectLidarToCamera (velo.h:329-375) for camera `cam`; results stay on the device */
int  velo_gpu_project(velo_gpu_ctx *ctx, int slot, int cam);
/* ring_count[n_rings]; proj: total x (x,y); valid: total x {x,y,z,1}, rings concatenated in order */
int  velo_gpu_project_download(velo_gpu_ctx *ctx, int slot, int cam, int *ring_count, float *proj, float *valid, int *total);

/* install an externally computed projection (the `projection` / `scans_valid` arguments of velo.h:377-383) for
 * (slot, cam): ring_count[n_rings of the slot], proj/valid ring-concatenated; every ring_count[s] <= ring length */
int  velo_gpu_projection_upload(velo_gpu_ctx *ctx, int slot, int cam, const int *ring_count, const float *proj, const float *valid);

/* featureDepthAssociation (velo.h:377-497): kp = F x (x,y) canonical; has_depth[F]; kpwd = n_hits x {x,y,z,1} */
int  velo_gpu_depth_assoc(velo_gpu_ctx *ctx, int slot, int cam, int set, const float *kp, int F,
                          int *has_depth, float *kpwd, int *n_hits);

/* install externally held association results for (slot, cam, set) — the keypoints / has_depth / keypoints_with_depth containers of
 * velo.h:601-606 as the caller holds them (kp = F x (x,y), has_depth[F] = -1 or an index < n_hits, kpwd = n_hits x {x,y,z,*}) —
 * so that velo_gpu_visual_residuals / velo_gpu_frame_to_frame can read them from device memory.  Used by the frameToFrame
 * adapter of velo_dropin.hpp, which receives these containers from main.cpp:388-405. */
int  velo_gpu_assoc_upload(velo_gpu_ctx *ctx, int slot, int cam, int set, const float *kp, int F,
                           const int *has_depth, const float *kpwd, int n_hits);

/* ICP block of frameToFrame (velo.h:800-895) at a supplied pose: transform_point (utility.h:97-103),
 * per-ring 1-NN + top-2 rings + third point + normal (velo.h:806-874), cost3DPD residual/Jacobian
 * (costfunctions.h:17-58) and the normal equations Ceres would form (velo.h:885-902).
 * slot_M = current frame (queries), slot_S = previous frame (targets + index).
 * corr may be NULL; otherwise capacity must be >= number of queries (sum over rings of ceil(len/icp_skip)).
 * neq: VELO_NEQ_STRIDE doubles. */
int  velo_gpu_icp_pass(velo_gpu_ctx *ctx, int slot_M, int slot_S, const double pose[6], int iter, int icp_skip,
                       velo_icp_corr *corr, int corr_capacity, int *n_queries, int *n_kept, double *neq);

/* The same block for ALL ICP passes of one frameToFrame call (velo.h:616,800: f2f_iterations x icp_iterations passes) at
 * supplied poses, in ONE launch of the fused kernel — the code path of the batched front end, where pass p+1 is seeded by the
 * correspondences of pass p.  poses[n_passes][6], iters[n_passes] (the `iter` of velo.h:829 for each pass), n_passes <=
 * max_icp_passes.  corr (nullable) receives the records of EVERY pass, [n_passes][corr_capacity]; neq [n_passes][VELO_NEQ_STRIDE].
 * Seeding only changes the order in which candidates are visited: each pass equals velo_gpu_icp_pass at its pose, bit for bit. */
int  velo_gpu_icp_passes(velo_gpu_ctx *ctx, int slot_M, int slot_S, const double *poses, const int *iters, int n_passes, int icp_skip,
                         velo_icp_corr *corr, int corr_capacity, int *n_queries, double *neq);

/* visual residual assembly of frameToFrame (velo.h:622-792) for all cameras at a supplied pose.
 * Frame1 = (slot1,set1) current, frame2 = (slot2,set2) previous: their keypoints / has_depth / kp_with_depth
 * are the device-resident results of velo_gpu_depth_assoc.  matches: per camera n_matches[c] pairs
 * (point1, point2) stored consecutively; lm_valid/lm_xyz (nullable) = landmarks_at_frame lookup per match
 * (velo.h:634-644), lm_xyz = {x,y,z,1} per match.
 * blocks (nullable, capacity 3 per match) receive the residual blocks in residualStats order. */
int  velo_gpu_visual_residuals(velo_gpu_ctx *ctx, int slot1, int set1, int slot2, int set2,
                               const int *n_matches, const int *matches, const int *lm_valid, const float *lm_xyz,
                               const double pose[6], int iter,
                               velo_vis_block *blocks, int block_capacity, int *n_blocks, double *neq);

/* frameToFrame (velo.h:598-919) with the per-pass ceres::Solve replaced by a device-resident Levenberg-Marquardt solve
 * (SURVEY.md §8(f1)): for iter = 1..f2f_iterations { freeze the visual blocks at the current transform (velo.h:622-792);
 * for icp_iter < icp_iterations { freeze the ICP correspondences at the current transform (velo.h:806-894); minimise the
 * robustified cost of the frozen blocks over the 6-vector } }.  The trust-region policy restates Ceres' defaults; agreement
 * with a Ceres build is "to solver tolerance" (Ceres is not available here), agreement with the oracle's identical
 * restatement is tested.  n_matches may be NULL (no visual terms); enable_icp = 0 drops the ICP blocks (main.cpp:43).
 * transform is in/out, like velo.h:611. */
#define VELO_MAX_SOLVES 16
typedef struct velo_f2f_report {
    int n_solves;
    int lm_iterations[VELO_MAX_SOLVES];   /* trial evaluations after the initial one */
    int accepted_steps[VELO_MAX_SOLVES];
    int reason[VELO_MAX_SOLVES];          /* 1 function tol, 2 gradient tol, 3 parameter tol, 4 radius, 5 no blocks, 6 max iterations */
    int n_blocks[VELO_MAX_SOLVES];
    double initial_cost[VELO_MAX_SOLVES], final_cost[VELO_MAX_SOLVES];
    double pose[VELO_MAX_SOLVES][6];      /* transform after each solve */
} velo_f2f_report;
int  velo_gpu_frame_to_frame(velo_gpu_ctx *ctx, int slot_M, int set1, int slot_S, int set2,
                             const int *n_matches, const int *matches, const int *lm_valid, const float *lm_xyz,
                             int enable_icp, int icp_skip, double transform[6], velo_f2f_report *report);

/* The same schedule for every frame pair (slot s, slot s-1) of a batch at once, s in [slot0 + (first_has_prev ? 0 : 1), slot0 + count):
 * the batch's ingest / index / projection / association stages must have run (velo_gpu_batch_run); frame1 = slot s with its tracked
 * keypoints (set 1), frame2 = slot s-1 with its detected keypoints (set 0), matches as uploaded with the batch, icp_skip from the
 * context's params, no landmarks.  All pairs solve side by side on the device (one Levenberg-Marquardt controller per pair);
 * this is the live, pose-dependent loop of velo.h:616-907 — each correspondence pass runs at the pose its pair has reached.
 * transforms [count][6] in/out (entry of a slot without a previous scan is left untouched), reports [count] nullable. */
int  velo_gpu_batch_frame_to_frame(velo_gpu_ctx *ctx, int slot0, int count, int first_has_prev, int enable_visual, int enable_icp,
                                   double *transforms, velo_f2f_report *reports);

/* residual types of the LAST f2f iteration of the preceding velo_gpu_frame_to_frame call, per camera and match: sel[cam *
 * max_matches + i] is a bit mask, 1 = 3D3D, 2 = 2D2D, 4 = 3D2D, 8 = 2D3D (blocks are added in that order, velo.h:662-789).  This is
 * what frameToFrame leaves in good_matches / residual_type (velo.h:624-625,690-692).  capacity >= num_cams * max_matches. */
int  velo_gpu_f2f_selection(velo_gpu_ctx *ctx, unsigned char *sel, int capacity);
/* util::pose_mat2vec (utility.h:67-82; despite its name: 6-vector -> 4x4): T = [R(angle-axis) | t], row-major 4x4. */
int  velo_pose_vec2mat(const double transform[6], double T[16]);

/* triangulatePoint (velo.h:1027-1130) for n_landmarks at once: each landmark is a 3-parameter problem
 * over its 3-D observations (triangulation3D, TrivialLoss) and 2-D observations (triangulation2D, Scaled(Cauchy)), listed in
 * the order the reference adds them (camera-major, frame ascending) in CSR form (off3/off2 have n_landmarks+1 entries).
 * camera_poses: n_frames x 6 (angle-axis, translation; velo.h:1030).  has_init/init_xyz (nullable) = `initial_guess` + `point`;
 * without an initial guess the start is (0,0,10) and the first 3-D observation alone initialises (velo.h:1041,1082-1085).
 * out_xyz: n_landmarks x 3 floats (velo.h:1127-1129); iterations (nullable): LM trial evaluations per landmark. */
typedef struct velo_tri_obs3 { int32_t frame; float x, y, z; } velo_tri_obs3;
typedef struct velo_tri_obs2 { int32_t frame, cam; float x, y; } velo_tri_obs2;
int  velo_gpu_triangulate(velo_gpu_ctx *ctx, int n_landmarks, const int *off3, const velo_tri_obs3 *obs3,
                          const int *off2, const velo_tri_obs2 *obs2, const double *camera_poses, int n_frames,
                          const float *init_xyz, const int *has_init, float *out_xyz, int *iterations);

/* matchFeatures (velo.h:499-550): brute-force Hamming 1-NN of every query descriptor among the train
 * descriptors (cv::BFMatcher(NORM_HAMMING)::match / cv::cuda::DescriptorMatcher, velo.h:517-531; ties -> lower train index),
 * then the reference's filter: keep (query, train) unless distance > max(1.5 * min_distance, match_thresh) (velo.h:536-548).
 * Descriptors are desc_bytes (multiple of 8, <= 64; FREAK = 64) bytes per row.  pairs: capacity n_query x 2 ints.
 * best_idx / best_dist (nullable, n_query each) receive the unfiltered 1-NN. */
int  velo_gpu_match_hamming(velo_gpu_ctx *ctx, const uint8_t *query, int n_query, const uint8_t *train, int n_train, int desc_bytes,
                            double match_thresh, int *pairs, int *n_pairs, int *best_idx, int *best_dist);

/* ---------------------------------------------------------------- batched path (throughput) */
/* A batch is `count` consecutive slots starting at slot0.  Inputs are concatenated with fixed strides:
 *   scans   [count][max_points][4] float,   n_points[count]      (KITTI .bin records {x,y,z,reflectance}, kitti.h:142;
 *           with scan_stride_floats == 3: [count][max_points][3], {x,y,z} only — loadPoints drops the reflectance anyway,
 *           kitti.h:145-148 — a quarter less to move over PCIe; only n_points[i] records of a scan are copied either way)
 *   kp      [count][VELO_NUM_KP_SETS][num_cams][max_features][2] float, n_kp[count][sets][cams]
 *   matches [count][num_cams][max_matches][2] int, n_matches[count][cams]   (frame pair slot s / slot s-1)
 *   poses   [count][n_passes][6] double, pass_iter[n_passes]  (iter value of each ICP pass)
 *   vis_poses [count][f2f_iterations][6] double
 * Pointers are HOST pointers (pinned preferred); upload copies them to device staging. */
typedef struct velo_batch_inputs {
    const float  *scans;     const int *n_points;
    const float  *kp;        const int *n_kp;
    const int    *matches;   const int *n_matches;
    const double *icp_poses; const int *pass_iter; int n_passes;
    const double *vis_poses; int n_vis_iters;
    int scan_stride_floats;          /* 0 or 4: KITTI float4 records; 3: xyz records */
} velo_batch_inputs;

int  velo_gpu_batch_upload(velo_gpu_ctx *ctx, int slot0, int count, const velo_batch_inputs *in);
/* stages bitmask */
enum { VELO_STAGE_INGEST = 1, VELO_STAGE_INDEX = 2, VELO_STAGE_PROJECT = 4, VELO_STAGE_ASSOC = 8,
       VELO_STAGE_ICP = 16, VELO_STAGE_VISUAL = 32, VELO_STAGE_ALL = 63 };
/* runs the selected stages for slots [slot0, slot0+count).  ICP/visual pair slot s with slot s-1 and are
 * skipped for the first slot of the batch when first_has_prev == 0. */
int  velo_gpu_batch_run(velo_gpu_ctx *ctx, int slot0, int count, int stages, int first_has_prev);
/* icp_neq [count][n_passes][VELO_NEQ_STRIDE], vis_neq [count][n_vis_iters][VELO_NEQ_STRIDE],
 * has_depth [count][sets][cams][max_features], n_hits [count][sets][cams]; any pointer may be NULL. Synchronises. */
int  velo_gpu_batch_download(velo_gpu_ctx *ctx, int slot0, int count, double *icp_neq, double *vis_neq,
                             int *has_depth, int *n_hits);
/* keypoints_with_depth of the batch (featureDepthAssociation's output cloud, velo.h:479-481): kpwd [count][sets][cams][max_features][4]
 * floats {x, y, z, 1}; the first n_hits[slot][set][cam] records of an image are valid and has_depth[k] indexes them.  Optional: the
 * batched frameToFrame consumes the cloud on the device.  (Per-query correspondence records exist per frame pair, velo_gpu_icp_passes:
 * 128 B x queries x passes is not a batch-sized output.)  Synchronises. */
int  velo_gpu_batch_download_kpwd(velo_gpu_ctx *ctx, int slot0, int count, float *kpwd);
/* the whole front end for a batch in one call: upload (pinned host buffers) -> all stages -> download, with the upload of
 * chunk c+1 overlapping the kernels of chunk c on a second stream (chunk = frames per chunk; 0 = a small first chunk — the only
 * upload nothing can hide — then geometric growth up to count/8, the factor (1.15 .. 2) adapted from the previous call so that a
 * chunk never takes longer to arrive than its predecessor takes to compute).  Slot slot0 is the halo scan when slot0 == 0 (no
 * frame pair for it).  Results do not depend on the chunking.  Synchronises. */
int  velo_gpu_batch_frontend(velo_gpu_ctx *ctx, int slot0, int count, const velo_batch_inputs *in, int chunk,
                             double *icp_neq, double *vis_neq, int *has_depth, int *n_hits);
/* the same call, also returning keypoints_with_depth (layout of velo_gpu_batch_download_kpwd): each chunk's cloud travels back on
 * its own stream as soon as the chunk's association stage is done, under the kernels of the later chunks */
int  velo_gpu_batch_frontend_kpwd(velo_gpu_ctx *ctx, int slot0, int count, const velo_batch_inputs *in, int chunk,
                                  double *icp_neq, double *vis_neq, int *has_depth, int *n_hits, float *kpwd);
/* number of kernel launches issued by this context since creation */
int  velo_gpu_launch_count(velo_gpu_ctx *ctx, int64_t *launches);
/* per-slot counts needed to state algorithmic bytes: n_points[count], n_rings[count],
 * proj_total[count][num_cams] (in-FOV survivors M), status[count] (velo_status per scan). Synchronises. */
int  velo_gpu_batch_counts(velo_gpu_ctx *ctx, int slot0, int count, int *n_points, int *n_rings, int *proj_total, int *status);

#ifdef __cplusplus
}
#endif
#endif /* VELO_GPU_H */
