// velo_dev.cuh — device-side data layout and kernel launchers of the B200-native VELO front end.
//
// Layout in HBM (one context = one GPU).  S = max_slots, N = max_points (padded to 128), R = max_rings,
// C = num_cams, F = max_features, MM = max_matches.  Everything is a dense [slot][...] array so that one
// launch covers a whole batch of frames (grid.y / grid.z = slot):
//   raw        float4 [S][N]          KITTI {x,y,z,reflectance} as uploaded            (kitti.h:121-152); or packed {x,y,z} records (raw_stride = 3)
//   flagbits   u32    [S][N/32]       ring-boundary flags (kitti.h:166)
//   ring_start int    [S][R+1]        ring r = pts[ring_start[r] .. ring_start[r+1])
//   pts        float4 [S][N]          ring-ordered cam-0 frame {x,y,z,1}               (kitti.h:154-185)
//   sorted     float4 [S][N]          per ring counting-sorted by azimuth bin, .w = index in ring (int bits)
//   cell_start int    [S][R][AZ+1]    start of (ring, azimuth bin) in `sorted` (slot-relative)
//   sec_box    float4 [S][R][SEC]     {elev lo, elev hi, range min, range max} of the ring inside one azimuth sector
//   mask_lo/hi u64    [S][SEC][EL][W] cumulative ring bit masks over elevation buckets (W = ceil(R/64) words)
//   rmask_lo/hi u64   [S][SEC][RG][W] cumulative ring bit masks over range buckets (piecewise-linear log2)
//   proj       float2 [S][C][N]       canonical projection, ring r at offset ring_start[r] (velo.h:366)
//   valid      float4 [S][C][N]       matching cam-0 points (velo.h:368)
//   proj_count int    [S][C][R],  proj_yrange float2 [S][C][R] (y range of the ring's projections, prunes the association)
//   kp         float2 [S][2][C][F]    keypoints (set 0 = detected, set 1 = tracked)
//   has_depth  int    [S][2][C][F],  kpwd float4 [S][2][C][F],  n_hits int [S][2][C]   (velo.h:377-497)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "velo_gpu.h"

#ifndef VELO_AZ_BINS
#define VELO_AZ_BINS 1024
#endif
#define VELO_SECTORS 64
#define VELO_BINS_PER_SECTOR (VELO_AZ_BINS / VELO_SECTORS)
#define VELO_MAX_RINGS_HARD 256
#define VELO_EL_BUCKETS 256         /* elevation buckets of the ring-mask tables */
#define VELO_EL_MIN (-0.47f)        /* rad; elevations outside [EL_MIN, EL_MAX] clamp to the edge buckets (still conservative) */
#define VELO_EL_MAX (0.10f)
#define VELO_RG_BUCKETS 256         /* range buckets of the ring-mask tables: 64 per octave from 2 m (0.8-1.6 % steps), 2..32 m, edge buckets clamp */
#define VELO_RG_MIN 2.0f
#ifndef VELO_RG_MANT_BITS
#define VELO_RG_MANT_BITS 6         /* measured 4/5/6 bits: 24.4 / 23.6 / 23.2 ms per 200 pairs x 6 passes */
#endif
#define VELO_IDX_BITS 20            /* index-in-ring bits of the neighbour key */
#define VELO_RING_BITS 12

enum {
    VK_INGEST_FLAGS = 0, VK_INGEST_RINGS, VK_INGEST_PERMUTE, VK_INDEX_BUILD, VK_PROJECT, VK_ASSOC_SEARCH,
    VK_ASSOC_COMPACT, VK_ICP_PASS, VK_NEQ_REDUCE, VK_VISUAL, VK_INDEX_MASKS, VK_SOLVE
};

// calibration packed for kernel parameters
struct DevCalib {
    float vtc[12];            // velo_to_cam rows 0..2 (kitti.h:100-105)
    float cam_t[VELO_MAX_CAMS][3];
    float fov[VELO_MAX_CAMS][4];   // min_x, max_x, min_y, max_y as the equivalent float thresholds (hazard H3)
    float assoc_thr;          // float equivalent of `(double)|dx| < depth_assoc_thresh`
    int   abs_truncates;
    int   num_cams;
};

struct DevBuffers {
    int S, N, R, C, F, MM, P;
    float4 *raw; int *raw_stride; uint32_t *flagbits; int *n_points; int *n_rings; int *ring_start; int *status;   // raw_stride[S]: floats per uploaded record (4 = KITTI, 3 = xyz)
    float4 *pts; float4 *sorted; int *cell_start; float4 *sec_box;
    unsigned long long *mask_lo, *mask_hi; int W;   // [S][SEC][EL_BUCKETS][W]: rings with bucket(elev lo) <= b / bucket(elev hi) >= b
    unsigned long long *rmask_lo, *rmask_hi;        // [S][SEC][RG_BUCKETS][W]: the same over range buckets
    float2 *proj; float4 *valid; int *proj_count; float2 *proj_yrange;
    float2 *kp; int *n_kp; int *has_depth; float4 *kpwd; int *n_hits; int *hit_tmp; float4 *kpwd_tmp;
    int *matches; int *n_matches;
};

// pose-dependent constants computed once on the host with the host libm (hazard H8)
struct PosePack {
    double w[3], t[3];        // angle-axis, translation
    double c, s;              // cos(theta), sin(theta)
    double u[3];              // w / theta
    double dR[27];            // dR[k][i][j] = d (R(w) e_j)_i / d w_k  (forward-mode, same branch as the rotation)
    int small_angle;          // theta^2 <= DBL_EPSILON branch of AngleAxisRotatePoint
    int pad;
};

#define VELO_MAX_PASSES 6        /* ICP passes per frame pair in one launch (f2f_iterations * icp_iterations = 6) */

struct IcpPass {
    PosePack pose;
    float thr_f;              // largest float f with (double)f <= correspondence_thresh_icp/iter^4 (velo.h:829)
    float thr_excl;           // smallest float above thr_f (exclusive bound of the candidate loop)
    int iter, pad;
};

// one frame pair (source = current scan, target = previous scan) with its supplied poses
struct IcpUnit {
    int src_slot, tgt_slot;
    int skip, n_pass;
    float norm_thr_f, pad0;   // smallest float f with (double)f >= icp_norm_condition (velo.h:873)
    double loss_a;            // loss_thresh_3DPD
    double weight;            // weight_3DPD
    IcpPass pass[VELO_MAX_PASSES];
};

struct VisUnit {
    int slot1, set1, slot2, set2;
    int iter, pad;
    double pose[6];
};

struct VisTun {
    double w3d2d, w2d2d, l3d2d, l2d2d, l3d3d, outlier;
    int en2d2d, en3d2d, abs_trunc, pad;
};

// state of the device-resident Levenberg-Marquardt solve (velo_solve.cu), one per frame pair being solved; lives in device memory
struct LmState {
    double x[6], xt[6], delta[6];     // accepted pose, trial pose, last step
    double H[21], g[6];               // robustified normal equations at x
    double cost, init_cost, model_change, radius, decrease_factor;
    double function_tolerance, gradient_tolerance, parameter_tolerance;
    int iter, done, phase, reason, accepted, max_iterations, n_blocks, pad;
};

// The visual kernel inside the device-resident solve.  Per unit u: sel_in / sel_out + u * sel_stride = frozen type masks per (camera,
// match) (fixed mode reads sel_in and applies no gate; free mode may record what it chose in sel_out); lm[u] supplies the pose
// (the accepted pose x when freezing the block list, else the trial pose xt) and the convergence flag.  All members may be null.
struct VisFixed { const unsigned char *sel_in; unsigned char *sel_out; const LmState *lm; int sel_stride; int use_accepted; };

// one frozen cost3DPD block of the device-resident solve: what velo.h:875-884 captured (normal, plane point, source point)
struct __align__(16) IcpFrozen { float n[3]; int src; float o[3]; int kept; };     // src = index into pts of the source slot

// per-match parity record of the visual kernel: up to 3 blocks
struct VisMatchOut { int n; int pad; velo_vis_block b[3]; };

struct Launcher {
    cudaStream_t stream;
    // profiling hook: called before/after each launch with the kernel class
    void (*pre)(void *user, int k);
    void (*post)(void *user, int k);
    void *user;
};

void launch_ingest(const Launcher &L, const DevBuffers &B, const DevCalib &cal, int slot0, int count);
void launch_index(const Launcher &L, const DevBuffers &B, const DevCalib &cal, int slot0, int count);
void launch_project(const Launcher &L, const DevBuffers &B, const DevCalib &cal, int slot0, int count);
void launch_assoc(const Launcher &L, const DevBuffers &B, const DevCalib &cal, int slot0, int count, int set0, int nsets, int cam0, int ncams);
// units: device array [n_units]; partial: [n_units][launch_icp_runs_cap(B.N)][VELO_MAX_PASSES][64] doubles; out: [n_units][out_stride_passes][VELO_NEQ_STRIDE];
// corr optional (single unit): records of its last pass, or with corr_stride > 0 of every pass ([pass][corr_stride]);
// frozen optional: compact records of the LAST pass per unit, [unit][frozen_stride]
int launch_icp_runs_cap(int max_points);
void launch_icp(const Launcher &L, const DevBuffers &B, const DevCalib &cal, const IcpUnit *units, int n_units, int n_pass, int ctas,
                double *partial, double *out, int out_stride_passes, velo_icp_corr *corr, int corr_stride = 0,
                IcpFrozen *frozen = nullptr, int frozen_stride = 0, bool stats = false);
void launch_visual(const Launcher &L, const DevBuffers &B, const DevCalib &cal, const VisUnit *units, int n_units, VisTun tun,
                   const int *lm_valid, const float4 *lm_xyz, double *partial, double *out, VisMatchOut *match_out, int ctas,
                   VisFixed fx = VisFixed{ nullptr, nullptr, nullptr, 0, 0 }, int *bad_flag = nullptr);

// device-resident solve of n frame pairs at once (unit u <-> S[u]); poses nullable (keep the accepted pose, restart the controller)
void launch_lm_init(const Launcher &L, LmState *S, int n, const double *d_poses, int max_iterations, int *n_done);
// frozen: [n][frozen_stride]; icp_out: [n][VELO_NEQ_STRIDE] of the correspondence pass ([58] = records written); units[u].src_slot = source scan;
// partial: [n][ctas][64]; out: [n][VELO_NEQ_STRIDE]
void launch_icp_eval(const Launcher &L, const DevBuffers &B, int n, const IcpFrozen *frozen, int frozen_stride, const double *icp_out, const IcpUnit *units,
                     const LmState *S, double loss_a, double weight, double *partial, int ctas, double *out);
// e_icp / e_vis: [n][VELO_NEQ_STRIDE] or null (term absent)
void launch_lm_step(const Launcher &L, LmState *S, int n, const double *e_icp, const double *e_vis, int *n_done);
// best: [nq] keys (distance << 32 | train index), all ones where the train set is empty
void launch_hamming(const Launcher &L, const unsigned long long *q, int nq, const unsigned long long *t, int nt, int words, unsigned long long *best, int sm_count);
void launch_triangulate(const Launcher &L, int n, const int *off3, const velo_tri_obs3 *obs3, const int *off2, const velo_tri_obs2 *obs2,
                        const double *poses, int n_frames, const DevCalib &cal, double loss_a, double weight,
                        const float *init_xyz, const int *has_init, float *out_xyz, int *iterations);
