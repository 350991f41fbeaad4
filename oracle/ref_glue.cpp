/*
 * ref_glue.cpp — C entry points around the reference's OWN hot-path source lines.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/build_ref.py into oracle/_ref/libvelo_ref.so, only where
 * /root/reference exists.  The `ref_*.inc` files included below are line ranges cut from the reference
 * headers at build time into a temporary directory (never into this repository):
 *     ref_utility_a.inc   utility.h:1-56      constants, util::linterpolate / subtract_assign / norm2 ...
 *     ref_utility_b.inc   utility.h:97-103    util::transform_point
 *     ref_kitti_consts    kitti.h:3-35        tunables
 *     ref_kitti_segment   kitti.h:154-185     segmentPoints
 *     ref_velo_enum       velo.h:3-8          ResidualType
 *     ref_velo_project    velo.h:329-375      projectLidarToCamera
 *     ref_velo_assoc      velo.h:377-497      featureDepthAssociation
 *     ref_velo_match      velo.h:499-550      matchFeatures (brute-force Hamming + min-distance filter)
 *     ref_velo_triangulate velo.h:1027-1130   triangulatePoint (block assembly; ceres::Solve = the LM stand-in of the shim)
 *     ref_velo_visual     velo.h:622-792      frameToFrame, visual residual assembly (loop body of one iter)
 *     ref_velo_icp_a/b    velo.h:806-874 / 875-894   frameToFrame, ICP correspondence + cost3DPD blocks
 * plus costfunctions.h included whole.  Third-party types come from ref_shim/velo_ref_shim.hpp.
 */
#include "velo_ref_shim.hpp"
#include "velo_gpu.h" /* POD records only */
#include <cstring>

/* ---- reference text: utility.h (class util is closed here because the slice stops inside it) */
#include "ref_utility_a.inc"
#include "ref_utility_b.inc"
};

/* ---- reference text: kitti.h tunables; the globals of kitti.h:37-51 that the slices read are declared here */
#include "ref_kitti_consts.inc"
std::vector<Eigen::Vector3f> cam_trans;
Eigen::Matrix4f velo_to_cam;
std::vector<double> min_x, max_x, min_y, max_y;
#include "ref_kitti_segment.inc"

/* ---- reference text: costfunctions.h (whole file) with the feature macros of main.cpp:44-45 */
#define ENABLE_2D2D
#define ENABLE_3D2D
#include "costfunctions.h"

#include "ref_velo_enum.inc"
#include "ref_velo_project.inc"
#include "ref_velo_assoc.inc"
#include "ref_velo_match.inc"
static int g_num_cams = 2;
#define num_cams g_num_cams
#include "ref_velo_triangulate.inc"
#undef num_cams

typedef pcl::PointCloud<pcl::PointXYZ> Cloud;

static void set_calib(const velo_gpu_calib *c) {
    cam_trans.clear(); min_x.clear(); max_x.clear(); min_y.clear(); max_y.clear();
    for (int cam = 0; cam < VELO_MAX_CAMS; cam++) {
        cam_trans.push_back(Eigen::Vector3f(c->cam_trans[cam][0], c->cam_trans[cam][1], c->cam_trans[cam][2]));
        min_x.push_back(c->min_x[cam]); max_x.push_back(c->max_x[cam]);
        min_y.push_back(c->min_y[cam]); max_y.push_back(c->max_y[cam]);
    }
    for (int i = 0; i < 16; i++) velo_to_cam.m[i] = c->velo_to_cam[i];
}
static std::vector<Cloud::Ptr> to_rings(const float *pts, const int *rs, int nr) {
    std::vector<Cloud::Ptr> v;
    for (int s = 0; s < nr; s++) {
        Cloud::Ptr c(new Cloud);
        for (int i = rs[s]; i < rs[s + 1]; i++) c->push_back(pcl::PointXYZ(pts[4 * i], pts[4 * i + 1], pts[4 * i + 2]));
        v.push_back(c);
    }
    return v;
}
static void neq_add(double *neq, const double *r, const double *J, int nr, const double rho[3]) {
    double s = 0; for (int i = 0; i < nr; i++) s += r[i] * r[i];
    int o = 0;
    for (int a = 0; a < 6; a++) for (int b = a; b < 6; b++, o++) {
        double h = 0; for (int i = 0; i < nr; i++) h += J[6 * i + a] * J[6 * i + b];
        neq[o] += rho[1] * h; neq[28 + o] += h;
    }
    for (int a = 0; a < 6; a++) {
        double g = 0; for (int i = 0; i < nr; i++) g += J[6 * i + a] * r[i];
        neq[21 + a] += rho[1] * g; neq[28 + 21 + a] += g;
    }
    neq[27] += 0.5 * rho[0]; neq[28 + 27] += 0.5 * s; neq[56] += 1; neq[57] += nr;
}

static int g_icp_skip = 200;

extern "C" {

int ref_segment(const float *xyzr, int n, const velo_gpu_calib *c, float *out_xyz1, int *ring_start, int max_rings) {
    set_calib(c);
    Cloud::Ptr cloud(new Cloud);
    for (int i = 0; i < n; i++) cloud->points.push_back(pcl::PointXYZ(xyzr[4 * i], xyzr[4 * i + 1], xyzr[4 * i + 2])); /* kitti.h:145 */
    std::vector<Cloud::Ptr> scans;
    segmentPoints(cloud, scans);
    int o = 0; ring_start[0] = 0;
    for (size_t s = 0; s < scans.size(); s++) {
        for (size_t i = 0; i < scans[s]->size(); i++, o++) {
            const pcl::PointXYZ &p = scans[s]->at(i);
            out_xyz1[4 * o] = p.x; out_xyz1[4 * o + 1] = p.y; out_xyz1[4 * o + 2] = p.z; out_xyz1[4 * o + 3] = 1.0f;
        }
        if ((int)s + 1 <= max_rings) ring_start[s + 1] = o;
    }
    return (int)scans.size();
}

int ref_project(const float *pts, const int *rs, int nr, const velo_gpu_calib *c, int cam, int *ring_count, float *proj, float *valid) {
    set_calib(c);
    std::vector<Cloud::Ptr> scans = to_rings(pts, rs, nr), scans_valid;
    std::vector<std::vector<cv::Point2f>> projection;
    projectLidarToCamera(scans, projection, scans_valid, cam);
    int o = 0;
    for (int s = 0; s < nr; s++) {
        ring_count[s] = (int)projection[s].size();
        for (size_t i = 0; i < projection[s].size(); i++, o++) {
            proj[2 * o] = projection[s][i].x; proj[2 * o + 1] = projection[s][i].y;
            const pcl::PointXYZ &p = scans_valid[s]->at(i);
            valid[4 * o] = p.x; valid[4 * o + 1] = p.y; valid[4 * o + 2] = p.z; valid[4 * o + 3] = 1.0f;
        }
    }
    return o;
}

int ref_depth_assoc(const float *valid, const float *proj, const int *ring_count, int nr, const float *kp, int F, int *has_depth_out, float *kpwd) {
    std::vector<Cloud::Ptr> scans; std::vector<std::vector<cv::Point2f>> projection;
    int o = 0;
    for (int s = 0; s < nr; s++) {
        Cloud::Ptr c(new Cloud); projection.emplace_back();
        for (int i = 0; i < ring_count[s]; i++, o++) {
            c->push_back(pcl::PointXYZ(valid[4 * o], valid[4 * o + 1], valid[4 * o + 2]));
            projection[s].push_back(cv::Point2f(proj[2 * o], proj[2 * o + 1]));
        }
        scans.push_back(c);
    }
    std::vector<cv::Point2f> keypoints;
    for (int k = 0; k < F; k++) keypoints.push_back(cv::Point2f(kp[2 * k], kp[2 * k + 1]));
    Cloud::Ptr kwd(new Cloud); std::vector<int> has_depth;
    featureDepthAssociation(scans, projection, keypoints, kwd, has_depth);
    for (int k = 0; k < F; k++) has_depth_out[k] = has_depth[k];
    for (size_t i = 0; i < kwd->size(); i++) { kpwd[4 * i] = kwd->at(i).x; kpwd[4 * i + 1] = kwd->at(i).y; kpwd[4 * i + 2] = kwd->at(i).z; kpwd[4 * i + 3] = 1.0f; }
    return (int)kwd->size();
}

int ref_transform_points(const float *xyz1, int n, const double pose[6], float *out) {
    for (int i = 0; i < n; i++) {
        pcl::PointXYZ p(xyz1[4 * i], xyz1[4 * i + 1], xyz1[4 * i + 2]);
        util::transform_point(p, pose);
        out[4 * i] = p.x; out[4 * i + 1] = p.y; out[4 * i + 2] = p.z; out[4 * i + 3] = 1.0f;
    }
    return 0;
}

/* functor evaluation through the reference's own templates: type VELO_RES_*, k = constructor doubles */
int ref_eval_functor(int type, const double *k, const double pose[6], double *r, double *J) {
    std::unique_ptr<ceres::CostFunction> c;
    switch (type) {
    case VELO_RES_3DPD: c.reset(new ceres::AutoDiffCostFunction<cost3DPD, 1, 6>(new cost3DPD(k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7], k[8]))); break;
    case VELO_RES_3D3D: c.reset(new ceres::AutoDiffCostFunction<cost3D3D, 3, 6>(new cost3D3D(k[0], k[1], k[2], k[3], k[4], k[5]))); break;
    case VELO_RES_3D2D: c.reset(new ceres::AutoDiffCostFunction<cost3D2D, 2, 6>(new cost3D2D(k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7]))); break;
    case VELO_RES_2D3D: c.reset(new ceres::AutoDiffCostFunction<cost2D3D, 2, 6>(new cost2D3D(k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7]))); break;
    case VELO_RES_2D2D: c.reset(new ceres::AutoDiffCostFunction<cost2D2D, 1, 6>(new cost2D2D(k[0], k[1], k[2], k[3], k[4], k[5], k[6]))); break;
    default: return -1;
    }
    c->Evaluate6(pose, r, J);
    return c->num_residuals();
}
/* plain-double evaluation of the functor (the `(*cost)(transform, residual_test)` calls of velo.h:672 etc.) */
int ref_eval_functor_plain(int type, const double *k, const double pose[6], double *r) {
    switch (type) {
    case VELO_RES_3DPD: { cost3DPD f(k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7], k[8]); f(pose, r); return 1; }
    case VELO_RES_3D3D: { cost3D3D f(k[0], k[1], k[2], k[3], k[4], k[5]); f(pose, r); return 3; }
    case VELO_RES_3D2D: { cost3D2D f(k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7]); f(pose, r); return 2; }
    case VELO_RES_2D3D: { cost2D3D f(k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7]); f(pose, r); return 2; }
    case VELO_RES_2D2D: { cost2D2D f(k[0], k[1], k[2], k[3], k[4], k[5], k[6]); f(pose, r); return 1; }
    }
    return -1;
}

/* velo.h:806-894 at a supplied pose; only queries that reach AddResidualBlock are recorded (kept == 1) */
int ref_icp_pass(const float *ptsM, const int *rsM, int nrM, const float *ptsS, const int *rsS, int nrS,
                 const double pose[6], int iter_in, int skip, velo_icp_corr *corr, int cap, double *neq) {
    std::vector<Cloud::Ptr> scans_M = to_rings(ptsM, rsM, nrM), scans_S = to_rings(ptsS, rsS, nrS);
    std::vector<pcl::KdTreeFLANN<pcl::PointXYZ>> kd_trees(scans_S.size());
    for (size_t i = 0; i < scans_S.size(); i++) kd_trees[i].setInputCloud(scans_S[i]);   /* lru.h:17-20 */
    double transform[6]; memcpy(transform, pose, sizeof(transform));
    const int iter = iter_in; const bool enable_icp = true;
    g_icp_skip = skip;
    ceres::Problem problem;
    std::vector<ceres::ResidualBlockId> icp_blocks;
    std::vector<velo_icp_corr> rec;
#define icp_skip g_icp_skip
#include "ref_velo_icp_a.inc"
                    { /* injected: remember the indices the reference just chose */
                        velo_icp_corr r; memset(&r, 0, sizeof(r));
                        r.src_ring = sm; r.src_idx = smi; r.np_s_i = np_s_i; r.np_i = np_i; r.np_s_j = np_s_j; r.np_j = np_j; r.np_k = np_k; r.kept = 1;
                        for (int q = 0; q < 3; q++) { r.normal[q] = N[q]; r.v0[q] = v0[q]; }
                        rec.push_back(r);
                    }
#include "ref_velo_icp_b.inc"
#undef icp_skip
    if (neq) memset(neq, 0, sizeof(double) * VELO_NEQ_STRIDE);
    for (size_t b = 0; b < problem.blocks.size(); b++) {
        double r[1], J[6], rho[3];
        problem.blocks[b].cost->Evaluate6(transform, r, J);
        problem.blocks[b].loss->Evaluate(r[0] * r[0], rho);
        rec[b].residual = r[0]; memcpy(rec[b].jacobian, J, sizeof(J));
        if (neq) neq_add(neq, r, J, 1, rho);
    }
    for (size_t i = 0; i < rec.size() && (int)i < cap; i++) corr[i] = rec[i];
    return (int)rec.size();
}

/* velo.h:622-792 for one `iter` at a supplied pose.  Arrays per camera with strides F / MM (see oracle_visual). */
int ref_visual(int ncam, int F, int MM, const float *kp1, const float *kp2, const int *hd1, const int *hd2,
               const float *kpwd1, const float *kpwd2, const int *n_matches, const int *matches_in,
               const int *lm_valid, const float *lm_xyz, const velo_gpu_calib *cal, const double pose[6], int iter_in,
               velo_vis_block *blocks, int cap, double *neq) {
    set_calib(cal);
    g_num_cams = ncam;
    const int frame1 = 1, frame2 = 0, iter = iter_in;
    std::vector<std::vector<std::pair<int, int>>> matches(ncam), good_matches(ncam);
    std::vector<std::vector<ResidualType>> residual_type(ncam);
    std::vector<std::vector<std::vector<cv::Point2f>>> keypoints(ncam, std::vector<std::vector<cv::Point2f>>(2));
    std::vector<std::vector<std::vector<int>>> keypoint_ids(ncam, std::vector<std::vector<int>>(2)), has_depth(ncam, std::vector<std::vector<int>>(2));
    std::vector<std::vector<Cloud::Ptr>> keypoints_with_depth(ncam, std::vector<Cloud::Ptr>(2));
    std::map<int, pcl::PointXYZ> landmarks_at_frame;
    for (int c = 0; c < ncam; c++) {
        for (int fr = 0; fr < 2; fr++) {
            const float *kp = fr == 1 ? kp1 + 2 * (size_t)c * F : kp2 + 2 * (size_t)c * F;
            const int *hd = fr == 1 ? hd1 + (size_t)c * F : hd2 + (size_t)c * F;
            const float *kw = fr == 1 ? kpwd1 + 4 * (size_t)c * F : kpwd2 + 4 * (size_t)c * F;
            keypoints_with_depth[c][fr].reset(new Cloud);
            int nh = 0;
            for (int i = 0; i < F; i++) {
                keypoints[c][fr].push_back(cv::Point2f(kp[2 * i], kp[2 * i + 1]));
                keypoint_ids[c][fr].push_back(c * 10000000 + i);
                has_depth[c][fr].push_back(hd[i]);
                if (hd[i] >= 0) nh = std::max(nh, hd[i] + 1);
            }
            for (int i = 0; i < nh; i++) keypoints_with_depth[c][fr]->push_back(pcl::PointXYZ(kw[4 * i], kw[4 * i + 1], kw[4 * i + 2]));
        }
        for (int i = 0; i < n_matches[c]; i++) {
            int p1 = matches_in[2 * ((size_t)c * MM + i)], p2 = matches_in[2 * ((size_t)c * MM + i) + 1];
            matches[c].push_back(std::make_pair(p1, p2));
            if (lm_valid && lm_valid[(size_t)c * MM + i]) {
                const float *l = lm_xyz + 4 * ((size_t)c * MM + i);
                landmarks_at_frame[c * 10000000 + p2] = pcl::PointXYZ(l[0], l[1], l[2]);
            }
        }
    }
    double transform[6]; memcpy(transform, pose, sizeof(transform));
    ceres::Problem problem;
#define num_cams g_num_cams
#include "ref_velo_visual.inc"
#undef num_cams
    if (neq) memset(neq, 0, sizeof(double) * VELO_NEQ_STRIDE);
    size_t b = 0;
    for (int c = 0; c < ncam; c++) {
        for (size_t k = 0; k < good_matches[c].size(); k++, b++) {
            double r[3], J[18], rho[3];
            int nr = problem.blocks[b].cost->num_residuals();
            problem.blocks[b].cost->Evaluate6(transform, r, J);
            double s = 0; for (int i = 0; i < nr; i++) s += r[i] * r[i];
            problem.blocks[b].loss->Evaluate(s, rho);
            if (neq) neq_add(neq, r, J, nr, rho);
            if ((int)b < cap) {
                velo_vis_block &o = blocks[b]; memset(&o, 0, sizeof(o));
                o.cam = c; o.type = (int)residual_type[c][k]; o.n_res = nr;
                o.match = -1;
                for (size_t m = 0; m < matches[c].size(); m++) if (matches[c][m] == good_matches[c][k]) { o.match = (int)m; break; }
                memcpy(o.residual, r, sizeof(double) * nr); memcpy(o.jacobian, J, sizeof(double) * 6 * nr);
            }
        }
    }
    return (int)b;
}

/* matchFeatures (velo.h:499-550) on two descriptor matrices; pairs capacity nq x 2 */
int ref_match_hamming(const unsigned char *q, int nq, const unsigned char *t, int nt, int bytes, int *pairs) {
    std::vector<std::vector<cv::Mat>> descriptors(1, std::vector<cv::Mat>(2));
    descriptors[0][0].rows = nq; descriptors[0][0].cols = bytes; descriptors[0][0].data.assign(q, q + (size_t)nq * bytes);
    descriptors[0][1].rows = nt; descriptors[0][1].cols = bytes; descriptors[0][1].data.assign(t, t + (size_t)nt * bytes);
    std::vector<std::pair<int, int>> matches;
    matchFeatures(descriptors, 0, 0, 0, 1, matches);
    for (size_t i = 0; i < matches.size(); i++) { pairs[2 * i] = matches[i].first; pairs[2 * i + 1] = matches[i].second; }
    return (int)matches.size();
}

/* triangulatePoint (velo.h:1027-1130) per landmark, CSR observations as in velo_gpu_triangulate */
int ref_triangulate(int n, const int *off3, const velo_tri_obs3 *obs3, const int *off2, const velo_tri_obs2 *obs2, const double *poses, int n_frames,
                    const velo_gpu_calib *cal, int ncam, const float *init_xyz, const int *has_init, float *out_xyz) {
    set_calib(cal);
    g_num_cams = ncam;
    std::vector<double[6]> camera_poses(n_frames);
    for (int f = 0; f < n_frames; f++) for (int i = 0; i < 6; i++) camera_poses[f][i] = poses[6 * f + i];
    for (int l = 0; l < n; l++) {
        /* the reference keys 3-D observations by camera too; the functor ignores the camera, so camera 0 carries them all in order */
        std::vector<std::map<int, cv::Point2f>> keypoint_obs2(ncam);
        std::vector<std::map<int, pcl::PointXYZ>> keypoint_obs3(ncam);
        for (int k = off3[l]; k < off3[l + 1]; k++) keypoint_obs3[0][obs3[k].frame] = pcl::PointXYZ(obs3[k].x, obs3[k].y, obs3[k].z);
        for (int k = off2[l]; k < off2[l + 1]; k++) keypoint_obs2[obs2[k].cam][obs2[k].frame] = cv::Point2f(obs2[k].x, obs2[k].y);
        pcl::PointXYZ point;
        const bool ig = has_init && has_init[l];
        if (ig) point = pcl::PointXYZ(init_xyz[3 * l], init_xyz[3 * l + 1], init_xyz[3 * l + 2]);
        triangulatePoint(keypoint_obs2, keypoint_obs3, camera_poses, point, ig);
        out_xyz[3 * l] = point.x; out_xyz[3 * l + 1] = point.y; out_xyz[3 * l + 2] = point.z;
    }
    return 0;
}

/* tunables as compiled from kitti.h:3-35, so tests can assert the defaults of velo_gpu_default_params */
int ref_constants(double *out) {
    double v[] = { (double)num_cams_actual, (double)corner_count, (double)icp_skip, (double)f2f_iterations, (double)icp_iterations,
                   weight_3D2D, weight_2D2D, weight_3DPD, loss_thresh_3D2D, loss_thresh_2D2D, loss_thresh_3DPD, loss_thresh_3D3D,
                   depth_assoc_thresh, outlier_reject, correspondence_thresh_icp, icp_norm_condition };
    for (size_t i = 0; i < sizeof(v) / sizeof(v[0]); i++) out[i] = v[i];
    return (int)(sizeof(v) / sizeof(v[0]));
}

} // extern "C"
