// velo_dropin.hpp — the reference's C++ call signatures (velo.h / kitti.h / lru.h) on top of the C ABI (velo_gpu.h).
//
// A maintainer of lichunshang/vision-enhanced-lidar-odometry includes this header INSTEAD of the function bodies it
// replaces and links libvelo_gpu.so; main.cpp's call sites (main.cpp:216,256,261,388-405,596,600) compile unchanged:
//
//   reference                                             here
//   ---------------------------------------------------   ------------------------------------------------------------
//   loadCalibration(dataset)              kitti.h:59      velo_dropin::loadCalibrationFromArrays(P, Tr, w, h)  (the file
//                                                         parsing stays on the host; only the math moved to the library)
//   ScanData(dataset, frame)              lru.h:12        velo_dropin::ScanData(xyzr, n, frame)  (same members: scans, _frame;
//                                                         `trees` became the device-resident index of a slot)
//   segmentPoints(cloud, scans)           kitti.h:154     inside ScanData (velo_gpu_scan_upload + download)
//   projectLidarToCamera(...)             velo.h:329      same name, same parameters
//   featureDepthAssociation(...)          velo.h:377      same name, same parameters
//   frameToFrame ICP block                velo.h:806-874  velo_dropin::icpCorrespondences(...) -> the (p, N, v0) triples that
//                                                         velo.h:875-891 wraps in cost3DPD blocks, or icpNormalEquations(...)
//
// PCL / OpenCV are used when their headers are available; otherwise layout-identical stand-ins are defined
// (pcl::PointXYZ = 16-byte {x,y,z,pad=1}, cv::Point2f = {x,y}, PointCloud::points contiguous), which is also how this
// header is compiled in this repository's tests (no PCL/OpenCV in the image).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "velo_gpu.h"

#if defined(__has_include)
#if __has_include(<pcl/point_types.h>) && __has_include(<pcl/point_cloud.h>)
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#define VELO_DROPIN_HAVE_PCL 1
#endif
#if __has_include(<opencv2/core/types.hpp>)
#include <opencv2/core/types.hpp>
#define VELO_DROPIN_HAVE_OPENCV 1
#endif
#endif

#ifndef VELO_DROPIN_HAVE_PCL
namespace pcl {
struct alignas(16) PointXYZ {
    float x, y, z, pad;
    PointXYZ() : x(0), y(0), z(0), pad(1.0f) {}
    PointXYZ(float a, float b, float c) : x(a), y(b), z(c), pad(1.0f) {}
};
template <class P> struct PointCloud {
    typedef std::shared_ptr<PointCloud<P>> Ptr;
    std::vector<P> points;
    size_t size() const { return points.size(); }
    const P &at(size_t i) const { return points.at(i); }
    P &at(size_t i) { return points.at(i); }
    void push_back(const P &p) { points.push_back(p); }
};
} // namespace pcl
#endif
#ifndef VELO_DROPIN_HAVE_OPENCV
namespace cv {
struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float a, float b) : x(a), y(b) {} };
} // namespace cv
#endif

static_assert(sizeof(pcl::PointXYZ) == 16, "pcl::PointXYZ must be a 16-byte float4 (lru.h:5)");
static_assert(sizeof(cv::Point2f) == 8, "cv::Point2f must be two floats");

namespace velo_dropin {

typedef pcl::PointCloud<pcl::PointXYZ> Cloud;

// ---- process-wide runtime: one context (the reference is single-threaded with global state, kitti.h:37-57)
struct Runtime {
    velo_gpu_ctx *ctx = nullptr;
    velo_gpu_params prm;
    velo_gpu_calib cal;
    int next_slot = 0;
    std::map<const void *, int> slot_of_scans;                       // &scans[0]->points[0]  -> slot
    std::map<const void *, std::pair<int, int>> proj_of_vector;      // &projection[0]       -> (slot, cam)
    static Runtime &get() { static Runtime r; return r; }
    void check(int rc, const char *what) {
        if (rc != VELO_OK) throw std::runtime_error(std::string(what) + ": " + velo_gpu_last_error(ctx));
    }
};

// kitti.h:59-108 with the calib.txt numbers already parsed (P = 4 x 3x4 row-major, Tr = 3x4 row-major).
inline void loadCalibrationFromArrays(const float P[48], const float Tr[12], int img_width, int img_height,
                                      const velo_gpu_params *params = nullptr, int device = 0) {
    Runtime &r = Runtime::get();
    if (r.ctx) { velo_gpu_destroy(r.ctx); r.ctx = nullptr; }
    if (params) r.prm = *params; else { velo_gpu_default_params(&r.prm); r.prm.max_slots = 8; }
    if (velo_gpu_calib_from_kitti(P, Tr, img_width, img_height, &r.cal) != VELO_OK) throw std::runtime_error("bad calibration");
    if (velo_gpu_create(device, &r.prm, &r.cal, &r.ctx) != VELO_OK) throw std::runtime_error(std::string("velo_gpu_create: ") + velo_gpu_last_error(nullptr));
    r.next_slot = 0; r.slot_of_scans.clear(); r.proj_of_vector.clear();
}

// lru.h:7-28.  `scans` is filled exactly as segmentPoints (kitti.h:154-185) fills it; `trees` is the slot's device index.
struct ScanData {
    std::vector<Cloud::Ptr> scans;
    int slot = -1;        // replaces std::vector<pcl::KdTreeFLANN<pcl::PointXYZ>> trees
    int _frame = -1;
    ScanData() {}
    // xyzr: the KITTI .bin contents (n x {x,y,z,reflectance}) that loadPoints (kitti.h:121-152) reads
    ScanData(const float *xyzr, int n, int frame) {
        Runtime &r = Runtime::get();
        slot = r.next_slot; r.next_slot = (r.next_slot + 1) % r.prm.max_slots;   // ring buffer of device scans (ScansLRU analogue)
        r.check(velo_gpu_scan_upload(r.ctx, slot, xyzr, n), "scan_upload");
        int np = 0, nr = 0;
        r.check(velo_gpu_scan_info(r.ctx, slot, &np, &nr), "scan_info");
        std::vector<pcl::PointXYZ> flat(np > 0 ? np : 1);
        std::vector<int> rs(nr + 1, 0);
        r.check(velo_gpu_scan_download(r.ctx, slot, reinterpret_cast<float *>(flat.data()), rs.data()), "scan_download");
        for (int s = 0; s < nr; s++) {
            Cloud::Ptr c(new Cloud);
            c->points.assign(flat.begin() + rs[s], flat.begin() + rs[s + 1]);
            scans.push_back(c);
        }
        _frame = frame;
        for (auto it = r.slot_of_scans.begin(); it != r.slot_of_scans.end();) it = (it->second == slot) ? r.slot_of_scans.erase(it) : ++it;
        if (!scans.empty()) r.slot_of_scans[scans[0].get()] = slot;
    }
    // loadPoints (kitti.h:121-152): read a KITTI velodyne .bin
    static ScanData fromFile(const std::string &path, int frame) {
        FILE *f = fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("cannot open " + path);
        std::vector<float> buf;
        float tmp[4096];
        size_t got;
        while ((got = fread(tmp, sizeof(float), 4096, f)) > 0) buf.insert(buf.end(), tmp, tmp + got);
        fclose(f);
        return ScanData(buf.data(), (int)(buf.size() / 4), frame);
    }
};

inline int slotOf(const std::vector<Cloud::Ptr> &scans, bool upload_if_unknown = true) {
    Runtime &r = Runtime::get();
    if (!scans.empty()) { auto it = r.slot_of_scans.find(scans[0].get()); if (it != r.slot_of_scans.end()) return it->second; }
    if (!upload_if_unknown) return -1;
    // scans that did not come from ScanData: flatten and install them (no re-segmentation)
    std::vector<pcl::PointXYZ> flat; std::vector<int> rs(1, 0);
    for (auto &c : scans) { flat.insert(flat.end(), c->points.begin(), c->points.end()); rs.push_back((int)flat.size()); }
    int slot = r.next_slot; r.next_slot = (r.next_slot + 1) % r.prm.max_slots;
    r.check(velo_gpu_scan_upload_rings(r.ctx, slot, reinterpret_cast<const float *>(flat.data()), rs.data(), (int)scans.size()), "scan_upload_rings");
    for (auto it = r.slot_of_scans.begin(); it != r.slot_of_scans.end();) it = (it->second == slot) ? r.slot_of_scans.erase(it) : ++it;
    if (!scans.empty()) r.slot_of_scans[scans[0].get()] = slot;
    return slot;
}

} // namespace velo_dropin

// ------------------------------------------------------------------ velo.h:329-334, same signature
inline void projectLidarToCamera(const std::vector<velo_dropin::Cloud::Ptr> &scans,
                                 std::vector<std::vector<cv::Point2f>> &projection,
                                 std::vector<velo_dropin::Cloud::Ptr> &scans_valid,
                                 const int cam) {
    using namespace velo_dropin;
    Runtime &r = Runtime::get();
    const int slot = slotOf(scans);
    r.check(velo_gpu_project(r.ctx, slot, cam), "project");
    int np = 0, nr = 0, total = 0;
    r.check(velo_gpu_scan_info(r.ctx, slot, &np, &nr), "scan_info");
    std::vector<int> rc(nr > 0 ? nr : 1);
    std::vector<cv::Point2f> p(np > 0 ? np : 1);
    std::vector<pcl::PointXYZ> v(np > 0 ? np : 1);
    r.check(velo_gpu_project_download(r.ctx, slot, cam, rc.data(), reinterpret_cast<float *>(p.data()), reinterpret_cast<float *>(v.data()), &total), "project_download");
    int o = 0;
    const size_t first = projection.size();
    for (int s = 0; s < nr; s++) {                                   // outputs are APPENDED (velo.h:341-343)
        projection.push_back(std::vector<cv::Point2f>(p.begin() + o, p.begin() + o + rc[s]));
        Cloud::Ptr c(new Cloud);
        c->points.assign(v.begin() + o, v.begin() + o + rc[s]);
        scans_valid.push_back(c);
        o += rc[s];
    }
    if (nr > 0) {
        r.proj_of_vector[&projection[first]] = std::make_pair(slot, cam);
        r.slot_of_scans[scans_valid[scans_valid.size() - nr].get()] = slot;   // featureDepthAssociation receives scans_valid
    }
}

// ------------------------------------------------------------------ velo.h:377-383, same signature
inline std::vector<int> featureDepthAssociation(const std::vector<velo_dropin::Cloud::Ptr> &scans,
                                                const std::vector<std::vector<cv::Point2f>> &projection,
                                                const std::vector<cv::Point2f> &keypoints,
                                                velo_dropin::Cloud::Ptr keypoints_with_depth,
                                                std::vector<int> &has_depth) {
    using namespace velo_dropin;
    Runtime &r = Runtime::get();
    int slot = -1, cam = 0;
    auto it = projection.empty() ? r.proj_of_vector.end() : r.proj_of_vector.find(&projection[0]);
    if (it != r.proj_of_vector.end()) { slot = it->second.first; cam = it->second.second; }
    else {
        // projection not produced by this library (or copied): install it on a slot holding `scans`
        slot = slotOf(scans);
        std::vector<int> rc; std::vector<cv::Point2f> p; std::vector<pcl::PointXYZ> v;
        for (size_t s = 0; s < projection.size(); s++) {
            rc.push_back((int)projection[s].size());
            p.insert(p.end(), projection[s].begin(), projection[s].end());
            v.insert(v.end(), scans[s]->points.begin(), scans[s]->points.begin() + projection[s].size());
        }
        r.check(velo_gpu_projection_upload(r.ctx, slot, 0, rc.data(), reinterpret_cast<const float *>(p.data()), reinterpret_cast<const float *>(v.data())), "projection_upload");
    }
    const int F = (int)keypoints.size();
    has_depth.assign(F, -1);
    std::vector<pcl::PointXYZ> kw(F > 0 ? F : 1);
    int nh = 0;
    r.check(velo_gpu_depth_assoc(r.ctx, slot, cam, 0, reinterpret_cast<const float *>(keypoints.data()), F, has_depth.data(),
                                 reinterpret_cast<float *>(kw.data()), &nh), "depth_assoc");
    for (int i = 0; i < nh; i++) keypoints_with_depth->push_back(kw[i]);
    return has_depth;
}

namespace velo_dropin {

// ------------------------------------------------------------------ velo.h:806-874 in one call.
// Returns the records of every query; record.kept == 1 are the ones velo.h:875-891 turns into cost3DPD blocks:
//     new cost3DPD(p.x, p.y, p.z, rec.normal[0..2], rec.v0[0..2])   with p = scans_M[rec.src_ring]->at(rec.src_idx)
inline std::vector<velo_icp_corr> icpCorrespondences(const std::vector<Cloud::Ptr> &scans_M, const std::vector<Cloud::Ptr> &scans_S,
                                                     const double transform[6], int iter, int icp_skip, double *neq = nullptr) {
    Runtime &r = Runtime::get();
    const int sm = slotOf(scans_M), ss = slotOf(scans_S);
    size_t cap = 0; for (auto &c : scans_M) cap += (c->size() + icp_skip - 1) / icp_skip;
    std::vector<velo_icp_corr> rec(cap > 0 ? cap : 1);
    int nq = 0, nk = 0;
    double tmp[VELO_NEQ_STRIDE];
    r.check(velo_gpu_icp_pass(r.ctx, sm, ss, transform, iter, icp_skip, rec.data(), (int)rec.size(), &nq, &nk, neq ? neq : tmp), "icp_pass");
    rec.resize(nq);
    return rec;
}

// Normal equations of the ICP term at `transform` (what ceres::Solve would form, velo.h:885-902): H (21 upper), g (6), cost.
inline void icpNormalEquations(const std::vector<Cloud::Ptr> &scans_M, const std::vector<Cloud::Ptr> &scans_S,
                               const double transform[6], int iter, int icp_skip, double neq[VELO_NEQ_STRIDE]) {
    Runtime &r = Runtime::get();
    int nq = 0, nk = 0;
    r.check(velo_gpu_icp_pass(r.ctx, slotOf(scans_M), slotOf(scans_S), transform, iter, icp_skip, nullptr, 0, &nq, &nk, neq), "icp_pass");
}

} // namespace velo_dropin
