// velo_dropin.hpp — the reference's C++ call signatures (velo.h / kitti.h / lru.h) on top of the C ABI (velo_gpu.h).
//
// A maintainer of lichunshang/vision-enhanced-lidar-odometry includes this header INSTEAD of the function bodies it
// replaces and links libvelo_gpu.so; main.cpp's call sites (main.cpp:216,256,261,388-405,596,600) compile unchanged:
//
//   reference                                             here
//   ---------------------------------------------------   ------------------------------------------------------------
//   loadCalibration(dataset)              kitti.h:59      velo_dropin::loadCalibration(dataset) (calib.txt through velo_kitti_load_calib)
//                                                         or loadCalibrationFromArrays(P, Tr, w, h)
//   ScanData(dataset, frame)              lru.h:12        same constructor (kittipath + dataset + "/velodyne/%06d.bin", kitti.h:57,125-127);
//                                                         same members: scans, _frame; `trees` became the device-resident index of a slot
//   ScansLRU::get(dataset, frame)         lru.h:31-61     same class, same policy (least recently used frame is deleted); a deleted
//                                                         ScanData gives its device slot back
//   segmentPoints(cloud, scans)           kitti.h:154     inside ScanData (velo_gpu_scan_upload + download)
//   projectLidarToCamera(...)             velo.h:329      same name, same parameters
//   featureDepthAssociation(...)          velo.h:377      same name, same parameters
//   frameToFrame(...)                     velo.h:598-614  same name, same parameter list (kd_trees is accepted and ignored); fills
//                                                         good_matches / residual_type like velo.h:624-625,690-692; the per-pass
//                                                         ceres::Solve is the device-resident LM solve of velo_gpu_frame_to_frame
//   frameToFrame ICP block only           velo.h:806-874  velo_dropin::icpCorrespondences(...) -> the (p, N, v0) triples that
//                                                         velo.h:875-891 wraps in cost3DPD blocks, or icpNormalEquations(...)
//
// How host containers find their device copies: a scan lives in a device SLOT.  ScanData owns its slot for its lifetime.  Any other
// cloud vector handed to these functions is looked up by the address of its first ring AND a fingerprint of its content (ring
// count, point count, first and last point); a stale or recycled address therefore never resolves to another scan's slot — the
// cloud is simply uploaded again into a scratch slot (least recently used scratch slot is recycled; a slot's projection records
// die with it).  Association results are never looked up by address: frameToFrame uploads the containers it is given.
//
// PCL / OpenCV / Eigen are used when their headers are available; otherwise layout-identical stand-ins are defined
// (pcl::PointXYZ = 16-byte {x,y,z,pad=1}, cv::Point2f = {x,y}, PointCloud::points contiguous, a column-major 4x4 double matrix),
// which is also how this header is compiled in this repository's tests (no PCL/OpenCV/Eigen in the image).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
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
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
#define VELO_DROPIN_HAVE_EIGEN 1
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
#ifdef VELO_DROPIN_HAVE_EIGEN
typedef Eigen::Matrix4d Matrix4d;
#else
struct Matrix4d {                                   // column-major like Eigen::Matrix4d
    double m[16];
    double &operator()(int i, int j) { return m[4 * j + i]; }
    double operator()(int i, int j) const { return m[4 * j + i]; }
};
#endif

// kitti.h:57: the dataset root; settable (the reference hard-codes the author's home directory)
inline std::string &kittipath() { static std::string p = "/home/dllu/kitti/dataset/sequences/"; return p; }

// what identifies the content of a vector of ring clouds / of a projection without reading all of it
struct Fingerprint {
    size_t rings = 0, points = 0;
    float first[3] = { 0, 0, 0 }, last[3] = { 0, 0, 0 };
    bool operator==(const Fingerprint &o) const { return rings == o.rings && points == o.points && !memcmp(first, o.first, sizeof(first)) && !memcmp(last, o.last, sizeof(last)); }
};
inline Fingerprint fingerprint(const std::vector<Cloud::Ptr> &scans) {
    Fingerprint f; f.rings = scans.size();
    bool have = false;
    for (auto &c : scans) {
        f.points += c->points.size();
        if (c->points.empty()) continue;
        if (!have) { const pcl::PointXYZ &p = c->points.front(); f.first[0] = p.x; f.first[1] = p.y; f.first[2] = p.z; have = true; }
        const pcl::PointXYZ &q = c->points.back(); f.last[0] = q.x; f.last[1] = q.y; f.last[2] = q.z;
    }
    return f;
}
inline Fingerprint fingerprint(const std::vector<std::vector<cv::Point2f>> &proj, size_t first_ring, size_t n_rings) {
    Fingerprint f; f.rings = n_rings;
    bool have = false;
    for (size_t s = first_ring; s < first_ring + n_rings && s < proj.size(); s++) {
        f.points += proj[s].size();
        if (proj[s].empty()) continue;
        if (!have) { f.first[0] = proj[s].front().x; f.first[1] = proj[s].front().y; have = true; }
        f.last[0] = proj[s].back().x; f.last[1] = proj[s].back().y;
    }
    return f;
}

// ---- process-wide runtime: one context (the reference is single-threaded with global state, kitti.h:37-57)
struct Runtime {
    struct Slot { bool used = false, owned = false; uint64_t gen = 0, stamp = 0; const void *key = nullptr; Fingerprint fp; };
    struct ProjRec { int slot, cam; uint64_t gen; Fingerprint fp; };
    velo_gpu_ctx *ctx = nullptr;
    velo_gpu_params prm;
    velo_gpu_calib cal;
    std::vector<Slot> slots;
    uint64_t clock = 0;
    std::map<const void *, int> slot_of_scans;            // address of the first ring cloud -> slot (verified by fingerprint)
    std::map<const void *, ProjRec> proj_of_vector;       // &projection[first]             -> (slot, cam) (verified by generation + fingerprint)
    static Runtime &get() { static Runtime r; return r; }
    void check(int rc, const char *what) {
        if (rc != VELO_OK) throw std::runtime_error(std::string(what) + ": " + velo_gpu_last_error(ctx));
    }
    void need_ctx() const { if (!ctx) throw std::runtime_error("velo_dropin: loadCalibration has not been called"); }
    // a slot stops describing its scan: forget every host container that pointed at it
    void retire(int slot) {
        for (auto it = slot_of_scans.begin(); it != slot_of_scans.end();) it = (it->second == slot) ? slot_of_scans.erase(it) : ++it;
        for (auto it = proj_of_vector.begin(); it != proj_of_vector.end();) it = (it->second.slot == slot) ? proj_of_vector.erase(it) : ++it;
        slots[slot].gen++; slots[slot].key = nullptr; slots[slot].fp = Fingerprint();
    }
    // a free slot, else the least recently used SCRATCH slot (slots owned by a live ScanData are never taken away)
    int acquire(bool owned) {
        need_ctx();
        int pick = -1;
        for (size_t i = 0; i < slots.size(); i++) if (!slots[i].used) { pick = (int)i; break; }
        if (pick < 0) for (size_t i = 0; i < slots.size(); i++) if (!slots[i].owned && (pick < 0 || slots[i].stamp < slots[pick].stamp)) pick = (int)i;
        if (pick < 0) throw std::runtime_error("velo_dropin: every device slot is owned by a live ScanData (raise velo_gpu_params.max_slots)");
        retire(pick);
        slots[pick].used = true; slots[pick].owned = owned; slots[pick].stamp = ++clock;
        return pick;
    }
    void release(int slot) { if (slot >= 0 && slot < (int)slots.size()) { retire(slot); slots[slot].used = false; slots[slot].owned = false; } }
    void bind(int slot, const std::vector<Cloud::Ptr> &scans) {
        if (scans.empty()) return;
        slots[slot].key = scans[0].get(); slots[slot].fp = fingerprint(scans);
        slot_of_scans[scans[0].get()] = slot;
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
    r.slots.assign(r.prm.max_slots, Runtime::Slot()); r.clock = 0; r.slot_of_scans.clear(); r.proj_of_vector.clear();
}
// loadCalibration(dataset) (kitti.h:59-108): kittipath + dataset + "/calib.txt"; the image size is what loadImage left in
// img_width / img_height (kitti.h:196-197, main.cpp:72-73)
inline void loadCalibration(const std::string &dataset, int img_width, int img_height, const velo_gpu_params *params = nullptr, int device = 0) {
    float P[48], Tr[12];
    const std::string path = kittipath() + dataset + "/calib.txt";
    if (velo_kitti_load_calib(path.c_str(), P, Tr) != VELO_OK) throw std::runtime_error("cannot parse " + path);
    loadCalibrationFromArrays(P, Tr, img_width, img_height, params, device);
}

// lru.h:7-28.  `scans` is filled exactly as segmentPoints (kitti.h:154-185) fills it; `trees` is the slot's device index.
struct ScanData {
    std::vector<Cloud::Ptr> scans;
    int slot = -1;        // replaces std::vector<pcl::KdTreeFLANN<pcl::PointXYZ>> trees
    int _frame = -1;
    ScanData() {}
    // lru.h:12: loadPoints (kitti.h:121-152) + segmentPoints + index build
    ScanData(const std::string dataset, const int frame) {
        char name[32];
        snprintf(name, sizeof(name), "%06d.bin", frame);                          // kitti.h:125-127
        const std::string path = kittipath() + dataset + "/velodyne/" + name;
        Runtime &r = Runtime::get();
        r.need_ctx();
        std::vector<float> buf((size_t)r.prm.max_points * 4);
        int n = 0;
        const int rc = velo_kitti_load_scan(path.c_str(), buf.data(), r.prm.max_points, &n);
        if (rc == VELO_ERR_CAPACITY) throw std::runtime_error(path + ": more points than velo_gpu_params.max_points");
        if (rc != VELO_OK) throw std::runtime_error("cannot read " + path);
        init(buf.data(), n, frame);
    }
    // xyzr: the KITTI .bin contents (n x {x,y,z,reflectance}) that loadPoints (kitti.h:121-152) reads
    ScanData(const float *xyzr, int n, int frame) { init(xyzr, n, frame); }
    ~ScanData() { if (slot >= 0 && Runtime::get().ctx) Runtime::get().release(slot); }
    ScanData(const ScanData &) = delete;
    ScanData &operator=(const ScanData &) = delete;
    ScanData(ScanData &&o) noexcept : scans(std::move(o.scans)), slot(o.slot), _frame(o._frame) { o.slot = -1; }
    // loadPoints (kitti.h:121-152) from an explicit path
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

private:
    void init(const float *xyzr, int n, int frame) {
        Runtime &r = Runtime::get();
        slot = r.acquire(true);
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
        r.bind(slot, scans);
    }
};

// lru.h:31-61, same policy: at most `size` scans, the least recently used one is deleted (and its device slot freed)
class ScansLRU {
    int size;
    std::list<ScanData *> times;
    std::unordered_map<int, std::list<ScanData *>::iterator> exists;

public:
    // lru.h:33 keeps 50; the device holds max_slots scans, two of which stay free for clouds that do not come from a ScanData
    explicit ScansLRU(int capacity = 50) : size(capacity) {
        const Runtime &r = Runtime::get();
        if (r.ctx && size > r.prm.max_slots - 2) size = r.prm.max_slots - 2 > 1 ? r.prm.max_slots - 2 : 1;
    }
    ~ScansLRU() { for (ScanData *sd : times) delete sd; }
    ScanData *get(const std::string dataset, const int frame) {
        if (exists.count(frame)) {                                                 // lru.h:42-47
            auto it = exists[frame];
            ScanData *sd = *it;
            times.erase(it);
            times.push_front(sd);
            exists[frame] = times.begin();
            return sd;
        }
        if ((int)times.size() >= size) {                                           // lru.h:52-57 (evict first: its slot is needed for the new scan)
            ScanData *old = times.back();
            exists.erase(old->_frame);
            delete old;
            times.pop_back();
        }
        ScanData *sd = new ScanData(dataset, frame);                               // lru.h:49-51
        times.push_front(sd);
        exists[frame] = times.begin();
        return sd;
    }
    int resident() const { return (int)times.size(); }
};

inline int slotOf(const std::vector<Cloud::Ptr> &scans, bool upload_if_unknown = true) {
    Runtime &r = Runtime::get();
    r.need_ctx();
    if (!scans.empty()) {
        auto it = r.slot_of_scans.find(scans[0].get());
        if (it != r.slot_of_scans.end()) {
            Runtime::Slot &s = r.slots[it->second];
            if (s.used && s.key == scans[0].get() && s.fp == fingerprint(scans)) { s.stamp = ++r.clock; return it->second; }
            r.slot_of_scans.erase(it);          // the address was recycled for another cloud
        }
    }
    if (!upload_if_unknown) return -1;
    // scans that did not come from a live ScanData: flatten and install them in a scratch slot (no re-segmentation)
    std::vector<pcl::PointXYZ> flat; std::vector<int> rs(1, 0);
    for (auto &c : scans) { flat.insert(flat.end(), c->points.begin(), c->points.end()); rs.push_back((int)flat.size()); }
    if (flat.empty()) flat.resize(1);
    const int slot = r.acquire(false);
    r.check(velo_gpu_scan_upload_rings(r.ctx, slot, reinterpret_cast<const float *>(flat.data()), rs.data(), (int)scans.size()), "scan_upload_rings");
    r.bind(slot, scans);
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
    // remember which device projection these host rows mirror (the usual caller passes empty vectors, so first == 0 and the
    // address survives the push_backs above only because it is taken afterwards)
    if (nr > 0) r.proj_of_vector[&projection[first]] = Runtime::ProjRec{ slot, cam, r.slots[slot].gen, fingerprint(projection, first, nr) };
}

// ------------------------------------------------------------------ velo.h:377-383, same signature
inline std::vector<int> featureDepthAssociation(const std::vector<velo_dropin::Cloud::Ptr> &scans,
                                                const std::vector<std::vector<cv::Point2f>> &projection,
                                                const std::vector<cv::Point2f> &keypoints,
                                                velo_dropin::Cloud::Ptr keypoints_with_depth,
                                                std::vector<int> &has_depth) {
    using namespace velo_dropin;
    Runtime &r = Runtime::get();
    r.need_ctx();
    int slot = -1, cam = 0;
    auto it = projection.empty() ? r.proj_of_vector.end() : r.proj_of_vector.find(&projection[0]);
    if (it != r.proj_of_vector.end()) {
        const Runtime::ProjRec &pr = it->second;
        // still the projection this library produced, on a slot that still holds the same scan?
        if (r.slots[pr.slot].used && r.slots[pr.slot].gen == pr.gen && pr.fp == fingerprint(projection, 0, projection.size())) { slot = pr.slot; cam = pr.cam; }
        else r.proj_of_vector.erase(it);
    }
    if (slot < 0) {
        // projection not produced by this library, copied, or its slot was recycled: install it on a scratch slot holding `scans`
        slot = slotOf(scans);
        std::vector<int> rc; std::vector<cv::Point2f> p; std::vector<pcl::PointXYZ> v;
        for (size_t s = 0; s < projection.size(); s++) {
            rc.push_back((int)projection[s].size());
            p.insert(p.end(), projection[s].begin(), projection[s].end());
            v.insert(v.end(), scans[s]->points.begin(), scans[s]->points.begin() + projection[s].size());
        }
        if (rc.empty()) rc.push_back(0);
        cam = 0;
        r.check(velo_gpu_projection_upload(r.ctx, slot, cam, rc.data(), reinterpret_cast<const float *>(p.data()), reinterpret_cast<const float *>(v.data())), "projection_upload");
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

// ------------------------------------------------------------------ velo.h:598-614, same parameter list
// KdTrees: whatever the caller keeps in ScanData::trees (ignored: the slot of scans_S carries the index).  ResidualTypeT: the
// reference's enum ResidualType (velo.h:3-8; same numbering as VELO_RES_*).  The keypoint / depth containers of frame1 and frame2
// are uploaded as given (a few hundred KB); scans are found by slotOf().  icp_skip comes from the runtime's velo_gpu_params.
template <class KdTrees, class ResidualTypeT, class IdT>
inline velo_dropin::Matrix4d frameToFrame(const std::vector<std::vector<std::pair<int, int>>> &matches,
                                          const std::vector<std::vector<std::vector<cv::Point2f>>> &keypoints,
                                          const std::vector<std::vector<std::vector<IdT>>> &keypoint_ids,
                                          const std::map<int, pcl::PointXYZ> &landmarks_at_frame,
                                          const std::vector<std::vector<velo_dropin::Cloud::Ptr>> &keypoints_with_depth,
                                          const std::vector<std::vector<std::vector<int>>> &has_depth,
                                          const std::vector<velo_dropin::Cloud::Ptr> &scans_M,
                                          const std::vector<velo_dropin::Cloud::Ptr> &scans_S,
                                          const KdTrees & /*kd_trees*/,
                                          const int frame1, const int frame2, double transform[6],
                                          std::vector<std::vector<std::pair<int, int>>> &good_matches,
                                          std::vector<std::vector<ResidualTypeT>> &residual_type,
                                          const bool enable_icp,
                                          velo_f2f_report *report = nullptr) {
    using namespace velo_dropin;
    Runtime &r = Runtime::get();
    r.need_ctx();
    const int C = r.prm.num_cams, MM = r.prm.max_matches;
    const int sM = slotOf(scans_M), sS = slotOf(scans_S);
    std::vector<int> nm(C, 0), flat, lmv;
    std::vector<float> lmx;
    for (int cam = 0; cam < C; cam++) {
        // frame1 = current frame: keypoint set 1 of slot M; frame2 = previous frame: set 0 of slot S (as in the batched path)
        const int fr[2] = { frame1, frame2 }, sl[2] = { sM, sS }, st[2] = { 1, 0 };
        for (int k = 0; k < 2; k++) {
            const std::vector<cv::Point2f> &kp = keypoints[cam][fr[k]];
            const std::vector<int> &hd = has_depth[cam][fr[k]];
            const Cloud &kw = *keypoints_with_depth[cam][fr[k]];
            if (hd.size() != kp.size()) throw std::runtime_error("frameToFrame: has_depth and keypoints differ in length");
            r.check(velo_gpu_assoc_upload(r.ctx, sl[k], cam, st[k], reinterpret_cast<const float *>(kp.data()), (int)kp.size(), hd.data(),
                                          reinterpret_cast<const float *>(kw.points.data()), (int)kw.points.size()), "assoc_upload");
        }
        const std::vector<std::pair<int, int>> &mc = matches[cam];
        if ((int)mc.size() > MM) throw std::runtime_error("frameToFrame: more matches than velo_gpu_params.max_matches");
        nm[cam] = (int)mc.size();
        for (size_t i = 0; i < mc.size(); i++) {
            flat.push_back(mc[i].first); flat.push_back(mc[i].second);
            // velo.h:630,634-644: a landmark of the matched keypoint's id overrides the lidar depth of frame2
            bool have = false;
            pcl::PointXYZ lp;
            if (mc[i].second >= 0 && (size_t)mc[i].second < keypoint_ids[cam][frame2].size()) {
                auto it = landmarks_at_frame.find((int)keypoint_ids[cam][frame2][mc[i].second]);
                if (it != landmarks_at_frame.end()) { have = true; lp = it->second; }
            }
            lmv.push_back(have ? 1 : 0);
            lmx.push_back(lp.x); lmx.push_back(lp.y); lmx.push_back(lp.z); lmx.push_back(1.0f);
        }
    }
    if (flat.empty()) { flat.resize(2, 0); lmv.resize(1, 0); lmx.resize(4, 0.f); }
    velo_f2f_report rep;
    r.check(velo_gpu_frame_to_frame(r.ctx, sM, 1, sS, 0, nm.data(), flat.data(), lmv.data(), lmx.data(), enable_icp ? 1 : 0, r.prm.icp_skip, transform, &rep),
            "frame_to_frame");
    if (report) *report = rep;
    // good_matches / residual_type as the last f2f iteration leaves them (velo.h:624-625 clears them per iteration; one entry
    // per residual block in the order 3D3D, 2D2D, 3D2D, 2D3D, velo.h:690-692,719-720,753-754,786-787)
    std::vector<unsigned char> sel((size_t)C * MM);
    r.check(velo_gpu_f2f_selection(r.ctx, sel.data(), (int)sel.size()), "f2f_selection");
    good_matches.resize(C > (int)good_matches.size() ? C : good_matches.size());
    residual_type.resize(C > (int)residual_type.size() ? C : residual_type.size());
    static const int bit_type[4][2] = { { 1, VELO_RES_3D3D }, { 2, VELO_RES_2D2D }, { 4, VELO_RES_3D2D }, { 8, VELO_RES_2D3D } };
    for (int cam = 0; cam < C; cam++) {
        good_matches[cam].clear(); residual_type[cam].clear();
        for (size_t i = 0; i < matches[cam].size(); i++)
            for (int b = 0; b < 4; b++)
                if (sel[(size_t)cam * MM + i] & bit_type[b][0]) {
                    residual_type[cam].push_back((ResidualTypeT)bit_type[b][1]);
                    good_matches[cam].push_back(matches[cam][i]);
                }
    }
    // util::pose_mat2vec(transform) (velo.h:918, utility.h:67-82)
    double T[16];
    velo_pose_vec2mat(transform, T);
    Matrix4d out;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out(i, j) = T[4 * i + j];
    return out;
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
