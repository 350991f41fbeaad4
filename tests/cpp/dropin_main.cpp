// dropin_main.cpp — exercises include/velo_dropin.hpp exactly the way main.cpp uses the reference functions
// (main.cpp:73,216,254-261,349-350,388-405): loadCalibration -> ScansLRU::get -> ScanData(dataset, frame) -> projectLidarToCamera ->
// featureDepthAssociation -> frameToFrame, on a KITTI-layout directory tree (<root>/<seq>/calib.txt, <root>/<seq>/velodyne/%06d.bin).
// usage: dropin_main <root> <seq> <img_w> <img_h>   (reads <root>/kpA<f>.bin kpB<f>.bin matches.bin pose.bin, writes <root>/out_*.bin)
#include <fstream>
#include <iostream>
#include "velo_dropin.hpp"

enum ResidualType { RESIDUAL_3D3D, RESIDUAL_3D2D, RESIDUAL_2D3D, RESIDUAL_2D2D };      // velo.h:3-8
struct FakeKdTree {};                                                                      // stands for pcl::KdTreeFLANN<pcl::PointXYZ> (lru.h:9)

template <class T> static std::vector<T> rd(const std::string &p) {
    std::ifstream f(p, std::ios::binary); if (!f) throw std::runtime_error("missing " + p);
    f.seekg(0, std::ios::end); size_t n = f.tellg(); f.seekg(0);
    std::vector<T> v(n / sizeof(T)); f.read(reinterpret_cast<char *>(v.data()), n); return v;
}
template <class T> static void wr(const std::string &p, const T *d, size_t n) { std::ofstream f(p, std::ios::binary); f.write(reinterpret_cast<const char *>(d), n * sizeof(T)); }
static std::vector<cv::Point2f> kps(const std::string &p, int cam, int ncam) {
    auto f = rd<float>(p);
    const size_t F = f.size() / 2 / ncam;
    std::vector<cv::Point2f> k(F);
    for (size_t i = 0; i < F; i++) k[i] = cv::Point2f(f[2 * (cam * F + i)], f[2 * (cam * F + i) + 1]);
    return k;
}
static void dump_projection(const std::string &dir, const std::string &tag, const std::vector<std::vector<cv::Point2f>> &projection) {
    std::vector<float> pj; std::vector<int> rc;
    for (auto &r : projection) { rc.push_back((int)r.size()); for (auto &p : r) { pj.push_back(p.x); pj.push_back(p.y); } }
    wr(dir + "out_rc" + tag + ".bin", rc.data(), rc.size());
    wr(dir + "out_proj" + tag + ".bin", pj.data(), pj.size());
}

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    const std::string dir = std::string(argv[1]) + "/", seq = argv[2];
    const int num_cams = 2;
    try {
        velo_dropin::kittipath() = dir;                                                                      // kitti.h:57
        velo_gpu_params prm;
        velo_gpu_default_params(&prm);
        prm.max_slots = 6; prm.icp_skip = 4; prm.max_features = 1500; prm.max_matches = 1500;
        velo_dropin::loadCalibration(seq, atoi(argv[3]), atoi(argv[4]), &prm);                               // main.cpp:73
        velo_dropin::ScansLRU lru(3);                                                                         // main.cpp:145; 3 resident scans here
        velo_dropin::ScanData *sd_prev = lru.get(seq, 0), *sd = lru.get(seq, 1);                             // main.cpp:216,349-350

        // per camera, per frame containers exactly as main.cpp keeps them (keypoints[cam][frame] ...)
        std::vector<std::vector<std::vector<cv::Point2f>>> keypoints(num_cams, std::vector<std::vector<cv::Point2f>>(2));
        std::vector<std::vector<std::vector<int>>> keypoint_ids(num_cams, std::vector<std::vector<int>>(2)), has_depth(num_cams, std::vector<std::vector<int>>(2));
        std::vector<std::vector<velo_dropin::Cloud::Ptr>> kp_with_depth(num_cams, std::vector<velo_dropin::Cloud::Ptr>(2));
        for (int cam = 0; cam < num_cams; cam++) {
            keypoints[cam][0] = kps(dir + "kpA0.bin", cam, num_cams);          // frame 0: detected set
            keypoints[cam][1] = kps(dir + "kpB1.bin", cam, num_cams);          // frame 1: tracked set
            for (int fr = 0; fr < 2; fr++) {
                keypoint_ids[cam][fr].resize(keypoints[cam][fr].size());
                for (size_t i = 0; i < keypoint_ids[cam][fr].size(); i++) keypoint_ids[cam][fr][i] = (int)i;
                std::vector<std::vector<cv::Point2f>> projection;                                             // main.cpp:254-256
                std::vector<velo_dropin::Cloud::Ptr> scans_valid;
                projectLidarToCamera((fr ? sd : sd_prev)->scans, projection, scans_valid, cam);
                kp_with_depth[cam][fr] = velo_dropin::Cloud::Ptr(new velo_dropin::Cloud);
                featureDepthAssociation(scans_valid, projection, keypoints[cam][fr], kp_with_depth[cam][fr], has_depth[cam][fr]);   // main.cpp:261
                if (fr == 1) {
                    const std::string c = std::to_string(cam);
                    dump_projection(dir, c, projection);
                    wr(dir + "out_hd" + c + ".bin", has_depth[cam][fr].data(), has_depth[cam][fr].size());
                    wr(dir + "out_kpwd" + c + ".bin", reinterpret_cast<const float *>(kp_with_depth[cam][fr]->points.data()), kp_with_depth[cam][fr]->points.size() * 4);
                    // the same association through the "projection came from somewhere else" path (copied containers)
                    if (cam == 0) {
                        auto projection2 = projection; auto valid2 = scans_valid;
                        for (auto &c2 : valid2) c2 = velo_dropin::Cloud::Ptr(new velo_dropin::Cloud(*c2));
                        velo_dropin::Cloud::Ptr kpwd2(new velo_dropin::Cloud); std::vector<int> hd2;
                        featureDepthAssociation(valid2, projection2, keypoints[cam][fr], kpwd2, hd2);
                        if (hd2 != has_depth[cam][fr]) { std::cerr << "copied-container path differs\n"; return 1; }
                    }
                }
            }
        }
        // ---- the ICP block alone (velo.h:806-874)
        auto pose = rd<double>(dir + "pose.bin");
        double neq[VELO_NEQ_STRIDE];
        auto rec = velo_dropin::icpCorrespondences(sd->scans, sd_prev->scans, pose.data(), 1, 5, neq);        // icp_skip = 5
        wr(dir + "out_corr.bin", rec.data(), rec.size());
        wr(dir + "out_neq.bin", neq, VELO_NEQ_STRIDE);

        // ---- frameToFrame with the reference's parameter list (main.cpp:388-405)
        auto mf = rd<int>(dir + "matches.bin");                                // [cam][MMfile][2] + n_matches[cam] at the end
        const size_t MMf = (mf.size() - num_cams) / 2 / num_cams;
        std::vector<std::vector<std::pair<int, int>>> matches(num_cams), good_matches(num_cams);
        std::vector<std::vector<ResidualType>> residual_type(num_cams);
        for (int cam = 0; cam < num_cams; cam++)
            for (int i = 0; i < mf[mf.size() - num_cams + cam]; i++) matches[cam].push_back(std::make_pair(mf[2 * (cam * MMf + i)], mf[2 * (cam * MMf + i) + 1]));
        std::map<int, pcl::PointXYZ> landmarks_at_frame;
        std::vector<FakeKdTree> trees;
        double transform[6] = { 0, 0, 0, 0, 0, 1 };                                                           // main.cpp:170
        velo_f2f_report rep;
        velo_dropin::Matrix4d dpose = frameToFrame(matches, keypoints, keypoint_ids, landmarks_at_frame, kp_with_depth, has_depth,
                                                   sd->scans, sd_prev->scans, trees, 1, 0, transform, good_matches, residual_type, true, &rep);
        wr(dir + "out_f2f_transform.bin", transform, 6);
        wr(dir + "out_f2f_report.bin", &rep, 1);
        double T[16];
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) T[4 * i + j] = dpose(i, j);
        wr(dir + "out_f2f_T.bin", T, 16);
        for (int cam = 0; cam < num_cams; cam++) {
            std::vector<int> gm;
            for (size_t i = 0; i < good_matches[cam].size(); i++) { gm.push_back(good_matches[cam][i].first); gm.push_back(good_matches[cam][i].second); gm.push_back((int)residual_type[cam][i]); }
            wr(dir + "out_f2f_good" + std::to_string(cam) + ".bin", gm.data(), gm.size());
        }
        // a landmark for one matched keypoint overrides its lidar depth (velo.h:634-644): the block list must change type there
        if (!matches[0].empty()) {
            landmarks_at_frame[keypoint_ids[0][0][matches[0][0].second]] = pcl::PointXYZ(0.5f, -0.2f, 9.0f);
            double t2[6] = { 0, 0, 0, 0, 0, 1 };
            frameToFrame(matches, keypoints, keypoint_ids, landmarks_at_frame, kp_with_depth, has_depth, sd->scans, sd_prev->scans, trees, 1, 0, t2, good_matches, residual_type, false);
            wr(dir + "out_f2f_lm_transform.bin", t2, 6);
        }

        // ---- ScansLRU (lru.h:31-61): 3 resident scans; frames 2, 3 evict 0 then 1; asking for 0 again reloads it into a freed slot
        const float f0x = sd_prev->scans[0]->points[0].x;
        lru.get(seq, 2); lru.get(seq, 3);
        if (lru.resident() != 3) { std::cerr << "LRU holds " << lru.resident() << " scans\n"; return 1; }
        velo_dropin::ScanData *again = lru.get(seq, 0);
        if (again->_frame != 0 || again->scans[0]->points[0].x != f0x) { std::cerr << "reloaded frame 0 differs\n"; return 1; }
        lru.get(seq, 4); lru.get(seq, 2);                                       // more churn: every slot has been reused by now
        {
            std::vector<std::vector<cv::Point2f>> projection; std::vector<velo_dropin::Cloud::Ptr> scans_valid;
            projectLidarToCamera(lru.get(seq, 0)->scans, projection, scans_valid, 0);
            dump_projection(dir, "_lru0", projection);
        }
        // ---- recycled addresses: a foreign cloud vector is projected, destroyed, and a DIFFERENT scan is built where it was
        for (int round = 0; round < 2; round++) {
            const std::vector<velo_dropin::Cloud::Ptr> &src = lru.get(seq, round == 0 ? 4 : 2)->scans;
            std::vector<velo_dropin::Cloud::Ptr> copy;
            for (auto &c : src) copy.push_back(velo_dropin::Cloud::Ptr(new velo_dropin::Cloud(*c)));
            std::vector<std::vector<cv::Point2f>> projection; std::vector<velo_dropin::Cloud::Ptr> scans_valid;
            projectLidarToCamera(copy, projection, scans_valid, 1);
            dump_projection(dir, std::string("_foreign") + (round == 0 ? "4" : "2"), projection);
        }
        std::cout << "dropin ok: rings " << sd->scans.size() << " queries " << rec.size() << " f2f solves " << rep.n_solves << std::endl;
    } catch (const std::exception &e) { std::cerr << "dropin failed: " << e.what() << std::endl; return 1; }
    return 0;
}
