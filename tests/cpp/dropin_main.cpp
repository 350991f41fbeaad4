// dropin_main.cpp — exercises include/velo_dropin.hpp exactly the way main.cpp uses the reference functions
// (main.cpp:216,254-261,388-405): ScanData -> projectLidarToCamera -> featureDepthAssociation -> ICP correspondences.
// usage: dropin_main <dir>   (reads calib.bin scan0.bin scan1.bin kp.bin pose.bin, writes out_*.bin)
#include <fstream>
#include <iostream>
#include "velo_dropin.hpp"

template <class T> static std::vector<T> rd(const std::string &p) {
    std::ifstream f(p, std::ios::binary); if (!f) throw std::runtime_error("missing " + p);
    f.seekg(0, std::ios::end); size_t n = f.tellg(); f.seekg(0);
    std::vector<T> v(n / sizeof(T)); f.read(reinterpret_cast<char *>(v.data()), n); return v;
}
template <class T> static void wr(const std::string &p, const T *d, size_t n) { std::ofstream f(p, std::ios::binary); f.write(reinterpret_cast<const char *>(d), n * sizeof(T)); }

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const std::string dir = std::string(argv[1]) + "/";
    try {
        auto cal = rd<float>(dir + "calib.bin");            // P[48], Tr[12], w, h
        velo_dropin::loadCalibrationFromArrays(cal.data(), cal.data() + 48, (int)cal[60], (int)cal[61]);
        auto s0 = rd<float>(dir + "scan0.bin"), s1 = rd<float>(dir + "scan1.bin");
        velo_dropin::ScanData sd_prev(s0.data(), (int)s0.size() / 4, 0), sd(s1.data(), (int)s1.size() / 4, 1);   // lru.h:12-28
        auto kpf = rd<float>(dir + "kp.bin");
        std::vector<cv::Point2f> keypoints(kpf.size() / 2);
        for (size_t i = 0; i < keypoints.size(); i++) keypoints[i] = cv::Point2f(kpf[2 * i], kpf[2 * i + 1]);
        for (int cam = 0; cam < 2; cam++) {
            std::vector<std::vector<cv::Point2f>> projection;                                                 // main.cpp:254-256
            std::vector<velo_dropin::Cloud::Ptr> scans_valid;
            projectLidarToCamera(sd.scans, projection, scans_valid, cam);
            velo_dropin::Cloud::Ptr kpwd(new velo_dropin::Cloud);
            std::vector<int> has_depth;
            featureDepthAssociation(scans_valid, projection, keypoints, kpwd, has_depth);                      // main.cpp:261
            std::vector<float> pj; std::vector<int> rc;
            for (auto &r : projection) { rc.push_back((int)r.size()); for (auto &p : r) { pj.push_back(p.x); pj.push_back(p.y); } }
            const std::string c = std::to_string(cam);
            wr(dir + "out_rc" + c + ".bin", rc.data(), rc.size());
            wr(dir + "out_proj" + c + ".bin", pj.data(), pj.size());
            wr(dir + "out_hd" + c + ".bin", has_depth.data(), has_depth.size());
            wr(dir + "out_kpwd" + c + ".bin", reinterpret_cast<const float *>(kpwd->points.data()), kpwd->points.size() * 4);
            // the same association through the "projection came from somewhere else" path (copied containers)
            auto projection2 = projection; auto valid2 = scans_valid;
            for (auto &c2 : valid2) c2 = velo_dropin::Cloud::Ptr(new velo_dropin::Cloud(*c2));
            velo_dropin::Cloud::Ptr kpwd2(new velo_dropin::Cloud); std::vector<int> hd2;
            if (cam == 0) {
                featureDepthAssociation(valid2, projection2, keypoints, kpwd2, hd2);
                if (hd2 != has_depth) { std::cerr << "copied-container path differs\n"; return 1; }
            }
        }
        auto pose = rd<double>(dir + "pose.bin");
        double neq[VELO_NEQ_STRIDE];
        auto rec = velo_dropin::icpCorrespondences(sd.scans, sd_prev.scans, pose.data(), 1, 5, neq);          // velo.h:806-874, icp_skip=5
        wr(dir + "out_corr.bin", rec.data(), rec.size());
        wr(dir + "out_neq.bin", neq, VELO_NEQ_STRIDE);
        std::cout << "dropin ok: rings " << sd.scans.size() << " queries " << rec.size() << std::endl;
    } catch (const std::exception &e) { std::cerr << "dropin failed: " << e.what() << std::endl; return 1; }
    return 0;
}
