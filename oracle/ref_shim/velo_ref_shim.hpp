/*
 * velo_ref_shim.hpp — minimal stand-ins for the third-party types the reference's hot-path source
 * lines use, so that those lines can be compiled VERBATIM from /root/reference (oracle/build_ref.py).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/velo_oracle.cpp header).  Nothing here is reference code: PCL,
 * FLANN, Eigen, OpenCV and Ceres are not in /root/reference and not installed; each stand-in restates
 * the published behaviour of the one call the reference makes, and is the
 * definition of parity for that call ("parity unpinned" for third-party arithmetic).
 */
#pragma once
#include <math.h>
#include <stdlib.h>
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <algorithm>
#include <iostream>
#include <fstream>
#include <iomanip>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

/* ------------------------------------------------------------------ Eigen */
namespace Eigen {
struct Vector3f {
    float v[3];
    Vector3f() : v{ 0, 0, 0 } {}
    Vector3f(float a, float b, float c) : v{ a, b, c } {}
    float &operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
    float &operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    Vector3f operator-(const Vector3f &o) const { return Vector3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vector3f cross(const Vector3f &o) const {
        return Vector3f(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
    }
    /* fixed-size redux order e0 + (e1 + e2) [recall Eigen redux_novec_unroller] */
    float squaredNorm() const { return v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]); }
    float norm() const { return std::sqrt(squaredNorm()); }
    Vector3f &operator/=(float s) { v[0] /= s; v[1] /= s; v[2] /= s; return *this; }
};
struct Matrix4f {
    float m[16];
    float &operator()(int i, int j) { return m[4 * i + j]; }
    float operator()(int i, int j) const { return m[4 * i + j]; }
};
template <class T> struct aligned_allocator : std::allocator<T> {};
} // namespace Eigen

/* ------------------------------------------------------------------ OpenCV */
namespace cv {
struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float a, float b) : x(a), y(b) {} };
/* descriptor matrix: one row per keypoint, `cols` bytes per row (CV_8U) */
struct Mat { int rows = 0, cols = 0; std::vector<unsigned char> data; };
struct DMatch { int queryIdx, trainIdx; float distance; };
enum { NORM_HAMMING = 6 };
/* cv::BFMatcher(NORM_HAMMING)::match: for every query row the train row with the smallest Hamming distance; the first
 * (lowest index) minimum wins [recall: batchDistance keeps strict improvements only] */
struct BFMatcher {
    explicit BFMatcher(int) {}
    void match(const Mat &q, const Mat &t, std::vector<DMatch> &out) const {
        out.clear();
        if (t.rows == 0) return;
        for (int i = 0; i < q.rows; i++) {
            int best = -1, bd = 1 << 30;
            for (int j = 0; j < t.rows; j++) {
                int d = 0;
                for (int k = 0; k < q.cols; k++) d += __builtin_popcount((unsigned)(q.data[(size_t)i * q.cols + k] ^ t.data[(size_t)j * t.cols + k]));
                if (d < bd) { bd = d; best = j; }
            }
            out.push_back(DMatch{ i, best, (float)bd });
        }
    }
};
} // namespace cv

/* ------------------------------------------------------------------ PCL */
namespace pcl {
struct PointXYZ {
    float x, y, z, pad;
    PointXYZ() : x(0), y(0), z(0), pad(1.0f) {}
    PointXYZ(float a, float b, float c) : x(a), y(b), z(c), pad(1.0f) {}
    Eigen::Vector3f getVector3fMap() const { return Eigen::Vector3f(x, y, z); }
};
template <class P> struct PointCloud {
    typedef std::shared_ptr<PointCloud<P>> Ptr;
    std::vector<P> points;
    size_t size() const { return points.size(); }
    const P &at(size_t i) const { return points.at(i); }
    P &at(size_t i) { return points.at(i); }
    void push_back(const P &p) { points.push_back(p); }
    const P &back() const { return points.back(); }
};
/* dense branch of PCL 1.7/1.8 transformPointCloud: left-to-right f32 [recall] */
template <class P> void transformPointCloud(const PointCloud<P> &in, PointCloud<P> &out, const Eigen::Matrix4f &t) {
    out.points.resize(in.points.size());
    for (size_t i = 0; i < in.points.size(); i++) {
        const P &p = in.points[i];
        out.points[i].x = static_cast<float>(t(0, 0) * p.x + t(0, 1) * p.y + t(0, 2) * p.z + t(0, 3));
        out.points[i].y = static_cast<float>(t(1, 0) * p.x + t(1, 1) * p.y + t(1, 2) * p.z + t(1, 3));
        out.points[i].z = static_cast<float>(t(2, 0) * p.x + t(2, 1) * p.y + t(2, 2) * p.z + t(2, 3));
        out.points[i].pad = 1.0f;
    }
}
/* exact 1-NN under flann::L2_Simple<float>; ties -> lowest index (north_star) */
template <class P> struct KdTreeFLANN {
    typename PointCloud<P>::Ptr cloud;
    void setInputCloud(const typename PointCloud<P>::Ptr &c) { cloud = c; }
    int nearestKSearch(const P &q, int k, std::vector<int> &ids, std::vector<float> &d2) const {
        if (!cloud || cloud->points.empty() || k != 1) return 0;
        int best = -1; float bd = std::numeric_limits<float>::infinity();
        for (size_t i = 0; i < cloud->points.size(); i++) {
            const P &p = cloud->points[i];
            float acc = 0, diff;
            diff = q.x - p.x; acc += diff * diff;
            diff = q.y - p.y; acc += diff * diff;
            diff = q.z - p.z; acc += diff * diff;
            if (acc < bd) { bd = acc; best = (int)i; }
        }
        ids[0] = best; d2[0] = bd;
        return 1;
    }
};
} // namespace pcl

/* ------------------------------------------------------------------ Ceres */
namespace ceres {
template <int N> struct Jet {
    double a; double v[N];
    Jet() : a(0) { for (int i = 0; i < N; i++) v[i] = 0; }
    Jet(double s) : a(s) { for (int i = 0; i < N; i++) v[i] = 0; }
    Jet(int s) : a(s) { for (int i = 0; i < N; i++) v[i] = 0; }
};
#define VJ template <int N> inline
VJ Jet<N> operator+(const Jet<N> &x, const Jet<N> &y) { Jet<N> r; r.a = x.a + y.a; for (int i = 0; i < N; i++) r.v[i] = x.v[i] + y.v[i]; return r; }
VJ Jet<N> operator-(const Jet<N> &x, const Jet<N> &y) { Jet<N> r; r.a = x.a - y.a; for (int i = 0; i < N; i++) r.v[i] = x.v[i] - y.v[i]; return r; }
VJ Jet<N> operator-(const Jet<N> &x) { Jet<N> r; r.a = -x.a; for (int i = 0; i < N; i++) r.v[i] = -x.v[i]; return r; }
VJ Jet<N> operator*(const Jet<N> &x, const Jet<N> &y) { Jet<N> r; r.a = x.a * y.a; for (int i = 0; i < N; i++) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
VJ Jet<N> operator/(const Jet<N> &x, const Jet<N> &y) { Jet<N> r; double inv = 1.0 / y.a; r.a = x.a * inv; for (int i = 0; i < N; i++) { r.v[i] = (x.v[i] - r.a * y.v[i]) * inv; } return r; }
VJ Jet<N> operator+(const Jet<N> &x, double s) { Jet<N> r = x; r.a += s; return r; }
VJ Jet<N> operator+(double s, const Jet<N> &x) { Jet<N> r = x; r.a += s; return r; }
VJ Jet<N> operator-(const Jet<N> &x, double s) { Jet<N> r = x; r.a -= s; return r; }
VJ Jet<N> operator-(double s, const Jet<N> &x) { Jet<N> r = -x; r.a += s; return r; }
VJ Jet<N> operator*(const Jet<N> &x, double s) { Jet<N> r; r.a = x.a * s; for (int i = 0; i < N; i++) r.v[i] = x.v[i] * s; return r; }
VJ Jet<N> operator*(double s, const Jet<N> &x) { return x * s; }
VJ Jet<N> operator/(const Jet<N> &x, double s) { return x * (1.0 / s); }
VJ Jet<N> &operator+=(Jet<N> &x, const Jet<N> &y) { x = x + y; return x; }
VJ Jet<N> &operator-=(Jet<N> &x, const Jet<N> &y) { x = x - y; return x; }
VJ Jet<N> &operator*=(Jet<N> &x, const Jet<N> &y) { x = x * y; return x; }
VJ Jet<N> &operator/=(Jet<N> &x, const Jet<N> &y) { x = x / y; return x; }
VJ bool operator>(const Jet<N> &x, const Jet<N> &y) { return x.a > y.a; }
VJ bool operator<(const Jet<N> &x, const Jet<N> &y) { return x.a < y.a; }
VJ Jet<N> sqrt(const Jet<N> &x) { Jet<N> r; r.a = std::sqrt(x.a); double d = 1.0 / (2.0 * r.a); for (int i = 0; i < N; i++) r.v[i] = x.v[i] * d; return r; }
VJ Jet<N> sin(const Jet<N> &x) { Jet<N> r; r.a = std::sin(x.a); double c = std::cos(x.a); for (int i = 0; i < N; i++) r.v[i] = c * x.v[i]; return r; }
VJ Jet<N> cos(const Jet<N> &x) { Jet<N> r; r.a = std::cos(x.a); double s = -std::sin(x.a); for (int i = 0; i < N; i++) r.v[i] = s * x.v[i]; return r; }
#undef VJ
inline double sqrt(double x) { return std::sqrt(x); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }

/* ceres/rotation.h AngleAxisRotatePoint */
template <typename T> inline void AngleAxisRotatePoint(const T angle_axis[3], const T pt[3], T result[3]) {
    const T theta2 = angle_axis[0] * angle_axis[0] + angle_axis[1] * angle_axis[1] + angle_axis[2] * angle_axis[2];
    if (theta2 > T(std::numeric_limits<double>::epsilon())) {
        const T theta = sqrt(theta2);
        const T costheta = cos(theta);
        const T sintheta = sin(theta);
        const T theta_inverse = T(1.0) / theta;
        const T w[3] = { angle_axis[0] * theta_inverse, angle_axis[1] * theta_inverse, angle_axis[2] * theta_inverse };
        const T w_cross_pt[3] = { w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0] };
        const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
        result[0] = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
        result[1] = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
        result[2] = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
    } else {
        const T w_cross_pt[3] = { angle_axis[1] * pt[2] - angle_axis[2] * pt[1], angle_axis[2] * pt[0] - angle_axis[0] * pt[2],
                                  angle_axis[0] * pt[1] - angle_axis[1] * pt[0] };
        result[0] = pt[0] + w_cross_pt[0];
        result[1] = pt[1] + w_cross_pt[1];
        result[2] = pt[2] + w_cross_pt[2];
    }
}

struct CostFunction {
    virtual ~CostFunction() {}
    virtual int num_residuals() const = 0;
    virtual int num_params() const = 0;
    virtual void Evaluate6(const double *x, double *r, double *J) const = 0; /* J row-major nres x num_params */
};
template <class F, int M, int NP> struct AutoDiffCostFunction : CostFunction {
    std::unique_ptr<F> f;
    explicit AutoDiffCostFunction(F *fn) : f(fn) {}
    int num_residuals() const override { return M; }
    int num_params() const override { return NP; }
    void Evaluate6(const double *x, double *r, double *J) const override {
        Jet<NP> xj[NP], rj[M];
        for (int i = 0; i < NP; i++) { xj[i] = Jet<NP>(x[i]); xj[i].v[i] = 1.0; }
        (*f)(xj, rj);
        for (int i = 0; i < M; i++) { r[i] = rj[i].a; for (int j = 0; j < NP; j++) J[NP * i + j] = rj[i].v[j]; }
    }
};
enum Ownership { DO_NOT_TAKE_OWNERSHIP, TAKE_OWNERSHIP };
struct LossFunction { virtual ~LossFunction() {} virtual void Evaluate(double s, double rho[3]) const = 0; };
/* loss_function.cc [recall] */
struct ArctanLoss : LossFunction {
    double a_, b_;
    explicit ArctanLoss(double a) : a_(a), b_(1 / (a * a)) {}
    void Evaluate(double s, double rho[3]) const override {
        const double sum = 1 + s * s * b_; const double inv = 1 / sum;
        rho[0] = a_ * atan2(s, a_); rho[1] = std::max(std::numeric_limits<double>::min(), inv); rho[2] = -2.0 * s * b_ * (inv * inv);
    }
};
struct CauchyLoss : LossFunction {
    double b_, c_;
    explicit CauchyLoss(double a) : b_(a * a), c_(1 / b_) {}
    void Evaluate(double s, double rho[3]) const override {
        const double sum = 1 + s * c_; const double inv = 1 / sum;
        rho[0] = b_ * log(sum); rho[1] = std::max(std::numeric_limits<double>::min(), inv); rho[2] = -c_ * (inv * inv);
    }
};
struct TrivialLoss : LossFunction {
    void Evaluate(double s, double rho[3]) const override { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
};
struct ScaledLoss : LossFunction {
    std::unique_ptr<LossFunction> rho_; double a_;
    ScaledLoss(LossFunction *rho, double a, Ownership) : rho_(rho), a_(a) {}
    void Evaluate(double s, double rho[3]) const override { rho_->Evaluate(s, rho); rho[0] *= a_; rho[1] *= a_; rho[2] *= a_; }
};
typedef int ResidualBlockId;
struct Problem {
    struct Options { bool enable_fast_removal = false; };
    struct Block { std::unique_ptr<CostFunction> cost; std::unique_ptr<LossFunction> loss; double *x; bool alive; };
    std::vector<Block> blocks;
    Problem() {}
    explicit Problem(const Options &) {}
    ResidualBlockId AddResidualBlock(CostFunction *c, LossFunction *l, double *x) {
        blocks.push_back(Block{ std::unique_ptr<CostFunction>(c), std::unique_ptr<LossFunction>(l), x, true });
        return (int)blocks.size() - 1;
    }
    void RemoveResidualBlock(ResidualBlockId id) { blocks[id].alive = false; }
};

/* ceres::Solve stand-in for problems with ONE parameter block of n <= 6 doubles: Levenberg-Marquardt trust region with Ceres'
 * documented defaults [recall] (initial radius 1e4, min/max_lm_diagonal 1e-6/1e32, min_relative_decrease 1e-3, radius update
 * 1/max(1/3, 1-(2rho-1)^3), halving with doubling factor on failure, function/gradient/parameter tolerance 1e-6/1e-10/1e-8,
 * 50 iterations).  Third-party behaviour: parity with a real Ceres build is "to solver tolerance". */
enum LinearSolverType { DENSE_QR, DENSE_SCHUR };
struct Solver { struct Options { LinearSolverType linear_solver_type = DENSE_QR; bool minimizer_progress_to_stdout = false; int num_threads = 1; };
                struct Summary { int iterations = 0; double initial_cost = 0, final_cost = 0; }; };
inline void Solve(const Solver::Options &, Problem *problem, Solver::Summary *summary) {
    int n = 0; double *x = nullptr;
    for (auto &b : problem->blocks) if (b.alive) { n = b.cost->num_params(); x = b.x; break; }
    if (!x) return;
    const int T = n * (n + 1) / 2;
    auto eval = [&](const double *xx, double &cost, double *H, double *g) {
        cost = 0; for (int i = 0; i < T; i++) H[i] = 0; for (int i = 0; i < n; i++) g[i] = 0;
        for (auto &b : problem->blocks) {
            if (!b.alive) continue;
            double r[8], J[48], rho[3]; const int nr = b.cost->num_residuals();
            b.cost->Evaluate6(xx, r, J);
            double s = 0; for (int i = 0; i < nr; i++) s += r[i] * r[i];
            b.loss->Evaluate(s, rho);
            int o = 0;
            for (int a = 0; a < n; a++) for (int c = a; c < n; c++, o++) { double h = 0; for (int i = 0; i < nr; i++) h += J[n * i + a] * J[n * i + c]; H[o] += rho[1] * h; }
            for (int a = 0; a < n; a++) { double t = 0; for (int i = 0; i < nr; i++) t += J[n * i + a] * r[i]; g[a] += rho[1] * t; }
            cost += 0.5 * rho[0];
        }
    };
    auto chol = [&](const double *A, const double *b, double *xo) -> bool {
        double M[6][6], L[6][6]; int o = 0;
        for (int i = 0; i < n; i++) for (int j = i; j < n; j++, o++) { M[i][j] = A[o]; M[j][i] = A[o]; }
        for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) {
            double s = M[i][j]; for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
            if (i == j) { if (!(s > 0.0)) return false; L[i][i] = std::sqrt(s); } else L[i][j] = s / L[j][j];
        }
        double y[6];
        for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[i][k] * y[k]; y[i] = s / L[i][i]; }
        for (int i = n - 1; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < n; k++) s -= L[k][i] * xo[k]; xo[i] = s / L[i][i]; }
        return true;
    };
    double cost, H[21], g[6], c2, H2[21], g2[6], xt[6], delta[6], radius = 1e4, dec = 2.0, mc = 0;
    bool done = false; int iter = 0;
    eval(x, cost, H, g);
    summary->initial_cost = cost;
    auto gsmall = [&]() { double m = 0; for (int i = 0; i < n; i++) m = std::max(m, std::fabs(g[i])); return m <= 1e-10; };
    auto propose = [&]() -> bool {
        double A[21], mg[6]; int o = 0;
        for (int i = 0; i < n; i++) for (int j = i; j < n; j++, o++) { A[o] = H[o]; if (i == j) A[o] += std::min(std::max(H[o], 1e-6), 1e32) / radius; }
        for (int i = 0; i < n; i++) mg[i] = -g[i];
        if (!chol(A, mg, delta)) return false;
        double M[6][6]; int k = 0;
        for (int i = 0; i < n; i++) for (int j = i; j < n; j++, k++) { M[i][j] = H[k]; M[j][i] = H[k]; }
        mc = 0;
        for (int i = 0; i < n; i++) { double hd = 0; for (int j = 0; j < n; j++) hd += M[i][j] * delta[j]; mc -= delta[i] * (g[i] + 0.5 * hd); }
        if (!(mc > 0.0)) return false;
        double nd = 0, nx = 0;
        for (int i = 0; i < n; i++) { nd += delta[i] * delta[i]; nx += x[i] * x[i]; xt[i] = x[i] + delta[i]; }
        if (std::sqrt(nd) <= 1e-8 * (std::sqrt(nx) + 1e-8)) done = true;
        return true;
    };
    auto next_trial = [&]() {
        if (!done && iter >= 50) done = true;
        while (!done && !propose()) { radius /= dec; dec *= 2.0; if (radius < 1e-32) done = true; }
    };
    if (gsmall()) done = true;
    next_trial();
    while (!done) {
        eval(xt, c2, H2, g2); iter++;
        const double rho = (cost - c2) / mc;
        if (rho > 1e-3) {
            for (int i = 0; i < n; i++) x[i] = xt[i];
            if (std::fabs(cost - c2) < 1e-6 * cost) done = true;
            const double t = 2.0 * rho - 1.0;
            radius = std::min(1e16, radius / std::max(1.0 / 3.0, 1.0 - t * t * t)); dec = 2.0;
            cost = c2; for (int i = 0; i < T; i++) H[i] = H2[i]; for (int i = 0; i < n; i++) g[i] = g2[i];
            if (gsmall()) done = true;
        } else { radius /= dec; dec *= 2.0; if (radius < 1e-32) done = true; }
        next_trial();
    }
    summary->iterations = iter; summary->final_cost = cost;
}
} // namespace ceres
