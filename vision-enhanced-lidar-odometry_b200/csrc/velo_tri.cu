// velo_tri.cu — batched landmark triangulation: triangulatePoint (velo.h:1027-1130) for thousands of
// landmarks at once, one thread per landmark.  Each landmark is an independent 3-parameter least-squares problem over its
// 3-D observations (triangulation3D, TrivialLoss, costfunctions.h:333-375) and 2-D observations (triangulation2D,
// Scaled(Cauchy(loss_thresh_3D2D), weight_3D2D), costfunctions.h:288-331), solved with the same Levenberg-Marquardt policy as
// velo_solve.cu (Ceres' documented defaults, restated; parity "to solver tolerance" against the oracle's identical restatement).
#include "velo_common.cuh"

namespace {
struct J3 { double a, v[3]; };
__device__ __forceinline__ J3 j3(double s) { return J3{ s, { 0.0, 0.0, 0.0 } }; }
__device__ __forceinline__ J3 operator+(const J3 &x, const J3 &y) { return J3{ x.a + y.a, { x.v[0] + y.v[0], x.v[1] + y.v[1], x.v[2] + y.v[2] } }; }
__device__ __forceinline__ J3 operator-(const J3 &x, const J3 &y) { return J3{ x.a - y.a, { x.v[0] - y.v[0], x.v[1] - y.v[1], x.v[2] - y.v[2] } }; }
__device__ __forceinline__ J3 operator*(const J3 &x, const J3 &y) { return J3{ x.a * y.a, { x.a * y.v[0] + x.v[0] * y.a, x.a * y.v[1] + x.v[1] * y.a, x.a * y.v[2] + x.v[2] * y.a } }; }
__device__ __forceinline__ J3 operator*(const J3 &x, double s) { return J3{ x.a * s, { x.v[0] * s, x.v[1] * s, x.v[2] * s } }; }

// ceres::AngleAxisRotatePoint with a CONSTANT angle-axis (the camera pose) applied to a dual-number point
__device__ void rot_const(const double w[3], const J3 p[3], J3 out[3]) {
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    if (th2 > 2.220446049250313e-16) {
        const double th = sqrt(th2), c = cos(th), s = sin(th), ith = 1.0 / th;
        const double u0 = w[0] * ith, u1 = w[1] * ith, u2 = w[2] * ith;
        const J3 x0 = p[2] * u1 - p[1] * u2, x1 = p[0] * u2 - p[2] * u0, x2 = p[1] * u0 - p[0] * u1;
        const J3 tmp = (p[0] * u0 + p[1] * u1 + p[2] * u2) * (1.0 - c);
        out[0] = p[0] * c + x0 * s + tmp * u0;
        out[1] = p[1] * c + x1 * s + tmp * u1;
        out[2] = p[2] * c + x2 * s + tmp * u2;
    } else {
        out[0] = p[0] + (p[2] * w[1] - p[1] * w[2]);
        out[1] = p[1] + (p[0] * w[2] - p[2] * w[0]);
        out[2] = p[2] + (p[1] * w[0] - p[0] * w[1]);
    }
}

struct TriProb {
    const velo_tri_obs3 *o3; int n3;
    const velo_tri_obs2 *o2; int n2;
    const double *poses; int n_frames;
    const float (*cam_t)[3];
    double loss_a, weight;
};

// cost, H (6 upper), g (3) of the problem at x; returns the number of blocks
__device__ int tri_eval(const TriProb &P, const double x[3], double &cost, double H[6], double g[3]) {
    cost = 0.0;
    for (int i = 0; i < 6; i++) H[i] = 0.0;
    for (int i = 0; i < 3; i++) g[i] = 0.0;
    J3 pt[3];
    for (int i = 0; i < 3; i++) { pt[i] = j3(x[i]); pt[i].v[i] = 1.0; }
    int nb = 0;
    auto add = [&](const J3 *r, int nr, double rho0, double rho1) {
        int o = 0;
        for (int a = 0; a < 3; a++) for (int b = a; b < 3; b++, o++) { double h = 0.0; for (int i = 0; i < nr; i++) h += r[i].v[a] * r[i].v[b]; H[o] += rho1 * h; }
        for (int a = 0; a < 3; a++) { double t = 0.0; for (int i = 0; i < nr; i++) t += r[i].v[a] * r[i].a; g[a] += rho1 * t; }
        cost += 0.5 * rho0; nb++;
    };
    for (int k = 0; k < P.n3; k++) {                                   // triangulation3D, costfunctions.h:357-371
        const velo_tri_obs3 ob = P.o3[k];
        if (ob.frame < 0 || ob.frame >= P.n_frames) continue;
        const double *cp = P.poses + 6 * (size_t)ob.frame;
        const double w[3] = { -cp[0], -cp[1], -cp[2] };
        J3 m0[3] = { pt[0] - j3(cp[3]), pt[1] - j3(cp[4]), pt[2] - j3(cp[5]) }, m[3], r[3];
        rot_const(w, m0, m);
        r[0] = m[0] - j3((double)ob.x); r[1] = m[1] - j3((double)ob.y); r[2] = m[2] - j3((double)ob.z);
        const double s = r[0].a * r[0].a + r[1].a * r[1].a + r[2].a * r[2].a;
        add(r, 3, s, 1.0);                                             // TrivialLoss, velo.h:1078
    }
    for (int k = 0; k < P.n2; k++) {                                   // triangulation2D, costfunctions.h:312-328
        const velo_tri_obs2 ob = P.o2[k];
        if (ob.frame < 0 || ob.frame >= P.n_frames || ob.cam < 0 || ob.cam >= VELO_MAX_CAMS) continue;
        const double *cp = P.poses + 6 * (size_t)ob.frame;
        const double w[3] = { -cp[0], -cp[1], -cp[2] };
        J3 m0[3] = { pt[0] - j3(cp[3]), pt[1] - j3(cp[4]), pt[2] - j3(cp[5]) }, m[3], r[2];
        rot_const(w, m0, m);
        m[0] = m[0] + j3((double)P.cam_t[ob.cam][0]); m[1] = m[1] + j3((double)P.cam_t[ob.cam][1]); m[2] = m[2] + j3((double)P.cam_t[ob.cam][2]);
        r[0] = m[0] - m[2] * (double)ob.x; r[1] = m[1] - m[2] * (double)ob.y;
        const double s = r[0].a * r[0].a + r[1].a * r[1].a;
        const double bb = P.loss_a * P.loss_a, cc = 1.0 / bb, sum = 1.0 + s * cc, inv = 1.0 / sum;   // Scaled(Cauchy), velo.h:1117-1121
        add(r, 2, P.weight * bb * log(sum), P.weight * fmax(2.2250738585072014e-308, inv));
    }
    return nb;
}

__device__ bool chol3_solve(const double A[6], const double b[3], double xo[3]) {
    const double l00s = A[0]; if (!(l00s > 0.0)) return false;
    const double l00 = sqrt(l00s), l10 = A[1] / l00, l20 = A[2] / l00;
    const double l11s = A[3] - l10 * l10; if (!(l11s > 0.0)) return false;
    const double l11 = sqrt(l11s), l21 = (A[4] - l20 * l10) / l11;
    const double l22s = A[5] - l20 * l20 - l21 * l21; if (!(l22s > 0.0)) return false;
    const double l22 = sqrt(l22s);
    const double y0 = b[0] / l00, y1 = (b[1] - l10 * y0) / l11, y2 = (b[2] - l20 * y0 - l21 * y1) / l22;
    xo[2] = y2 / l22; xo[1] = (y1 - l21 * xo[2]) / l11; xo[0] = (y0 - l10 * xo[1] - l20 * xo[2]) / l00;
    return true;
}

// Levenberg-Marquardt on the 3-vector (same policy as k_lm_step); returns the number of trial evaluations
__device__ int tri_solve(const TriProb &P, double x[3], int max_iterations) {
    double cost, H[6], g[3], c2, H2[6], g2[3], xt[3], delta[3], radius = 1e4, dec = 2.0, mc = 0.0;
    bool done = false;
    int iter = 0;
    if (tri_eval(P, x, cost, H, g) == 0) return 0;
    auto gradient_small = [&]() { return fmax(fabs(g[0]), fmax(fabs(g[1]), fabs(g[2]))) <= 1e-10; };
    auto propose = [&]() -> bool {
        double A[6] = { H[0], H[1], H[2], H[3], H[4], H[5] }, mg[3] = { -g[0], -g[1], -g[2] };
        A[0] += fmin(fmax(H[0], 1e-6), 1e32) / radius; A[3] += fmin(fmax(H[3], 1e-6), 1e32) / radius; A[5] += fmin(fmax(H[5], 1e-6), 1e32) / radius;
        if (!chol3_solve(A, mg, delta)) return false;
        const double hd0 = H[0] * delta[0] + H[1] * delta[1] + H[2] * delta[2], hd1 = H[1] * delta[0] + H[3] * delta[1] + H[4] * delta[2],
                     hd2 = H[2] * delta[0] + H[4] * delta[1] + H[5] * delta[2];
        mc = -(delta[0] * (g[0] + 0.5 * hd0) + delta[1] * (g[1] + 0.5 * hd1) + delta[2] * (g[2] + 0.5 * hd2));
        if (!(mc > 0.0)) return false;
        double nd = 0.0, nx = 0.0;
        for (int i = 0; i < 3; i++) { nd += delta[i] * delta[i]; nx += x[i] * x[i]; xt[i] = x[i] + delta[i]; }
        if (sqrt(nd) <= 1e-8 * (sqrt(nx) + 1e-8)) done = true;
        return true;
    };
    auto next_trial = [&]() {
        if (!done && iter >= max_iterations) done = true;
        while (!done && !propose()) { radius /= dec; dec *= 2.0; if (radius < 1e-32) done = true; }
    };
    if (gradient_small()) done = true;
    next_trial();
    while (!done) {
        tri_eval(P, xt, c2, H2, g2);
        iter++;
        const double rho = (cost - c2) / mc;
        if (rho > 1e-3) {
            for (int i = 0; i < 3; i++) x[i] = xt[i];
            if (fabs(cost - c2) < 1e-6 * cost) done = true;
            const double t = 2.0 * rho - 1.0;
            radius = fmin(1e16, radius / fmax(1.0 / 3.0, 1.0 - t * t * t)); dec = 2.0;
            cost = c2;
            for (int i = 0; i < 6; i++) H[i] = H2[i];
            for (int i = 0; i < 3; i++) g[i] = g2[i];
            if (gradient_small()) done = true;
        } else { radius /= dec; dec *= 2.0; if (radius < 1e-32) done = true; }
        next_trial();
    }
    return iter;
}
} // namespace

struct TriCal { float cam_t[VELO_MAX_CAMS][3]; };

__global__ void __launch_bounds__(128) k_triangulate(int L, const int *__restrict__ off3, const velo_tri_obs3 *__restrict__ obs3,
                                                     const int *__restrict__ off2, const velo_tri_obs2 *__restrict__ obs2,
                                                     const double *__restrict__ poses, int n_frames, TriCal cal, double loss_a, double weight,
                                                     const float *__restrict__ init_xyz, const int *__restrict__ has_init,
                                                     float *__restrict__ out_xyz, int *__restrict__ iterations) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    TriProb P;
    P.o3 = obs3 + off3[l]; P.n3 = off3[l + 1] - off3[l];
    P.o2 = obs2 + off2[l]; P.n2 = off2[l + 1] - off2[l];
    P.poses = poses; P.n_frames = n_frames; P.cam_t = cal.cam_t; P.loss_a = loss_a; P.weight = weight;
    double x[3] = { 0.0, 0.0, 10.0 };                                  // velo.h:1041
    const bool init = has_init && has_init[l];
    if (init) { x[0] = init_xyz[3 * l]; x[1] = init_xyz[3 * l + 1]; x[2] = init_xyz[3 * l + 2]; }
    int it = 0;
    if (!init && P.n3 > 0) {                                           // velo.h:1082-1085: first 3-D observation alone initialises
        TriProb P1 = P; P1.n3 = 1; P1.n2 = 0;
        it += tri_solve(P1, x, 50);
    }
    it += tri_solve(P, x, 50);                                         // velo.h:1126
    out_xyz[3 * l] = (float)x[0]; out_xyz[3 * l + 1] = (float)x[1]; out_xyz[3 * l + 2] = (float)x[2];   // velo.h:1127-1129
    if (iterations) iterations[l] = it;
}

void launch_triangulate(const Launcher &L, int n, const int *off3, const velo_tri_obs3 *obs3, const int *off2, const velo_tri_obs2 *obs2,
                        const double *poses, int n_frames, const DevCalib &cal, double loss_a, double weight,
                        const float *init_xyz, const int *has_init, float *out_xyz, int *iterations) {
    if (n <= 0) return;
    TriCal tc;
    for (int c = 0; c < VELO_MAX_CAMS; c++) for (int i = 0; i < 3; i++) tc.cam_t[c][i] = cal.cam_t[c][i];
    if (L.pre) L.pre(L.user, VK_SOLVE);
    k_triangulate<<<(n + 127) / 128, 128, 0, L.stream>>>(n, off3, obs3, off2, obs2, poses, n_frames, tc, loss_a, weight, init_xyz, has_init, out_xyz, iterations);
    if (L.post) L.post(L.user, VK_SOLVE);
}
