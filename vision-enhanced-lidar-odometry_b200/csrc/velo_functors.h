// velo_functors.h — the costfunctions.h functors (cost3DPD :17-58, cost3D3D :60-90, cost3D2D :92-130, cost2D3D :132-172,
// cost2D2D :174-220) with their Jacobians, written for one pose shared by many residual blocks.
//
// ceres::AutoDiffCostFunction pushes 6 dual-number partials through AngleAxisRotatePoint for every block (sqrt, sin, cos and
// ~400 flops).  All blocks of a (frame pair, iteration) share ONE pose, and R(w) p is linear in p, so the rotation and its
// derivative are computed once per pose — by the same dual numbers through the same branch of the rotation formula, on the three
// unit vectors — and a block only applies them:   R(w) p = sum_j p_j R e_j,   d(R(w) p)/dw_k = sum_j p_j d(R e_j)/dw_k.
// That is the autodiff derivative up to rounding (checked against the oracle's dual numbers to 1e-12, tests/test_functors_host.py).
//
// Plain C++ (host and device): compiled by nvcc into the kernels and by g++ into the CPU test.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define VELO_HD __host__ __device__ __forceinline__
#define VELO_UNROLL_ALL _Pragma("unroll")
#else
#define VELO_HD inline
#define VELO_UNROLL_ALL
#endif

// ------------------------------------------------------------------------------------------------ 6-partial dual numbers
struct DJ { double a; double v[6]; };
VELO_HD DJ dj(double s) { DJ r; r.a = s; for (int i = 0; i < 6; i++) r.v[i] = 0.0; return r; }
VELO_HD DJ operator+(const DJ &x, const DJ &y) { DJ r; r.a = x.a + y.a; for (int i = 0; i < 6; i++) r.v[i] = x.v[i] + y.v[i]; return r; }
VELO_HD DJ operator-(const DJ &x, const DJ &y) { DJ r; r.a = x.a - y.a; for (int i = 0; i < 6; i++) r.v[i] = x.v[i] - y.v[i]; return r; }
VELO_HD DJ operator-(const DJ &x) { DJ r; r.a = -x.a; for (int i = 0; i < 6; i++) r.v[i] = -x.v[i]; return r; }
VELO_HD DJ operator*(const DJ &x, const DJ &y) { DJ r; r.a = x.a * y.a; for (int i = 0; i < 6; i++) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
VELO_HD DJ operator/(const DJ &x, const DJ &y) { DJ r; const double inv = 1.0 / y.a; r.a = x.a * inv; for (int i = 0; i < 6; i++) r.v[i] = (x.v[i] - r.a * y.v[i]) * inv; return r; }
VELO_HD DJ operator*(const DJ &x, double s) { DJ r; r.a = x.a * s; for (int i = 0; i < 6; i++) r.v[i] = x.v[i] * s; return r; }
VELO_HD DJ operator*(double s, const DJ &x) { return x * s; }
VELO_HD DJ operator+(const DJ &x, double s) { DJ r = x; r.a += s; return r; }
VELO_HD DJ operator-(const DJ &x, double s) { DJ r = x; r.a -= s; return r; }
VELO_HD DJ jsqrt(const DJ &x) { DJ r; r.a = sqrt(x.a); const double d = 1.0 / (2.0 * r.a); for (int i = 0; i < 6; i++) r.v[i] = x.v[i] * d; return r; }
VELO_HD DJ jsin(const DJ &x) { DJ r; r.a = sin(x.a); const double c = cos(x.a); for (int i = 0; i < 6; i++) r.v[i] = c * x.v[i]; return r; }
VELO_HD DJ jcos(const DJ &x) { DJ r; r.a = cos(x.a); const double s = -sin(x.a); for (int i = 0; i < 6; i++) r.v[i] = s * x.v[i]; return r; }

// ceres::AngleAxisRotatePoint on dual numbers
VELO_HD void rot(const DJ w[3], const DJ p[3], DJ out[3]) {
    const DJ th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    if (th2.a > 2.220446049250313e-16) {
        const DJ th = jsqrt(th2), c = jcos(th), s = jsin(th), ith = dj(1.0) / th;
        const DJ u0 = w[0] * ith, u1 = w[1] * ith, u2 = w[2] * ith;
        const DJ x0 = u1 * p[2] - u2 * p[1], x1 = u2 * p[0] - u0 * p[2], x2 = u0 * p[1] - u1 * p[0];
        const DJ tmp = (u0 * p[0] + u1 * p[1] + u2 * p[2]) * (dj(1.0) - c);
        out[0] = p[0] * c + x0 * s + u0 * tmp;
        out[1] = p[1] * c + x1 * s + u1 * tmp;
        out[2] = p[2] * c + x2 * s + u2 * tmp;
    } else {
        out[0] = p[0] + (w[1] * p[2] - w[2] * p[1]);
        out[1] = p[1] + (w[2] * p[0] - w[0] * p[2]);
        out[2] = p[2] + (w[0] * p[1] - w[1] * p[0]);
    }
}

// ------------------------------------------------------------------------------------------------ rotation of one pose as a linear map
// R[3*i+j] = (R(s*w) e_j)_i,  dR[9*k+3*i+j] = d (R(s*w) e_j)_i / d w_k   (s = +1, or -1 for cost2D3D's inverse motion; the
// derivative is with respect to the POSE's w either way, i.e. the chain-rule sign is already in dR).
struct RotPack { double R[9]; double dR[27]; };

// column j (0..2) of the pack: one dual-number rotation of the unit vector e_j
VELO_HD void rotpack_column(const double pose[6], bool inverse, int j, RotPack *P) {
    DJ w[3], e[3], o[3];
    VELO_UNROLL_ALL
    for (int k = 0; k < 3; k++) { w[k] = dj(inverse ? -pose[k] : pose[k]); w[k].v[k] = inverse ? -1.0 : 1.0; e[k] = dj(k == j ? 1.0 : 0.0); }
    rot(w, e, o);
    VELO_UNROLL_ALL
    for (int i = 0; i < 3; i++) { P->R[3 * i + j] = o[i].a; for (int k = 0; k < 3; k++) P->dR[9 * k + 3 * i + j] = o[i].v[k]; }
}
VELO_HD void rotpack_make(const double pose[6], bool inverse, RotPack *P) { for (int j = 0; j < 3; j++) rotpack_column(pose, inverse, j, P); }

// m = R p and Jm[i][k] = d m_i / d w_k for a constant point p
VELO_HD void rot_apply(const RotPack &P, const double p[3], double m[3], double Jm[3][3]) {
    VELO_UNROLL_ALL
    for (int i = 0; i < 3; i++) {
        m[i] = P.R[3 * i] * p[0] + P.R[3 * i + 1] * p[1] + P.R[3 * i + 2] * p[2];
        VELO_UNROLL_ALL
        for (int k = 0; k < 3; k++) Jm[i][k] = P.dR[9 * k + 3 * i] * p[0] + P.dR[9 * k + 3 * i + 1] * p[1] + P.dR[9 * k + 3 * i + 2] * p[2];
    }
}

// Every functor: k = the reference constructor's doubles in order, t = pose[3..6], r = residuals, J = row-major n_res x 6.

// cost3DPD (costfunctions.h:40-53): k = {p[3], n[3], o[3]};  M = R p; M += t - o; r = M . n
VELO_HD void lin3dpd(const double *k, const RotPack &P, const double *t, double *r, double *J) {
    double m[3], Jm[3][3];
    rot_apply(P, k, m, Jm);
    const double m0 = m[0] + (t[0] - k[6]), m1 = m[1] + (t[1] - k[7]), m2 = m[2] + (t[2] - k[8]);
    r[0] = m0 * k[3] + m1 * k[4] + m2 * k[5];
    VELO_UNROLL_ALL
    for (int q = 0; q < 3; q++) { J[q] = Jm[0][q] * k[3] + Jm[1][q] * k[4] + Jm[2][q] * k[5]; J[3 + q] = k[3 + q]; }
}
// cost3D3D (costfunctions.h:77-86): k = {m[3], s[3]};  r = R m + t - s
VELO_HD void lin3d3d(const double *k, const RotPack &P, const double *t, double *r, double *J) {
    double m[3], Jm[3][3];
    rot_apply(P, k, m, Jm);
    VELO_UNROLL_ALL
    for (int i = 0; i < 3; i++) {
        r[i] = m[i] + t[i] - k[3 + i];
        VELO_UNROLL_ALL
        for (int q = 0; q < 3; q++) { J[6 * i + q] = Jm[i][q]; J[6 * i + 3 + q] = (i == q) ? 1.0 : 0.0; }
    }
}
// cost3D2D (costfunctions.h:111-126): k = {m[3], s[2], tc[3]};  M = R m + (t + tc); r = (M0 - sx M2, M1 - sy M2)
VELO_HD void lin3d2d(const double *k, const RotPack &P, const double *t, double *r, double *J) {
    double m[3], Jm[3][3];
    rot_apply(P, k, m, Jm);
    const double M0 = m[0] + (t[0] + k[5]), M1 = m[1] + (t[1] + k[6]), M2 = m[2] + (t[2] + k[7]);
    r[0] = M0 - k[3] * M2; r[1] = M1 - k[4] * M2;
    VELO_UNROLL_ALL
    for (int q = 0; q < 3; q++) { J[q] = Jm[0][q] - k[3] * Jm[2][q]; J[6 + q] = Jm[1][q] - k[4] * Jm[2][q]; }
    J[3] = 1.0; J[4] = 0.0; J[5] = -k[3];
    J[9] = 0.0; J[10] = 1.0; J[11] = -k[4];
}
// cost2D3D (costfunctions.h:151-168): k = {m[3], s[2], tc[3]};  M = R(-w) (m - t) + tc; r = (M0 - sx M2, M1 - sy M2).  Pinv = pack of -w
VELO_HD void lin2d3d(const double *k, const RotPack &Pinv, const double *t, double *r, double *J) {
    const double p[3] = { k[0] - t[0], k[1] - t[1], k[2] - t[2] };
    double m[3], Jm[3][3];
    rot_apply(Pinv, p, m, Jm);
    const double M0 = m[0] + k[5], M1 = m[1] + k[6], M2 = m[2] + k[7];
    r[0] = M0 - k[3] * M2; r[1] = M1 - k[4] * M2;
    VELO_UNROLL_ALL
    for (int q = 0; q < 3; q++) {
        J[q] = Jm[0][q] - k[3] * Jm[2][q]; J[6 + q] = Jm[1][q] - k[4] * Jm[2][q];
        // d M_i / d t_q = -R(-w)[i][q]
        J[3 + q] = -Pinv.R[q] + k[3] * Pinv.R[6 + q]; J[9 + q] = -Pinv.R[3 + q] + k[4] * Pinv.R[6 + q];
    }
}
// cost2D2D (costfunctions.h:192-216): k = {m[2], s[2], tc[3]};  M = R (mx, my, 1); tt = normalise(-R tc + t + tc);
// r = M . (s x tt) written out as in the reference.  The part after the two rotations stays on dual numbers.
VELO_HD void lin2d2d(const double *k, const RotPack &P, const double *t, double *r, double *J) {
    const double p[3] = { k[0], k[1], 1.0 };
    double m[3], Jm[3][3], b[3], Jb[3][3];
    rot_apply(P, p, m, Jm);
    rot_apply(P, k + 4, b, Jb);
    DJ M[3], tt[3];
    VELO_UNROLL_ALL
    for (int i = 0; i < 3; i++) {
        M[i] = dj(m[i]); tt[i] = dj(-b[i] + t[i] + k[4 + i]);
        VELO_UNROLL_ALL
        for (int q = 0; q < 3; q++) { M[i].v[q] = Jm[i][q]; tt[i].v[q] = -Jb[i][q]; }
        tt[i].v[3 + i] = 1.0;
    }
    const DJ tn = jsqrt(tt[0] * tt[0] + tt[1] * tt[1] + tt[2] * tt[2]);
    const DJ tx = tt[0] / tn, ty = tt[1] / tn, tz = tt[2] / tn;
    const double sx = k[2], sy = k[3];
    const DJ res = M[0] * ((-sy) * tz + ty) + M[1] * (sx * tz - tx) + M[2] * ((-sx) * ty + sy * tx);
    r[0] = res.a;
    VELO_UNROLL_ALL
    for (int q = 0; q < 6; q++) J[q] = res.v[q];
}
