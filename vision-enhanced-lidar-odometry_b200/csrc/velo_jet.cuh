// velo_jet.cuh — 6-partial forward-mode dual numbers on the device and the costfunctions.h functors written on them
// (what ceres::AutoDiffCostFunction evaluates): shared by the visual residual kernel and the fixed-block evaluation of the
// device-resident solve.
#pragma once
#include "velo_common.cuh"

struct DJ { double a; double v[6]; };
__device__ __forceinline__ DJ dj(double s) { DJ r; r.a = s; for (int i = 0; i < 6; i++) r.v[i] = 0.0; return r; }
__device__ __forceinline__ DJ operator+(const DJ &x, const DJ &y) { DJ r; r.a = x.a + y.a; for (int i = 0; i < 6; i++) r.v[i] = x.v[i] + y.v[i]; return r; }
__device__ __forceinline__ DJ operator-(const DJ &x, const DJ &y) { DJ r; r.a = x.a - y.a; for (int i = 0; i < 6; i++) r.v[i] = x.v[i] - y.v[i]; return r; }
__device__ __forceinline__ DJ operator-(const DJ &x) { DJ r; r.a = -x.a; for (int i = 0; i < 6; i++) r.v[i] = -x.v[i]; return r; }
__device__ __forceinline__ DJ operator*(const DJ &x, const DJ &y) { DJ r; r.a = x.a * y.a; for (int i = 0; i < 6; i++) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
__device__ __forceinline__ DJ operator/(const DJ &x, const DJ &y) { DJ r; const double inv = 1.0 / y.a; r.a = x.a * inv; for (int i = 0; i < 6; i++) r.v[i] = (x.v[i] - r.a * y.v[i]) * inv; return r; }
__device__ __forceinline__ DJ operator*(const DJ &x, double s) { DJ r; r.a = x.a * s; for (int i = 0; i < 6; i++) r.v[i] = x.v[i] * s; return r; }
__device__ __forceinline__ DJ operator*(double s, const DJ &x) { return x * s; }
__device__ __forceinline__ DJ operator+(const DJ &x, double s) { DJ r = x; r.a += s; return r; }
__device__ __forceinline__ DJ operator-(const DJ &x, double s) { DJ r = x; r.a -= s; return r; }
__device__ __forceinline__ DJ jsqrt(const DJ &x) { DJ r; r.a = sqrt(x.a); const double d = 1.0 / (2.0 * r.a); for (int i = 0; i < 6; i++) r.v[i] = x.v[i] * d; return r; }
__device__ __forceinline__ DJ jsin(const DJ &x) { DJ r; r.a = sin(x.a); const double c = cos(x.a); for (int i = 0; i < 6; i++) r.v[i] = c * x.v[i]; return r; }
__device__ __forceinline__ DJ jcos(const DJ &x) { DJ r; r.a = cos(x.a); const double s = -sin(x.a); for (int i = 0; i < 6; i++) r.v[i] = s * x.v[i]; return r; }

// ceres::AngleAxisRotatePoint on dual numbers
__device__ inline void rot(const DJ w[3], const DJ p[3], DJ out[3]) {
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

__device__ inline void f3d3d(const double *k, const DJ *x, DJ *r) {           // costfunctions.h:77-86
    DJ p[3] = { dj(k[0]), dj(k[1]), dj(k[2]) }, m[3];
    rot(x, p, m);
    r[0] = m[0] + x[3] - k[3]; r[1] = m[1] + x[4] - k[4]; r[2] = m[2] + x[5] - k[5];
}
__device__ inline void f3d2d(const double *k, const DJ *x, DJ *r) {           // costfunctions.h:111-126
    DJ p[3] = { dj(k[0]), dj(k[1]), dj(k[2]) }, m[3];
    rot(x, p, m);
    m[0] = m[0] + (x[3] + k[5]); m[1] = m[1] + (x[4] + k[6]); m[2] = m[2] + (x[5] + k[7]);
    r[0] = m[0] - k[3] * m[2]; r[1] = m[1] - k[4] * m[2];
}
__device__ inline void f2d3d(const double *k, const DJ *x, DJ *r) {           // costfunctions.h:151-168
    DJ w[3] = { -x[0], -x[1], -x[2] };
    DJ p[3] = { dj(k[0]) - x[3], dj(k[1]) - x[4], dj(k[2]) - x[5] }, m[3];
    rot(w, p, m);
    m[0] = m[0] + k[5]; m[1] = m[1] + k[6]; m[2] = m[2] + k[7];
    r[0] = m[0] - k[3] * m[2]; r[1] = m[1] - k[4] * m[2];
}
__device__ inline void f2d2d(const double *k, const DJ *x, DJ *r) {           // costfunctions.h:192-216
    DJ p[3] = { dj(k[0]), dj(k[1]), dj(1.0) }, m[3];
    rot(x, p, m);
    DJ b[3] = { dj(k[4]), dj(k[5]), dj(k[6]) }, tt[3];
    rot(x, b, tt);
    DJ tx = -tt[0] + x[3] + k[4], ty = -tt[1] + x[4] + k[5], tz = -tt[2] + x[5] + k[6];
    const DJ tn = jsqrt(tx * tx + ty * ty + tz * tz);
    tx = tx / tn; ty = ty / tn; tz = tz / tn;
    const double sx = k[2], sy = k[3];
    r[0] = m[0] * ((-sy) * tz + ty) + m[1] * (sx * tz - tx) + m[2] * ((-sx) * ty + sy * tx);
}


__device__ inline void f3dpd(const double *k, const DJ *x, DJ *r) {      // costfunctions.h:40-53, k = {p[3], n[3], o[3]}
    DJ p[3] = { dj(k[0]), dj(k[1]), dj(k[2]) }, m[3];
    rot(x, p, m);
    m[0] = m[0] + (x[3] - k[6]); m[1] = m[1] + (x[4] - k[7]); m[2] = m[2] + (x[5] - k[8]);
    r[0] = m[0] * k[3] + m[1] * k[4] + m[2] * k[5];
}
