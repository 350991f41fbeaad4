// velo_visual.cu — stage 4 for the camera terms: the per-match residual selection / outlier gating of
// frameToFrame (velo.h:622-792) with cost3D3D / cost3D2D / cost2D3D / cost2D2D (costfunctions.h:60-220) evaluated
// by 6-partial forward-mode dual numbers (what ceres::AutoDiffCostFunction does), the Arctan / Scaled losses
// (velo.h:688,714-717,748-751,781-784) and the per-(frame, iteration) normal equations.
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
__device__ void rot(const DJ w[3], const DJ p[3], DJ out[3]) {
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

__device__ void f3d3d(const double *k, const DJ *x, DJ *r) {           // costfunctions.h:77-86
    DJ p[3] = { dj(k[0]), dj(k[1]), dj(k[2]) }, m[3];
    rot(x, p, m);
    r[0] = m[0] + x[3] - k[3]; r[1] = m[1] + x[4] - k[4]; r[2] = m[2] + x[5] - k[5];
}
__device__ void f3d2d(const double *k, const DJ *x, DJ *r) {           // costfunctions.h:111-126
    DJ p[3] = { dj(k[0]), dj(k[1]), dj(k[2]) }, m[3];
    rot(x, p, m);
    m[0] = m[0] + (x[3] + k[5]); m[1] = m[1] + (x[4] + k[6]); m[2] = m[2] + (x[5] + k[7]);
    r[0] = m[0] - k[3] * m[2]; r[1] = m[1] - k[4] * m[2];
}
__device__ void f2d3d(const double *k, const DJ *x, DJ *r) {           // costfunctions.h:151-168
    DJ w[3] = { -x[0], -x[1], -x[2] };
    DJ p[3] = { dj(k[0]) - x[3], dj(k[1]) - x[4], dj(k[2]) - x[5] }, m[3];
    rot(w, p, m);
    m[0] = m[0] + k[5]; m[1] = m[1] + k[6]; m[2] = m[2] + k[7];
    r[0] = m[0] - k[3] * m[2]; r[1] = m[1] - k[4] * m[2];
}
__device__ void f2d2d(const double *k, const DJ *x, DJ *r) {           // costfunctions.h:192-216
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

struct Blk { int type, nres; double r[3], J[18], rho0, rho1; };

__device__ __forceinline__ void loss_arctan(double a, double w, double s, double &rho0, double &rho1) {
    const double b = 1.0 / (a * a), sum = 1.0 + s * s * b, inv = 1.0 / sum;
    rho0 = w * a * atan2(s, a); rho1 = w * fmax(2.2250738585072014e-308, inv);
}
__device__ __forceinline__ void take(Blk &o, int type, int nres, const DJ *r) {
    o.type = type; o.nres = nres;
    for (int i = 0; i < nres; i++) { o.r[i] = r[i].a; for (int j = 0; j < 6; j++) o.J[6 * i + j] = r[i].v[j]; }
}

#define VIS_THREADS 128
// grid = (ctas, n_units); threads stride over the (camera, match) pairs of the unit
__global__ void __launch_bounds__(VIS_THREADS) k_visual(DevBuffers B, DevCalib cal, const VisUnit *__restrict__ units, VisTun tn,
                                                        const int *__restrict__ lm_valid, const float4 *__restrict__ lm_xyz,
                                                        double *__restrict__ partial, VisMatchOut *__restrict__ mout) {
    __shared__ double s_rows[VIS_THREADS / 32][32 * NEQ_ROW];
    __shared__ double s_red[(VIS_THREADS / 32) * 56];
    __shared__ int s_cnt[2];
    const VisUnit U = units[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int C = cal.num_cams, MM = B.MM, iter = U.iter;
    if (tid == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
    __syncthreads();
    int nm[VELO_MAX_CAMS], tot = 0;
    for (int c = 0; c < VELO_MAX_CAMS; c++) { nm[c] = c < C ? B.n_matches[(size_t)U.slot1 * B.C + c] : 0; tot += nm[c]; }
    DJ x[6];
    for (int i = 0; i < 6; i++) { x[i] = dj(U.pose[i]); x[i].v[i] = 1.0; }
    double acc = 0.0, raw = 0.0;
    int nblk = 0, nres = 0;
    const int per = (((tot + gridDim.x - 1) / gridDim.x) + 31) & ~31;
    const int e0 = blockIdx.x * per, e1 = min(tot, e0 + per);
    for (int eb = e0; eb < e1; eb += VIS_THREADS) {
        const int e = eb + tid;
        Blk blk[3]; int nb = 0;
        if (e < e1) {
            int cam = 0, i = e;
            while (i >= nm[cam]) { i -= nm[cam]; cam++; }
            const size_t s1 = ((size_t)U.slot1 * VELO_NUM_KP_SETS + U.set1) * B.C + cam, s2 = ((size_t)U.slot2 * VELO_NUM_KP_SETS + U.set2) * B.C + cam;
            const int *mt = B.matches + 2 * (((size_t)U.slot1 * B.C + cam) * MM + i);
            const int p1 = mt[0], p2 = mt[1];
            const int h1 = B.has_depth[s1 * B.F + p1], h2 = B.has_depth[s2 * B.F + p2];
            bool d1 = h1 != -1, d2 = h2 != -1;                                     // velo.h:631-632
            float4 q1 = make_float4(0, 0, 0, 0), q2 = make_float4(0, 0, 0, 0);
            const size_t lmi = (size_t)cam * MM + i;
            if (lm_valid && lm_valid[lmi]) { q2 = lm_xyz[lmi]; d2 = true; }        // velo.h:634-644
            else if (d2) q2 = B.kpwd[s2 * B.F + h2];
            if (d1) q1 = B.kpwd[s1 * B.F + h1];
            const float2 u1 = B.kp[s1 * B.F + p1], u2 = B.kp[s2 * B.F + p2];
            const double t0 = cal.cam_t[cam][0], t1 = cal.cam_t[cam][1], t2 = cal.cam_t[cam][2];
            DJ r[3];
            bool go = true;
            if (d1 && d2) {                                                        // velo.h:662-693
                const double k[6] = { q1.x, q1.y, q1.z, q2.x, q2.y, q2.z };
                f3d3d(k, x, r);
                const double s = r[0].a * r[0].a + r[1].a * r[1].a + r[2].a * r[2].a;
                const double lim = tn.l3d3d * tn.outlier / iter * tn.l3d3d * tn.outlier / iter;
                if (iter > 1 && s > lim) go = false;
                else { take(blk[nb], VELO_RES_3D3D, 3, r); loss_arctan(tn.l3d3d, 1.0, s, blk[nb].rho0, blk[nb].rho1); nb++; }
            }
            if (go && !d1 && !d2 && tn.en2d2d) {                                   // velo.h:694-722
                const double k[7] = { u1.x, u1.y, u2.x, u2.y, t0, t1, t2 };
                f2d2d(k, x, r);
                const double av = tn.abs_trunc ? (double)abs((int)r[0].a) : fabs(r[0].a);   // hazard H1 (velo.h:709)
                if (iter > 1 && av > tn.l2d2d * tn.outlier / iter) go = false;
                else { take(blk[nb], VELO_RES_2D2D, 1, r); loss_arctan(tn.l2d2d, tn.w2d2d, r[0].a * r[0].a, blk[nb].rho0, blk[nb].rho1); nb++; }
            }
            if (go && tn.en3d2d) {
                const double lim = tn.l3d2d * tn.outlier / iter * tn.l3d2d * tn.outlier / iter;
                if (d1) {                                                          // velo.h:724-756
                    const double k[8] = { q1.x, q1.y, q1.z, u2.x, u2.y, t0, t1, t2 };
                    f3d2d(k, x, r);
                    const double s = r[0].a * r[0].a + r[1].a * r[1].a;
                    if (iter > 1 && s > lim) go = false;
                    else { take(blk[nb], VELO_RES_3D2D, 2, r); loss_arctan(tn.l3d2d, tn.w3d2d, s, blk[nb].rho0, blk[nb].rho1); nb++; }
                }
                if (go && d2) {                                                    // velo.h:757-789
                    const double k[8] = { q2.x, q2.y, q2.z, u1.x, u1.y, t0, t1, t2 };
                    f2d3d(k, x, r);
                    const double s = r[0].a * r[0].a + r[1].a * r[1].a;
                    if (iter > 1 && s > lim) go = false;
                    else { take(blk[nb], VELO_RES_2D3D, 2, r); loss_arctan(tn.l3d2d, tn.w3d2d, s, blk[nb].rho0, blk[nb].rho1); nb++; }
                }
            }
            if (mout) {
                VisMatchOut &o = mout[(size_t)cam * MM + i];
                o.n = nb;
                for (int b = 0; b < nb; b++) {
                    velo_vis_block &vb = o.b[b];
                    vb.cam = cam; vb.match = i; vb.type = blk[b].type; vb.n_res = blk[b].nres;
                    for (int q = 0; q < 3; q++) vb.residual[q] = q < blk[b].nres ? blk[b].r[q] : 0.0;
                    for (int q = 0; q < 18; q++) vb.jacobian[q] = q < 6 * blk[b].nres ? blk[b].J[q] : 0.0;
                }
            }
        }
        for (int b = 0; b < 3; b++) {
            for (int row = 0; row < 3; row++) {
                const bool v = (b < nb) && (row < blk[b].nres);
                warp_accum(s_rows[wid], blk[b].J + 6 * row, v ? blk[b].r[row] : 0.0, v ? blk[b].rho1 : 0.0,
                           (v && row == 0) ? 0.5 * blk[b].rho0 : 0.0, v, lane, acc, raw);
            }
            if (b < nb) { nblk++; nres += blk[b].nres; }
        }
    }
    for (int o = 16; o > 0; o >>= 1) { nblk += __shfl_down_sync(FULL, nblk, o); nres += __shfl_down_sync(FULL, nres, o); }
    if (lane == 0) { atomicAdd(&s_cnt[0], nblk); atomicAdd(&s_cnt[1], nres); }
    double *pout = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 64;
    block_neq_finish(s_red, acc, raw, pout);
    if (tid == 0) { pout[56] = (double)s_cnt[0]; pout[57] = (double)s_cnt[1]; pout[58] = (double)(e1 > e0 ? e1 - e0 : 0); }
}

__global__ void k_neq_reduce_vis(const double *__restrict__ partial, double *__restrict__ out, int ctas) {
    const int u = blockIdx.x, t = threadIdx.x;
    if (t >= VELO_NEQ_STRIDE) return;
    double s = 0.0;
    if (t < 59) for (int c = 0; c < ctas; c++) s += partial[((size_t)u * ctas + c) * 64 + t];
    out[(size_t)u * VELO_NEQ_STRIDE + t] = s;
}

void launch_visual(const Launcher &L, const DevBuffers &B, const DevCalib &cal, const VisUnit *units, int n_units, VisTun tun,
                   const int *lm_valid, const float4 *lm_xyz, double *partial, double *out, VisMatchOut *match_out, int ctas) {
    if (n_units <= 0) return;
    dim3 g(ctas, n_units);
    if (L.pre) L.pre(L.user, VK_VISUAL);
    k_visual<<<g, VIS_THREADS, 0, L.stream>>>(B, cal, units, tun, lm_valid, lm_xyz, partial, match_out);
    if (L.post) L.post(L.user, VK_VISUAL);
    if (L.pre) L.pre(L.user, VK_NEQ_REDUCE);
    k_neq_reduce_vis<<<n_units, 64, 0, L.stream>>>(partial, out, ctas);
    if (L.post) L.post(L.user, VK_NEQ_REDUCE);
}
