// velo_visual.cu — stage 4 for the camera terms: the per-match residual selection / outlier gating of
// frameToFrame (velo.h:622-792) with cost3D3D / cost3D2D / cost2D3D / cost2D2D (costfunctions.h:60-220) evaluated
// by 6-partial forward-mode dual numbers (what ceres::AutoDiffCostFunction does), the Arctan / Scaled losses
// (velo.h:688,714-717,748-751,781-784) and the per-(frame, iteration) normal equations.
#include "velo_common.cuh"
#include "velo_jet.cuh"

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
                                                        double *__restrict__ partial, VisMatchOut *__restrict__ mout, VisFixed fx, int *__restrict__ bad_flag) {
    __shared__ double s_rows[VIS_THREADS / 32][NEQ_STAGE];
    __shared__ double s_red[(VIS_THREADS / 32) * 56];
    __shared__ int s_cnt[2];
    const VisUnit U = units[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int C = cal.num_cams, MM = B.MM, iter = U.iter;
    if (fx.done && *fx.done) return;                 // device-side solver already converged: nothing to evaluate
    if (tid == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
    __syncthreads();
    int nm[VELO_MAX_CAMS], tot = 0;
    for (int c = 0; c < VELO_MAX_CAMS; c++) { nm[c] = c < C ? B.n_matches[(size_t)U.slot1 * B.C + c] : 0; tot += nm[c]; }
    DJ x[6];
    for (int i = 0; i < 6; i++) { x[i] = dj(fx.pose ? fx.pose[i] : U.pose[i]); x[i].v[i] = 1.0; }
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
            // caller-supplied indices: a pair that points outside its keypoint set adds no block and raises the error flag
            const bool in_range = (unsigned)p1 < (unsigned)B.n_kp[s1] && (unsigned)p2 < (unsigned)B.n_kp[s2];
            if (!in_range && bad_flag) atomicOr(bad_flag, 1);
            if (in_range) {
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
            // Free mode: the reference's selection + iter>1 outlier gates (velo.h:662-789); the chosen types are optionally
            // recorded as a bit mask (1 = 3D3D, 2 = 2D2D, 4 = 3D2D, 8 = 2D3D).  Fixed mode (fx.sel_in): the block list was
            // frozen by an earlier call (what ceres::Solve sees: AddResidualBlock happened before), so no gate is applied.
            const unsigned fixed = fx.sel_in ? (unsigned)fx.sel_in[lmi] | 0x100u : 0u;
            unsigned chosen = 0;
            bool go = true;
            if (fixed ? (fixed & 1u) : (d1 && d2)) {                               // velo.h:662-693
                const double k[6] = { q1.x, q1.y, q1.z, q2.x, q2.y, q2.z };
                f3d3d(k, x, r);
                const double s = r[0].a * r[0].a + r[1].a * r[1].a + r[2].a * r[2].a;
                const double lim = tn.l3d3d * tn.outlier / iter * tn.l3d3d * tn.outlier / iter;
                if (!fixed && iter > 1 && s > lim) go = false;
                else { take(blk[nb], VELO_RES_3D3D, 3, r); loss_arctan(tn.l3d3d, 1.0, s, blk[nb].rho0, blk[nb].rho1); nb++; chosen |= 1u; }
            }
            if (fixed ? (fixed & 2u) : (go && !d1 && !d2 && tn.en2d2d)) {          // velo.h:694-722
                const double k[7] = { u1.x, u1.y, u2.x, u2.y, t0, t1, t2 };
                f2d2d(k, x, r);
                const double av = tn.abs_trunc ? (double)abs((int)r[0].a) : fabs(r[0].a);   // hazard H1 (velo.h:709)
                if (!fixed && iter > 1 && av > tn.l2d2d * tn.outlier / iter) go = false;
                else { take(blk[nb], VELO_RES_2D2D, 1, r); loss_arctan(tn.l2d2d, tn.w2d2d, r[0].a * r[0].a, blk[nb].rho0, blk[nb].rho1); nb++; chosen |= 2u; }
            }
            if (fixed || (go && tn.en3d2d)) {
                const double lim = tn.l3d2d * tn.outlier / iter * tn.l3d2d * tn.outlier / iter;
                if (fixed ? (fixed & 4u) : d1) {                                   // velo.h:724-756
                    const double k[8] = { q1.x, q1.y, q1.z, u2.x, u2.y, t0, t1, t2 };
                    f3d2d(k, x, r);
                    const double s = r[0].a * r[0].a + r[1].a * r[1].a;
                    if (!fixed && iter > 1 && s > lim) go = false;
                    else { take(blk[nb], VELO_RES_3D2D, 2, r); loss_arctan(tn.l3d2d, tn.w3d2d, s, blk[nb].rho0, blk[nb].rho1); nb++; chosen |= 4u; }
                }
                if (fixed ? (fixed & 8u) : (go && d2)) {                           // velo.h:757-789
                    const double k[8] = { q2.x, q2.y, q2.z, u1.x, u1.y, t0, t1, t2 };
                    f2d3d(k, x, r);
                    const double s = r[0].a * r[0].a + r[1].a * r[1].a;
                    if (!fixed && iter > 1 && s > lim) go = false;
                    else { take(blk[nb], VELO_RES_2D3D, 2, r); loss_arctan(tn.l3d2d, tn.w3d2d, s, blk[nb].rho0, blk[nb].rho1); nb++; chosen |= 8u; }
                }
            }
            if (fx.sel_out) fx.sel_out[lmi] = (unsigned char)chosen;
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
                   const int *lm_valid, const float4 *lm_xyz, double *partial, double *out, VisMatchOut *match_out, int ctas, VisFixed fx, int *bad_flag) {
    if (n_units <= 0) return;
    dim3 g(ctas, n_units);
    if (L.pre) L.pre(L.user, VK_VISUAL);
    k_visual<<<g, VIS_THREADS, 0, L.stream>>>(B, cal, units, tun, lm_valid, lm_xyz, partial, match_out, fx, bad_flag);
    if (L.post) L.post(L.user, VK_VISUAL);
    if (L.pre) L.pre(L.user, VK_NEQ_REDUCE);
    k_neq_reduce_vis<<<n_units, 64, 0, L.stream>>>(partial, out, ctas);
    if (L.post) L.post(L.user, VK_NEQ_REDUCE);
}
