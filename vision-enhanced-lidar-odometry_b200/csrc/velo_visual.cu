// velo_visual.cu — stage 4 for the camera terms: the per-match residual selection / outlier gating of
// frameToFrame (velo.h:622-792) with cost3D3D / cost3D2D / cost2D3D / cost2D2D (costfunctions.h:60-220), the Arctan / Scaled
// losses (velo.h:688,714-717,748-751,781-784) and the per-(frame, iteration) normal equations.
//
// All blocks of a unit share one pose, so the rotation and its derivative are formed once per CTA (velo_functors.h: the same
// dual numbers ceres::AutoDiffCostFunction would push through every block, applied to the three unit vectors) and a block costs a
// 3x3 map instead of a sqrt / sin / cos autodiff chain.  A warp walks its 32 matches through the four residual types in the
// reference's order (3D3D, 2D2D, 3D2D, 2D3D — the gate of one type decides whether the later ones are evaluated, H11); after each
// type the rows of the 32 lanes go into the 6x6 sums as an FP64-MMA X^T W X (neq_mma_rows), one m8n8k4 chain per residual row.
#include "velo_common.cuh"
#include "velo_jet.cuh"

__device__ __forceinline__ void loss_arctan(double a, double w, double s, double &rho0, double &rho1) {
    const double b = 1.0 / (a * a), sum = 1.0 + s * s * b, inv = 1.0 / sum;
    rho0 = w * a * atan2(s, a); rho1 = w * fmax(2.2250738585072014e-308, inv);
}
__device__ __forceinline__ void emit_block(VisMatchOut *mo, int &nb, int cam, int match, int type, int nres, const double *r, const double *J) {
    if (mo) {
        velo_vis_block &vb = mo->b[nb];
        vb.cam = cam; vb.match = match; vb.type = type; vb.n_res = nres;
#pragma unroll
        for (int q = 0; q < 3; q++) vb.residual[q] = q < nres ? r[q] : 0.0;
#pragma unroll
        for (int q = 0; q < 18; q++) vb.jacobian[q] = q < 6 * nres ? J[q] : 0.0;
    }
    nb++;
}

#define VIS_THREADS 128
// grid = (ctas, n_units); threads stride over the (camera, match) pairs of the unit
#ifndef VIS_MIN_BLOCKS
#define VIS_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(VIS_THREADS, VIS_MIN_BLOCKS) k_visual(DevBuffers B, DevCalib cal, const VisUnit *__restrict__ units, VisTun tn,
                                                        const int *__restrict__ lm_valid, const float4 *__restrict__ lm_xyz,
                                                        double *__restrict__ partial, VisMatchOut *__restrict__ mout, VisFixed fx, int *__restrict__ bad_flag) {
    __shared__ double s_rows[VIS_THREADS / 32][NEQ_STAGE];
    __shared__ double s_red[(VIS_THREADS / 32) * 56];
    __shared__ int s_cnt[2];
    __shared__ RotPack s_rp[2];        // [0]: R(w), [1]: R(-w) with derivatives w.r.t. the pose's w (cost2D3D)
    __shared__ double s_pose[6];
    const VisUnit U = units[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int C = cal.num_cams, MM = B.MM, iter = U.iter;
    const LmState *lm = fx.lm ? fx.lm + blockIdx.y : nullptr;
    if (lm && !fx.use_accepted && lm->done) return;  // device-side solver of this unit already converged: nothing to evaluate
    if (tid == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
    if (tid < 6) s_pose[tid] = lm ? (fx.use_accepted ? lm->x[tid] : lm->xt[tid]) : U.pose[tid];
    const unsigned char *sel_in = fx.sel_in ? fx.sel_in + (size_t)blockIdx.y * fx.sel_stride : nullptr;
    unsigned char *sel_out = fx.sel_out ? fx.sel_out + (size_t)blockIdx.y * fx.sel_stride : nullptr;
    __syncthreads();
    if (tid < 6) rotpack_column(s_pose, tid >= 3, tid % 3, &s_rp[tid >= 3]);
    __syncthreads();
    const RotPack &P = s_rp[0], &Pinv = s_rp[1];
    const double *t = s_pose + 3;
    int nm[VELO_MAX_CAMS], tot = 0;
    for (int c = 0; c < VELO_MAX_CAMS; c++) { nm[c] = c < C ? B.n_matches[(size_t)U.slot1 * B.C + c] : 0; tot += nm[c]; }
    double cr0 = 0.0, cr1 = 0.0, cw0 = 0.0, cw1 = 0.0, cost_half = 0.0;
    int nblk = 0, nres = 0;
    double *S = s_rows[wid];
    const int per = (((tot + gridDim.x - 1) / gridDim.x) + 31) & ~31;
    const int e0 = blockIdx.x * per, e1 = min(tot, e0 + per);
    for (int eb = e0; eb < e1; eb += VIS_THREADS) {      // warp-uniform trip count: every lane takes part in the MMA steps
        const int e = eb + tid;
        bool ok = e < e1, d1 = false, d2 = false, go = true;
        int cam = 0, i = 0, nb = 0;
        unsigned fixed = 0u, chosen = 0u;
        float4 q1 = make_float4(0, 0, 0, 0), q2 = make_float4(0, 0, 0, 0);
        float2 u1 = make_float2(0, 0), u2 = make_float2(0, 0);
        size_t lmi = 0;
        if (ok) {
            i = e;
            while (i >= nm[cam]) { i -= nm[cam]; cam++; }
            const size_t s1 = ((size_t)U.slot1 * VELO_NUM_KP_SETS + U.set1) * B.C + cam, s2 = ((size_t)U.slot2 * VELO_NUM_KP_SETS + U.set2) * B.C + cam;
            const int *mt = B.matches + 2 * (((size_t)U.slot1 * B.C + cam) * MM + i);
            const int p1 = mt[0], p2 = mt[1];
            lmi = (size_t)cam * MM + i;
            // caller-supplied indices: a pair that points outside its keypoint set adds no block and raises the error flag
            ok = (unsigned)p1 < (unsigned)B.n_kp[s1] && (unsigned)p2 < (unsigned)B.n_kp[s2];
            if (!ok && bad_flag) atomicOr(bad_flag, 1);
            if (ok) {
                const int h1 = B.has_depth[s1 * B.F + p1], h2 = B.has_depth[s2 * B.F + p2];
                d1 = h1 != -1; d2 = h2 != -1;                                      // velo.h:631-632
                if (lm_valid && lm_valid[lmi]) { q2 = lm_xyz[lmi]; d2 = true; }    // velo.h:634-644
                else if (d2) q2 = B.kpwd[s2 * B.F + h2];
                if (d1) q1 = B.kpwd[s1 * B.F + h1];
                u1 = B.kp[s1 * B.F + p1]; u2 = B.kp[s2 * B.F + p2];
                // Free mode: the reference's selection + iter>1 outlier gates (velo.h:662-789); the chosen types are optionally
                // recorded as a bit mask (1 = 3D3D, 2 = 2D2D, 4 = 3D2D, 8 = 2D3D).  Fixed mode (fx.sel_in): the block list was
                // frozen by an earlier call (what ceres::Solve sees: AddResidualBlock happened before), so no gate is applied.
                fixed = sel_in ? (unsigned)sel_in[lmi] | 0x100u : 0u;
            }
        }
        VisMatchOut *mo = (mout && ok) ? &mout[lmi] : nullptr;
        const double t0 = cal.cam_t[cam][0], t1 = cal.cam_t[cam][1], t2 = cal.cam_t[cam][2];
        double r[3], J[18], rho0 = 0.0, rho1 = 0.0;
        bool emit;
        // ---- 3D3D, velo.h:662-693
        emit = false;
        if (ok && (fixed ? (fixed & 1u) != 0u : (d1 && d2))) {
            const double k[6] = { q1.x, q1.y, q1.z, q2.x, q2.y, q2.z };
            lin3d3d(k, P, t, r, J);
            const double s = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
            const double lim = tn.l3d3d * tn.outlier / iter * tn.l3d3d * tn.outlier / iter;
            if (!fixed && iter > 1 && s > lim) go = false;
            else { loss_arctan(tn.l3d3d, 1.0, s, rho0, rho1); emit = true; chosen |= 1u; emit_block(mo, nb, cam, i, VELO_RES_3D3D, 3, r, J); }
        }
        if (__any_sync(FULL, emit)) {
#pragma unroll
            for (int row = 0; row < 3; row++) neq_mma_rows(S, lane, J + 6 * row, r[row], rho1, emit, cr0, cr1, cw0, cw1);
        }
        if (emit) { cost_half += 0.5 * rho0; nblk++; nres += 3; }
        // ---- 2D2D, velo.h:694-722
        emit = false;
        if (ok && (fixed ? (fixed & 2u) != 0u : (go && !d1 && !d2 && tn.en2d2d))) {
            const double k[7] = { u1.x, u1.y, u2.x, u2.y, t0, t1, t2 };
            lin2d2d(k, P, t, r, J);
            const double av = tn.abs_trunc ? (double)abs((int)r[0]) : fabs(r[0]);   // hazard H1 (velo.h:709)
            if (!fixed && iter > 1 && av > tn.l2d2d * tn.outlier / iter) go = false;
            else { loss_arctan(tn.l2d2d, tn.w2d2d, r[0] * r[0], rho0, rho1); emit = true; chosen |= 2u; emit_block(mo, nb, cam, i, VELO_RES_2D2D, 1, r, J); }
        }
        if (__any_sync(FULL, emit)) neq_mma_rows(S, lane, J, r[0], rho1, emit, cr0, cr1, cw0, cw1);
        if (emit) { cost_half += 0.5 * rho0; nblk++; nres += 1; }
        const double lim2 = tn.l3d2d * tn.outlier / iter * tn.l3d2d * tn.outlier / iter;
        // ---- 3D2D, velo.h:724-756
        emit = false;
        if (ok && (fixed ? (fixed & 4u) != 0u : (go && tn.en3d2d && d1))) {
            const double k[8] = { q1.x, q1.y, q1.z, u2.x, u2.y, t0, t1, t2 };
            lin3d2d(k, P, t, r, J);
            const double s = r[0] * r[0] + r[1] * r[1];
            if (!fixed && iter > 1 && s > lim2) go = false;
            else { loss_arctan(tn.l3d2d, tn.w3d2d, s, rho0, rho1); emit = true; chosen |= 4u; emit_block(mo, nb, cam, i, VELO_RES_3D2D, 2, r, J); }
        }
        if (__any_sync(FULL, emit)) {
#pragma unroll
            for (int row = 0; row < 2; row++) neq_mma_rows(S, lane, J + 6 * row, r[row], rho1, emit, cr0, cr1, cw0, cw1);
        }
        if (emit) { cost_half += 0.5 * rho0; nblk++; nres += 2; }
        // ---- 2D3D, velo.h:757-789
        emit = false;
        if (ok && (fixed ? (fixed & 8u) != 0u : (go && tn.en3d2d && d2))) {
            const double k[8] = { q2.x, q2.y, q2.z, u1.x, u1.y, t0, t1, t2 };
            lin2d3d(k, Pinv, t, r, J);
            const double s = r[0] * r[0] + r[1] * r[1];
            if (!fixed && iter > 1 && s > lim2) go = false;
            else { loss_arctan(tn.l3d2d, tn.w3d2d, s, rho0, rho1); emit = true; chosen |= 8u; emit_block(mo, nb, cam, i, VELO_RES_2D3D, 2, r, J); }
        }
        if (__any_sync(FULL, emit)) {
#pragma unroll
            for (int row = 0; row < 2; row++) neq_mma_rows(S, lane, J + 6 * row, r[row], rho1, emit, cr0, cr1, cw0, cw1);
        }
        if (emit) { cost_half += 0.5 * rho0; nblk++; nres += 2; }
        if (ok && sel_out) sel_out[lmi] = (unsigned char)chosen;
        if (mo) mo->n = nb;
    }
    for (int o = 16; o > 0; o >>= 1) { nblk += __shfl_down_sync(FULL, nblk, o); nres += __shfl_down_sync(FULL, nres, o); }
    if (lane == 0) { atomicAdd(&s_cnt[0], nblk); atomicAdd(&s_cnt[1], nres); }
    double *pout = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 64;
    block_neq_finish_mma(s_red, cr0, cr1, cw0, cw1, cost_half, pout);
    if (tid == 0) { pout[56] = (double)s_cnt[0]; pout[57] = (double)s_cnt[1]; pout[58] = (double)(e1 > e0 ? e1 - e0 : 0); }
}

__global__ void k_neq_reduce_vis(const double *__restrict__ partial, double *__restrict__ out, int ctas, const LmState *__restrict__ lm, int use_accepted) {
    const int u = blockIdx.x, t = threadIdx.x;
    if (t >= VELO_NEQ_STRIDE) return;
    if (lm && !use_accepted && lm[u].done) return;   // the partials of a converged unit are stale; its last sums stay
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
    k_neq_reduce_vis<<<n_units, 64, 0, L.stream>>>(partial, out, ctas, fx.lm, fx.use_accepted);
    if (L.post) L.post(L.user, VK_NEQ_REDUCE);
}
