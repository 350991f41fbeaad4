// velo_solve.cu — device-resident replacement of the per-pass ceres::Solve of frameToFrame (velo.h:897-902):
// the residual blocks are frozen (the visual selection of the current f2f iteration + the ICP correspondences of the current
// ICP pass, exactly what AddResidualBlock had put into the problem), every Levenberg-Marquardt iterate re-evaluates them at
// the trial pose on the GPU (k_icp_eval_fixed + k_visual in fixed mode), and a one-thread controller kernel does the
// 6x6 damped solve, the step acceptance and the trust-region update in device memory, so a whole solve is a stream of
// launches with no host round trip until its result is read.  Every kernel works on n frame pairs at once (unit u <-> LmState[u]):
// one controller thread per pair, converged pairs drop out of the evaluation launches.
//
// The trust-region policy restates Ceres' documented defaults [recall; Ceres is not in /root/reference, parity for this row is
// "to solver tolerance" and defined by the CPU checker's identical restatement]: Levenberg-Marquardt,
// initial_trust_region_radius 1e4, min/max_lm_diagonal 1e-6/1e32, min_relative_decrease 1e-3,
// radius /= max(1/3, 1-(2 rho-1)^3) on success, radius /= 2,4,8.. on failure, function/gradient/parameter tolerances
// 1e-6 / 1e-10 / 1e-8, max_num_iterations 50.
#include "velo_jet.cuh"
#include <algorithm>

// ------------------------------------------------------------------------------------------------ fixed 3DPD blocks
// one thread per frozen record of the unit's last correspondence pass (kept == 1 => a cost3DPD block with p, n, o frozen); grid = (ctas, n)
__global__ void __launch_bounds__(256) k_icp_eval_fixed(DevBuffers B, const IcpFrozen *__restrict__ frozen, int frozen_stride, const double *__restrict__ icp_out,
                                                        const IcpUnit *__restrict__ units, const LmState *__restrict__ lm,
                                                        double loss_a, double weight, double *__restrict__ partial) {
    __shared__ double s_rows[8][NEQ_STAGE];
    __shared__ double s_red[8 * 56];
    __shared__ int s_cnt;
    __shared__ RotPack s_rp;
    __shared__ double s_pose[6];
    const int u = blockIdx.y;
    if (lm[u].done) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_cnt = 0;
    if (tid < 6) s_pose[tid] = lm[u].xt[tid];
    __syncthreads();
    if (tid < 3) rotpack_column(s_pose, false, tid, &s_rp);          // rotation + derivative of the trial pose, once per CTA
    __syncthreads();
    const float4 *pts = B.pts + (size_t)units[u].src_slot * B.N;
    const IcpFrozen *fz = frozen + (size_t)u * frozen_stride;
    double cr0 = 0.0, cr1 = 0.0, cw0 = 0.0, cw1 = 0.0, cost_half = 0.0;
    int nk = 0;
    const int cap = min(frozen_stride, max(0, (int)icp_out[(size_t)u * VELO_NEQ_STRIDE + 58]));   // the records the correspondence pass wrote
    const int per = (((cap + gridDim.x - 1) / gridDim.x) + 31) & ~31;
    const int q0 = blockIdx.x * per, q1 = min(cap, q0 + per);
    for (int qb = q0; qb < q1; qb += blockDim.x) {              // warp-uniform trip count
        const int q = qb + tid;
        bool kept = false;
        double J[6] = { 0, 0, 0, 0, 0, 0 }, res = 0.0, rho1 = 0.0;
        if (q < q1) {
            const IcpFrozen c = fz[q];
            if (c.kept == 1) {
                const float4 p = pts[c.src];                    // written by k_icp_pass for this source slot: in range by construction
                const double k[9] = { p.x, p.y, p.z, c.n[0], c.n[1], c.n[2], c.o[0], c.o[1], c.o[2] };
                lin3dpd(k, s_rp, s_pose + 3, &res, J);
                const double bb = loss_a * loss_a, cc = 1.0 / bb, sum = 1.0 + res * res * cc, inv = 1.0 / sum;   // Scaled(Cauchy), velo.h:885-891
                rho1 = weight * fmax(2.2250738585072014e-308, inv);
                cost_half += 0.5 * weight * bb * log(sum);
                kept = true; nk++;
            }
        }
        if (__any_sync(FULL, kept)) neq_mma_rows(s_rows[wid], lane, J, res, rho1, kept, cr0, cr1, cw0, cw1);
    }
    for (int o = 16; o > 0; o >>= 1) nk += __shfl_down_sync(FULL, nk, o);
    if (lane == 0 && nk) atomicAdd(&s_cnt, nk);
    double *pout = partial + ((size_t)u * gridDim.x + blockIdx.x) * 64;
    block_neq_finish_mma(s_red, cr0, cr1, cw0, cw1, cost_half, pout);
    if (tid == 0) { pout[56] = (double)s_cnt; pout[57] = (double)s_cnt; pout[58] = 0.0; }
}

__global__ void k_eval_reduce(const double *__restrict__ partial, int ctas, double *__restrict__ out, const LmState *__restrict__ lm) {
    const int u = blockIdx.x, t = threadIdx.x;
    if (lm[u].done || t >= VELO_NEQ_STRIDE) return;
    double s = 0.0;
    if (t < 59) for (int c = 0; c < ctas; c++) s += partial[((size_t)u * ctas + c) * 64 + t];
    out[(size_t)u * VELO_NEQ_STRIDE + t] = s;
}

// ------------------------------------------------------------------------------------------------ LM controller
__device__ bool chol6_solve(const double *A /*21 upper*/, const double *b, double *xo) {
    double M[6][6], L[6][6];
    int o = 0;
    for (int i = 0; i < 6; i++) for (int j = i; j < 6; j++, o++) { M[i][j] = A[o]; M[j][i] = A[o]; }
    for (int i = 0; i < 6; i++) for (int j = 0; j <= i; j++) {
        double s = M[i][j];
        for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
        if (i == j) { if (!(s > 0.0)) return false; L[i][i] = sqrt(s); }
        else L[i][j] = s / L[j][j];
    }
    double y[6];
    for (int i = 0; i < 6; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[i][k] * y[k]; y[i] = s / L[i][i]; }
    for (int i = 5; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < 6; k++) s -= L[k][i] * xo[k]; xo[i] = s / L[i][i]; }
    return true;
}
__device__ void sym6_mul(const double *H, const double *v, double *o) {
    double M[6][6]; int k = 0;
    for (int i = 0; i < 6; i++) for (int j = i; j < 6; j++, k++) { M[i][j] = H[k]; M[j][i] = H[k]; }
    for (int i = 0; i < 6; i++) { double s = 0.0; for (int j = 0; j < 6; j++) s += M[i][j] * v[j]; o[i] = s; }
}
// propose the next trial step from (H, g, radius); returns false when the damped system is not usable
__device__ bool lm_propose(LmState *S) {
    double A[21], mg[6];
    int o = 0;
    for (int i = 0; i < 6; i++) for (int j = i; j < 6; j++, o++) {
        A[o] = S->H[o];
        if (i == j) A[o] += fmin(fmax(S->H[o], 1e-6), 1e32) / S->radius;           // LevenbergMarquardtStrategy: diag clamp / radius
    }
    for (int i = 0; i < 6; i++) mg[i] = -S->g[i];
    if (!chol6_solve(A, mg, S->delta)) return false;
    double Hd[6]; sym6_mul(S->H, S->delta, Hd);
    double mc = 0.0;
    for (int i = 0; i < 6; i++) mc -= S->delta[i] * (S->g[i] + 0.5 * Hd[i]);       // model_cost_change
    S->model_change = mc;
    if (!(mc > 0.0)) return false;
    double nd = 0.0, nx = 0.0;
    for (int i = 0; i < 6; i++) { nd += S->delta[i] * S->delta[i]; nx += S->x[i] * S->x[i]; S->xt[i] = S->x[i] + S->delta[i]; }
    if (sqrt(nd) <= S->parameter_tolerance * (sqrt(nx) + S->parameter_tolerance)) { S->done = 1; S->reason = 3; }
    return true;
}

// poses == nullptr: keep the accepted pose of the previous solve and restart the controller (the next frozen block list starts there)
__global__ void k_lm_init(LmState *Sall, int n, const double *poses, int max_iterations, int *n_done) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u == 0) *n_done = 0;
    if (u >= n) return;
    LmState *S = Sall + u;
    for (int i = 0; i < 6; i++) { if (poses) S->x[i] = poses[6 * (size_t)u + i]; S->xt[i] = S->x[i]; S->delta[i] = 0.0; }
    S->radius = 1e4; S->decrease_factor = 2.0; S->cost = 0.0; S->init_cost = 0.0; S->model_change = 0.0;
    S->function_tolerance = 1e-6; S->gradient_tolerance = 1e-10; S->parameter_tolerance = 1e-8;
    S->iter = 0; S->done = 0; S->phase = 0; S->reason = 0; S->accepted = 0; S->max_iterations = max_iterations; S->n_blocks = 0;
}

// one controller step per unit: consumes the evaluation at the trial pose (sum of the ICP and the visual normal equations)
__global__ void k_lm_step(LmState *Sall, int n, const double *__restrict__ e_icp_all, const double *__restrict__ e_vis_all, int *n_done) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    LmState *S = Sall + u;
    if (S->done) return;
    const double *e_icp = e_icp_all ? e_icp_all + (size_t)u * VELO_NEQ_STRIDE : nullptr, *e_vis = e_vis_all ? e_vis_all + (size_t)u * VELO_NEQ_STRIDE : nullptr;
    double H[21], g[6];
    for (int i = 0; i < 21; i++) H[i] = (e_icp ? e_icp[i] : 0.0) + (e_vis ? e_vis[i] : 0.0);
    for (int i = 0; i < 6; i++) g[i] = (e_icp ? e_icp[21 + i] : 0.0) + (e_vis ? e_vis[21 + i] : 0.0);
    const double cost = (e_icp ? e_icp[27] : 0.0) + (e_vis ? e_vis[27] : 0.0);
    bool take = false;
    if (S->phase == 0) {                                   // evaluation at the starting point
        S->phase = 1; S->init_cost = cost; S->n_blocks = (int)((e_icp ? e_icp[56] : 0.0) + (e_vis ? e_vis[56] : 0.0));
        take = true;
        if (S->n_blocks == 0) { S->cost = cost; S->done = 1; S->reason = 5; atomicAdd(n_done, 1); return; }
    } else {
        S->iter++;
        const double rho = (S->cost - cost) / S->model_change;
        if (rho > 1e-3) {                                  // successful step
            take = true; S->accepted++;
            for (int i = 0; i < 6; i++) S->x[i] = S->xt[i];
            if (fabs(S->cost - cost) < S->function_tolerance * S->cost) { S->done = 1; S->reason = 1; }
            const double t = 2.0 * rho - 1.0;
            S->radius = fmin(1e16, S->radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
            S->decrease_factor = 2.0;
        } else {                                           // rejected: shrink the trust region
            S->radius /= S->decrease_factor; S->decrease_factor *= 2.0;
            if (S->radius < 1e-32) { S->done = 1; S->reason = 4; }
        }
    }
    if (take) {
        S->cost = cost;
        for (int i = 0; i < 21; i++) S->H[i] = H[i];
        double gmax = 0.0;
        for (int i = 0; i < 6; i++) { S->g[i] = g[i]; gmax = fmax(gmax, fabs(g[i])); }
        if (gmax <= S->gradient_tolerance) { S->done = 1; S->reason = 2; }
    }
    if (!S->done && S->iter >= S->max_iterations) { S->done = 1; S->reason = 6; }
    // next trial point (retry with smaller radii while the damped system is unusable)
    while (!S->done && !lm_propose(S)) {
        S->radius /= S->decrease_factor; S->decrease_factor *= 2.0;
        if (S->radius < 1e-32) { S->done = 1; S->reason = 4; }
    }
    if (S->done) atomicAdd(n_done, 1);
}

void launch_lm_init(const Launcher &L, LmState *S, int n, const double *d_poses, int max_iterations, int *n_done) {
    if (L.pre) L.pre(L.user, VK_SOLVE);
    k_lm_init<<<(n + 63) / 64, 64, 0, L.stream>>>(S, n, d_poses, max_iterations, n_done);
    if (L.post) L.post(L.user, VK_SOLVE);
}
void launch_icp_eval(const Launcher &L, const DevBuffers &B, int n, const IcpFrozen *frozen, int frozen_stride, const double *icp_out, const IcpUnit *units,
                     const LmState *S, double loss_a, double weight, double *partial, int ctas, double *out) {
    if (L.pre) L.pre(L.user, VK_SOLVE);
    k_icp_eval_fixed<<<dim3(ctas, n), 256, 0, L.stream>>>(B, frozen, frozen_stride, icp_out, units, S, loss_a, weight, partial);
    k_eval_reduce<<<n, 64, 0, L.stream>>>(partial, ctas, out, S);
    if (L.post) L.post(L.user, VK_SOLVE);
}
void launch_lm_step(const Launcher &L, LmState *S, int n, const double *e_icp, const double *e_vis, int *n_done) {
    if (L.pre) L.pre(L.user, VK_SOLVE);
    k_lm_step<<<(n + 63) / 64, 64, 0, L.stream>>>(S, n, e_icp, e_vis, n_done);
    if (L.post) L.post(L.user, VK_SOLVE);
}

// ------------------------------------------------------------------------------------------------ f4: Hamming matcher
// velo.h:517-531: for every query descriptor the train descriptor with the smallest Hamming distance (ties -> lower index).
// One thread per query, descriptor in registers; the train descriptors stream through shared memory in tiles that every
// thread of the CTA reads as broadcasts.  The train set is additionally cut into gridDim.y slices so that a few thousand queries
// fill the machine (3000 queries alone are 24 CTAs on 148 SMs); a slice's winner goes into best[query] as the 64-bit key
// (distance << 32 | index) with atomicMin — the minimum of the keys is the smallest distance and, among equals, the lowest
// index, whatever the order of the slices.  Integer work, bit-exact.  best[] must be preset to all ones.
#define HAM_THREADS 128
#define HAM_TILE 128
__global__ void __launch_bounds__(HAM_THREADS) k_hamming_nn(const unsigned long long *__restrict__ q, int nq, const unsigned long long *__restrict__ t, int nt,
                                                            int words, unsigned long long *__restrict__ best) {
    __shared__ unsigned long long s_t[HAM_TILE * 8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = (((nt + gridDim.y - 1) / gridDim.y) + HAM_TILE - 1) / HAM_TILE * HAM_TILE;
    const int j0 = blockIdx.y * per, j1 = min(nt, j0 + per);
    unsigned long long d[8];
#pragma unroll
    for (int w = 0; w < 8; w++) d[w] = (i < nq && w < words) ? q[(size_t)i * words + w] : 0ull;
    int bi = -1, bd = 0x7fffffff;
    for (int t0 = j0; t0 < j1; t0 += HAM_TILE) {
        const int n = min(HAM_TILE, j1 - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < n * words; k += blockDim.x) s_t[(k / words) * 8 + (k % words)] = t[(size_t)t0 * words + k];
        __syncthreads();
        for (int j = 0; j < n; j++) {
            int dist = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) if (w < words) dist += __popcll(d[w] ^ s_t[j * 8 + w]);
            if (dist < bd) { bd = dist; bi = t0 + j; }
        }
    }
    if (i < nq && bi >= 0) atomicMin(&best[i], ((unsigned long long)(unsigned)bd << 32) | (unsigned)bi);
}
void launch_hamming(const Launcher &L, const unsigned long long *q, int nq, const unsigned long long *t, int nt, int words, unsigned long long *best, int sm_count) {
    if (nq <= 0) return;
    const int gx = (nq + HAM_THREADS - 1) / HAM_THREADS;
    int gy = (2 * sm_count + gx - 1) / gx;                       // about two CTAs per SM in total
    gy = std::max(1, std::min(gy, (nt + HAM_TILE - 1) / HAM_TILE));
    cudaMemsetAsync(best, 0xFF, (size_t)nq * sizeof(unsigned long long), L.stream);
    if (L.pre) L.pre(L.user, VK_SOLVE);
    k_hamming_nn<<<dim3(gx, gy), HAM_THREADS, 0, L.stream>>>(q, nq, t, nt, words, best);
    if (L.post) L.post(L.user, VK_SOLVE);
}
