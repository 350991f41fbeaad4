// velo_common.cuh — small device helpers shared by the kernel translation units.
#pragma once
#include "velo_dev.cuh"
#include <math_constants.h>

#define FULL 0xffffffffu
typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ float d2f(float ax, float ay, float az, float bx, float by, float bz) {
    // subtract_assign + norm2 (utility.h:35-39,51-53) == flann::L2_Simple<float>: ((dx*dx)+dy*dy)+dz*dz
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, v, o); if (lane >= o) v += t; }
    return v;
}
// exclusive scan over a block of up to 1024 threads; s_w: >= 33 ints of shared memory
__device__ __forceinline__ int block_excl_scan(int v, int *s_w, int &total) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int inc = warp_incl_scan(v, lane);
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int t = lane < nw ? s_w[lane] : 0;
        int ti = warp_incl_scan(t, lane);
        s_w[lane] = ti - t;
        if (lane == 31) s_w[32] = ti;
    }
    __syncthreads();
    int r = s_w[wid] + inc - v;
    total = s_w[32];
    __syncthreads();
    return r;
}
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__device__ __forceinline__ void idx_frame(const DevCalib &cal, float x, float y, float z, float &vx, float &vy, float &vz) {
    float dx = x - cal.vtc[3], dy = y - cal.vtc[7], dz = z - cal.vtc[11];
    vx = cal.vtc[0] * dx + cal.vtc[4] * dy + cal.vtc[8] * dz;
    vy = cal.vtc[1] * dx + cal.vtc[5] * dy + cal.vtc[9] * dz;
    vz = cal.vtc[2] * dx + cal.vtc[6] * dy + cal.vtc[10] * dz;
}
__device__ __forceinline__ int el_bucket(float el) {   // monotone non-decreasing in el (required for conservative masks)
    int b = (int)floorf((el - VELO_EL_MIN) * (VELO_EL_BUCKETS / (VELO_EL_MAX - VELO_EL_MIN)));
    return min(max(b, 0), VELO_EL_BUCKETS - 1);
}
__device__ __forceinline__ int rg_bucket(float rho) {   // monotone non-decreasing in rho
    int b = (int)floorf(log2f(fmaxf(rho, VELO_RG_MIN) * (1.0f / VELO_RG_MIN)) * VELO_RG_PER_OCTAVE);
    return min(max(b, 0), VELO_RG_BUCKETS - 1);
}
__device__ __forceinline__ int az_bin(float az) {
    int b = (int)((az + CUDART_PI_F) * (VELO_AZ_BINS / (2.0f * CUDART_PI_F)));
    return min(max(b, 0), VELO_AZ_BINS - 1);
}


// ------------------------------------------------------------------------------------------------ a13: normal equations
// Row staging for the warp-cooperative accumulation of  H += rho' J^T J,  g += rho' J^T r,  cost += rho/2
// (what ceres::Solve forms for the single 6-vector block, velo.h:897-902).
// Each lane deposits one residual ROW {J[6], r, rho', rho/2 (only on the first row of a block)}; lane l < 28 then
// owns one of the 28 sums and walks the deposited rows in lane order: fixed order => run-to-run deterministic.
#define NEQ_ROW 9
static __constant__ int c_pa[28] = { 0,0,0,0,0,0, 1,1,1,1,1, 2,2,2,2, 3,3,3, 4,4, 5,  0,1,2,3,4,5, 6 };
static __constant__ int c_pb[28] = { 0,1,2,3,4,5, 1,2,3,4,5, 2,3,4,5, 3,4,5, 4,5, 5,  6,6,6,6,6,6, 6 };

__device__ __forceinline__ void warp_accum(double *s_rows, const double J[6], double r, double rho1, double rho0h, bool valid,
                                           int lane, double &acc, double &raw) {
    __syncwarp();
    unsigned mask = __ballot_sync(FULL, valid);
    if (valid) {
        double *row = s_rows + lane * NEQ_ROW;
#pragma unroll
        for (int i = 0; i < 6; i++) row[i] = J[i];
        row[6] = r; row[7] = rho1; row[8] = rho0h;
    }
    __syncwarp();
    if (lane < 28) {
        const int a = c_pa[lane], b = c_pb[lane];
        for (unsigned m = mask; m; m &= m - 1) {
            const double *rq = s_rows + (__ffs(m) - 1) * NEQ_ROW;
            const double p = rq[a] * rq[b];
            if (lane == 27) { raw = fma(0.5, p, raw); acc += rq[8]; }     // explicit fma: the sums are not parity-critical (1e-4)
            else { raw += p; acc = fma(rq[7], p, acc); }
        }
    }
}

// CTA-level finish: lanes' (acc, raw) of every warp -> partial[0..55]; fixed warp order.  s_red: [warps][56]
__device__ __forceinline__ void block_neq_finish(double *s_red, double acc, double raw, double *partial) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane < 28) { s_red[wid * 56 + lane] = acc; s_red[wid * 56 + 28 + lane] = raw; }
    __syncthreads();
    if (threadIdx.x < 56) {
        double s = 0.0;
        for (int w = 0; w < nw; w++) s += s_red[w * 56 + threadIdx.x];
        partial[threadIdx.x] = s;
    }
}
