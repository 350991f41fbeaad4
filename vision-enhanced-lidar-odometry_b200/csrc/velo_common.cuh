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
// the same distance with x,y carried in one packed f32x2 register pair (sm_100 FADD2 / FMUL2: two IEEE-rounded f32 operations
// per issued instruction, results bit-identical to the scalar form); mxy = pack(bx, by)
__device__ __forceinline__ u64 pack2f(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float d2f_xy2(const float4 &a, u64 mxy, float bz) {
    u64 dxy, sq; float sx, sy;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dxy) : "l"(pack2f(a.x, a.y)), "l"(mxy));
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(sq) : "l"(dxy));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(sx), "=f"(sy) : "l"(sq));
    const float dz = __fsub_rn(SORTED_Z(a), bz);
    return __fadd_rn(__fadd_rn(sx, sy), __fmul_rn(dz, dz));
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
// atan2 for the *pruning* geometry of a query (azimuth / elevation in the index frame): odd minimax polynomial of degree 11 on
// [0,1] + octant fix-up, |error| < 2e-6 rad over all quadrants (checked against f64 atan2 on 2.4e7 points, tools/check_atan2.py);
// asin_ub() pads every window by 1e-5 rad for it.  About a third of the instructions of atan2f.  Index *construction* keeps atan2f.
__device__ __forceinline__ float atan2_q(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mn = fminf(ax, ay), mx = fmaxf(ax, ay);
    const float a = (mx > 0.f) ? __fdividef(mn, mx) : 0.f;
    const float s = a * a;
    float p = -0.011719132173485577f;
    p = p * s + 0.05264734316289409f; p = p * s - 0.11642647568056484f; p = p * s + 0.1935403746008874f;
    p = p * s - 0.3326228279880411f; p = p * s + 0.9999772191230201f;
    float r = p * a;
    if (ay > ax) r = 1.57079632679489662f - r;
    if (x < 0.f) r = 3.14159265358979324f - r;
    return copysignf(r, y);
}
__device__ __forceinline__ int el_bucket(float el) {   // monotone non-decreasing in el (required for conservative masks)
    int b = (int)floorf((el - VELO_EL_MIN) * (VELO_EL_BUCKETS / (VELO_EL_MAX - VELO_EL_MIN)));
    return min(max(b, 0), VELO_EL_BUCKETS - 1);
}
// monotone non-decreasing in rho (all the masks need).  A float's exponent and top mantissa bits are a piecewise-linear log2:
// VELO_RG_MANT_BITS mantissa bits = 64 buckets per octave from VELO_RG_MIN, four integer instructions instead of log2f.
__device__ __forceinline__ int rg_bucket(float rho) {
    const int b = (__float_as_int(fmaxf(rho, VELO_RG_MIN)) - __float_as_int(VELO_RG_MIN)) >> (23 - VELO_RG_MANT_BITS);
    return min(b, VELO_RG_BUCKETS - 1);
}
__device__ __forceinline__ int az_bin(float az) {
    int b = (int)((az + CUDART_PI_F) * (VELO_AZ_BINS / (2.0f * CUDART_PI_F)));
    return min(max(b, 0), VELO_AZ_BINS - 1);
}


// ------------------------------------------------------------------------------------------------ a13: normal equations
// Row staging for the warp-cooperative accumulation of  H += rho' J^T J,  g += rho' J^T r,  cost += rho/2
// (what ceres::Solve forms for the single 6-vector block, velo.h:897-902).
// Each lane deposits one residual ROW {J[6], r, rho', rho/2 (only on the first row of a block), valid flag} component-major
// (stride 33 doubles: conflict-free for the deposit and for the walk).  Lane l then walks the deposited rows in lane order with
// the same three loads + DMUL/DADD/DFMA whatever it owns: fixed order => run-to-run deterministic.  Lane l < 27 owns H/g sum l
// (robust in acc, unweighted in raw); lane 27 owns acc = sum of rho/2; lane 28 owns raw = sum of r^2 (the unweighted cost is
// half of it, see neq_store).  (k_icp_pass, where this reduction is 15 % of the kernel, uses the FP64 tensor pipe instead.)
#define NEQ_COMP 10
#define NEQ_CS 33
#define NEQ_STAGE (NEQ_COMP * NEQ_CS)                 /* doubles of the per-warp staging area (s_rows[warps][NEQ_STAGE]) */
static __constant__ int c_pa[32] = { 0,0,0,0,0,0, 1,1,1,1,1, 2,2,2,2, 3,3,3, 4,4, 5,  0,1,2,3,4,5, 8, 6, 0,0,0 };
static __constant__ int c_pb[32] = { 0,1,2,3,4,5, 1,2,3,4,5, 2,3,4,5, 3,4,5, 4,5, 5,  6,6,6,6,6,6, 9, 6, 0,0,0 };

__device__ __forceinline__ void warp_accum(double *s_rows, const double J[6], double r, double rho1, double rho0h, bool valid,
                                           int lane, double &acc, double &raw) {
    __syncwarp();
    if (valid) {
        double *row = s_rows + lane;
#pragma unroll
        for (int i = 0; i < 6; i++) row[i * NEQ_CS] = J[i];
        row[6 * NEQ_CS] = r; row[7 * NEQ_CS] = rho1; row[8 * NEQ_CS] = rho0h; row[9 * NEQ_CS] = 1.0;
    }
    __syncwarp();
    const double *pa = s_rows + c_pa[lane] * NEQ_CS, *pb = s_rows + c_pb[lane] * NEQ_CS, *pw = s_rows + (lane == 27 ? 9 : 7) * NEQ_CS;
    for (unsigned m = __ballot_sync(FULL, valid); m; m &= m - 1) {
        const int i = __ffs(m) - 1;
        const double p = pa[i] * pb[i];
        raw += p;
        acc = fma(pw[i], p, acc);                   // explicit fma: the sums are not parity-critical (1e-4)
    }
}
// where lane `lane`'s (acc, raw) go in a 56-slot record {28 robust, 28 unweighted}
__device__ __forceinline__ void neq_store(double *rec56, int lane, double acc, double raw, bool add) {
    if (lane < 28) rec56[lane] = (add ? rec56[lane] : 0.0) + acc;
    if (lane < 27) rec56[28 + lane] = (add ? rec56[28 + lane] : 0.0) + raw;
    if (lane == 28) rec56[55] = (add ? rec56[55] : 0.0) + 0.5 * raw;
}

// CTA-level finish: lanes' (acc, raw) of every warp -> partial[0..55]; fixed warp order.  s_red: [warps][56]
__device__ __forceinline__ void block_neq_finish(double *s_red, double acc, double raw, double *partial) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    neq_store(s_red + wid * 56, lane, acc, raw, false);
    __syncthreads();
    if (threadIdx.x < 56) {
        double s = 0.0;
        for (int w = 0; w < nw; w++) s += s_red[w * 56 + threadIdx.x];
        partial[threadIdx.x] = s;
    }
}
