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
    const float dz = __fsub_rn(a.z, bz);
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

// ---- pruning geometry (index frame, azimuth / elevation / range of a point).  None of it decides an output bit: it only sizes and
// places conservative search windows, whose pads (asin_ub, mask_query, INDEX_EL_PAD) cover its error.  So, unlike the parity-critical
// arithmetic, it uses fused multiply-adds and the 2-ulp hardware square root / division.
__device__ __forceinline__ float sqrt_ap(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ void idx_frame(const DevCalib &cal, float x, float y, float z, float &vx, float &vy, float &vz) {
    const float dx = x - cal.vtc[3], dy = y - cal.vtc[7], dz = z - cal.vtc[11];
    vx = fmaf(cal.vtc[0], dx, fmaf(cal.vtc[4], dy, cal.vtc[8] * dz));
    vy = fmaf(cal.vtc[1], dx, fmaf(cal.vtc[5], dy, cal.vtc[9] * dz));
    vz = fmaf(cal.vtc[2], dx, fmaf(cal.vtc[6], dy, cal.vtc[10] * dz));
}
// atan2 for the *pruning* geometry (azimuth / elevation in the index frame): odd minimax polynomial of degree 11 on [0,1] + octant
// fix-up, |error| < 2.5e-6 rad over all quadrants (checked against f64 atan2, tools/check_atan2.py, tests/test_abi_host.py);
// asin_ub() pads every window by 1e-5 rad for it.  About a third of the instructions of atan2f.
__device__ __forceinline__ float atan2_q(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mn = fminf(ax, ay), mx = fmaxf(ax, ay);
    const float a = (mx > 0.f) ? __fdividef(mn, mx) : 0.f;
    const float s = a * a;
    float p = -0.011719132173485577f;
    p = fmaf(p, s, 0.05264734316289409f); p = fmaf(p, s, -0.11642647568056484f); p = fmaf(p, s, 0.1935403746008874f);
    p = fmaf(p, s, -0.3326228279880411f); p = fmaf(p, s, 0.9999772191230201f);
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
// H += rho' J^T J,  g += rho' J^T r,  cost += rho/2  — what ceres::Solve forms for the single 6-vector block (velo.h:897-902;
// SURVEY.md A.3), robustified and unweighted.
#define NEQ_STAGE (9 * 36 + 6)                        /* doubles of the per-warp staging area (s_rows[warps][NEQ_STAGE]) */

// On the FP64 tensor pipe (the form k_icp_pass uses inline): X^T W X of one residual row per lane, X = [J, r, 0]
// (8 columns), as eight m8n8k4 steps; a lane reads ONE staged element per step — it is both its A and its B fragment entry — plus the
// row weight.  S = per-warp staging (>= 9 * 36 doubles).  The lane's accumulator fragments persist across calls: lane (fr, fk) =
// (lane >> 2, lane & 3) holds C[fr][2 fk] and C[fr][2 fk + 1] of the raw (cr) and the weighted (cw) product.  Fixed order => deterministic.
__device__ __forceinline__ void neq_mma_rows(double *S, int lane, const double J[6], double res, double wgt, bool valid,
                                             double &cr0, double &cr1, double &cw0, double &cw1) {
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 6; c++) S[c * 36 + lane] = valid ? J[c] : 0.0;
    S[6 * 36 + lane] = valid ? res : 0.0; S[7 * 36 + lane] = 0.0; S[8 * 36 + lane] = valid ? wgt : 0.0;
    __syncwarp();
    const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const double x = S[fr * 36 + 4 * t + fk], w = S[8 * 36 + 4 * t + fk];
        const double xw = x * w;
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(cr0), "+d"(cr1) : "d"(x), "d"(x));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(cw0), "+d"(cw1) : "d"(x), "d"(xw));
    }
}
// the warp's fragments -> a 56-slot record {H upper row-major 21, g 6, cost | the same unweighted}; cost_half = the warp's sum of rho/2
// (valid in every lane).  rec56 must be zero beforehand; every slot is written by exactly one lane.
__device__ __forceinline__ void neq_mma_store(double *rec56, int lane, double cr0, double cr1, double cw0, double cw1, double cost_half) {
    const int fr = lane >> 2, c0 = 2 * (lane & 3), c1 = c0 + 1;
    if (fr < 6) {
        const int base = fr * 6 - (fr * (fr - 1)) / 2 - fr;
        if (c0 >= fr && c0 < 6) { rec56[base + c0] = cw0; rec56[28 + base + c0] = cr0; }
        if (c1 >= fr && c1 < 6) { rec56[base + c1] = cw1; rec56[28 + base + c1] = cr1; }
        if (c0 == 6) { rec56[21 + fr] = cw0; rec56[28 + 21 + fr] = cr0; }
    } else if (fr == 6 && c0 == 6) { rec56[27] = cost_half; rec56[55] = 0.5 * cr0; }
}
// CTA-level finish of the tensor-pipe form: per-warp records summed in warp order -> partial[0..55].  s_red: [warps][56]
__device__ __forceinline__ void block_neq_finish_mma(double *s_red, double cr0, double cr1, double cw0, double cw1, double cost_half_lane, double *partial) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cost_half_lane += __shfl_xor_sync(FULL, cost_half_lane, o);
    for (int i = lane; i < 56; i += 32) s_red[wid * 56 + i] = 0.0;
    __syncwarp();
    neq_mma_store(s_red + wid * 56, lane, cr0, cr1, cw0, cw1, cost_half_lane);
    __syncthreads();
    if (threadIdx.x < 56) {
        double s = 0.0;
        for (int w = 0; w < nw; w++) s += s_red[w * 56 + threadIdx.x];
        partial[threadIdx.x] = s;
    }
}
