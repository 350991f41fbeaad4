// velo_kernels.cu — hand-written sm_100a kernels of the VELO front end (SURVEY.md §8 a3–a13).
//
// Compiled with -fmad=false; in addition every index-determining float expression uses the explicit
// round-to-nearest intrinsics (__fadd_rn/__fmul_rn/__fdiv_rn, never contracted) in exactly the operation
// order of the reference's FMA-free x86-64 build (hazards H2/H3).  Nothing here is GEMM shaped: the work is
// streaming / gather / neighbour search over SoA-of-float4 buffers, bounded by HBM and L1/L2 behaviour.
#include "velo_dev.cuh"

#include "velo_common.cuh"

// ------------------------------------------------------------------------------------------------ a3: ingest
// kitti.h:164-176: flag(i) = i>0 && x_i>0 && (y_i>0)!=(y_{i-1}>0) on the UNtransformed points.
__global__ void __launch_bounds__(256) k_ingest_flags(DevBuffers B, int slot0) {
    const int slot = slot0 + blockIdx.y;
    const int n = B.n_points[slot];
    if ((int)(blockIdx.x * blockDim.x) >= n) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const float4 *raw = B.raw + (size_t)slot * B.N;
    const float *rawf = reinterpret_cast<const float *>(raw);
    const int st = B.raw_stride[slot];
    float x = 0.f, y = 0.f;
    if (i < n) {
        if (st == 4) { float2 xy = *reinterpret_cast<const float2 *>(raw + i); x = xy.x; y = xy.y; }
        else { x = rawf[3 * (size_t)i]; y = rawf[3 * (size_t)i + 1]; }
    }
    float py = __shfl_up_sync(FULL, y, 1);
    if (lane == 0 && i > 0 && i < n) py = rawf[(size_t)st * (i - 1) + 1];
    bool f = (i > 0) && (i < n) && (x > 0.f) && ((y > 0.f) != (py > 0.f));
    unsigned m = __ballot_sync(FULL, f);
    if (lane == 0 && i < n) B.flagbits[(size_t)slot * (B.N / 32) + (i >> 5)] = m;
}

// ring id = running count of flags; ring_start[ring] = position of the flagged point (kitti.h:169-175)
__global__ void __launch_bounds__(1024) k_ingest_rings(DevBuffers B, int slot0) {
    __shared__ int s_w[33];
    const int slot = slot0 + blockIdx.x, tid = threadIdx.x;
    const int n = B.n_points[slot];
    const int W = (n + 31) >> 5, wpt = (W + blockDim.x - 1) / blockDim.x;
    const uint32_t *bits = B.flagbits + (size_t)slot * (B.N / 32);
    int *rs = B.ring_start + (size_t)slot * (B.R + 1);
    int w0 = min(W, tid * wpt), w1 = min(W, w0 + wpt), cnt = 0;
    for (int w = w0; w < w1; w++) cnt += __popc(bits[w]);
    int total;
    int run = block_excl_scan(cnt, s_w, total);
    for (int w = w0; w < w1; w++) {
        uint32_t b = bits[w];
        while (b) {
            int bit = __ffs(b) - 1; b &= b - 1;
            ++run;
            if (run <= B.R) rs[run] = w * 32 + bit;
        }
    }
    __syncthreads();
    if (tid == 0) {
        int nr = n > 0 ? total + 1 : 0;
        int st = 0;
        if (nr > B.R) { st = VELO_ERR_CAPACITY; nr = 0; }   // hazard H12: ring count is data dependent
        rs[0] = 0;
        if (nr > 0) rs[nr] = n;
        B.n_rings[slot] = nr;
        B.status[slot] = st;
    }
}

// velo_to_cam transform (kitti.h:162, PCL dense branch: left-to-right f32) + reverse/half-rotate reorder (kitti.h:178-183)
__global__ void __launch_bounds__(256) k_ingest_permute(DevBuffers B, DevCalib cal, int slot0) {
    __shared__ int s_rs[VELO_MAX_RINGS_HARD + 1];
    const int slot = slot0 + blockIdx.y;
    const int n = B.n_points[slot], nr = B.n_rings[slot];
    if ((int)(blockIdx.x * blockDim.x) >= n || nr == 0) return;
    const int *rs = B.ring_start + (size_t)slot * (B.R + 1);
    for (int i = threadIdx.x; i <= nr; i += blockDim.x) s_rs[i] = rs[i];
    __syncthreads();
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    int lo = 0, hi = nr - 1;               // largest r with s_rs[r] <= d
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (s_rs[mid] <= d) lo = mid; else hi = mid - 1; }
    const int r0 = s_rs[lo], L = s_rs[lo + 1] - r0, i = d - r0;
    const int src = r0 + (L - 1 - (i + L / 2) % L);
    float4 p;
    if (B.raw_stride[slot] == 4) p = __ldg(B.raw + (size_t)slot * B.N + src);
    else { const float *q = reinterpret_cast<const float *>(B.raw + (size_t)slot * B.N) + 3 * (size_t)src; p = make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), 0.f); }
    float4 o;
    o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(cal.vtc[0], p.x), __fmul_rn(cal.vtc[1], p.y)), __fmul_rn(cal.vtc[2], p.z)), cal.vtc[3]);
    o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(cal.vtc[4], p.x), __fmul_rn(cal.vtc[5], p.y)), __fmul_rn(cal.vtc[6], p.z)), cal.vtc[7]);
    o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(cal.vtc[8], p.x), __fmul_rn(cal.vtc[9], p.y)), __fmul_rn(cal.vtc[10], p.z)), cal.vtc[11]);
    o.w = 1.0f;
    B.pts[(size_t)slot * B.N + d] = o;
}

// ------------------------------------------------------------------------------------------------ a4: neighbour index
// Index frame = velodyne-like frame (rotation part of velo_to_cam transposed): rings are cones about its z axis,
// so a ring has a narrow elevation interval and is nearly sorted in azimuth.  Used ONLY for pruning; all
// distances / decisions are evaluated on the cam-0 coordinates exactly as the reference does.
// one CTA per (ring, slot): counting sort of the ring by azimuth bin + per-sector elevation intervals.
// Replaces the 64 KdTreeFLANN::setInputCloud calls of lru.h:17-20.
// Angles come from atan2_q (|error| < 2e-6 rad, velo_common.cuh) instead of atan2f (a third of the instructions; the kernel was
// instruction bound on three atan2f per point): the azimuth only decides a BIN, and the query's window is padded by 1e-5 rad for
// exactly this (asin_ub); the elevation only feeds the sector boxes, which are widened by INDEX_EL_PAD here.  Bins are kept in
// registers between the histogram pass and the scatter pass.
#define INDEX_EL_PAD 4e-6f
#define INDEX_KEEP 8          /* points per thread whose bin stays in a register (rings up to 2048 points) */
__device__ __forceinline__ int index_bin_of(const DevCalib &cal, const float4 &p, float &el, float &rho) {
    float vx, vy, vz; idx_frame(cal, p.x, p.y, p.z, vx, vy, vz);
    const float dxy2 = fmaf(vx, vx, vy * vy);
    el = atan2_q(vz, sqrt_ap(dxy2)); rho = sqrt_ap(fmaf(vz, vz, dxy2));
    return az_bin(atan2_q(vy, vx));
}
__global__ void __launch_bounds__(256) k_index_build(DevBuffers B, DevCalib cal, int slot0) {
    __shared__ int s_hist[VELO_AZ_BINS];
    __shared__ int s_cur[VELO_AZ_BINS];
    __shared__ int s_lo[VELO_SECTORS], s_hi[VELO_SECTORS], s_rlo[VELO_SECTORS], s_rhi[VELO_SECTORS];
    __shared__ int s_w[33];
    const int slot = slot0 + blockIdx.y, ring = blockIdx.x, tid = threadIdx.x;
    if (ring >= B.n_rings[slot]) return;
    const int *rs = B.ring_start + (size_t)slot * (B.R + 1);
    const int r0 = rs[ring], L = rs[ring + 1] - r0;
    const float4 *pts = B.pts + (size_t)slot * B.N + r0;
    for (int i = tid; i < VELO_AZ_BINS; i += blockDim.x) { s_hist[i] = 0; s_cur[i] = 0; }
    if (tid < VELO_SECTORS) { s_lo[tid] = s_rlo[tid] = f2ord(CUDART_INF_F); s_hi[tid] = s_rhi[tid] = f2ord(-CUDART_INF_F); }
    __syncthreads();
    int bins[INDEX_KEEP];
    for (int i0 = 0; i0 < L; i0 += 256 * INDEX_KEEP) {               // (one trip for every real ring)
#pragma unroll
        for (int k = 0; k < INDEX_KEEP; k++) {
            const int i = i0 + tid + 256 * k;
            int b = -1, e = 0, rr = 0;
            if (i < L) {
                float el, rho;
                b = index_bin_of(cal, pts[i], el, rho);
                e = f2ord(el); rr = f2ord(rho);
                atomicAdd(&s_hist[b], 1);
            }
            if (i0 == 0) bins[k] = b;
            // sector bounds: the 32 lanes hold consecutive points of the ring, i.e. one sector or two neighbouring ones.  Each of the
            // (at most two) sectors present at the ends of the warp is reduced with full-mask REDUX (lanes outside contribute the
            // neutral element) and written by one lane; a lane in neither (noise in the point order) writes for itself.
            {
                const int sec = (i < L) ? b / VELO_BINS_PER_SECTOR : -1;
                const unsigned act = __ballot_sync(FULL, i < L);
                if (act) {
                    const int secA = __shfl_sync(FULL, sec, __ffs(act) - 1), secB = __shfl_sync(FULL, sec, 31 - __clz(act));
#pragma unroll
                    for (int g = 0; g < 2; g++) {
                        const int sg = g == 0 ? secA : secB;
                        if (g == 1 && secB == secA) break;
                        const bool mine = sec == sg;
                        const int elo = __reduce_min_sync(FULL, mine ? e : 0x7fffffff), ehi = __reduce_max_sync(FULL, mine ? e : (int)0x80000000);
                        const int rlo = __reduce_min_sync(FULL, mine ? rr : 0x7fffffff), rhi = __reduce_max_sync(FULL, mine ? rr : (int)0x80000000);
                        if ((tid & 31) == 0) { atomicMin(&s_lo[sg], elo); atomicMax(&s_hi[sg], ehi); atomicMin(&s_rlo[sg], rlo); atomicMax(&s_rhi[sg], rhi); }
                    }
                    if (i < L && sec != secA && sec != secB) { atomicMin(&s_lo[sec], e); atomicMax(&s_hi[sec], e); atomicMin(&s_rlo[sec], rr); atomicMax(&s_rhi[sec], rr); }
                }
            }
        }
    }
    __syncthreads();
    // exclusive scan of the AZ bins with 256 threads (AZ/256 consecutive bins each)
    constexpr int PER = VELO_AZ_BINS / 256;
    int av[PER], sum = 0, total;
#pragma unroll
    for (int k = 0; k < PER; k++) { av[k] = s_hist[PER * tid + k]; sum += av[k]; }
    int ex = block_excl_scan(sum, s_w, total);
#pragma unroll
    for (int k = 0; k < PER; k++) { s_hist[PER * tid + k] = ex; ex += av[k]; }
    __syncthreads();
    int *cs = B.cell_start + ((size_t)slot * B.R + ring) * (VELO_AZ_BINS + 1);
    for (int i = tid; i < VELO_AZ_BINS; i += blockDim.x) cs[i] = r0 + s_hist[i];
    if (tid == 0) cs[VELO_AZ_BINS] = r0 + L;
    if (tid < VELO_SECTORS) {                                          // (an empty sector keeps lo = +inf > hi = -inf)
        const float lo = ord2f(s_lo[tid]), hi = ord2f(s_hi[tid]);
        B.sec_box[((size_t)slot * B.R + ring) * VELO_SECTORS + tid] = make_float4(lo - INDEX_EL_PAD, hi + INDEX_EL_PAD, ord2f(s_rlo[tid]), ord2f(s_rhi[tid]));
    }
    float4 *sorted = B.sorted + (size_t)slot * B.N + r0;
    for (int i0 = 0; i0 < L; i0 += 256 * INDEX_KEEP) {
#pragma unroll
        for (int k = 0; k < INDEX_KEEP; k++) {
            const int i = i0 + tid + 256 * k;
            if (i < L) {
                float4 p = pts[i];
                float el, rho;
                const int b = (i0 == 0) ? bins[k] : index_bin_of(cal, p, el, rho);
                const int pos = s_hist[b] + atomicAdd(&s_cur[b], 1);
                p.w = __int_as_float(i);
                sorted[pos] = p;
            }
        }
    }
}

// Ring-mask tables: for (slot, sector, elevation bucket b)  mask_lo = { rings r : bucket(lo_r) <= b },  mask_hi = { r : bucket(hi_r) >= b }.
// The rings whose elevation interval in that sector can intersect [e0, e1] are a subset of mask_lo[bucket(e1)] & mask_hi[bucket(e0)]
// (bucket() is monotone), which turns the per-query 64-ring scan into two 8-byte loads per sector; the same over range buckets.
// One CTA per (sector, slot), one thread per bucket: every ring sets its bit in the one bucket each of its four bounds falls in
// (shared-memory atomicOr), then a prefix-OR (lo tables) / suffix-OR (hi tables) over the 256 buckets gives the cumulative masks.
__device__ __forceinline__ unsigned long long block_or_scan_256(unsigned long long v, unsigned long long *s_w, int tid) {
    // inclusive OR-scan over 256 threads (8 warps); s_w: 8 words
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(FULL, v, o); if (lane >= o) v |= t; }
    if (lane == 31) s_w[wid] = v;
    __syncthreads();
    unsigned long long pre = 0ull;
    for (int k = 0; k < wid; k++) pre |= s_w[k];
    __syncthreads();
    return v | pre;
}
__global__ void __launch_bounds__(VELO_EL_BUCKETS) k_index_masks(DevBuffers B, int slot0) {
    static_assert(VELO_EL_BUCKETS == 256 && VELO_RG_BUCKETS == 256, "one thread per bucket of either table, 8 warps");
    __shared__ unsigned long long s_set[4][VELO_EL_BUCKETS];   // bits of the rings whose bound falls into bucket b: elev lo, elev hi, range min, range max
    __shared__ unsigned long long s_w[8];
    const int slot = slot0 + blockIdx.y, sec = blockIdx.x, b = threadIdx.x;
    const int nr = B.n_rings[slot];
    const float4 *se = B.sec_box + (size_t)slot * B.R * VELO_SECTORS + sec;
    const size_t o = (((size_t)slot * VELO_SECTORS + sec) * VELO_EL_BUCKETS + b) * B.W;
    for (int w = 0; w < B.W; w++) {
#pragma unroll
        for (int t = 0; t < 4; t++) s_set[t][b] = 0ull;
        __syncthreads();
        const int r = w * 64 + b;
        if (b < 64 && r < nr) {
            const float4 e = se[(size_t)r * VELO_SECTORS];
            if (e.x <= e.y) {                                   // sector not empty for this ring
                const unsigned long long bit = 1ull << b;
                atomicOr(&s_set[0][el_bucket(e.x)], bit); atomicOr(&s_set[1][el_bucket(e.y)], bit);
                atomicOr(&s_set[2][rg_bucket(e.z)], bit); atomicOr(&s_set[3][rg_bucket(e.w)], bit);
            }
        }
        __syncthreads();
        const int rb = VELO_EL_BUCKETS - 1 - b;               // suffix-OR = prefix-OR over the reversed index
        const unsigned long long lo = block_or_scan_256(s_set[0][b], s_w, b);
        const unsigned long long hi = block_or_scan_256(s_set[1][rb], s_w, b);
        const unsigned long long rlo = block_or_scan_256(s_set[2][b], s_w, b);
        const unsigned long long rhi = block_or_scan_256(s_set[3][rb], s_w, b);
        B.mask_lo[o + w] = lo; B.rmask_lo[o + w] = rlo;
        const size_t orv = (((size_t)slot * VELO_SECTORS + sec) * VELO_EL_BUCKETS + rb) * B.W;
        B.mask_hi[orv + w] = hi; B.rmask_hi[orv + w] = rhi;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ a5: projection
// One warp per (ring, slot), all cameras.  Lanes project + FOV-test 32 points at a time (velo.h:346-349, exact IEEE operations).
// The reference's sequential occlusion stack (velo.h:351-368) pops or skips only when a point's canonical x is smaller than the
// x on top of the stack; a ring sweeps the image left to right, so for almost every 32-point chunk the survivors' x are
// non-decreasing (also against the current top) and the whole chunk is pushed by one ballot-compacted store.  Any chunk that
// breaks the order is replayed point by point, warp-uniformly, with the reference's exact pop / skip / push sequence.  The stack
// IS the output array, so a pop only re-reads the new top (hazards H6/H7); the stack state is warp-uniform in registers.
#ifndef PROJ_WARPS
#define PROJ_WARPS 4
#endif
#ifndef PROJ_STAGES
#define PROJ_STAGES 8         /* 32-point chunks in flight per warp (cp.async ring in shared memory, no register cost) */
#endif
template <int NC>
__global__ void __launch_bounds__(PROJ_WARPS * 32) k_project(DevBuffers B, DevCalib cal, int slot0) {
    __shared__ float4 s_ring[PROJ_WARPS][PROJ_STAGES][32];
    const int slot = slot0 + blockIdx.y, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ring = blockIdx.x * PROJ_WARPS + wid;
    if (ring >= B.n_rings[slot]) return;
    const int *rs = B.ring_start + (size_t)slot * (B.R + 1);
    const int r0 = rs[ring], L = rs[ring + 1] - r0;
    const float4 *pts = B.pts + (size_t)slot * B.N + r0;
    const int C = cal.num_cams;
    const unsigned lt = (1u << lane) - 1u;
    int depth[NC]; float top_x[NC], top_z[NC];
    float ylo[NC], yhi[NC];       // per lane: y range of everything this lane ever pushed (superset of the final stack)
#pragma unroll
    for (int cam = 0; cam < NC; cam++) { depth[cam] = 0; top_x[cam] = 0.f; top_z[cam] = 0.f; ylo[cam] = CUDART_INF_F; yhi[cam] = -CUDART_INF_F; }
    // each lane copies and later reads only its own element of a stage, so the ring needs no warp synchronisation
    const unsigned s_base = (unsigned)__cvta_generic_to_shared(&s_ring[wid][0][lane]);
#pragma unroll
    for (int st = 0; st < PROJ_STAGES - 1; st++) {
        if (st * 32 + lane < L) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_base + st * 512u), "l"(pts + st * 32 + lane) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    int stage = 0;
    for (int c0 = 0; c0 < L; c0 += 32) {
        const int i = c0 + lane;
        {
            const int ahead = c0 + (PROJ_STAGES - 1) * 32 + lane;
            const int st = (stage == 0) ? PROJ_STAGES - 1 : stage - 1;
            if (ahead < L) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_base + st * 512u), "l"(pts + ahead) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.wait_group %0;" ::"n"(PROJ_STAGES - 1) : "memory");
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < L) p = s_ring[wid][stage][lane];
        stage = (stage + 1 == PROJ_STAGES) ? 0 : stage + 1;
#pragma unroll
        for (int cam = 0; cam < NC; cam++) {
            if (cam >= C) break;
            const float tz = cal.cam_t[cam][2];
            const float ppx = __fadd_rn(p.x, cal.cam_t[cam][0]), ppy = __fadd_rn(p.y, cal.cam_t[cam][1]), ppz = __fadd_rn(p.z, tz);        // velo.h:346
            // cheap conservative pre-test (multiplications, padded): most 32-point chunks are entirely behind or beside the camera
            const float tol = 1e-4f * (fabsf(ppx) + fabsf(ppy) + ppz);
            const bool maybe = (i < L) && (ppz > 0.f) && (ppx >= cal.fov[cam][0] * ppz - tol) && (ppx <= cal.fov[cam][1] * ppz + tol) &&
                               (ppy >= cal.fov[cam][2] * ppz - tol) && (ppy <= cal.fov[cam][3] * ppz + tol);
            if (!__any_sync(FULL, maybe)) continue;
            const float cx = __fdiv_rn(ppx, ppz), cy = __fdiv_rn(ppy, ppz);                                                              // velo.h:347
            const bool in = (i < L) && (ppz > 0.f) && (cx >= cal.fov[cam][0]) && (cx < cal.fov[cam][1]) && (cy >= cal.fov[cam][2]) && (cy < cal.fov[cam][3]);
            unsigned rem = __ballot_sync(FULL, in);
            if (rem == 0u) continue;
            float2 *proj = B.proj + ((size_t)slot * B.C + cam) * B.N + r0;
            float4 *valid = B.valid + ((size_t)slot * B.C + cam) * B.N + r0;
            if (in) { ylo[cam] = fminf(ylo[cam], cy); yhi[cam] = fmaxf(yhi[cam], cy); }     // superset of what is pushed: still a valid bound
            while (rem) {
                // survivors not left of their predecessor (the stack top for the first one) cause no pop and no skip
                const unsigned below = rem & lt;
                const float pcx = __shfl_sync(FULL, cx, below ? 31 - __clz(below) : 0);
                const bool mine = (rem >> lane) & 1u;
                const bool bad = mine && (below ? (cx < pcx) : (depth[cam] > 0 && cx < top_x[cam]));
                const unsigned viol = __ballot_sync(FULL, bad);
                const unsigned run = viol ? (rem & ((1u << (__ffs(viol) - 1)) - 1u)) : rem;      // ordered prefix
                if (run) {                                                                        // pushed as one compacted store (velo.h:366-368)
                    if ((run >> lane) & 1u) {
                        const int pos = depth[cam] + __popc(run & lt);
                        proj[pos] = make_float2(cx, cy);
                        valid[pos] = make_float4(p.x, p.y, p.z, 1.0f);
                    }
                    const int last = 31 - __clz(run);
                    top_x[cam] = __shfl_sync(FULL, cx, last); top_z[cam] = __shfl_sync(FULL, ppz, last);
                    depth[cam] += __popc(run);
                    rem &= ~run;
                }
                if (viol) {
                    // the first out-of-order survivor takes the reference's scalar path; every lane runs it, lane b owns the stores
                    const int b = __ffs(viol) - 1;
                    rem &= ~(1u << b);
                    const float bx = __shfl_sync(FULL, cx, b), bz = __shfl_sync(FULL, ppz, b);
                    while (depth[cam] > 0 && bx < top_x[cam] && bz < top_z[cam]) {             // velo.h:351-358: pop occluded
                        depth[cam]--;
                        if (depth[cam] > 0) {
                            __syncwarp();                                                       // the new top may have been stored by another lane
                            top_x[cam] = proj[depth[cam] - 1].x; top_z[cam] = __fadd_rn(valid[depth[cam] - 1].z, tz);
                        }
                    }
                    if (depth[cam] > 0 && bx < top_x[cam] && bz > top_z[cam]) continue;          // velo.h:360-365: skip occluded
                    if (lane == b) {
                        proj[depth[cam]] = make_float2(cx, cy);                                  // velo.h:366-368
                        valid[depth[cam]] = make_float4(p.x, p.y, p.z, 1.0f);
                    }
                    depth[cam]++; top_x[cam] = bx; top_z[cam] = bz;
                }
            }
        }
    }
#pragma unroll
    for (int cam = 0; cam < NC; cam++) {
        if (cam >= C) break;
        float lo = ylo[cam], hi = yhi[cam];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(FULL, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(FULL, hi, o)); }
        if (lane == 0) {
            B.proj_count[((size_t)slot * B.C + cam) * B.R + ring] = depth[cam];
            B.proj_yrange[((size_t)slot * B.C + cam) * B.R + ring] = make_float2(lo, hi);
        }
    }
}

// ------------------------------------------------------------------------------------------------ a6/a7: depth association
__device__ __forceinline__ float lerp1(float p1, float p2, float start, float end, float mid) {   // utility.h:20-29
    float a = __fdiv_rn(__fsub_rn(mid, start), __fsub_rn(end, start));
    float b = __fsub_rn(1.0f, a);
    return __fadd_rn(__fmul_rn(p1, b), __fmul_rn(p2, a));
}
__device__ __forceinline__ float3 lerp3(float4 p1, float4 p2, float start, float end, float mid) { // utility.h:7-19
    float a = __fdiv_rn(__fsub_rn(mid, start), __fsub_rn(end, start));
    float b = __fsub_rn(1.0f, a);
    return make_float3(__fadd_rn(__fmul_rn(p1.x, b), __fmul_rn(p2.x, a)), __fadd_rn(__fmul_rn(p1.y, b), __fmul_rn(p2.y, a)),
                       __fadd_rn(__fmul_rn(p1.z, b), __fmul_rn(p2.z, a)));
}
__device__ __forceinline__ bool width_ok(float dx, const DevCalib &cal) {   // hazard H1 (velo.h:416,419)
    if (cal.abs_truncates) return abs((int)dx) == 0;                        // (double)abs((int)dx) < 0.015  <=>  (int)dx == 0
    return fabsf(dx) < cal.assoc_thr;
}

// The ring loop, bracket search and bilinear patch of velo.h:390-492 (H4/H5), one thread per keypoint.
// One CTA per (camera, slot) serves every keypoint set of that image: the projections of all rings (x,y pairs, ~140 KB for a
// KITTI frame) are first staged in shared memory, because the searches are scattered 8-byte reads.
//  * Which rings a keypoint has to search is read from a table.  A hit on the ring pair (s-1, s) needs
//    (proj[s][mid].y > kp.y) != (proj[s-1][last].y > kp.y) (velo.h:414-415): impossible when, over the points that can be those
//    bracket ends, both rings lie above the keypoint (y > kp.y) or both lie not-above.  A ring whose two pairs are both
//    impossible cannot influence the result, so it is not searched and `last` is -1 after it, exactly as after a ring without
//    a bracket (searching a ring that was not needed is what the reference does anyway, so any superset is exact).
//    1-D table (always): per bucket of the image height, the rings needed judging by each ring's whole y range.
//    2-D table (<= 64 rings, float `abs`): per (y bucket, x zone).  The width gate (velo.h:416-421) only lets a bracket hit if its
//    two ends are closer than depth_assoc_thresh in x, and the keypoint lies between them, so both ends are within the threshold
//    of kp.x: only a ring's points inside the zone widened by the threshold count for its y range.  Rings are smooth curves in
//    the image, so this leaves the two or three rings around the keypoint instead of the dozen whose full y range covers it.
//  * The reference's binary search (velo.h:404-412) finds an index mid with proj[mid].x <= kp.x < proj[mid+1].x.  On a ring whose
//    projected x never decreases (every ring, unless the occlusion stack pushed a z-tie, velo.h:360-366) that index is unique —
//    the last point with x <= kp.x — so ANY search finds the reference's bracket.  Such rings get a 128-bucket x table while they
//    are staged (start index of each bucket: points in lower buckets are strictly left of the keypoint, points in higher ones
//    strictly right, because the bucket function is monotone) and the search is two table reads plus a scan of the keypoint's own
//    bucket (~2 points) instead of 9 dependent steps.  A ring with a decreasing x (H4), a NaN keypoint, or a projection that does
//    not fit in shared memory takes the reference's exact sequence of mid points.
// Staging, the x tables and the zone ranges are built with one warp per ring (lanes stride over the ring's points).
#ifndef ASSOC_THREADS
#define ASSOC_THREADS 1024        /* measured 512 -> 1024: 0.303 -> 0.252 ms per 200 frames (more keypoints in flight per staged image) */
#endif
#define ASSOC_DYN_BYTES (200 * 1024)   /* dynamic shared memory: staged (x,y) pairs + the x tables + the 2-D ring table (+ 24 KB static <= 227 KB) */
#define ASSOC_YB 256                   /* keypoint-y buckets of the needed-ring tables */
#define ASSOC_NB 128                   /* x buckets per ring */
#define ASSOC_ZONES 16                 /* x zones of the 2-D needed-ring table */
#define ASSOC_MAXW (VELO_MAX_RINGS_HARD / 64)
__device__ __forceinline__ int assoc_xbucket(float x, float xmin, float xscale) {   // monotone non-decreasing in x (NaN -> 0)
    const int b = (int)((x - xmin) * xscale);
    return min(max(b, 0), ASSOC_NB - 1);
}
__device__ __forceinline__ int assoc_zone(float x, float xmin, float zscale) {      // monotone non-decreasing in x
    const int z = (int)((x - xmin) * zscale);
    return min(max(z, 0), ASSOC_ZONES - 1);
}
__device__ __forceinline__ int assoc_ybucket(float y, float ymin, float yscale) { return min(max((int)((y - ymin) * yscale), 0), ASSOC_YB - 1); }
__global__ void __launch_bounds__(ASSOC_THREADS) k_assoc_search(DevBuffers B, DevCalib cal, int slot0, int set0, int nsets, int cam0) {
    extern __shared__ float2 s_proj[];
    __shared__ int s_off[VELO_MAX_RINGS_HARD + 1];
    __shared__ int s_rs[VELO_MAX_RINGS_HARD + 1];
    __shared__ int s_cnt[VELO_MAX_RINGS_HARD];
    __shared__ float2 s_yr[VELO_MAX_RINGS_HARD];
    __shared__ float2 s_pair[VELO_MAX_RINGS_HARD + 1];   // pair (p-1, p): a keypoint can hit it only if s_pair[p].x <= kp.y < s_pair[p].y
    __shared__ int s_mono[VELO_MAX_RINGS_HARD];
    __shared__ unsigned long long s_need[ASSOC_YB + 1][ASSOC_MAXW];   // [ASSOC_YB]: every searchable ring (keypoints outside the image)
    __shared__ int2 s_zr[64][ASSOC_ZONES];                            // y range (ordered ints) of ring s inside zone z widened by the width gate
    __shared__ int s_w[33];
    const int cam = cam0 + blockIdx.x, slot = slot0 + blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
    const int nr = B.n_rings[slot], NW = (nr + 63) >> 6;
    const int *rs = B.ring_start + (size_t)slot * (B.R + 1);
    const int *pc = B.proj_count + ((size_t)slot * B.C + cam) * B.R;
    const float2 *yr = B.proj_yrange + ((size_t)slot * B.C + cam) * B.R;
    const float2 *proj = B.proj + ((size_t)slot * B.C + cam) * B.N;
    const float4 *valid = B.valid + ((size_t)slot * B.C + cam) * B.N;
    int total = 0;
    for (int i0 = 0; i0 < nr; i0 += blockDim.x) {                 // ring offsets in the staged array (exclusive scan of the counts)
        const int i = i0 + tid;
        const int c = i < nr ? pc[i] : 0;
        if (i < nr) { s_rs[i] = rs[i]; s_cnt[i] = c; s_yr[i] = yr[i]; s_mono[i] = 1; }
        int t;
        const int ex = block_excl_scan(c, s_w, t);
        if (i < nr) s_off[i] = total + ex;
        total += t;
    }
    if (tid == 0) s_off[nr] = total;
    for (int i = tid; i < 64 * ASSOC_ZONES; i += blockDim.x) (&s_zr[0][0])[i] = make_int2(f2ord(CUDART_INF_F), f2ord(-CUDART_INF_F));
    for (int i = tid; i < (ASSOC_YB + 1) * ASSOC_MAXW; i += blockDim.x) (&s_need[0][0])[i] = 0ull;
    __syncthreads();
    const size_t b_pts = (size_t)total * sizeof(float2), b_lut = ((size_t)nr * (ASSOC_NB + 1) * sizeof(unsigned short) + 15) & ~(size_t)15;
    unsigned short *s_lut = reinterpret_cast<unsigned short *>(s_proj + total);
    unsigned long long *s_need2 = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(s_proj) + ((b_pts + b_lut + 15) & ~(size_t)15));   // [ASSOC_YB][ASSOC_ZONES]
    const bool fits = b_pts + b_lut + 16 <= ASSOC_DYN_BYTES;
    const bool use2d = fits && nr <= 64 && !cal.abs_truncates &&
                       b_pts + b_lut + 32 + (size_t)ASSOC_YB * ASSOC_ZONES * sizeof(unsigned long long) <= ASSOC_DYN_BYTES;
    const float xmin = cal.fov[cam][0], xspan = cal.fov[cam][1] - cal.fov[cam][0], xscale = ASSOC_NB / xspan, zscale = ASSOC_ZONES / xspan;
    const float ymin = cal.fov[cam][2], ymax = cal.fov[cam][3], yscale = ASSOC_YB / (ymax - ymin);
    const float zpad = cal.assoc_thr * 1.01f + 1e-6f;             // the width gate: bracket ends lie within assoc_thr of the keypoint in x
    if (use2d) for (int i = tid; i < ASSOC_YB * ASSOC_ZONES; i += blockDim.x) s_need2[i] = 0ull;
    if (fits) {
        // one warp per ring: 8-byte cp.async per projection (no register round trip; all rings in flight together)
        for (int s = wid; s < nr; s += nwarps) {
            const float2 *src = proj + s_rs[s];
            const unsigned dst = (unsigned)__cvta_generic_to_shared(s_proj + s_off[s]);
            const int cnt = s_cnt[s];
            for (int i = lane; i < cnt; i += 32) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * i), "l"(src + i) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // x tables: lut[s][b] = first index of ring s whose bucket is >= b (points are bucket-sorted on a ring with non-decreasing x);
        // a decreasing x marks the ring for the exact search.
        for (int s = wid; s < nr; s += nwarps) {
            const float2 *ps = s_proj + s_off[s];
            unsigned short *lut = s_lut + s * (ASSOC_NB + 1);
            const int cnt = s_cnt[s];
            float xprev = 0.f; int bprev = -1;                               // last point of the previous 32 (lane 31's)
            int bfirst = ASSOC_NB, blast = ASSOC_NB;                         // (an empty ring: every entry is 0 = cnt)
            for (int i0 = 0; i0 < cnt; i0 += 32) {
                const int i = i0 + lane;
                const bool in = i < cnt;
                const float xi = in ? ps[i].x : 0.f;
                const int bi = in ? assoc_xbucket(xi, xmin, xscale) : ASSOC_NB;
                float xp = __shfl_up_sync(FULL, xi, 1); int bp = __shfl_up_sync(FULL, bi, 1);
                if (lane == 0) { xp = xprev; bp = bprev; }
                if (in && i > 0) {                                           // the gap between two neighbours: usually empty or one bucket
                    if (xp > xi) s_mono[s] = 0;
                    for (int b = bp + 1; b <= bi; b++) lut[b] = (unsigned short)i;
                }
                if (i0 == 0) bfirst = __shfl_sync(FULL, bi, 0);
                blast = __shfl_sync(FULL, bi, min(31, cnt - 1 - i0));
                xprev = __shfl_sync(FULL, xi, 31); bprev = __shfl_sync(FULL, bi, 31);
            }
            // the buckets up to the first point's and beyond the last point's (most of the table for a ring that crosses only part of
            // the image) are filled by all lanes instead of by the first / last point's lane alone
            for (int b = lane; b <= bfirst; b += 32) lut[b] = 0;
            for (int b = blast + 1 + lane; b <= ASSOC_NB; b += 32) lut[b] = (unsigned short)cnt;
        }
        __syncthreads();
        // zone ranges: one thread per (ring, zone).  On a ring with non-decreasing x the points whose x lies in the zone widened by
        // the width gate are one index span, read off the x table (buckets are monotone in x, so the span of the buckets that
        // the widened zone touches covers them); any other ring counts with its whole y range.
        if (use2d) {
            for (int t = tid; t < nr * ASSOC_ZONES; t += blockDim.x) {
                const int s = t / ASSOC_ZONES, z = t % ASSOC_ZONES;
                float lo = CUDART_INF_F, hi = -CUDART_INF_F;
                if (s_mono[s]) {
                    const float zlo = xmin + z / zscale - zpad, zhi = xmin + (z + 1) / zscale + zpad;
                    const unsigned short *lut = s_lut + s * (ASSOC_NB + 1);
                    const float2 *ps = s_proj + s_off[s];
                    const int i0 = z == 0 ? 0 : lut[assoc_xbucket(zlo, xmin, xscale)], i1 = z == ASSOC_ZONES - 1 ? s_cnt[s] : lut[assoc_xbucket(zhi, xmin, xscale) + 1];
                    for (int i = i0; i < i1; i++) { const float y = ps[i].y; lo = fminf(lo, y); hi = fmaxf(hi, y); }
                } else { lo = s_yr[s].x; hi = s_yr[s].y; }
                s_zr[s][z] = make_int2(f2ord(lo), f2ord(hi));
            }
        }
    }
    // 1-D table.  Pair (a, b) is impossible for a keypoint iff (ymin_a > y && ymin_b > y) || (ymax_a <= y && ymax_b <= y), i.e. possible
    // iff min(ymin_a, ymin_b) <= y < max(ymax_a, ymax_b); ring s is needed iff pair (s-1, s) or pair (s, s+1) is possible.
    for (int p = tid; p <= nr; p += blockDim.x) {
        float2 pr = make_float2(CUDART_INF_F, -CUDART_INF_F);        // pairs with a ring that does not exist are impossible
        if (p > 0 && p < nr) pr = make_float2(fminf(s_yr[p - 1].x, s_yr[p].x), fmaxf(s_yr[p - 1].y, s_yr[p].y));
        s_pair[p] = pr;
    }
    __syncthreads();
    for (int bkt = wid; bkt <= ASSOC_YB; bkt += nwarps) {               // a warp per bucket, a lane per ring: the mask is a ballot
        const float y0 = ymin + bkt / yscale - 1e-4f, y1 = ymin + (bkt + 1) / yscale + 1e-4f;   // bucket edges, padded
#pragma unroll
        for (int w = 0; w < ASSOC_MAXW; w++) {
            unsigned long long m = 0ull;
            if (64 * w < nr) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int s2 = 64 * w + 32 * h + lane;
                    bool need = false;
                    if (s2 < nr && s_cnt[s2] > 1) {                                              // rings with <= 1 points: velo.h:400-403
                        const float2 pa = s_pair[s2], pb = s_pair[s2 + 1];
                        need = bkt == ASSOC_YB || (pa.x <= y1 && y0 < pa.y) || (pb.x <= y1 && y0 < pb.y);
                    }
                    m |= (unsigned long long)__ballot_sync(FULL, need) << (32 * h);
                }
            }
            if (lane == 0) s_need[bkt][w] = m;
        }
    }
    // 2-D table: one thread per (pair, zone) marks both rings of the pair in the y buckets the pair can be hit from in that zone
    if (use2d) {
        for (int t = tid; t < nr * ASSOC_ZONES; t += blockDim.x) {
            const int p = t / ASSOC_ZONES + 1, z = t % ASSOC_ZONES;      // pair (p-1, p), p = 1 .. nr-1
            if (p >= nr || s_cnt[p - 1] <= 1 || s_cnt[p] <= 1) continue;
            const int2 ra = s_zr[p - 1][z], rb = s_zr[p][z];
            if (ra.x > ra.y || rb.x > rb.y) continue;                    // a ring without points near the zone cannot close a bracket there
            const float lo = fminf(ord2f(ra.x), ord2f(rb.x)), hi = fmaxf(ord2f(ra.y), ord2f(rb.y));
            if (!(hi >= ymin) || !(lo < ymax)) continue;
            const int b0 = assoc_ybucket(lo - 2e-4f, ymin, yscale), b1 = assoc_ybucket(hi + 2e-4f, ymin, yscale);
            const unsigned long long bits = (1ull << (p - 1)) | (1ull << p);
            for (int bkt = b0; bkt <= b1; bkt++) atomicOr(&s_need2[bkt * ASSOC_ZONES + z], bits);
        }
    }
    __syncthreads();
    const float2 *base = fits ? s_proj : proj;
    const int *off = fits ? s_off : s_rs;
    for (int si = 0; si < nsets; si++) {
        const size_t sc = ((size_t)slot * VELO_NUM_KP_SETS + (set0 + si)) * B.C + cam;
        const int F = B.n_kp[sc];
        for (int k = tid; k < F; k += blockDim.x) {
            const float2 kp = B.kp[sc * B.F + k];
            const bool y_in = kp.y >= ymin && kp.y < ymax, kx_ok = kp.x == kp.x;
            const int bkt = y_in ? assoc_ybucket(kp.y, ymin, yscale) : ASSOC_YB;
            const int xb = assoc_xbucket(kp.x, xmin, xscale);
            const bool two_d = use2d && y_in && kx_ok;
            int last = -1, prev_s = -2, hit = 0;
            int h_s = 0, h_mid = 0, h_last = 0;          // the bracket of the hit; its interpolation (global gathers) runs after the search
            float2 pa = make_float2(0.f, 0.f), pb = pa;  // bracket of the previous ring
            for (int w = 0; w < NW && !hit; w++) {
                unsigned long long m = s_need[bkt][w];
                if (two_d) m &= s_need2[bkt * ASSOC_ZONES + assoc_zone(kp.x, xmin, zscale)];
                for (; m && !hit; m &= m - 1) {
                    const int s = (w << 6) + __ffsll((long long)m) - 1;
                    if (s != prev_s + 1) last = -1;                                  // a ring that was not searched lies in between
                    prev_s = s;
                    const int cnt = s_cnt[s];
                    const float2 *ps = base + off[s];
                    int mid = 0;
                    bool found = false;
                    float2 a = make_float2(0.f, 0.f), b = a;
                    if (fits && s_mono[s] && kx_ok) {
                        const unsigned short *lut = s_lut + s * (ASSOC_NB + 1);
                        int j = lut[xb];
                        const int j1 = lut[xb + 1];
                        // the keypoint's own bucket (~2 points): x does not decrease along the ring, so "x <= kp.x" holds for a prefix of the
                        // bucket and counting it four points at a time is one instruction stream for the whole warp
                        for (;;) {
                            const int j0 = j;
#pragma unroll
                            for (int t = 0; t < 4; t++) j += (int)((j0 + t < j1) & (ps[max(min(j0 + t, j1 - 1), 0)].x <= kp.x));
                            if (j != j0 + 4) break;
                        }
                        mid = j - 1;
                        found = mid >= 0 && mid <= cnt - 2;
                        if (found) { a = ps[mid]; b = ps[mid + 1]; }
                    } else {
                        // velo.h:404-412 with the same sequence of mid points
                        int lo = 0, hi = cnt - 2;
                        while (lo <= hi && !found) {
                            mid = (lo + hi) >> 1;
                            a = ps[mid]; b = ps[mid + 1];
                            const bool left = a.x > kp.x, right = !left && (b.x <= kp.x);
                            hi = left ? mid - 1 : hi; lo = right ? mid + 1 : lo;
                            found = !left && !right;
                        }
                    }
                    if (found) {
                        if (last != -1 && ((a.y > kp.y) != (pa.y > kp.y)) && width_ok(__fsub_rn(a.x, b.x), cal) && width_ok(__fsub_rn(pa.x, pb.x), cal)) {
                            hit = 1; h_s = s; h_mid = mid; h_last = last;            // velo.h:413-422
                        }
                        last = mid; pa = a; pb = b;                                  // velo.h:483
                    } else last = -1;                                                // velo.h:487-489
                }
            }
            B.hit_tmp[sc * B.F + k] = hit;
            if (hit) {
                const float2 *ps = base + off[h_s], *pq = base + off[h_s - 1];
                const float2 a = ps[h_mid], b = ps[h_mid + 1], c = pq[h_last], d = pq[h_last + 1];
                const float4 *vs = valid + s_rs[h_s], *vq = valid + s_rs[h_s - 1];
                const float3 i1 = lerp3(vs[h_mid], vs[h_mid + 1], a.x, b.x, kp.x);       // velo.h:445-450
                const float3 i2 = lerp3(vq[h_last], vq[h_last + 1], c.x, d.x, kp.x);     // velo.h:451-456
                const float i1y = lerp1(a.y, b.y, a.x, b.x, kp.x);                        // velo.h:457-462
                const float i2y = lerp1(c.y, d.y, c.x, d.x, kp.x);                        // velo.h:463-468
                const float3 r = lerp3(make_float4(i1.x, i1.y, i1.z, 0.f), make_float4(i2.x, i2.y, i2.z, 0.f), i1y, i2y, kp.y); // velo.h:470-475
                B.kpwd_tmp[sc * B.F + k] = make_float4(r.x, r.y, r.z, 1.0f);
            }
        }
    }
}

// stable compaction in keypoint order (hazard H13): has_depth[k] = running hit count, kpwd appended (velo.h:479-481)
__global__ void __launch_bounds__(256) k_assoc_compact(DevBuffers B, int slot0, int set0, int nsets, int cam0) {
    __shared__ int s_w[33];
    const int cam = cam0 + blockIdx.x;
    const int slot = slot0 + blockIdx.y / nsets, set = set0 + blockIdx.y % nsets;
    const size_t sc = ((size_t)slot * VELO_NUM_KP_SETS + set) * B.C + cam;
    const int F = B.n_kp[sc];
    const int *hit = B.hit_tmp + sc * B.F;
    int base = 0;
    for (int k0 = 0; k0 < F; k0 += blockDim.x) {
        const int k = k0 + threadIdx.x;
        int h = (k < F) ? hit[k] : 0, total;
        int ex = block_excl_scan(h, s_w, total);
        if (k < F) {
            B.has_depth[sc * B.F + k] = h ? base + ex : -1;
            if (h) B.kpwd[sc * B.F + base + ex] = B.kpwd_tmp[sc * B.F + k];
        }
        base += total;
    }
    if (threadIdx.x == 0) B.n_hits[sc] = base;
}

// ------------------------------------------------------------------------------------------------ launchers
#define PRE(k) do { if (L.pre) L.pre(L.user, (k)); } while (0)
#define POST(k) do { if (L.post) L.post(L.user, (k)); } while (0)

void launch_ingest(const Launcher &L, const DevBuffers &B, const DevCalib &cal, int slot0, int count) {
    dim3 g((B.N + 255) / 256, count);
    PRE(VK_INGEST_FLAGS); k_ingest_flags<<<g, 256, 0, L.stream>>>(B, slot0); POST(VK_INGEST_FLAGS);
    PRE(VK_INGEST_RINGS); k_ingest_rings<<<count, 1024, 0, L.stream>>>(B, slot0); POST(VK_INGEST_RINGS);
    PRE(VK_INGEST_PERMUTE); k_ingest_permute<<<g, 256, 0, L.stream>>>(B, cal, slot0); POST(VK_INGEST_PERMUTE);
}
void launch_index(const Launcher &L, const DevBuffers &B, const DevCalib &cal, int slot0, int count) {
    dim3 g(B.R, count);
    PRE(VK_INDEX_BUILD); k_index_build<<<g, 256, 0, L.stream>>>(B, cal, slot0); POST(VK_INDEX_BUILD);
    dim3 g2(VELO_SECTORS, count);
    PRE(VK_INDEX_MASKS); k_index_masks<<<g2, VELO_EL_BUCKETS, 0, L.stream>>>(B, slot0); POST(VK_INDEX_MASKS);
}
void launch_project(const Launcher &L, const DevBuffers &B, const DevCalib &cal, int slot0, int count) {
    dim3 g((B.R + PROJ_WARPS - 1) / PROJ_WARPS, count);
    PRE(VK_PROJECT);
    if (cal.num_cams <= 2) k_project<2><<<g, PROJ_WARPS * 32, 0, L.stream>>>(B, cal, slot0);
    else k_project<VELO_MAX_CAMS><<<g, PROJ_WARPS * 32, 0, L.stream>>>(B, cal, slot0);
    POST(VK_PROJECT);
}
void launch_assoc(const Launcher &L, const DevBuffers &B, const DevCalib &cal, int slot0, int count, int set0, int nsets, int cam0, int ncams) {
    dim3 g(ncams, count);
    cudaFuncSetAttribute(k_assoc_search, cudaFuncAttributeMaxDynamicSharedMemorySize, ASSOC_DYN_BYTES);
    PRE(VK_ASSOC_SEARCH); k_assoc_search<<<g, ASSOC_THREADS, ASSOC_DYN_BYTES, L.stream>>>(B, cal, slot0, set0, nsets, cam0); POST(VK_ASSOC_SEARCH);
    dim3 g2(ncams, count * nsets);
    PRE(VK_ASSOC_COMPACT); k_assoc_compact<<<g2, 256, 0, L.stream>>>(B, slot0, set0, nsets, cam0); POST(VK_ASSOC_COMPACT);
}
