// velo_icp.cu — stages 3+4 for the lidar term: transform_point (utility.h:97-103), the per-ring 1-NN / top-2 ring /
// third-point / normal logic of velo.h:806-874, cost3DPD residual + Jacobian (costfunctions.h:17-58) and the 6x6
// normal equations (velo.h:885-902), fused in one kernel.
//
// The reference asks 64 kd-trees for one nearest neighbour each and then keeps the two best rings.  That is
// equivalent to: over all target points within the distance threshold, ordered by the key (d2, ring, index),
//     i = the smallest key,        j = the smallest key whose ring differs from ring(i).
// (d2 as f32 bits orders like the value; ties go to the lower ring, then the lower point index — hazards H9 and the
// north-star tie rule.)  A min over keys is order independent, so any exhaustive enumeration gives bit-identical
// indices.  The enumeration is pruned, conservatively, with the ring x azimuth organisation of the scan: a ring can only
// contain a point within b of the query q if, inside the azimuth window asin(b/|q_xy|), its elevation interval comes
// within asin(b/|q|) of q's elevation and its range interval within b of q's range (both read from cumulative bucket
// masks, two loads each); b = current bound on sqrt(d2_j), seeded by the previous pass of the same frame pair.
// Measured design rule (tools/exp_icp.sh, tools/tune_*.sh): distance evaluations are cheap, control flow and extra
// rounds are not.
#include "velo_common.cuh"

#ifndef ICP_THREADS
#define ICP_THREADS 256
#endif
#ifndef ICP_MIN_BLOCKS
#define ICP_MIN_BLOCKS 3
#endif
#ifndef ICP_SCAN_CHUNK
#define ICP_SCAN_CHUNK (1 << 30)   /* measured (tools/tune_chunk.sh): capping the candidates per round at 8/16/32 is 31/15/7 % slower */
#endif
#ifndef ICP_UNROLL
#define ICP_UNROLL 4           /* candidate loads in flight per lane */
#endif
#ifndef ICP_TIGHT
#define ICP_TIGHT 0.03f        /* elevation tolerance (rad) up to which a query's ring set is taken in one mask level */
#endif
#ifndef ICP_LEV0
#define ICP_LEV0 0.013f        /* first elevation-tolerance level of a query that is not tight (about two ring spacings; 0.0065 measured 1 % slower) */
#endif
#ifndef ICP_LEVMUL
#define ICP_LEVMUL 4.0f        /* growth of the tolerance from level to level (2.0 measured 1 % slower) */
#endif
#ifndef ICP_CELLSEED
#define ICP_CELLSEED 1         /* cell seeds: 0 off, 1 first pass of a frame pair, 3 whenever the previous pass left no pair (measured 1 % slower than 1) */
#endif
#ifndef ICP_CELLSEED_TOL
#define ICP_CELLSEED_TOL 0.006f /* elevation tolerance (rad) of the rings a cell seed is taken from (about one ring spacing wide in total) */
#endif
#define VELO_STR_(x) #x
#define VELO_UNROLL(n) _Pragma(VELO_STR_(unroll n))
#define KEY_INF 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ u64 make_key(float d2, int ring, int idx) {
    return ((u64)__float_as_uint(d2) << 32) | (u64)(((unsigned)ring << VELO_IDX_BITS) | (unsigned)idx);
}
__device__ __forceinline__ int key_ring(u64 k) { return (int)(((unsigned)k) >> VELO_IDX_BITS); }
__device__ __forceinline__ int key_idx(u64 k) { return (int)(((unsigned)k) & ((1u << VELO_IDX_BITS) - 1u)); }
__device__ __forceinline__ float key_d2(u64 k) { return __uint_as_float((unsigned)(k >> 32)); }

struct Window { int b0, b1; bool wrapped, full; float gam, half; };

// upper bound of asin(x) for 0 <= x < 1:  asin(x) <= tan(asin(x)) = x / sqrt(1 - x^2)  (a few instructions instead of asinf;
// only used to size pruning windows, where "too large" is always safe)
__device__ __forceinline__ float asin_ub(float x) {
    if (!(x < 0.999f)) return 4.0f;
    return x * rsqrtf(1.0f - x * x) * (1.0f + 4e-6f) + 3e-5f;     // 2e-5 float slack + 1e-5 for atan2_q
}
// bins covered by [az - half, az + half]
__device__ __forceinline__ void set_bins(Window &w, float az, float half) {
    w.half = half; w.wrapped = false; w.full = false; w.b0 = 0; w.b1 = VELO_AZ_BINS - 1;
    const float lo = az - half, hi = az + half;
    if (!(half < CUDART_PI_F)) { w.full = true; return; }
    if (lo < -CUDART_PI_F) { w.b0 = az_bin(lo + 2.0f * CUDART_PI_F); w.b1 = az_bin(hi); w.wrapped = true; }
    else if (hi > CUDART_PI_F) { w.b0 = az_bin(lo); w.b1 = az_bin(hi - 2.0f * CUDART_PI_F); w.wrapped = true; }
    else { w.b0 = az_bin(lo); w.b1 = az_bin(hi); }
    if (w.wrapped && w.b0 <= w.b1 + 1) { w.full = true; w.wrapped = false; w.b0 = 0; w.b1 = VELO_AZ_BINS - 1; }
}
// azimuth window / elevation tolerance for bound d2b around a query with azimuth az, xy-range D, range rho
__device__ __forceinline__ Window make_window(float d2b, float az, float D, float rho) {
    Window w;
    const float b = sqrt_ap(d2b) * (1.0f + 1e-5f) + 1e-6f;
    w.gam = (b < rho) ? asin_ub(__fdividef(b, rho)) : 4.0f;
    set_bins(w, az, (b < D) ? asin_ub(__fdividef(b, D)) : 4.0f);
    return w;
}

// nearest candidate of one ring inside sorted[start, end): smallest d2, ties -> lower index in ring.  The running best is the
// pair (bits of d2, index) compared as one unsigned 64-bit number (d2 >= 0, so its bits order like its value): branch-free,
// whereas a "rare" improvement branch diverges in almost every iteration once 32 lanes share the loop.
// bd starts at the smallest float above the threshold with index 0, so the comparison also applies the threshold (velo.h:829).
__device__ __forceinline__ void scan_range(const float4 *__restrict__ sorted, int start, int end, float mx, float my, float mz, u64 &best, int &ncand) {
    ncand += max(end - start, 0);
    const u64 mxy = pack2f(mx, my);
    VELO_UNROLL(ICP_UNROLL)
    for (int p = start; p < end; p++) {
        const float4 c = __ldg(sorted + p);
        const float d2 = d2f_xy2(c, mxy, mz);
        const u64 k = ((u64)__float_as_uint(d2) << 32) | (u64)(unsigned)__float_as_int(c.w);
        best = min(best, k);
    }
}
__device__ __forceinline__ u64 scan_init(float thr_excl) { return (u64)__float_as_uint(thr_excl) << 32; }
__device__ __forceinline__ bool scan_found(u64 best, float thr_excl) { return (unsigned)(best >> 32) < __float_as_uint(thr_excl); }
__device__ __forceinline__ u64 scan_key(u64 best, int s) { return (best & 0xFFFFFFFF00000000ull) | (u64)(((unsigned)s << VELO_IDX_BITS) | (unsigned)best); }
// Candidate rings (64-ring word `word`) for a query at elevation el / range rho with search radius b: a ring qualifies in a
// sector of the window if its elevation interval comes within w.gam of el AND its range interval within b of rho
// (|q-p| >= |rq-rp| and |q-p| >= rq sin(angle), angle >= elevation gap).  Both tests are two loads from cumulative bucket
// masks (bucket functions are monotone, windows padded => conservative); no per-ring test is needed afterwards.
struct MaskQ { int e0, e1, r0, r1; };
__device__ __forceinline__ MaskQ mask_query(float el, float gam, float rho, float b) {
    MaskQ q;
    q.e0 = el_bucket(el - gam); q.e1 = el_bucket(el + gam);
    const float pad = b * (1.0f + 1e-5f) + 1e-5f * rho + 1e-6f;
    q.r0 = rg_bucket(rho - pad); q.r1 = rg_bucket(rho + pad);
    return q;
}
__device__ __forceinline__ u64 ring_mask(const u64 *__restrict__ mlo, const u64 *__restrict__ mhi, const u64 *__restrict__ rlo, const u64 *__restrict__ rhi,
                                         int W, int word, const Window &w, const MaskQ &q, bool use_range) {
    int sa = w.b0 / VELO_BINS_PER_SECTOR, sb = w.b1 / VELO_BINS_PER_SECTOR;
    if (w.full || (w.wrapped && sa == sb)) { sa = 0; sb = VELO_SECTORS - 1; }   // a wrapped window whose ends share a sector covers all the others
    u64 m = 0ull;
    for (int sec = sa;; sec = (sec + 1) & (VELO_SECTORS - 1)) {
        const int base = sec * VELO_EL_BUCKETS;
        u64 t = __ldg(mlo + (base + q.e1) * W + word) & __ldg(mhi + (base + q.e0) * W + word);
        if (use_range) t &= __ldg(rlo + (base + q.r1) * W + word) & __ldg(rhi + (base + q.r0) * W + word);
        m |= t;
        if (sec == sb) break;
    }
    return m;
}

// grid = (ctas per unit, n_units).  A unit is one frame pair with up to VELO_MAX_PASSES supplied poses (the ICP passes of
// frameToFrame, velo.h:616,800).  One thread = one query point at a time, looping over the passes: the correspondence of
// pass p (two real target points) seeds pass p+1 with an immediately tight bound, so only the first pass needs the probe.
// Work distribution: the unit's queries are cut into BLOCKS of ICP_WARPS x ICP_RUN_CHUNKS chunks of 32 consecutive queries; RUN w
// of a block is its chunks w, w + ICP_WARPS, w + 2 ICP_WARPS, ... (so the warps of a CTA that work on the runs of one block
// sweep neighbouring queries at the same time and share L1 lines).  A CTA owns a contiguous range of blocks and its warps take
// runs from a shared counter (search cost per query varies 10x; static striding left 13 % of the warp time waiting at the CTA's
// end).  Each run's sums go to its own record of `partial`, indexed by the run's number in the unit, and k_neq_reduce adds the
// records in run order: the result does not depend on which warp took which run.
#ifndef ICP_RUN_CHUNKS
#define ICP_RUN_CHUNKS 2            /* measured at the bench size: 2 chunks per run 1.4 % faster than 4 (finer balancing at the CTA's end), 8 is 2 % slower */
#endif
#define ICP_WARPS (ICP_THREADS / 32)
#define ICP_BLOCK_QUERIES (32 * ICP_RUN_CHUNKS * ICP_WARPS)
__host__ __device__ inline int icp_blocks(int queries) { return (queries + ICP_BLOCK_QUERIES - 1) / ICP_BLOCK_QUERIES; }
__host__ __device__ inline int icp_runs_cap(int max_points) { return icp_blocks(max_points) * ICP_WARPS; }

// RECORDS = false: the throughput instantiation of the batched front end (no per-query records are written; 1.3 % faster than
// carrying the dead record code)
// W1 = true: at most 64 rings, i.e. ring masks of ONE 64-bit word: the word loops and the visited-mask array disappear at compile
// time (the array otherwise lives in local memory because it is indexed by a runtime word number; measured 4 % of the kernel).
// STATS = true: candidates / rings / mask bits per pass are counted into slots 60..62 of the records (1.6 % of the kernel; off by default,
// velo_gpu_search_stats_enable)
// Shared memory of a CTA.  RINGS = ring capacity + 1 of the instantiation (65 for a context of at most 64 rings): with the
// staging area below that keeps three resident CTAs under 132 KB, i.e. one shared-memory carve-out step lower = 32 KB more L1.
#define ICP_STAGE (8 * 36)               /* per warp: 6 Jacobian columns + residual + weight, 32 rows padded to 36 */
template <int RINGS>
struct IcpShared {
    int q[RINGS];                        // query prefix per source ring
    int rsM[RINGS];
    int rsS[RINGS];                      // ring starts of the target scan (seeds and the third point read them every pass)
    double rows[ICP_THREADS / 32][ICP_STAGE];
    double acc[ICP_THREADS / 32][VELO_MAX_PASSES][56];               // sums of the warp's current run (accumulating in the run's global record
                                                                     // with store + REDG.ADD.F64 instead frees 21.5 KB = one more carve-out step of L1: measured equal, 82.9 vs 82.5 ms)
    unsigned stat[ICP_THREADS / 32][VELO_MAX_PASSES][5];             // per warp and run: no atomics
    IcpPass pass[VELO_MAX_PASSES];
    int next;
    double loss[4];                       // loss constants of the unit: a^2, 1/a^2, w a^2/2, w
};
template <bool RECORDS, bool W1, bool STATS, int RINGS>
__device__ __forceinline__ void icp_pass_body(IcpShared<RINGS> &sh, const DevBuffers &B, const DevCalib &cal, const IcpUnit *__restrict__ units,
                                              double *__restrict__ partial, int runs_cap, velo_icp_corr *__restrict__ corr, int corr_stride,
                                              IcpFrozen *__restrict__ frozen, int frozen_stride) {
    int (&s_q)[RINGS] = sh.q;
    int (&s_rsM)[RINGS] = sh.rsM;
    int (&s_rsS)[RINGS] = sh.rsS;
    double (&s_rows)[ICP_THREADS / 32][ICP_STAGE] = sh.rows;
    double (&s_acc)[ICP_THREADS / 32][VELO_MAX_PASSES][56] = sh.acc;
    unsigned (&s_stat)[ICP_THREADS / 32][VELO_MAX_PASSES][5] = sh.stat;
    IcpPass (&s_pass)[VELO_MAX_PASSES] = sh.pass;
    int &s_next = sh.next;
    double (&s_loss)[4] = sh.loss;
    const IcpUnit &U = units[blockIdx.y];
    if (threadIdx.x == 0) { s_next = 0; const double bb = U.loss_a * U.loss_a; s_loss[0] = bb; s_loss[1] = 1.0 / bb; s_loss[2] = 0.5 * U.weight * bb; s_loss[3] = U.weight; }
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int NP = U.n_pass;
    if (U.src_slot < 0) return;   // unit without a previous scan: contributes nothing (k_neq_reduce writes its zeros)
    const int nrM = B.n_rings[U.src_slot], nrS = B.n_rings[U.tgt_slot];
    const int *rsM = B.ring_start + (size_t)U.src_slot * (B.R + 1);
    const int *rsS = B.ring_start + (size_t)U.tgt_slot * (B.R + 1);
    const int skip = U.skip;
    for (int i = tid; i < (int)(NP * sizeof(IcpPass) / sizeof(double)); i += blockDim.x)
        reinterpret_cast<double *>(s_pass)[i] = reinterpret_cast<const double *>(U.pass)[i];
    for (int i = tid; i <= B.n_rings[U.tgt_slot]; i += blockDim.x) s_rsS[i] = rsS[i];
#define RS_S(i) s_rsS[i]
    for (int i = tid; i < (ICP_THREADS / 32) * VELO_MAX_PASSES * 56; i += blockDim.x) (&s_acc[0][0][0])[i] = 0.0;
    for (int i = tid; i < (ICP_THREADS / 32) * VELO_MAX_PASSES * 5; i += blockDim.x) (&s_stat[0][0][0])[i] = 0u;
    if (tid == 0) {
        int q = 0;
        for (int s = 0; s < nrM; s++) { s_q[s] = q; int r0 = rsM[s], L = rsM[s + 1] - r0; s_rsM[s] = r0; q += (L + skip - 1) / skip; }
        s_q[nrM] = q;
    }
    __syncthreads();
    const int Q = s_q[nrM];
    const int nblk = icp_blocks(Q), blk_per = (nblk + gridDim.x - 1) / gridDim.x;
    const int blk0 = blockIdx.x * blk_per, blk1 = min(nblk, blk0 + blk_per);
    const int q1 = Q;
    double *pbase = partial + (size_t)blockIdx.y * runs_cap * VELO_MAX_PASSES * 64;
    const float4 *ptsM = B.pts + (size_t)U.src_slot * B.N;
    const float4 *ptsS = B.pts + (size_t)U.tgt_slot * B.N;
    const float4 *sorted = B.sorted + (size_t)U.tgt_slot * B.N;
    const int *csS = B.cell_start + (size_t)U.tgt_slot * B.R * (VELO_AZ_BINS + 1);
    const int W = W1 ? 1 : B.W, WS = B.W;      // words visited / words per table entry (a context for 128 rings holds 64-ring scans too)
    const u64 *mloS = B.mask_lo + (size_t)U.tgt_slot * VELO_SECTORS * VELO_EL_BUCKETS * WS;
    const u64 *mhiS = B.mask_hi + (size_t)U.tgt_slot * VELO_SECTORS * VELO_EL_BUCKETS * WS;
    const u64 *rloS = B.rmask_lo + (size_t)U.tgt_slot * VELO_SECTORS * VELO_RG_BUCKETS * WS;
    const u64 *rhiS = B.rmask_hi + (size_t)U.tgt_slot * VELO_SECTORS * VELO_RG_BUCKETS * WS;

    for (;;) {
        int run = 0;
        if (lane == 0) run = atomicAdd(&s_next, 1);
        run = __shfl_sync(FULL, run, 0);
        const int blk = blk0 + run / ICP_WARPS, rw = run % ICP_WARPS;
        if (blk >= blk1) break;
        int nq = 0;
      for (int ci = 0; ci < ICP_RUN_CHUNKS; ci++) {
        const int qb = blk * ICP_BLOCK_QUERIES + (rw + ci * ICP_WARPS) * 32;
        if (qb >= q1) break;
        nq += min(q1 - qb, 32);
        const int q = qb + lane;
        const bool active = q < q1;
        int sm = 0, smi = 0;
        float4 pm = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {                                   // (sm, smi) of this query: velo.h:806-807
            int lo = 0, hi = nrM - 1;
            while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (s_q[mid] <= q) lo = mid; else hi = mid - 1; }
            sm = lo; smi = (q - s_q[sm]) * skip;
            pm = __ldg(ptsM + s_rsM[sm] + smi);
        }
        const double x0 = pm.x, x1 = pm.y, x2 = pm.z;
        u64 pki = KEY_INF, pkj = KEY_INF;               // correspondence of the previous pass (ring / index parts are the seeds)

        for (int ps = 0; ps < NP; ps++) {
            const IcpPass &P = s_pass[ps];
            const float thr_f = P.thr_f, thr_excl = P.thr_excl;
            bool kept = false;
            double J[6] = { 0, 0, 0, 0, 0, 0 }, res = 0.0, rho1 = 0.0, rho0h = 0.0;
            int st_exh = 0, st_rings = 0, st_mask = 0;
            // (all 32 lanes run the pass; lanes without a query are born finished so that the warp votes below stay uniform)
            // util::transform_point (utility.h:97-103) = ceres::AngleAxisRotatePoint in f64, op for op (hazard H8)
#define ICP_ROTATE(y0, y1, y2) \
            if (!P.pose.small_angle) { \
                const double c0 = __dsub_rn(__dmul_rn(P.pose.u[1], x2), __dmul_rn(P.pose.u[2], x1)); \
                const double c1 = __dsub_rn(__dmul_rn(P.pose.u[2], x0), __dmul_rn(P.pose.u[0], x2)); \
                const double c2 = __dsub_rn(__dmul_rn(P.pose.u[0], x1), __dmul_rn(P.pose.u[1], x0)); \
                const double dot = __dadd_rn(__dadd_rn(__dmul_rn(P.pose.u[0], x0), __dmul_rn(P.pose.u[1], x1)), __dmul_rn(P.pose.u[2], x2)); \
                const double tmp = __dmul_rn(dot, __dsub_rn(1.0, P.pose.c)); \
                y0 = __dadd_rn(__dadd_rn(__dmul_rn(x0, P.pose.c), __dmul_rn(c0, P.pose.s)), __dmul_rn(P.pose.u[0], tmp)); \
                y1 = __dadd_rn(__dadd_rn(__dmul_rn(x1, P.pose.c), __dmul_rn(c1, P.pose.s)), __dmul_rn(P.pose.u[1], tmp)); \
                y2 = __dadd_rn(__dadd_rn(__dmul_rn(x2, P.pose.c), __dmul_rn(c2, P.pose.s)), __dmul_rn(P.pose.u[2], tmp)); \
            } else { \
                y0 = __dadd_rn(x0, __dsub_rn(__dmul_rn(P.pose.w[1], x2), __dmul_rn(P.pose.w[2], x1))); \
                y1 = __dadd_rn(x1, __dsub_rn(__dmul_rn(P.pose.w[2], x0), __dmul_rn(P.pose.w[0], x2))); \
                y2 = __dadd_rn(x2, __dsub_rn(__dmul_rn(P.pose.w[0], x1), __dmul_rn(P.pose.w[1], x0))); \
            }
            float mx, my, mz;
            {
                double y0, y1, y2;
                ICP_ROTATE(y0, y1, y2)
                mx = __double2float_rn(__dadd_rn(y0, P.pose.t[0]));
                my = __double2float_rn(__dadd_rn(y1, P.pose.t[1]));
                mz = __double2float_rn(__dadd_rn(y2, P.pose.t[2]));
            }

            // ---- pruning geometry of the query in the index frame of the target scan
            float vx, vy, vz; idx_frame(cal, mx, my, mz, vx, vy, vz);
            const float dxy2 = fmaf(vx, vx, vy * vy), D = sqrt_ap(dxy2), rho = sqrt_ap(fmaf(vz, vz, dxy2));
            const float az = atan2_q(vy, vx), el = atan2_q(vz, D);

            u64 ki = KEY_INF, kj = KEY_INF;
            // Seeds only BOUND the search: two real target points of different rings within the threshold => the runner-up is at most as
            // far as the farther of them.  The search below starts from empty (ki, kj), visits every ring once and finds the seed points
            // again (they lie inside the window their own distance defines), so ring results merge without any same-ring case.
            float seed_bound = thr_f;
            if (active) {
                if (pkj != KEY_INF) {                   // the pair the previous pass of this frame pair chose
                    const float4 ca = __ldg(ptsS + RS_S(key_ring(pki)) + key_idx(pki)), cb = __ldg(ptsS + RS_S(key_ring(pkj)) + key_idx(pkj));
                    const float m = fmaxf(d2f(ca.x, ca.y, ca.z, mx, my, mz), d2f(cb.x, cb.y, cb.z, mx, my, mz));
                    if (m <= thr_f) seed_bound = m;
                }
                // Cell seeds (first pass of a frame pair, no pair yet): the target point stored for the query's own azimuth bin in the
                // two rings nearest its elevation — two cell lookups and two points instead of a search that starts from the full
                // threshold radius (measured: 180 -> 81 candidates per first-pass query).
                if ((ICP_CELLSEED == 1 && ps == 0) || (ICP_CELLSEED == 3 && pkj == KEY_INF)) {
                    const int b = az_bin(az);
                    const int base = (b / VELO_BINS_PER_SECTOR) * VELO_EL_BUCKETS;
                    const int eb0 = el_bucket(el - ICP_CELLSEED_TOL), eb1 = el_bucket(el + ICP_CELLSEED_TOL);
                    int s1 = -1, s2 = -1;
                    for (int wd = 0; wd < W && s2 < 0; wd++) {
                        u64 m = __ldg(mloS + (size_t)(base + eb1) * WS + wd) & __ldg(mhiS + (size_t)(base + eb0) * WS + wd);
                        if (m != 0ull && s1 < 0) { s1 = wd * 64 + __ffsll((long long)m) - 1; m &= m - 1; }
                        if (m != 0ull && s1 >= 0) s2 = wd * 64 + __ffsll((long long)m) - 1;
                    }
                    if (s1 >= 0 && s2 < 0) s2 = (s1 + 1 < nrS) ? s1 + 1 : s1 - 1;
                    if (s1 >= 0 && s2 >= 0) {
                        const int *csa = csS + s1 * (VELO_AZ_BINS + 1) + b, *csb = csS + s2 * (VELO_AZ_BINS + 1) + b;
                        const int a0 = __ldg(csa), a1 = __ldg(csa + 1), b0 = __ldg(csb), b1 = __ldg(csb + 1);
                        if (a1 > a0 && b1 > b0) {
                            // nearest point of either cell (a cell holds about two points; the order inside a cell is not fixed, a minimum is)
                            float da = CUDART_INF_F, db = CUDART_INF_F;
                            for (int p = a0; p < a1; p++) { const float4 c = __ldg(sorted + p); da = fminf(da, d2f(c.x, c.y, c.z, mx, my, mz)); }
                            for (int p = b0; p < b1; p++) { const float4 c = __ldg(sorted + p); db = fminf(db, d2f(c.x, c.y, c.z, mx, my, mz)); }
                            const float m = fmaxf(da, db);
                            if (m < seed_bound) seed_bound = m;
                        }
                    }
                }
            }
            // Exhaustive search: every ring that can hold a point within the current bound on d2_j (velo.h:825-848) is visited exactly
            // once, nearest elevation first (levels of growing tolerance; a single level when the bound is already tight); the bound,
            // the azimuth window and the elevation tolerance shrink whenever the runner-up improves.  (A separate probe phase that
            // first looked only at the query's own azimuth bin of the nearest rings, to start with a tight bound, was measured and
            // removed: without it the unseeded pass evaluates 164 instead of 74 candidates per query and the kernel is 7 % faster.)
            // Written as a warp-synchronous "advance / scan" loop: lanes first advance (cheap ring tests) until each holds a
            // candidate range, then all of them scan together, so the distance loop runs converged.
            float bound = seed_bound;
            Window w = make_window(bound, az, D, rho);
            {
                u64 V[4] = { 0ull, 0ull, 0ull, 0ull };
                u64 m = 0ull;
                int word = 0, p0 = 0, e0 = 0, p1 = 0, e1 = 0, s_cur = 0;
                u64 best = scan_init(thr_excl);
                float lev = 0.f, gcur = 0.f;
#ifdef EXP_NO_PHASE2
                bool started = false, have = false, fin = true;
#else
                bool started = false, have = false, fin = !active;
#endif
                const bool tight = w.gam <= ICP_TIGHT;   // seeded query: one mask level, window kept for the whole pass
                for (;;) {
                    while (!have && !fin) {
                        if (m == 0ull) {                                   // next (level, word)
                            if (started && word + 1 < W) word++;
                            else {
                                if (started && !(gcur < w.gam)) { fin = true; break; }   // every ring within the tolerance was visited
                                lev = started ? lev * ICP_LEVMUL : (tight ? 8.0f : ICP_LEV0);
                                started = true; word = 0; gcur = fminf(lev, w.gam);
                            }
                            m = ring_mask(mloS, mhiS, rloS, rhiS, WS, word, w, mask_query(el, gcur, rho, sqrt_ap(bound)), true) & ~V[word];
                            V[word] |= m;
                            continue;
                        }
                        const int s = word * 64 + __ffsll((long long)m) - 1; m &= m - 1; st_mask++;
                        const int *cs = csS + s * (VELO_AZ_BINS + 1);
                        if (!w.wrapped) { p0 = __ldg(cs + w.b0); e0 = __ldg(cs + w.b1 + 1); p1 = 0; e1 = 0; }
                        else { p0 = __ldg(cs + w.b0); e0 = __ldg(cs + VELO_AZ_BINS); p1 = __ldg(cs); e1 = __ldg(cs + w.b1 + 1); }
                        s_cur = s; have = true; best = scan_init(thr_excl);
                    }
                    if (!__any_sync(FULL, have)) break;
                    if (have) {
                        // at most ICP_SCAN_CHUNK candidates per round: lanes with long ranges continue in the next round while
                        // the others already advance to their next ring, which keeps the distance loop's trip counts uniform
                        if (p0 >= e0) { p0 = p1; e0 = e1; p1 = 0; e1 = 0; }
                        const int ec = min(e0, p0 + ICP_SCAN_CHUNK);
                        scan_range(sorted, p0, ec, mx, my, mz, best, st_exh);
                        p0 = ec;
                        if (p0 >= e0 && p1 >= e1) {                          // ring finished
                            st_rings++; have = false;
                            if (scan_found(best, thr_excl)) {
                                const u64 oj = kj;
                                {   // every ring is visited once: plain insertion into the two smallest keys, branch-free
                                    const u64 k = scan_key(best, s_cur), lo = min(k, ki), hi = max(k, ki);
                                    ki = lo; kj = min(kj, hi);
                                }
                                if (kj != oj) { bound = fminf(seed_bound, key_d2(kj)); if (!tight) w = make_window(bound, az, D, rho); }
                            }
                        }
                    }
                }
            }
            if (active) {
                pki = ki; pkj = kj;

                velo_icp_corr rec;
                rec.src_ring = sm; rec.src_idx = smi; rec.np_s_i = -1; rec.np_i = 0; rec.np_s_j = -1; rec.np_j = 0; rec.np_k = -1; rec.kept = 0;
                rec.normal[0] = rec.normal[1] = rec.normal[2] = 0.f; rec.v0[0] = rec.v0[1] = rec.v0[2] = 0.f; rec.residual = 0.0;
                if (ki != KEY_INF) { rec.np_s_i = key_ring(ki); rec.np_i = key_idx(ki); }
                if (kj != KEY_INF) { rec.np_s_j = key_ring(kj); rec.np_j = key_idx(kj); }
#ifdef EXP_NO_EPILOGUE
                if (false) {
#else
                if (ki != KEY_INF && kj != KEY_INF) {                        // velo.h:849-851
#endif
                    const int si = rec.np_s_i, ni = rec.np_i, sj = rec.np_s_j, nj = rec.np_j;
                    const int ri0 = RS_S(si), Ln = RS_S(si + 1) - ri0;
                    const int k1 = (ni + 1 == Ln) ? 0 : ni + 1, k2 = (ni == 0) ? Ln - 1 : ni - 1;   // (np_i +- 1) mod n, velo.h:852-854
                    const float4 a1 = __ldg(ptsS + ri0 + k1), a2 = __ldg(ptsS + ri0 + k2);
                    const float n1 = d2f(a1.x, a1.y, a1.z, mx, my, mz), n2 = d2f(a2.x, a2.y, a2.z, mx, my, mz);
                    const int nk = (n1 < n2) ? k1 : k2;                      // velo.h:859-863
                    rec.np_k = nk;
                    const float4 v0 = __ldg(ptsS + ri0 + ni), v1 = __ldg(ptsS + RS_S(sj) + nj), v2 = (n1 < n2) ? a1 : a2;
                    // Eigen::Vector3f (v1-v0).cross(v2-v0), norm(), operator/= (velo.h:868-874)
                    const float ax = __fsub_rn(v1.x, v0.x), ay = __fsub_rn(v1.y, v0.y), az3 = __fsub_rn(v1.z, v0.z);
                    const float bx = __fsub_rn(v2.x, v0.x), by = __fsub_rn(v2.y, v0.y), bz = __fsub_rn(v2.z, v0.z);
                    float nx = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az3, by));
                    float ny = __fsub_rn(__fmul_rn(az3, bx), __fmul_rn(ax, bz));
                    float nz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
                    const float nn = __fsqrt_rn(__fadd_rn(__fmul_rn(nx, nx), __fadd_rn(__fmul_rn(ny, ny), __fmul_rn(nz, nz))));
                    rec.v0[0] = v0.x; rec.v0[1] = v0.y; rec.v0[2] = v0.z;
                    if (nn < U.norm_thr_f) { rec.kept = 2; }                 // velo.h:873
                    else {
                        nx = __fdiv_rn(nx, nn); ny = __fdiv_rn(ny, nn); nz = __fdiv_rn(nz, nn);
                        rec.normal[0] = nx; rec.normal[1] = ny; rec.normal[2] = nz;
                        // cost3DPD (costfunctions.h:40-53): M = R p; M += t - o; r = M . n
                        const double dnx = nx, dny = ny, dnz = nz;
                        // (the rotated point is formed again here rather than kept in six registers across the whole search)
                        double y0, y1, y2;
                        ICP_ROTATE(y0, y1, y2)
                        const double m0 = y0 + (P.pose.t[0] - (double)v0.x), m1 = y1 + (P.pose.t[1] - (double)v0.y), m2 = y2 + (P.pose.t[2] - (double)v0.z);
                        // (f64 residual / Jacobian rows are compared at 1e-5 relative and agree to ~1e-15 either way: explicit fused multiply-adds)
                        res = fma(m0, dnx, fma(m1, dny, m2 * dnz));
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            const double *d = P.pose.dR + 9 * k;
                            const double e0 = fma(d[0], x0, fma(d[1], x1, d[2] * x2)), e1 = fma(d[3], x0, fma(d[4], x1, d[5] * x2)), e2 = fma(d[6], x0, fma(d[7], x1, d[8] * x2));
                            J[k] = fma(e0, dnx, fma(e1, dny, e2 * dnz));
                        }
                        J[3] = dnx; J[4] = dny; J[5] = dnz;
                        // ScaledLoss(CauchyLoss(a), w) (velo.h:885-891)
                        const double sum = 1.0 + res * res * s_loss[1], inv = 1.0 / sum;
                        rho1 = s_loss[3] * fmax(2.2250738585072014e-308, inv);
                        rho0h = s_loss[2] * log(sum);
                        rec.kept = 1; rec.residual = res;
                        kept = true;
                    }
                }
                if (RECORDS && frozen && ps == NP - 1) {      // compact record for the device-resident solve (two 16-byte stores)
                    IcpFrozen f;
                    f.n[0] = rec.normal[0]; f.n[1] = rec.normal[1]; f.n[2] = rec.normal[2]; f.src = s_rsM[sm] + smi;
                    f.o[0] = rec.v0[0]; f.o[1] = rec.v0[1]; f.o[2] = rec.v0[2]; f.kept = rec.kept;
                    frozen[(size_t)blockIdx.y * frozen_stride + q] = f;
                }
                if (RECORDS && corr && (corr_stride > 0 || ps == NP - 1)) {      // corr_stride > 0: the records of EVERY pass, [pass][corr_stride]
#pragma unroll
                    for (int k = 0; k < 6; k++) rec.jacobian[k] = J[k];
#ifdef VELO_ICP_DEBUG   /* tools/icp_debug_hist.py: per-query search statistics instead of the Jacobian */
                    rec.jacobian[3] = 0.0; rec.jacobian[4] = st_exh; rec.jacobian[5] = st_rings + 1000.0 * st_mask;
#endif
                    corr[(size_t)(corr_stride > 0 ? ps : 0) * corr_stride + q] = rec;
                }
            }
            // ---- normal equations of this pass: rows of the 32 lanes -> the 28+28 sums of the (pass, warp) record
#ifndef EXP_NO_ACCUM
            // X^T W X of the warp's 32 rows (X = [J, r, 0], 8 columns) on the FP64 tensor pipe: eight m8n8k4 steps, each lane
            // reads ONE staged element per step (it is both its A and its B fragment entry) plus the row weight; fixed order =>
            // run-to-run deterministic.  (Measured: 1.3 ms of 27 faster per 200 frame pairs than walking the rows on the FP64 ALU.)
            {
                double *S = s_rows[wid];
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 6; c++) S[c * 36 + lane] = kept ? J[c] : 0.0;
                S[6 * 36 + lane] = kept ? res : 0.0; S[7 * 36 + lane] = kept ? rho1 : 0.0;     // (column 7 of X is zero and is not staged)
                __syncwarp();
                const int fr = lane >> 2, fk = lane & 3;
                double cr0 = 0.0, cr1 = 0.0, cw0 = 0.0, cw1 = 0.0;
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const double wgt = S[7 * 36 + 4 * t + fk], x = fr < 7 ? S[fr * 36 + 4 * t + fk] : 0.0;
                    const double xw = x * wgt;
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(cr0), "+d"(cr1) : "d"(x), "d"(x));
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(cw0), "+d"(cw1) : "d"(x), "d"(xw));
                }
                double ch = kept ? rho0h : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ch += __shfl_xor_sync(FULL, ch, o);
                // lane holds C[fr][2 fk], C[fr][2 fk + 1]; upper triangle -> record slots (H row-major upper, then g, then cost)
                double *rec = s_acc[wid][ps];
                const int c0 = 2 * fk, c1 = c0 + 1;
                if (fr < 6) {
                    const int base = fr * 6 - (fr * (fr - 1)) / 2 - fr;
                    if (c0 >= fr && c0 < 6) { rec[base + c0] += cw0; rec[28 + base + c0] += cr0; }
                    if (c1 >= fr && c1 < 6) { rec[base + c1] += cw1; rec[28 + base + c1] += cr1; }
                    if (c0 == 6) { rec[21 + fr] += cw0; rec[28 + 21 + fr] += cr0; }
                } else if (fr == 6 && c0 == 6) { rec[27] += ch; rec[55] += 0.5 * cr0; }
            }
#endif
            {   // kept count of the pass; with STATS also the search statistics (one REDUX per counter, lane 0 adds them to the warp's own counters)
                const unsigned r_kept = __popc(__ballot_sync(FULL, kept));
                if (STATS) {
                    const unsigned r_exh = __reduce_add_sync(FULL, (unsigned)st_exh);
                    const unsigned r_rings = __reduce_add_sync(FULL, (unsigned)st_rings), r_mask = __reduce_add_sync(FULL, (unsigned)st_mask);
                    if (lane == 0) { unsigned *st = s_stat[wid][ps]; st[0] += r_kept; st[2] += r_exh; st[3] += r_rings; st[4] += r_mask; }
                } else if (lane == 0) s_stat[wid][ps][0] += r_kept;
            }
        }
      }
        // flush the run: record [run number in the unit][pass] = {56 sums, kept, kept, queries, 4 search counters}; then clear
        __syncwarp();
        {
            double *po = pbase + (size_t)(blk * ICP_WARPS + rw) * VELO_MAX_PASSES * 64;
            for (int i = lane; i < NP * 64; i += 32) {
                const int ps = i >> 6, t = i & 63;
                double v = 0.0;
                if (t < 56) { v = s_acc[wid][ps][t]; s_acc[wid][ps][t] = 0.0; }
                else if (t == 58) v = (double)nq;
                else if (t < 63) { const int k = (t <= 57) ? 0 : t - 58; v = (double)s_stat[wid][ps][k]; }
                po[i] = v;
            }
            __syncwarp();
            for (int i = lane; i < NP * 5; i += 32) (&s_stat[wid][0][0])[i] = 0u;
            __syncwarp();
        }
    }
}

// The kernel proper: a frame pair whose target scan has at most 64 rings takes the single-word body even in a context created for more
// (the drop-in default is 128 rings, because kitti.h:166-173 can split a sweep into a few more than 64); uniform per CTA.
// (ONE = true: a context for at most 64 rings — only the single-word body is compiled in, which keeps its register allocation to itself:
// 0.6 % at the bench configuration.)
template <bool RECORDS, bool STATS, bool ONE>
__global__ void __launch_bounds__(ICP_THREADS, ICP_MIN_BLOCKS) k_icp_pass(DevBuffers B, DevCalib cal, const IcpUnit *__restrict__ units,
                                                          double *__restrict__ partial, int runs_cap, velo_icp_corr *__restrict__ corr, int corr_stride,
                                                          IcpFrozen *__restrict__ frozen, int frozen_stride) {
    constexpr int RINGS = (ONE ? 64 : VELO_MAX_RINGS_HARD) + 1;
    __shared__ IcpShared<RINGS> sh;
    const int tgt = units[blockIdx.y].tgt_slot;
    if (ONE || (tgt >= 0 && B.n_rings[tgt] <= 64)) icp_pass_body<RECORDS, true, STATS, RINGS>(sh, B, cal, units, partial, runs_cap, corr, corr_stride, frozen, frozen_stride);
    else if (!ONE) icp_pass_body<RECORDS, false, STATS, RINGS>(sh, B, cal, units, partial, runs_cap, corr, corr_stride, frozen, frozen_stride);
}

// fixed-order sum of the per-run records of one (unit, pass): out[unit][pass][0..62]
__global__ void __launch_bounds__(256) k_neq_reduce(DevBuffers B, const IcpUnit *__restrict__ units, const double *__restrict__ partial, int runs_cap,
                                                    double *__restrict__ out, int out_stride_passes) {
    __shared__ double s_part[4][64];
    __shared__ int s_runs;
    const int u = blockIdx.x, ps = blockIdx.y, t = threadIdx.x & 63, g = threadIdx.x >> 6;
    const IcpUnit &U = units[u];
    if (threadIdx.x == 0) {
        int q = 0;
        if (U.src_slot >= 0 && ps < U.n_pass) {
            const int *rs = B.ring_start + (size_t)U.src_slot * (B.R + 1);
            const int nr = B.n_rings[U.src_slot];
            for (int s = 0; s < nr; s++) q += (rs[s + 1] - rs[s] + U.skip - 1) / U.skip;
        }
        s_runs = icp_blocks(q) * ICP_WARPS;
    }
    __syncthreads();
    const int runs = s_runs;
    const double *p = partial + ((size_t)u * runs_cap * VELO_MAX_PASSES + ps) * 64 + t;
    double s = 0.0;                                      // four interleaved partial sums (run % 4), combined in a fixed order
    for (int r = g; r < runs; r += 4) s += p[(size_t)r * VELO_MAX_PASSES * 64];
    s_part[g][t] = s;
    __syncthreads();
    if (threadIdx.x < VELO_NEQ_STRIDE) {
        double v = 0.0;
        if (t < 63) v = ((s_part[0][t] + s_part[1][t]) + s_part[2][t]) + s_part[3][t];
        out[((size_t)u * out_stride_passes + ps) * VELO_NEQ_STRIDE + threadIdx.x] = v;
    }
}

int launch_icp_runs_cap(int max_points) { return icp_runs_cap(max_points); }

void launch_icp(const Launcher &L, const DevBuffers &B, const DevCalib &cal, const IcpUnit *units, int n_units, int n_pass, int ctas,
                double *partial, double *out, int out_stride_passes, velo_icp_corr *corr, int corr_stride, IcpFrozen *frozen, int frozen_stride, bool stats) {
    if (n_units <= 0) return;
    const int runs_cap = icp_runs_cap(B.N);
    dim3 g(ctas, n_units);
    if (L.pre) L.pre(L.user, VK_ICP_PASS);
    const bool rec = corr || frozen;
    // (The single-word kernel's three resident CTAs fit the 132 KB shared-memory carve-out, which the driver selects on its own; measured
    // with the carve-out forced: 132 KB 82.5 ms, 164 KB 86.1 / 88.3 ms, 228 KB 91.8 ms per 1000 pairs x 6 passes — L1 capacity matters.)
#define ICP_LAUNCH(R_, S_) do { \
                                 if (B.W == 1) k_icp_pass<R_, S_, true><<<g, ICP_THREADS, 0, L.stream>>>(B, cal, units, partial, runs_cap, rec ? corr : nullptr, rec ? corr_stride : 0, rec ? frozen : nullptr, rec ? frozen_stride : 0); \
                                 else k_icp_pass<R_, S_, false><<<g, ICP_THREADS, 0, L.stream>>>(B, cal, units, partial, runs_cap, rec ? corr : nullptr, rec ? corr_stride : 0, rec ? frozen : nullptr, rec ? frozen_stride : 0); } while (0)
    if (stats) { if (rec) ICP_LAUNCH(true, true); else ICP_LAUNCH(false, true); }
    else { if (rec) ICP_LAUNCH(true, false); else ICP_LAUNCH(false, false); }
#undef ICP_LAUNCH
    if (L.post) L.post(L.user, VK_ICP_PASS);
    dim3 g2(n_units, n_pass);
    if (L.pre) L.pre(L.user, VK_NEQ_REDUCE);
    k_neq_reduce<<<g2, 256, 0, L.stream>>>(B, units, partial, runs_cap, out, out_stride_passes);
    if (L.post) L.post(L.user, VK_NEQ_REDUCE);
}
