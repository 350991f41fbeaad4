/*
 * velo_synth.c — seeded synthetic KITTI-shaped data for the VELO front end.
 *
 * Host-only C (no CUDA, no oracle): it produces the INPUTS that both the CUDA
 * path and the CPU oracle consume, following the generator spec of
 * SURVEY.md §8(d):
 *   - HDL-64E model: 64 lasers (+2.0°..-8.33° in 1/3° steps, -8.83°..-24.33°
 *     in 1/2° steps), 2083 azimuth steps CCW from +x (velodyne frame: x fwd,
 *     y left, z up), emitted ring-major so the ring-boundary test of the
 *     reference (kitti.h:166) fires once per ring; stored as the KITTI .bin
 *     float4 {x,y,z,reflectance} that kitti.h:121-152 reads.
 *   - scene: ground plane + axis-aligned boxes (buildings, cars, poles),
 *     periodic along the driving direction so any frame number is valid.
 *   - calibration: KITTI-odometry-like P0..P3 / Tr.
 *   - features: two keypoint sets per (frame, camera) in canonical camera
 *     coordinates (velo.h:10-17): set A = "detected in this frame", set B =
 *     "set A of the previous frame tracked into this frame"; identity matches
 *     B_t[i] <-> A_{t-1}[i] feed the visual residuals (velo.h:622-792).
 *   - poses: ground-truth relative motion (cam-0 frame, angle-axis +
 *     translation, maps frame t points into frame t-1) plus a perturbation,
 *     i.e. the "supplied pose" of the throughput benchmark.
 *
 * Everything is a pure function of (seed, frame, rig, F).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SYN_LASERS 64
#define SYN_AZ 2083
#define SYN_PERIOD 48.0
#define SYN_MAX_RANGE 80.0
#define SYN_MIN_RANGE 2.5
#define SYN_GROUND_Y 1.68 /* world y (down) of the ground plane, cam-0 at y=0 */

/* ---------------- PRNG: splitmix64 -> uniform doubles ---------------- */
typedef struct { uint64_t s; } syn_rng;
static inline uint64_t syn_next(syn_rng *r) {
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline double syn_u01(syn_rng *r) { return (double)(syn_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline double syn_sym(syn_rng *r, double a) { return (2.0 * syn_u01(r) - 1.0) * a; }
static inline syn_rng syn_seed(uint64_t seed, uint64_t stream, uint64_t frame) {
    syn_rng r; r.s = seed * 0xD1342543DE82EF95ULL + stream * 0x2545F4914F6CDD1DULL + frame * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
    syn_next(&r); syn_next(&r);
    return r;
}

/* ---------------- small linear algebra (double) ---------------- */
typedef struct { double R[9]; double t[3]; } syn_se3; /* x_out = R x + t, row-major R */

static void aa_to_R(const double w[3], double R[9]) {
    double th2 = w[0]*w[0] + w[1]*w[1] + w[2]*w[2];
    if (th2 < 1e-24) { R[0]=1;R[1]=-w[2];R[2]=w[1];R[3]=w[2];R[4]=1;R[5]=-w[0];R[6]=-w[1];R[7]=w[0];R[8]=1; return; }
    double th = sqrt(th2), c = cos(th), s = sin(th), k = 1.0 - c;
    double x = w[0]/th, y = w[1]/th, z = w[2]/th;
    R[0]=c+x*x*k;   R[1]=x*y*k-z*s; R[2]=x*z*k+y*s;
    R[3]=y*x*k+z*s; R[4]=c+y*y*k;   R[5]=y*z*k-x*s;
    R[6]=z*x*k-y*s; R[7]=z*y*k+x*s; R[8]=c+z*z*k;
}
static void R_to_aa(const double R[9], double w[3]) {
    /* small rotations only (generator poses): robust enough via the skew part */
    double sx = 0.5*(R[7]-R[5]), sy = 0.5*(R[2]-R[6]), sz = 0.5*(R[3]-R[1]);
    double s = sqrt(sx*sx+sy*sy+sz*sz);
    double c = 0.5*(R[0]+R[4]+R[8]-1.0);
    double th = atan2(s, c);
    double f = (s > 1e-12) ? th/s : 1.0;
    w[0]=sx*f; w[1]=sy*f; w[2]=sz*f;
}
static void se3_apply(const syn_se3 *T, const double x[3], double y[3]) {
    for (int i=0;i<3;i++) y[i] = T->R[3*i]*x[0] + T->R[3*i+1]*x[1] + T->R[3*i+2]*x[2] + T->t[i];
}
static void se3_rot(const syn_se3 *T, const double x[3], double y[3]) {
    for (int i=0;i<3;i++) y[i] = T->R[3*i]*x[0] + T->R[3*i+1]*x[1] + T->R[3*i+2]*x[2];
}
static void se3_mul(const syn_se3 *A, const syn_se3 *B, syn_se3 *C) { /* C = A*B */
    syn_se3 o;
    for (int i=0;i<3;i++) for (int j=0;j<3;j++) {
        o.R[3*i+j] = A->R[3*i]*B->R[j] + A->R[3*i+1]*B->R[3+j] + A->R[3*i+2]*B->R[6+j];
    }
    se3_apply(A, B->t, o.t);
    *C = o;
}
static void se3_inv(const syn_se3 *A, syn_se3 *B) {
    syn_se3 o;
    for (int i=0;i<3;i++) for (int j=0;j<3;j++) o.R[3*i+j] = A->R[3*j+i];
    for (int i=0;i<3;i++) o.t[i] = -(o.R[3*i]*A->t[0] + o.R[3*i+1]*A->t[1] + o.R[3*i+2]*A->t[2]);
    *B = o;
}

/* ---------------- calibration ---------------- */
static const double SYN_FX = 718.856, SYN_CX = 607.1928, SYN_CY = 185.2157;
static const double SYN_TR[12] = {
    4.2768e-04, -9.99967e-01, -8.0845e-03, -1.19846e-02,
   -7.21063e-03, 8.0812e-03, -9.99941e-01, -5.40398e-02,
    9.99974e-01, 4.8595e-04, -7.2069e-03, -2.92197e-01 };

/* rig 0: KITTI (P0,P1 grey stereo, P2,P3 colour); rig 1: off-road, 4 cams with lateral offsets */
void velo_synth_calib(int rig, float P[48], float Tr[12], int *img_w, int *img_h) {
    static const double tx_kitti[4] = {0.0, -386.1448, 45.38225, -337.2877};
    static const double off_m[4]    = {0.0, -0.537, 0.30, -0.30}; /* metres, rig 1 */
    for (int c=0;c<4;c++) {
        float *p = P + 12*c;
        memset(p, 0, 12*sizeof(float));
        p[0] = (float)SYN_FX; p[2] = (float)SYN_CX; p[5] = (float)SYN_FX; p[6] = (float)SYN_CY; p[10] = 1.0f;
        p[3] = (float)(rig == 0 ? tx_kitti[c] : off_m[c]*SYN_FX);
    }
    for (int i=0;i<12;i++) Tr[i] = (float)SYN_TR[i];
    *img_w = 1241; *img_h = 376;
}
static double syn_cam_tx(int rig, int cam) { /* K^-1 * P[:,3], x component (metres) */
    static const double tx_kitti[4] = {0.0, -386.1448, 45.38225, -337.2877};
    static const double off_m[4]    = {0.0, -0.537, 0.30, -0.30};
    return rig == 0 ? tx_kitti[cam]/SYN_FX : off_m[cam];
}

/* ---------------- scene ---------------- */
typedef struct { double lo[3], hi[3]; } syn_box; /* world axes: x right, y down, z forward */
#define SYN_BOXES_PER_PERIOD 20
static int syn_scene_period(uint64_t seed, syn_box *b) {
    syn_rng r = syn_seed(seed, 777, 0);
    int n = 0;
    /* buildings: 3 per side, gaps of a few metres between them */
    for (int side=-1; side<=1; side+=2) {
        double z = syn_u01(&r)*2.0;
        for (int k=0;k<3;k++) {
            double len = 11.0 + syn_u01(&r)*3.0, gap = 1.5 + syn_u01(&r)*2.0;
            double x0 = 8.0 + syn_u01(&r)*3.0, depth = 8.0, h = 6.0 + syn_u01(&r)*6.0;
            syn_box q;
            q.lo[0] = side>0 ? x0 : -(x0+depth); q.hi[0] = side>0 ? x0+depth : -x0;
            q.lo[1] = SYN_GROUND_Y - h; q.hi[1] = SYN_GROUND_Y;
            q.lo[2] = z; q.hi[2] = z+len;
            b[n++] = q; z += len + gap;
        }
    }
    /* parked cars: 4 per side */
    for (int side=-1; side<=1; side+=2) {
        for (int k=0;k<4;k++) {
            double z = 2.0 + k*11.5 + syn_u01(&r)*4.0, x0 = 3.2 + syn_u01(&r)*0.8;
            syn_box q;
            q.lo[0] = side>0 ? x0 : -(x0+1.8); q.hi[0] = side>0 ? x0+1.8 : -x0;
            q.lo[1] = SYN_GROUND_Y - 1.5; q.hi[1] = SYN_GROUND_Y;
            q.lo[2] = z; q.hi[2] = z+4.2;
            b[n++] = q;
        }
    }
    /* poles: 3 per side */
    for (int side=-1; side<=1; side+=2) {
        for (int k=0;k<3;k++) {
            double z = 5.0 + k*15.0 + syn_u01(&r)*6.0, x0 = 5.6 + syn_u01(&r)*1.0;
            syn_box q;
            q.lo[0] = side>0 ? x0 : -(x0+0.3); q.hi[0] = side>0 ? x0+0.3 : -x0;
            q.lo[1] = SYN_GROUND_Y - 5.0; q.hi[1] = SYN_GROUND_Y;
            q.lo[2] = z; q.hi[2] = z+0.3;
            b[n++] = q;
        }
    }
    return n; /* 6 + 8 + 6 = 20 */
}

/* nearest hit of ray o + s*d (|d|=1) with ground + periodic boxes; returns s or -1 */
static double syn_raycast(const syn_box *boxes, int nb, const double o[3], const double d[3], double smax) {
    double best = smax;
    int hit = 0;
    if (d[1] > 1e-9) { /* ground: y = GROUND_Y (y points down) */
        double s = (SYN_GROUND_Y - o[1]) / d[1];
        if (s > 0 && s < best) { best = s; hit = 1; }
    }
    double inv[3];
    for (int a=0;a<3;a++) inv[a] = (fabs(d[a]) > 1e-12) ? 1.0/d[a] : 1e12 * (d[a] < 0 ? -1.0 : 1.0);
    /* periods that the ray can reach within best */
    double z0 = o[2], z1 = o[2] + d[2]*best;
    if (z0 > z1) { double t = z0; z0 = z1; z1 = t; }
    int k0 = (int)floor(z0 / SYN_PERIOD) - 1, k1 = (int)floor(z1 / SYN_PERIOD);
    for (int k=k0; k<=k1; k++) {
        double zoff = k * SYN_PERIOD;
        for (int i=0;i<nb;i++) {
            const syn_box *q = &boxes[i];
            double tmin = 0.0, tmax = best;
            int ok = 1;
            for (int a=0;a<3;a++) {
                double lo = q->lo[a] + (a==2 ? zoff : 0.0), hi = q->hi[a] + (a==2 ? zoff : 0.0);
                double t1 = (lo - o[a]) * inv[a], t2 = (hi - o[a]) * inv[a];
                if (t1 > t2) { double t = t1; t1 = t2; t2 = t; }
                if (t1 > tmin) tmin = t1;
                if (t2 < tmax) tmax = t2;
                if (tmin > tmax) { ok = 0; break; }
            }
            if (ok && tmin > 1e-6 && tmin < best) { best = tmin; hit = 1; }
        }
    }
    return hit ? best : -1.0;
}

/* ---------------- trajectory ---------------- */
/* cam-0 world pose of frame t: forward drift 1 m/frame + small seeded wobble */
static void syn_cam_pose(uint64_t seed, int64_t frame, syn_se3 *C) {
    syn_rng r = syn_seed(seed, 11, (uint64_t)(frame + 1000000));
    double w[3] = { syn_sym(&r, 0.010), syn_sym(&r, 0.010), syn_sym(&r, 0.010) };
    aa_to_R(w, C->R);
    C->t[0] = syn_sym(&r, 0.025);
    C->t[1] = syn_sym(&r, 0.025);
    C->t[2] = (double)frame * 1.0 + syn_sym(&r, 0.10);
}
static void syn_velo_to_cam(syn_se3 *V) {
    for (int i=0;i<3;i++) { for (int j=0;j<3;j++) V->R[3*i+j] = SYN_TR[4*i+j]; V->t[i] = SYN_TR[4*i+3]; }
}

/* ground-truth relative pose: p_{t-1} = R(w) p_t + tau, cam-0 frame. out[0:3]=w, out[3:6]=tau */
void velo_synth_pose(uint64_t seed, int frame, double out[6]) {
    syn_se3 C1, C0, C0i, T;
    syn_cam_pose(seed, frame, &C1);
    syn_cam_pose(seed, (int64_t)frame - 1, &C0);
    se3_inv(&C0, &C0i);
    se3_mul(&C0i, &C1, &T);
    R_to_aa(T.R, out);
    out[3] = T.t[0]; out[4] = T.t[1]; out[5] = T.t[2];
}
/* "supplied pose" for ICP pass `pass`: truth + shrinking perturbation (mimics solver iterates) */
void velo_synth_pose_guess(uint64_t seed, int frame, int pass, double out[6]) {
    velo_synth_pose(seed, frame, out);
    syn_rng r = syn_seed(seed, 23 + (uint64_t)pass, (uint64_t)frame);
    double s = 1.0 / (double)(1 + pass);
    for (int i=0;i<3;i++) out[i] += syn_sym(&r, 0.004) * s;
    for (int i=3;i<6;i++) out[i] += syn_sym(&r, 0.04) * s;
}

/* the same with a selectable spread.  spread 0: the tight guesses above.  spread 1 ("spec"): pass 0 is the
 * reference's own start (0,0,0,0,0,1) (main.cpp:170); pass p >= 1 is the truth + w ~ U(+-0.02 rad)^3, t ~ (U(+-0.05), U(+-0.05),
 * U(+-0.2)) m, shrinking 1/(1+p) like solver iterates do. */
void velo_synth_pose_guess_spread(uint64_t seed, int frame, int pass, int spread, double out[6]) {
    if (spread == 0) { velo_synth_pose_guess(seed, frame, pass, out); return; }
    if (pass == 0) { out[0] = out[1] = out[2] = out[3] = out[4] = 0.0; out[5] = 1.0; return; }
    velo_synth_pose(seed, frame, out);
    syn_rng r = syn_seed(seed, 41 + (uint64_t)pass, (uint64_t)frame);
    double s = 1.0 / (double)(1 + pass);
    for (int i=0;i<3;i++) out[i] += syn_sym(&r, 0.02) * s;
    out[3] += syn_sym(&r, 0.05) * s; out[4] += syn_sym(&r, 0.05) * s; out[5] += syn_sym(&r, 0.2) * s;
}

/* ---------------- lidar scan ---------------- */
static double syn_elev_deg(int k) { return k < 32 ? 2.0 - k/3.0 : -8.83 - (k-32)/2.0; }

/* writes up to max_points float4 {x,y,z,refl} (velodyne frame), returns count */
int velo_synth_scan(uint64_t seed, int frame, float *xyzr, int max_points) {
    syn_box boxes[SYN_BOXES_PER_PERIOD];
    int nb = syn_scene_period(seed, boxes);
    syn_se3 C, V, W; /* W: velodyne coords -> world coords = C * velo_to_cam */
    syn_cam_pose(seed, frame, &C);
    syn_velo_to_cam(&V);
    se3_mul(&C, &V, &W);
    double zero[3] = {0,0,0}, o[3];
    se3_apply(&W, zero, o);
    syn_rng r = syn_seed(seed, 5, (uint64_t)frame);
    int n = 0;
    const double PI = 3.14159265358979323846;
    for (int k=0;k<SYN_LASERS;k++) {
        double el = syn_elev_deg(k) * PI/180.0, ce = cos(el), se = sin(el);
        double phase = 0.25 + 0.5*syn_u01(&r);
        for (int j=0;j<SYN_AZ;j++) {
            double az = (j + phase) * (2.0*PI/SYN_AZ);
            double dv[3] = { ce*cos(az), ce*sin(az), se }, dw[3];
            se3_rot(&W, dv, dw);
            double noise = syn_sym(&r, 0.03);
            double refl = syn_u01(&r);
            double drop = syn_u01(&r);
            double s = syn_raycast(boxes, nb, o, dw, SYN_MAX_RANGE);
            if (s < SYN_MIN_RANGE || drop < 0.08) continue; /* no-return rays: lands at N = 120k +- 2% */
            s += noise;
            if (n >= max_points) return n;
            xyzr[4*n+0] = (float)(dv[0]*s);
            xyzr[4*n+1] = (float)(dv[1]*s);
            xyzr[4*n+2] = (float)(dv[2]*s);
            xyzr[4*n+3] = (float)refl;
            n++;
        }
    }
    return n;
}

/* ---------------- features ---------------- */
/* 3-D anchor (cam-0 frame of `frame`) of feature i of set A in camera cam; returns 0 if the ray hits nothing */
static int syn_feature_anchor(uint64_t seed, int frame, int rig, int cam, int i,
                              const syn_box *boxes, int nb, const syn_se3 *C,
                              double uv[2], double pc[3]) {
    syn_rng r = syn_seed(seed, 100 + (uint64_t)cam, ((uint64_t)frame << 20) + (uint64_t)i);
    uv[0] = syn_u01(&r) * 1241.0;
    uv[1] = syn_u01(&r) * 376.0;
    double tx = syn_cam_tx(rig, cam);
    /* camera centre in cam-0 coords: pixel = K (p + t)  =>  centre = -t */
    double oc[3] = { -tx, 0.0, 0.0 };
    double dc[3] = { (uv[0]-SYN_CX)/SYN_FX, (uv[1]-SYN_CY)/SYN_FX, 1.0 };
    double nrm = sqrt(dc[0]*dc[0]+dc[1]*dc[1]+dc[2]*dc[2]);
    for (int a=0;a<3;a++) dc[a] /= nrm;
    double ow[3], dw[3];
    se3_apply(C, oc, ow);
    se3_rot(C, dc, dw);
    double s = syn_raycast(boxes, nb, ow, dw, 60.0);
    if (s < 1.0) return 0;
    for (int a=0;a<3;a++) pc[a] = oc[a] + dc[a]*s;
    return 1;
}

static inline void syn_pix2canon(double u, double v, float *out) {
    /* float restatement of velo.h:10-17 for the synthetic K (K^-1 = [1/f 0 -cx/f; 0 1/f -cy/f; 0 0 1]) */
    float fi = 1.0f/(float)SYN_FX;
    float kx = -(float)SYN_CX*fi, ky = -(float)SYN_CY*fi;
    float fu = (float)u, fv = (float)v;
    float p0 = fi*fu + kx, p1 = fi*fv + ky, p2 = 1.0f;
    out[0] = p0/p2; out[1] = p1/p2;
}

/*
 * kpA: [ncam][F][2] canonical keypoints "detected" in `frame`
 * kpB: [ncam][F][2] canonical keypoints = set A of frame-1 tracked into `frame` (+0.3 px noise)
 * matchB: [ncam][F] 1 where B[i] <-> A_{frame-1}[i] is a usable identity match
 */
void velo_synth_features(uint64_t seed, int frame, int rig, int ncam, int F,
                         float *kpA, float *kpB, int *matchB) {
    syn_box boxes[SYN_BOXES_PER_PERIOD];
    int nb = syn_scene_period(seed, boxes);
    syn_se3 C1, C0, C1i, T10; /* T10: frame-1 cam coords -> frame cam coords */
    syn_cam_pose(seed, frame, &C1);
    syn_cam_pose(seed, (int64_t)frame - 1, &C0);
    se3_inv(&C1, &C1i);
    se3_mul(&C1i, &C0, &T10);
    for (int cam=0; cam<ncam; cam++) {
        double tx = syn_cam_tx(rig, cam);
        for (int i=0;i<F;i++) {
            double uv[2], pc[3];
            syn_feature_anchor(seed, frame, rig, cam, i, boxes, nb, &C1, uv, pc);
            syn_pix2canon(uv[0], uv[1], kpA + 2*((size_t)cam*F + i));
            /* tracked set */
            double uv0[2], p0[3], p1[3];
            int ok = syn_feature_anchor(seed, frame-1, rig, cam, i, boxes, nb, &C0, uv0, p0);
            syn_rng r = syn_seed(seed, 200 + (uint64_t)cam, ((uint64_t)frame << 20) + (uint64_t)i);
            double u = syn_u01(&r)*1241.0, v = syn_u01(&r)*376.0; /* fallback: unmatched clutter */
            int m = 0;
            if (ok) {
                se3_apply(&T10, p0, p1);
                double X = p1[0] + tx, Y = p1[1], Z = p1[2];
                if (Z > 0.5) {
                    double uu = SYN_FX*X/Z + SYN_CX + syn_sym(&r, 0.3);
                    double vv = SYN_FX*Y/Z + SYN_CY + syn_sym(&r, 0.3);
                    if (uu >= 0 && uu < 1241.0 && vv >= 0 && vv < 376.0) { u = uu; v = vv; m = 1; }
                }
            }
            syn_pix2canon(u, v, kpB + 2*((size_t)cam*F + i));
            matchB[(size_t)cam*F + i] = m;
        }
    }
}

#ifdef __cplusplus
}
#endif
