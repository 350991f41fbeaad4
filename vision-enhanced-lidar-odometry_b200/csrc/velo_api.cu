// velo_api.cu — host side of libvelo_gpu.so: context, device memory, C-ABI entry points (include/velo_gpu.h).
// No CPU compute path exists here: every stage is a kernel launch; without a CUDA device create() fails.
#include "velo_dev.cuh"
#include <math.h>
#include <float.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <algorithm>
#include <vector>

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_create_error;

struct ProfRec { int k; cudaEvent_t a, b; };

struct velo_gpu_ctx {
    int device = 0, sm_count = 148;
    cudaStream_t stream = nullptr, copy_stream = nullptr, stream2 = nullptr;   // stream2: every other chunk of batch_frontend
    cudaStream_t d2h_stream = nullptr;                                         // keypoints_with_depth of finished chunks travel back while later chunks compute
    cudaStream_t launch_stream = nullptr;                                        // stream of the launch being profiled
    std::vector<cudaEvent_t> chunk_ev;
    cudaEvent_t fe_up0 = nullptr, fe_up1 = nullptr, fe_c0 = nullptr, fe_c1 = nullptr;   // batch_frontend: upload / whole-call timing of the last call
    float fe_growth = 2.0f;                                                             // chunk growth factor derived from it
    velo_gpu_params prm;
    velo_gpu_calib cal;
    DevCalib dcal;
    DevBuffers B;
    std::string err;
    // ICP / visual unit staging
    IcpUnit *h_icp_units = nullptr, *d_icp_units = nullptr;
    VisUnit *h_vis_units = nullptr, *d_vis_units = nullptr;
    double *d_icp_partial = nullptr, *d_icp_out = nullptr, *d_vis_partial = nullptr, *d_vis_out = nullptr;
    int icp_partial_ctas = 0, vis_ctas = 0;
    // the single-frame entry points stage through their own unit / partial / out records (index S resp. S*V), never through the
    // batch path's slot records
    int icp_scratch = 0, vis_scratch = 0;
    int *d_flags = nullptr;         // [0]: a visual match index was out of range
    int batch_passes = 0, batch_vis = 0;
    velo_icp_corr *d_corr = nullptr;
    VisMatchOut *d_mout = nullptr;
    int *d_lm_valid = nullptr; float4 *d_lm_xyz = nullptr;
    // device-resident solve of up to f2f_cap frame pairs at once (buffers grown on demand, velo_gpu_frame_to_frame needs 1)
    int f2f_cap = 0, f2f_last_n = 0;
    LmState *d_f2f_lm = nullptr, *h_f2f_lm = nullptr;
    IcpUnit *d_f2f_icp_units = nullptr, *h_f2f_icp_units = nullptr;
    VisUnit *d_f2f_vis_units = nullptr, *h_f2f_vis_units = nullptr;
    IcpFrozen *d_f2f_frozen = nullptr;
    unsigned char *d_f2f_sel = nullptr;
    double *d_f2f_poses = nullptr, *d_f2f_icp_partial = nullptr, *d_f2f_icp_out = nullptr, *d_f2f_eval_partial = nullptr, *d_f2f_eval_out = nullptr,
           *d_f2f_vis_partial = nullptr, *d_f2f_vis_out = nullptr;
    int *d_f2f_ndone = nullptr;
    // Hamming matcher scratch (grown on demand)
    unsigned long long *d_hq = nullptr, *d_ht = nullptr, *d_hbest = nullptr; size_t ham_cap_q = 0, ham_cap_t = 0;
    // triangulation scratch (one arena, grown on demand)
    char *d_tri = nullptr; size_t tri_cap = 0;
    std::vector<int> h_npoints;     // host copy of n_points per slot (single-frame path)
    std::vector<int> h_stride;      // staging of raw_stride (must outlive the asynchronous copy)
    // timing
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    bool profile = false;
    bool search_stats = false;     // k_icp_pass also counts candidates / rings per pass (velo_gpu_search_stats_enable)
    std::vector<cudaEvent_t> ev_pool;
    std::vector<ProfRec> recs;
    cudaEvent_t cur_a = nullptr;
    int64_t launches = 0;
    std::vector<void *> allocs;
};

static int fail(velo_gpu_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, VELO_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

// ------------------------------------------------------------------------------------------------ float thresholds (hazard H3)
// smallest float f with (double)f >= d:   for float v,  (double)v <  d  <=>  v <  f   and   (double)v >= d  <=>  v >= f
static float ceil_to_float(double d) { float f = (float)d; if ((double)f < d) f = nextafterf(f, INFINITY); return f; }
// largest float f with (double)f <= d:    for float v,  (double)v >  d  <=>  v >  f
static float floor_to_float(double d) { float f = (float)d; if ((double)f > d) f = nextafterf(f, -INFINITY); return f; }

// ------------------------------------------------------------------------------------------------ a1: calibration (kitti.h:59-108)
// Eigen::Matrix3f::inverse() (cofactor form) and 3-term fixed-size reductions e0 + (e1 + e2) [Eigen, SURVEY.md §8 a1]
static inline float s3(float a, float b, float c) { return a + (b + c); }
static inline float cofac(const float *m, int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[3 * i1 + j1] * m[3 * i2 + j2] - m[3 * i1 + j2] * m[3 * i2 + j1];
}
static void inverse3(const float *m, float *r) {
    const float c0 = cofac(m, 0, 0), c1 = cofac(m, 1, 0), c2 = cofac(m, 2, 0);
    const float det = s3(c0 * m[0], c1 * m[3], c2 * m[6]), id = 1.0f / det;
    r[0] = c0 * id; r[1] = c1 * id; r[2] = c2 * id;
    for (int row = 1; row < 3; row++) for (int col = 0; col < 3; col++) r[3 * row + col] = cofac(m, col, row) * id;
}
static void matvec3(const float *m, const float *v, float *o) {
    for (int i = 0; i < 3; i++) o[i] = s3(m[3 * i] * v[0], m[3 * i + 1] * v[1], m[3 * i + 2] * v[2]);
}

extern "C" int velo_gpu_abi_version(void) { return VELO_GPU_ABI_VERSION; }

extern "C" int velo_gpu_default_params(velo_gpu_params *p) {
    if (!p) return VELO_ERR_INVALID_ARG;
    memset(p, 0, sizeof(*p));
    p->num_cams = 2; p->icp_skip = 200; p->f2f_iterations = 2; p->icp_iterations = 3;     // kitti.h:3,8-10
    p->enable_2d2d = 1; p->enable_3d2d = 1; p->abs_truncates = 0;                          // main.cpp:44-45
    p->weight_3D2D = 10; p->weight_2D2D = 500; p->weight_3DPD = 1;                         // kitti.h:20-22
    p->loss_thresh_3D2D = 0.01; p->loss_thresh_2D2D = 0.00002; p->loss_thresh_3DPD = 0.1; p->loss_thresh_3D3D = 0.04; // kitti.h:23-26
    p->depth_assoc_thresh = 0.015; p->outlier_reject = 5.0; p->correspondence_thresh_icp = 0.5; p->icp_norm_condition = 1e-5; // kitti.h:28-32
    p->max_slots = 4; p->max_points = 131072; p->max_rings = 128; p->max_features = 3000;  // kitti.h:7
    p->max_matches = 3000; p->max_icp_passes = 6; p->ctas_per_icp_unit = 0;
    return VELO_OK;
}

extern "C" int velo_gpu_calib_from_kitti(const float P[48], const float Tr[12], int w, int h, velo_gpu_calib *c) {
    if (!P || !Tr || !c) return VELO_ERR_INVALID_ARG;
    memset(c, 0, sizeof(*c));
    for (int cam = 0; cam < VELO_MAX_CAMS; cam++) {
        const float *p = P + 12 * cam;
        const float K[9] = { p[0], p[1], p[2], p[4], p[5], p[6], p[8], p[9], p[10] };
        const float Kt[3] = { p[3], p[7], p[11] };
        float Ki[9], t[3];
        inverse3(K, Ki);
        matvec3(Ki, Kt, t);                                       // cam_trans = K^-1 * P[:,3]
        memcpy(c->cam_K[cam], K, sizeof(K)); memcpy(c->cam_Kinv[cam], Ki, sizeof(Ki));
        for (int i = 0; i < 3; i++) c->cam_trans[cam][i] = t[i];
        const float lo[3] = { 0.f, 0.f, 1.f }, hi[3] = { (float)w, (float)h, 1.f };
        float a[3], b[3];
        matvec3(Ki, lo, a); matvec3(Ki, hi, b);
        c->min_x[cam] = a[0] / a[2]; c->min_y[cam] = a[1] / a[2];
        c->max_x[cam] = b[0] / b[2]; c->max_y[cam] = b[1] / b[2];
    }
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c->velo_to_cam[4 * i + j] = (i < 3) ? Tr[4 * i + j] : (j == 3 ? 1.f : 0.f);
    c->img_width = w; c->img_height = h;
    return VELO_OK;
}

extern "C" int velo_pixel2canonical(const velo_gpu_calib *c, int cam, const float *pix, int n, float *out) {  // velo.h:10-17
    if (!c || cam < 0 || cam >= VELO_MAX_CAMS || (n > 0 && (!pix || !out))) return VELO_ERR_INVALID_ARG;
    for (int i = 0; i < n; i++) { const float v[3] = { pix[2 * i], pix[2 * i + 1], 1.f }; float p[3]; matvec3(c->cam_Kinv[cam], v, p); out[2 * i] = p[0] / p[2]; out[2 * i + 1] = p[1] / p[2]; }
    return VELO_OK;
}
extern "C" int velo_canonical2pixel(const velo_gpu_calib *c, int cam, const float *can, int n, float *out) {  // velo.h:19-26
    if (!c || cam < 0 || cam >= VELO_MAX_CAMS || (n > 0 && (!can || !out))) return VELO_ERR_INVALID_ARG;
    for (int i = 0; i < n; i++) { const float v[3] = { can[2 * i], can[2 * i + 1], 1.f }; float p[3]; matvec3(c->cam_K[cam], v, p); out[2 * i] = p[0] / p[2]; out[2 * i + 1] = p[1] / p[2]; }
    return VELO_OK;
}

// ------------------------------------------------------------------------------------------------ KITTI wire formats (host only)
extern "C" int velo_kitti_load_calib(const char *path, float P[48], float Tr[12]) {      // kitti.h:61-105
    if (!path || !P || !Tr) return VELO_ERR_INVALID_ARG;
    FILE *f = fopen(path, "r");
    if (!f) return VELO_ERR_INVALID_ARG;
    char label[64];
    int ok = 1;
    for (int cam = 0; cam < VELO_MAX_CAMS && ok; cam++) {
        if (fscanf(f, "%63s", label) != 1) { ok = 0; break; }                              // calib_stream >> P;
        for (int i = 0; i < 12; i++) if (fscanf(f, "%f", &P[12 * cam + i]) != 1) { ok = 0; break; }
    }
    if (ok && fscanf(f, "%63s", label) != 1) ok = 0;                                       // "Tr:"
    for (int i = 0; i < 12 && ok; i++) if (fscanf(f, "%f", &Tr[i]) != 1) ok = 0;
    fclose(f);
    return ok ? VELO_OK : VELO_ERR_INVALID_ARG;
}

extern "C" int velo_kitti_load_scan(const char *path, float *xyzr, int max_points, int *n) { // kitti.h:121-152
    if (!path || !n || max_points < 0 || (max_points > 0 && !xyzr)) return VELO_ERR_INVALID_ARG;
    FILE *f = fopen(path, "rb");
    if (!f) return VELO_ERR_INVALID_ARG;
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    const long total = bytes / (long)(4 * sizeof(float));                                   // fread(...)/4
    const long take = total < max_points ? total : max_points;
    const size_t got = take > 0 ? fread(xyzr, 4 * sizeof(float), (size_t)take, f) : 0;
    fclose(f);
    *n = (int)total;
    if ((long)got != take) return VELO_ERR_INVALID_ARG;
    return total > max_points ? VELO_ERR_CAPACITY : VELO_OK;
}

extern "C" int velo_kitti_format_pose(const double T[16], char *buf, int buflen) {         // kitti.h:202-216
    if (!T || !buf || buflen < 1) return VELO_ERR_INVALID_ARG;
    int o = 0;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) {
        const int w = snprintf(buf + o, (size_t)(buflen - o), "%g ", T[4 * i + j]);       // std::ostream default: %g, precision 6
        if (w < 0 || w >= buflen - o) return VELO_ERR_CAPACITY;
        o += w;
    }
    return VELO_OK;
}

// ------------------------------------------------------------------------------------------------ pose constants (hazard H8)
namespace {
struct J3 { double a, v[3]; };
inline J3 j3(double s) { return J3{ s, { 0, 0, 0 } }; }
inline J3 operator+(J3 x, J3 y) { return J3{ x.a + y.a, { x.v[0] + y.v[0], x.v[1] + y.v[1], x.v[2] + y.v[2] } }; }
inline J3 operator-(J3 x, J3 y) { return J3{ x.a - y.a, { x.v[0] - y.v[0], x.v[1] - y.v[1], x.v[2] - y.v[2] } }; }
inline J3 operator*(J3 x, J3 y) { return J3{ x.a * y.a, { x.a * y.v[0] + x.v[0] * y.a, x.a * y.v[1] + x.v[1] * y.a, x.a * y.v[2] + x.v[2] * y.a } }; }
inline J3 operator/(J3 x, J3 y) { const double inv = 1.0 / y.a, q = x.a * inv; return J3{ q, { (x.v[0] - q * y.v[0]) * inv, (x.v[1] - q * y.v[1]) * inv, (x.v[2] - q * y.v[2]) * inv } }; }
inline J3 jsqrt(J3 x) { const double r = sqrt(x.a), d = 1.0 / (2.0 * r); return J3{ r, { x.v[0] * d, x.v[1] * d, x.v[2] * d } }; }
inline J3 jsin(J3 x) { const double c = cos(x.a); return J3{ sin(x.a), { c * x.v[0], c * x.v[1], c * x.v[2] } }; }
inline J3 jcos(J3 x) { const double s = -sin(x.a); return J3{ cos(x.a), { s * x.v[0], s * x.v[1], s * x.v[2] } }; }

// AngleAxisRotatePoint on 3-partial dual numbers: derivative of R(w) e_j w.r.t. w
void rotate_j3(const J3 w[3], const J3 p[3], J3 out[3]) {
    const J3 th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    if (th2.a > DBL_EPSILON) {
        const J3 th = jsqrt(th2), c = jcos(th), s = jsin(th), ith = j3(1.0) / th;
        const J3 u[3] = { w[0] * ith, w[1] * ith, w[2] * ith };
        const J3 x[3] = { u[1] * p[2] - u[2] * p[1], u[2] * p[0] - u[0] * p[2], u[0] * p[1] - u[1] * p[0] };
        const J3 tmp = (u[0] * p[0] + u[1] * p[1] + u[2] * p[2]) * (j3(1.0) - c);
        for (int i = 0; i < 3; i++) out[i] = p[i] * c + x[i] * s + u[i] * tmp;
    } else {
        out[0] = p[0] + (w[1] * p[2] - w[2] * p[1]);
        out[1] = p[1] + (w[2] * p[0] - w[0] * p[2]);
        out[2] = p[2] + (w[0] * p[1] - w[1] * p[0]);
    }
}

void make_pose_pack(const double pose[6], PosePack *P) {
    memset(P, 0, sizeof(*P));
    for (int i = 0; i < 3; i++) { P->w[i] = pose[i]; P->t[i] = pose[3 + i]; }
    const double th2 = pose[0] * pose[0] + pose[1] * pose[1] + pose[2] * pose[2];
    P->small_angle = !(th2 > DBL_EPSILON);
    if (!P->small_angle) {
        const double th = sqrt(th2);
        P->c = cos(th); P->s = sin(th);
        const double ith = 1.0 / th;
        for (int i = 0; i < 3; i++) P->u[i] = pose[i] * ith;
    } else { P->c = 1.0; P->s = 0.0; }
    J3 w[3];
    for (int k = 0; k < 3; k++) { w[k] = j3(pose[k]); w[k].v[k] = 1.0; }
    for (int j = 0; j < 3; j++) {
        J3 e[3] = { j3(j == 0), j3(j == 1), j3(j == 2) }, o[3];
        rotate_j3(w, e, o);
        for (int k = 0; k < 3; k++) for (int i = 0; i < 3; i++) P->dR[9 * k + 3 * i + j] = o[i].v[k];
    }
}
} // namespace

// ------------------------------------------------------------------------------------------------ profiling hooks
static const char *k_names[VELO_NUM_KERNELS] = { "ingest_flags", "ingest_rings", "ingest_permute", "index_build", "project_occlude",
                                                 "assoc_search", "assoc_compact", "icp_pass", "neq_reduce", "visual_residuals", "index_masks", "lm_solve" };
extern "C" const char *velo_gpu_kernel_name(int k) { return (k >= 0 && k < VELO_NUM_KERNELS) ? k_names[k] : ""; }

static cudaEvent_t get_event(velo_gpu_ctx *c) {
    if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
static void prof_pre(void *u, int k) {
    velo_gpu_ctx *c = (velo_gpu_ctx *)u; c->launches++;
    if (!c->profile) return;
    c->cur_a = get_event(c); cudaEventRecord(c->cur_a, c->launch_stream);
}
static void prof_post(void *u, int k) {
    velo_gpu_ctx *c = (velo_gpu_ctx *)u;
    if (!c->profile) return;
    cudaEvent_t b = get_event(c); cudaEventRecord(b, c->launch_stream);
    c->recs.push_back(ProfRec{ k, c->cur_a, b });
}
static Launcher launcher_on(velo_gpu_ctx *c, cudaStream_t st) { c->launch_stream = st; return Launcher{ st, prof_pre, prof_post, c }; }
static Launcher launcher(velo_gpu_ctx *c) { return launcher_on(c, c->stream); }

// ------------------------------------------------------------------------------------------------ lifecycle
template <class T> static cudaError_t dalloc(velo_gpu_ctx *c, T **p, size_t n) {
    cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
    if (e == cudaSuccess) { c->allocs.push_back(*p); e = cudaMemsetAsync(*p, 0, n * sizeof(T), c->stream); }
    return e;
}

extern "C" int velo_gpu_create(int device, const velo_gpu_params *prm, const velo_gpu_calib *cal, velo_gpu_ctx **out) {
    velo_gpu_ctx *ctx = nullptr;
    if (!prm || !cal || !out) return fail(nullptr, VELO_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (prm->num_cams < 1 || prm->num_cams > VELO_MAX_CAMS || prm->max_rings < 1 || prm->max_rings > VELO_MAX_RINGS_HARD ||
        prm->max_points < 1 || prm->max_points > (1 << VELO_IDX_BITS) || prm->max_slots < 1 || prm->max_features < 1 ||
        prm->max_matches < 1 || prm->max_icp_passes < 1 || prm->icp_skip < 1)
        return fail(nullptr, VELO_ERR_INVALID_ARG, "parameter out of range (num_cams 1..4, max_rings 1..256, max_points 1..2^20, ...)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, VELO_ERR_NO_DEVICE, std::string("no CUDA device (") + cudaGetErrorString(e) + "); libvelo_gpu has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, VELO_ERR_INVALID_ARG, "device index out of range");
    cudaDeviceProp dp;
    if (cudaGetDeviceProperties(&dp, device) != cudaSuccess || dp.major != 10)
        return fail(nullptr, VELO_ERR_NO_DEVICE, "device is not sm_100 (Blackwell B200); kernels are built for sm_100a only");
    ctx = new velo_gpu_ctx();
    ctx->device = device; ctx->prm = *prm; ctx->cal = *cal; ctx->sm_count = dp.multiProcessorCount;
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { std::string m = std::string(#call) + ": " + cudaGetErrorString(e_); velo_gpu_destroy(ctx); return fail(nullptr, VELO_ERR_CUDA, m); } } while (0)
    CKC(cudaSetDevice(device));
    CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CKC(cudaEventCreate(&ctx->t0)); CKC(cudaEventCreate(&ctx->t1));
    DevBuffers &B = ctx->B;
    memset(&B, 0, sizeof(B));
    B.S = prm->max_slots; B.N = (prm->max_points + 127) & ~127; B.R = prm->max_rings; B.C = prm->num_cams;
    B.F = prm->max_features; B.MM = prm->max_matches; B.P = prm->max_icp_passes;
    const size_t S = B.S, N = B.N, R = B.R, C = B.C, F = B.F, MM = B.MM, P = B.P;
    CKC(dalloc(ctx, &B.raw, S * N)); CKC(dalloc(ctx, &B.raw_stride, S)); CKC(dalloc(ctx, &B.flagbits, S * (N / 32)));
    CKC(dalloc(ctx, &B.n_points, S)); CKC(dalloc(ctx, &B.n_rings, S)); CKC(dalloc(ctx, &B.ring_start, S * (R + 1))); CKC(dalloc(ctx, &B.status, S));
    CKC(dalloc(ctx, &B.pts, S * N)); CKC(dalloc(ctx, &B.sorted, S * N));
    CKC(dalloc(ctx, &B.cell_start, S * R * (VELO_AZ_BINS + 1))); CKC(dalloc(ctx, &B.sec_box, S * R * VELO_SECTORS));
    B.W = (B.R + 63) / 64;
    CKC(dalloc(ctx, &B.mask_lo, S * VELO_SECTORS * VELO_EL_BUCKETS * (size_t)B.W)); CKC(dalloc(ctx, &B.mask_hi, S * VELO_SECTORS * VELO_EL_BUCKETS * (size_t)B.W));
    CKC(dalloc(ctx, &B.rmask_lo, S * VELO_SECTORS * VELO_RG_BUCKETS * (size_t)B.W)); CKC(dalloc(ctx, &B.rmask_hi, S * VELO_SECTORS * VELO_RG_BUCKETS * (size_t)B.W));
    CKC(dalloc(ctx, &B.proj, S * C * N)); CKC(dalloc(ctx, &B.valid, S * C * N)); CKC(dalloc(ctx, &B.proj_count, S * C * R)); CKC(dalloc(ctx, &B.proj_yrange, S * C * R));
    const size_t SK = S * VELO_NUM_KP_SETS * C;
    CKC(dalloc(ctx, &B.kp, SK * F)); CKC(dalloc(ctx, &B.n_kp, SK)); CKC(dalloc(ctx, &B.has_depth, SK * F)); CKC(dalloc(ctx, &B.kpwd, SK * F));
    CKC(dalloc(ctx, &B.n_hits, SK)); CKC(dalloc(ctx, &B.hit_tmp, SK * F)); CKC(dalloc(ctx, &B.kpwd_tmp, SK * F));
    CKC(dalloc(ctx, &B.matches, S * C * MM * 2)); CKC(dalloc(ctx, &B.n_matches, S * C));
    // units / normal equations
    ctx->icp_partial_ctas = 296;
    if (prm->max_icp_passes > VELO_MAX_PASSES) { velo_gpu_destroy(ctx); return fail(nullptr, VELO_ERR_INVALID_ARG, "max_icp_passes exceeds VELO_MAX_PASSES (6)"); }
    const size_t n_vis_batch = S * (size_t)(prm->f2f_iterations > 0 ? prm->f2f_iterations : 1);
    const size_t n_icp_units = S + 1, n_vis_units = n_vis_batch + 1;
    ctx->icp_scratch = (int)S; ctx->vis_scratch = (int)n_vis_batch;
    CKC(cudaMallocHost((void **)&ctx->h_icp_units, n_icp_units * sizeof(IcpUnit)));
    CKC(cudaMallocHost((void **)&ctx->h_vis_units, n_vis_units * sizeof(VisUnit)));
    CKC(dalloc(ctx, &ctx->d_icp_units, n_icp_units)); CKC(dalloc(ctx, &ctx->d_vis_units, n_vis_units));
    const size_t part = n_icp_units * (size_t)launch_icp_runs_cap((int)N);     // one record per run of queries (velo_icp.cu)
    CKC(dalloc(ctx, &ctx->d_icp_partial, part * VELO_MAX_PASSES * 64)); CKC(dalloc(ctx, &ctx->d_icp_out, n_icp_units * P * VELO_NEQ_STRIDE));
    ctx->vis_ctas = 0;
    const size_t vpart = n_vis_batch * 4 + 64;      // batch: 4 CTAs per unit; scratch: up to 64
    CKC(dalloc(ctx, &ctx->d_vis_partial, vpart * 64)); CKC(dalloc(ctx, &ctx->d_vis_out, n_vis_units * VELO_NEQ_STRIDE));
    CKC(dalloc(ctx, &ctx->d_corr, N * P)); CKC(dalloc(ctx, &ctx->d_mout, C * MM)); CKC(dalloc(ctx, &ctx->d_flags, 4));
    CKC(dalloc(ctx, &ctx->d_lm_valid, C * MM)); CKC(dalloc(ctx, &ctx->d_lm_xyz, C * MM));
    ctx->h_npoints.assign(S, 0); ctx->h_stride.assign(S, 4);
    // calibration for the device
    DevCalib &d = ctx->dcal; memset(&d, 0, sizeof(d));
    for (int i = 0; i < 12; i++) d.vtc[i] = cal->velo_to_cam[i];
    for (int c = 0; c < VELO_MAX_CAMS; c++) {
        for (int i = 0; i < 3; i++) d.cam_t[c][i] = cal->cam_trans[c][i];
        d.fov[c][0] = ceil_to_float(cal->min_x[c]); d.fov[c][1] = ceil_to_float(cal->max_x[c]);
        d.fov[c][2] = ceil_to_float(cal->min_y[c]); d.fov[c][3] = ceil_to_float(cal->max_y[c]);
    }
    d.assoc_thr = ceil_to_float(prm->depth_assoc_thresh); d.abs_truncates = prm->abs_truncates; d.num_cams = prm->num_cams;
    CKC(cudaStreamSynchronize(ctx->stream));
#undef CKC
    *out = ctx;
    return VELO_OK;
}

static void f2f_free(velo_gpu_ctx *ctx) {
    void *dev[] = { ctx->d_f2f_lm, ctx->d_f2f_icp_units, ctx->d_f2f_vis_units, ctx->d_f2f_frozen, ctx->d_f2f_sel, ctx->d_f2f_poses, ctx->d_f2f_icp_partial,
                    ctx->d_f2f_icp_out, ctx->d_f2f_eval_partial, ctx->d_f2f_eval_out, ctx->d_f2f_vis_partial, ctx->d_f2f_vis_out, ctx->d_f2f_ndone };
    for (void *p : dev) if (p) cudaFree(p);
    if (ctx->h_f2f_lm) cudaFreeHost(ctx->h_f2f_lm);
    if (ctx->h_f2f_icp_units) cudaFreeHost(ctx->h_f2f_icp_units);
    if (ctx->h_f2f_vis_units) cudaFreeHost(ctx->h_f2f_vis_units);
    ctx->d_f2f_lm = nullptr; ctx->h_f2f_lm = nullptr; ctx->d_f2f_icp_units = nullptr; ctx->h_f2f_icp_units = nullptr; ctx->d_f2f_vis_units = nullptr;
    ctx->h_f2f_vis_units = nullptr; ctx->d_f2f_frozen = nullptr; ctx->d_f2f_sel = nullptr; ctx->d_f2f_poses = nullptr; ctx->d_f2f_icp_partial = nullptr;
    ctx->d_f2f_icp_out = nullptr; ctx->d_f2f_eval_partial = nullptr; ctx->d_f2f_eval_out = nullptr; ctx->d_f2f_vis_partial = nullptr; ctx->d_f2f_vis_out = nullptr;
    ctx->d_f2f_ndone = nullptr; ctx->f2f_cap = 0;
}

extern "C" int velo_gpu_destroy(velo_gpu_ctx *ctx) {
    if (!ctx) return VELO_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (void *p : ctx->allocs) cudaFree(p);
    if (ctx->d_hq) { cudaFree(ctx->d_hq); cudaFree(ctx->d_hbest); }
    if (ctx->d_ht) cudaFree(ctx->d_ht);
    if (ctx->d_tri) cudaFree(ctx->d_tri);
    f2f_free(ctx);
    if (ctx->h_icp_units) cudaFreeHost(ctx->h_icp_units);
    if (ctx->h_vis_units) cudaFreeHost(ctx->h_vis_units);
    for (auto &r : ctx->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->t0) cudaEventDestroy(ctx->t0);
    if (ctx->t1) cudaEventDestroy(ctx->t1);
    for (auto e : ctx->chunk_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : { ctx->fe_up0, ctx->fe_up1, ctx->fe_c0, ctx->fe_c1 }) if (e) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return VELO_OK;
}

extern "C" const char *velo_gpu_last_error(const velo_gpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
extern "C" int velo_gpu_sync(velo_gpu_ctx *ctx) { if (!ctx) return VELO_ERR_INVALID_ARG; CK(cudaSetDevice(ctx->device)); CK(cudaStreamSynchronize(ctx->stream)); CK(cudaGetLastError()); return VELO_OK; }
extern "C" int velo_gpu_device_name(velo_gpu_ctx *ctx, char *buf, int n) {
    if (!ctx || !buf || n < 1) return VELO_ERR_INVALID_ARG;
    cudaDeviceProp dp; CK(cudaGetDeviceProperties(&dp, ctx->device));
    snprintf(buf, n, "%s", dp.name); return VELO_OK;
}
extern "C" int velo_gpu_host_alloc(void **p, uint64_t bytes) { if (!p) return VELO_ERR_INVALID_ARG; return cudaMallocHost(p, bytes ? bytes : 1) == cudaSuccess ? VELO_OK : VELO_ERR_CUDA; }
extern "C" int velo_gpu_host_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? VELO_OK : VELO_ERR_CUDA; }
extern "C" int velo_gpu_timer_begin(velo_gpu_ctx *ctx) { if (!ctx) return VELO_ERR_INVALID_ARG; CK(cudaSetDevice(ctx->device)); CK(cudaEventRecord(ctx->t0, ctx->stream)); return VELO_OK; }
extern "C" int velo_gpu_timer_end(velo_gpu_ctx *ctx, float *ms) {
    if (!ctx || !ms) return VELO_ERR_INVALID_ARG;
    CK(cudaEventRecord(ctx->t1, ctx->stream)); CK(cudaEventSynchronize(ctx->t1)); CK(cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
    return VELO_OK;
}
extern "C" int velo_gpu_profile_enable(velo_gpu_ctx *ctx, int on) { if (!ctx) return VELO_ERR_INVALID_ARG; ctx->profile = on != 0; return VELO_OK; }
extern "C" int velo_gpu_search_stats_enable(velo_gpu_ctx *ctx, int on) { if (!ctx) return VELO_ERR_INVALID_ARG; ctx->search_stats = on != 0; return VELO_OK; }
extern "C" int velo_gpu_profile_reset(velo_gpu_ctx *ctx) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->recs) { ctx->ev_pool.push_back(r.a); ctx->ev_pool.push_back(r.b); }
    ctx->recs.clear();
    return VELO_OK;
}
extern "C" int velo_gpu_profile_read(velo_gpu_ctx *ctx, float ms[VELO_NUM_KERNELS], int launches[VELO_NUM_KERNELS]) {
    if (!ctx || !ms || !launches) return VELO_ERR_INVALID_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < VELO_NUM_KERNELS; i++) { ms[i] = 0.f; launches[i] = 0; }
    for (auto &r : ctx->recs) { float t = 0.f; CK(cudaEventElapsedTime(&t, r.a, r.b)); ms[r.k] += t; launches[r.k]++; }
    return VELO_OK;
}
extern "C" int velo_gpu_launch_count(velo_gpu_ctx *ctx, int64_t *n) { if (!ctx || !n) return VELO_ERR_INVALID_ARG; *n = ctx->launches; return VELO_OK; }

static int check_slot(velo_gpu_ctx *ctx, int slot) { return (slot >= 0 && slot < ctx->B.S) ? VELO_OK : fail(ctx, VELO_ERR_INVALID_ARG, "slot out of range"); }
static int check_cam(velo_gpu_ctx *ctx, int cam) { return (cam >= 0 && cam < ctx->B.C) ? VELO_OK : fail(ctx, VELO_ERR_INVALID_ARG, "camera out of range"); }
static int slot_status(velo_gpu_ctx *ctx, int slot) {
    int st = 0;
    CK(cudaMemcpyAsync(&st, ctx->B.status + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (st != 0) return fail(ctx, st, "scan has more rings than max_rings (kitti.h:166-173 ring count is data dependent)");
    return VELO_OK;
}

// ------------------------------------------------------------------------------------------------ single-frame path
extern "C" int velo_gpu_scan_upload(velo_gpu_ctx *ctx, int slot, const float *xyzr, int n) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (check_slot(ctx, slot)) return VELO_ERR_INVALID_ARG;
    if (n < 0 || (n > 0 && !xyzr)) return fail(ctx, VELO_ERR_INVALID_ARG, "bad scan pointer / size");
    if (n > ctx->prm.max_points) return fail(ctx, VELO_ERR_CAPACITY, "scan has more points than max_points");
    CK(cudaSetDevice(ctx->device));
    const DevBuffers &B = ctx->B;
    if (n > 0) CK(cudaMemcpyAsync(B.raw + (size_t)slot * B.N, xyzr, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(B.n_points + slot, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    const int four = 4;
    CK(cudaMemcpyAsync(B.raw_stride + slot, &four, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(B.proj_count + (size_t)slot * B.C * B.R, 0, (size_t)B.C * B.R * sizeof(int), ctx->stream));   // projections of the old scan are void
    CK(cudaStreamSynchronize(ctx->stream)); // &n is a stack variable
    ctx->h_npoints[slot] = n;
    Launcher L = launcher(ctx);
    launch_ingest(L, B, ctx->dcal, slot, 1);
    launch_index(L, B, ctx->dcal, slot, 1);
    CK(cudaGetLastError());
    return VELO_OK;
}

extern "C" int velo_gpu_scan_upload_rings(velo_gpu_ctx *ctx, int slot, const float *xyz1, const int *ring_start, int n_rings) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (check_slot(ctx, slot)) return VELO_ERR_INVALID_ARG;
    if (n_rings < 0 || !ring_start || ring_start[0] != 0) return fail(ctx, VELO_ERR_INVALID_ARG, "bad ring_start");
    if (n_rings > ctx->B.R) return fail(ctx, VELO_ERR_CAPACITY, "more rings than max_rings");
    for (int s = 0; s < n_rings; s++) if (ring_start[s + 1] < ring_start[s]) return fail(ctx, VELO_ERR_INVALID_ARG, "ring_start must be non-decreasing");
    const int n = ring_start[n_rings];
    if (n > ctx->prm.max_points) return fail(ctx, VELO_ERR_CAPACITY, "scan has more points than max_points");
    if (n > 0 && !xyz1) return fail(ctx, VELO_ERR_INVALID_ARG, "null points");
    CK(cudaSetDevice(ctx->device));
    const DevBuffers &B = ctx->B;
    const int zero = 0;
    if (n > 0) CK(cudaMemcpyAsync(B.pts + (size_t)slot * B.N, xyz1, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(B.ring_start + (size_t)slot * (B.R + 1), ring_start, (size_t)(n_rings + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(B.n_points + slot, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(B.n_rings + slot, &n_rings, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(B.status + slot, &zero, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(B.proj_count + (size_t)slot * B.C * B.R, 0, (size_t)B.C * B.R * sizeof(int), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->h_npoints[slot] = n;
    launch_index(launcher(ctx), B, ctx->dcal, slot, 1);
    CK(cudaGetLastError());
    return VELO_OK;
}

extern "C" int velo_gpu_scan_info(velo_gpu_ctx *ctx, int slot, int *n_points, int *n_rings) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (check_slot(ctx, slot)) return VELO_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    int st = slot_status(ctx, slot); if (st) return st;
    int np = 0, nr = 0;
    CK(cudaMemcpyAsync(&np, ctx->B.n_points + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&nr, ctx->B.n_rings + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (n_points) *n_points = np;
    if (n_rings) *n_rings = nr;
    return VELO_OK;
}

extern "C" int velo_gpu_scan_download(velo_gpu_ctx *ctx, int slot, float *xyz1, int *ring_start) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    int np, nr; int st = velo_gpu_scan_info(ctx, slot, &np, &nr); if (st) return st;
    const DevBuffers &B = ctx->B;
    if (xyz1 && np > 0) CK(cudaMemcpyAsync(xyz1, B.pts + (size_t)slot * B.N, (size_t)np * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    if (ring_start) CK(cudaMemcpyAsync(ring_start, B.ring_start + (size_t)slot * (B.R + 1), (size_t)(nr + 1) * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ring_start && nr == 0) ring_start[0] = 0;
    return VELO_OK;
}

extern "C" int velo_gpu_project(velo_gpu_ctx *ctx, int slot, int cam) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (check_slot(ctx, slot) || check_cam(ctx, cam)) return VELO_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    // the projection kernel handles all cameras of the rig in one pass over the scan (main.cpp:254-256 projects per camera)
    launch_project(launcher(ctx), ctx->B, ctx->dcal, slot, 1);
    CK(cudaGetLastError());
    return VELO_OK;
}

extern "C" int velo_gpu_project_download(velo_gpu_ctx *ctx, int slot, int cam, int *ring_count, float *proj, float *valid, int *total) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (check_cam(ctx, cam)) return VELO_ERR_INVALID_ARG;
    int np, nr; int st = velo_gpu_scan_info(ctx, slot, &np, &nr); if (st) return st;
    const DevBuffers &B = ctx->B;
    std::vector<int> rc(nr + 1, 0), rs(nr + 1, 0);
    if (nr > 0) {
        CK(cudaMemcpyAsync(rc.data(), B.proj_count + ((size_t)slot * B.C + cam) * B.R, nr * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(rs.data(), B.ring_start + (size_t)slot * (B.R + 1), (nr + 1) * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    int o = 0;
    for (int s = 0; s < nr; s++) {
        if (ring_count) ring_count[s] = rc[s];
        if (rc[s] > 0) {
            if (proj) CK(cudaMemcpyAsync(proj + 2 * (size_t)o, B.proj + ((size_t)slot * B.C + cam) * B.N + rs[s], rc[s] * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
            if (valid) CK(cudaMemcpyAsync(valid + 4 * (size_t)o, B.valid + ((size_t)slot * B.C + cam) * B.N + rs[s], rc[s] * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
        }
        o += rc[s];
    }
    CK(cudaStreamSynchronize(ctx->stream));
    if (total) *total = o;
    return VELO_OK;
}

extern "C" int velo_gpu_projection_upload(velo_gpu_ctx *ctx, int slot, int cam, const int *ring_count, const float *proj, const float *valid) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (check_cam(ctx, cam) || !ring_count) return fail(ctx, VELO_ERR_INVALID_ARG, "bad camera / ring_count");
    int np, nr; int st = velo_gpu_scan_info(ctx, slot, &np, &nr); if (st) return st;
    const DevBuffers &B = ctx->B;
    std::vector<int> rs(nr + 1, 0);
    if (nr > 0) { CK(cudaMemcpyAsync(rs.data(), B.ring_start + (size_t)slot * (B.R + 1), (nr + 1) * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream)); }
    size_t o = 0;
    for (int s = 0; s < nr; s++) {
        if (ring_count[s] < 0 || ring_count[s] > rs[s + 1] - rs[s]) return fail(ctx, VELO_ERR_INVALID_ARG, "ring_count exceeds the ring length");
        if (ring_count[s] > 0) {
            if (!proj || !valid) return fail(ctx, VELO_ERR_INVALID_ARG, "null projection arrays");
            CK(cudaMemcpyAsync(B.proj + ((size_t)slot * B.C + cam) * B.N + rs[s], proj + 2 * o, ring_count[s] * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(B.valid + ((size_t)slot * B.C + cam) * B.N + rs[s], valid + 4 * o, ring_count[s] * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
        }
        o += ring_count[s];
    }
    std::vector<float2> yr(nr > 0 ? nr : 1);
    o = 0;
    for (int s = 0; s < nr; s++) {
        float lo = INFINITY, hi = -INFINITY;
        for (int i = 0; i < ring_count[s]; i++) { const float y = proj[2 * (o + i) + 1]; lo = fminf(lo, y); hi = fmaxf(hi, y); }
        yr[s] = make_float2(lo, hi); o += ring_count[s];
    }
    if (nr > 0) {
        CK(cudaMemcpyAsync(B.proj_count + ((size_t)slot * B.C + cam) * B.R, ring_count, nr * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(B.proj_yrange + ((size_t)slot * B.C + cam) * B.R, yr.data(), nr * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return VELO_OK;
}

extern "C" int velo_gpu_depth_assoc(velo_gpu_ctx *ctx, int slot, int cam, int set, const float *kp, int F, int *has_depth, float *kpwd, int *n_hits) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (check_slot(ctx, slot) || check_cam(ctx, cam)) return VELO_ERR_INVALID_ARG;
    if (set < 0 || set >= VELO_NUM_KP_SETS || F < 0 || (F > 0 && !kp)) return fail(ctx, VELO_ERR_INVALID_ARG, "bad keypoint set / pointer");
    if (F > ctx->B.F) return fail(ctx, VELO_ERR_CAPACITY, "more keypoints than max_features");
    CK(cudaSetDevice(ctx->device));
    const DevBuffers &B = ctx->B;
    const size_t sc = ((size_t)slot * VELO_NUM_KP_SETS + set) * B.C + cam;
    if (F > 0) CK(cudaMemcpyAsync(B.kp + sc * B.F, kp, (size_t)F * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(B.n_kp + sc, &F, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    launch_assoc(launcher(ctx), B, ctx->dcal, slot, 1, set, 1, cam, 1);
    CK(cudaGetLastError());
    int nh = 0;
    CK(cudaMemcpyAsync(&nh, B.n_hits + sc, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (has_depth && F > 0) CK(cudaMemcpyAsync(has_depth, B.has_depth + sc * B.F, (size_t)F * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (kpwd && nh > 0) { CK(cudaMemcpyAsync(kpwd, B.kpwd + sc * B.F, (size_t)nh * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream)); }
    if (n_hits) *n_hits = nh;
    return VELO_OK;
}

extern "C" int velo_gpu_assoc_upload(velo_gpu_ctx *ctx, int slot, int cam, int set, const float *kp, int F, const int *has_depth, const float *kpwd, int n_hits) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (check_slot(ctx, slot) || check_cam(ctx, cam)) return VELO_ERR_INVALID_ARG;
    if (set < 0 || set >= VELO_NUM_KP_SETS || F < 0 || n_hits < 0 || n_hits > F || (F > 0 && (!kp || !has_depth)) || (n_hits > 0 && !kpwd))
        return fail(ctx, VELO_ERR_INVALID_ARG, "bad association arrays");
    if (F > ctx->B.F) return fail(ctx, VELO_ERR_CAPACITY, "more keypoints than max_features");
    for (int i = 0; i < F; i++) if (has_depth[i] < -1 || has_depth[i] >= n_hits) return fail(ctx, VELO_ERR_INVALID_ARG, "has_depth entry outside keypoints_with_depth");
    CK(cudaSetDevice(ctx->device));
    const DevBuffers &B = ctx->B;
    const size_t sc = ((size_t)slot * VELO_NUM_KP_SETS + set) * B.C + cam;
    if (F > 0) {
        CK(cudaMemcpyAsync(B.kp + sc * B.F, kp, (size_t)F * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(B.has_depth + sc * B.F, has_depth, (size_t)F * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (n_hits > 0) CK(cudaMemcpyAsync(B.kpwd + sc * B.F, kpwd, (size_t)n_hits * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(B.n_kp + sc, &F, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(B.n_hits + sc, &n_hits, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));      // F / n_hits are stack variables; the caller's arrays may go away
    return VELO_OK;
}

static void init_icp_unit(const velo_gpu_ctx *ctx, IcpUnit *u, int src, int tgt, int skip) {
    memset(u, 0, sizeof(*u));
    u->src_slot = src; u->tgt_slot = tgt; u->skip = skip; u->n_pass = 0;
    u->norm_thr_f = ceil_to_float(ctx->prm.icp_norm_condition);                            // velo.h:873
    u->loss_a = ctx->prm.loss_thresh_3DPD; u->weight = ctx->prm.weight_3DPD;               // velo.h:885-891
}
static void add_icp_pass(const velo_gpu_ctx *ctx, IcpUnit *u, const double pose[6], int iter) {
    IcpPass *p = &u->pass[u->n_pass++];
    const double thr = ctx->prm.correspondence_thresh_icp / iter / iter / iter / iter;     // velo.h:829
    p->thr_f = floor_to_float(thr);
    p->thr_excl = nextafterf(p->thr_f, INFINITY);
    p->iter = iter;
    make_pose_pack(pose, &p->pose);
}
static int auto_ctas(const velo_gpu_ctx *ctx, int n_units, int cap) {
    int c = ctx->prm.ctas_per_icp_unit;
    if (c <= 0) { c = (ctx->sm_count * 6 * 10 + n_units - 1) / n_units; if (c < 8) c = 8; }   // ~10 waves of 6 CTAs/SM
    if (c > cap) c = cap;
    return c < 1 ? 1 : c;
}

// scratch records of the single-frame entry points (never the batch path's slot records)
static IcpUnit *sc_h_icp(velo_gpu_ctx *c) { return c->h_icp_units + c->icp_scratch; }
static IcpUnit *sc_d_icp(velo_gpu_ctx *c) { return c->d_icp_units + c->icp_scratch; }
static double *sc_icp_partial(velo_gpu_ctx *c) { return c->d_icp_partial + (size_t)c->icp_scratch * launch_icp_runs_cap(c->B.N) * VELO_MAX_PASSES * 64; }
static double *sc_icp_out(velo_gpu_ctx *c) { return c->d_icp_out + (size_t)c->icp_scratch * c->B.P * VELO_NEQ_STRIDE; }
static VisUnit *sc_h_vis(velo_gpu_ctx *c) { return c->h_vis_units + c->vis_scratch; }
static VisUnit *sc_d_vis(velo_gpu_ctx *c) { return c->d_vis_units + c->vis_scratch; }
static double *sc_vis_partial(velo_gpu_ctx *c) { return c->d_vis_partial + (size_t)c->vis_scratch * 4 * 64; }
static double *sc_vis_out(velo_gpu_ctx *c) { return c->d_vis_out + (size_t)c->vis_scratch * VELO_NEQ_STRIDE; }

// one frame pair, n_passes supplied poses, ONE launch of the fused kernel (pass p+1 seeded by pass p) — the code path of the batch
extern "C" int velo_gpu_icp_passes(velo_gpu_ctx *ctx, int slot_M, int slot_S, const double *poses, const int *iters, int n_passes, int icp_skip,
                                   velo_icp_corr *corr, int corr_capacity, int *n_queries, double *neq) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (check_slot(ctx, slot_M) || check_slot(ctx, slot_S)) return VELO_ERR_INVALID_ARG;
    if (!poses || !iters || icp_skip < 1) return fail(ctx, VELO_ERR_INVALID_ARG, "bad poses / iters / icp_skip");
    if (n_passes < 1 || n_passes > ctx->B.P) return fail(ctx, VELO_ERR_CAPACITY, "n_passes exceeds max_icp_passes");
    for (int p = 0; p < n_passes; p++) if (iters[p] < 1) return fail(ctx, VELO_ERR_INVALID_ARG, "iter must be >= 1");
    CK(cudaSetDevice(ctx->device));
    int st = slot_status(ctx, slot_M); if (st) return st;
    st = slot_status(ctx, slot_S); if (st) return st;
    IcpUnit *u = sc_h_icp(ctx);
    init_icp_unit(ctx, u, slot_M, slot_S, icp_skip);
    for (int p = 0; p < n_passes; p++) add_icp_pass(ctx, u, poses + 6 * p, iters[p]);
    CK(cudaMemcpyAsync(sc_d_icp(ctx), u, sizeof(IcpUnit), cudaMemcpyHostToDevice, ctx->stream));
    const int ctas = auto_ctas(ctx, 1, ctx->icp_partial_ctas);
    // every query writes its record in every pass, so the records need no clearing
    launch_icp(launcher(ctx), ctx->B, ctx->dcal, sc_d_icp(ctx), 1, n_passes, ctas, sc_icp_partial(ctx), sc_icp_out(ctx), ctx->B.P,
               corr ? ctx->d_corr : nullptr, corr ? ctx->B.N : 0, nullptr, 0, ctx->search_stats);
    CK(cudaGetLastError());
    std::vector<double> out((size_t)n_passes * VELO_NEQ_STRIDE);
    CK(cudaMemcpyAsync(out.data(), sc_icp_out(ctx), out.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));     // also: the scratch unit is host memory reused by the next call
    const int nq = (int)out[58];
    if (corr) {
        if (nq > corr_capacity) return fail(ctx, VELO_ERR_CAPACITY, "corr_capacity smaller than the number of queries");
        if (nq > 0) CK(cudaMemcpy2DAsync(corr, (size_t)corr_capacity * sizeof(velo_icp_corr), ctx->d_corr, (size_t)ctx->B.N * sizeof(velo_icp_corr),
                                         (size_t)nq * sizeof(velo_icp_corr), n_passes, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (n_queries) *n_queries = nq;
    if (neq) memcpy(neq, out.data(), out.size() * sizeof(double));
    return VELO_OK;
}

extern "C" int velo_gpu_icp_pass(velo_gpu_ctx *ctx, int slot_M, int slot_S, const double pose[6], int iter, int icp_skip,
                                 velo_icp_corr *corr, int corr_capacity, int *n_queries, int *n_kept, double *neq) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (!pose || iter < 1) return fail(ctx, VELO_ERR_INVALID_ARG, "bad pose / iter");
    double out[VELO_NEQ_STRIDE];
    const int rc = velo_gpu_icp_passes(ctx, slot_M, slot_S, pose, &iter, 1, icp_skip, corr, corr_capacity, n_queries, out);
    if (rc) return rc;
    if (n_kept) *n_kept = (int)out[56];
    if (neq) memcpy(neq, out, sizeof(out));
    return VELO_OK;
}

static VisTun vis_tun(const velo_gpu_ctx *ctx) {
    const velo_gpu_params &p = ctx->prm;
    return VisTun{ p.weight_3D2D, p.weight_2D2D, p.loss_thresh_3D2D, p.loss_thresh_2D2D, p.loss_thresh_3D3D, p.outlier_reject,
                   p.enable_2d2d, p.enable_3d2d, p.abs_truncates, 0 };
}

extern "C" int velo_gpu_visual_residuals(velo_gpu_ctx *ctx, int slot1, int set1, int slot2, int set2, const int *n_matches, const int *matches,
                                         const int *lm_valid, const float *lm_xyz, const double pose[6], int iter,
                                         velo_vis_block *blocks, int block_capacity, int *n_blocks, double *neq) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (check_slot(ctx, slot1) || check_slot(ctx, slot2)) return VELO_ERR_INVALID_ARG;
    if (!n_matches || !pose || iter < 1 || set1 < 0 || set1 >= VELO_NUM_KP_SETS || set2 < 0 || set2 >= VELO_NUM_KP_SETS)
        return fail(ctx, VELO_ERR_INVALID_ARG, "bad argument");
    CK(cudaSetDevice(ctx->device));
    const DevBuffers &B = ctx->B;
    const int C = B.C, MM = B.MM;
    int tot = 0;
    for (int c = 0; c < C; c++) { if (n_matches[c] < 0 || n_matches[c] > MM) return fail(ctx, VELO_ERR_CAPACITY, "more matches than max_matches"); tot += n_matches[c]; }
    if (tot > 0 && !matches) return fail(ctx, VELO_ERR_INVALID_ARG, "null matches");
    // matches arrive concatenated per camera; the device layout is [cam][MM]
    int off = 0;
    for (int c = 0; c < C; c++) {
        if (n_matches[c] > 0) {
            CK(cudaMemcpyAsync(B.matches + 2 * (((size_t)slot1 * C + c) * MM), matches + 2 * (size_t)off, (size_t)n_matches[c] * 2 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            if (lm_valid) {
                CK(cudaMemcpyAsync(ctx->d_lm_valid + (size_t)c * MM, lm_valid + off, (size_t)n_matches[c] * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
                CK(cudaMemcpyAsync(ctx->d_lm_xyz + (size_t)c * MM, lm_xyz + 4 * (size_t)off, (size_t)n_matches[c] * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
            }
        }
        off += n_matches[c];
    }
    CK(cudaMemcpyAsync(B.n_matches + (size_t)slot1 * C, n_matches, C * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    VisUnit *u = sc_h_vis(ctx);
    memset(u, 0, sizeof(*u));
    u->slot1 = slot1; u->set1 = set1; u->slot2 = slot2; u->set2 = set2; u->iter = iter;
    memcpy(u->pose, pose, 6 * sizeof(double));
    CK(cudaMemcpyAsync(sc_d_vis(ctx), u, sizeof(VisUnit), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_mout, 0, (size_t)C * MM * sizeof(VisMatchOut), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    const int ctas = 32;
    launch_visual(launcher(ctx), B, ctx->dcal, sc_d_vis(ctx), 1, vis_tun(ctx), lm_valid ? ctx->d_lm_valid : nullptr, lm_valid ? ctx->d_lm_xyz : nullptr,
                  sc_vis_partial(ctx), sc_vis_out(ctx), ctx->d_mout, ctas, VisFixed{ nullptr, nullptr, nullptr, 0, 0 }, ctx->d_flags);
    CK(cudaGetLastError());
    double out[VELO_NEQ_STRIDE];
    int bad = 0;
    CK(cudaMemcpyAsync(out, sc_vis_out(ctx), sizeof(out), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&bad, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<VisMatchOut> mo((size_t)C * MM);
    CK(cudaMemcpyAsync(mo.data(), ctx->d_mout, mo.size() * sizeof(VisMatchOut), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (bad) return fail(ctx, VELO_ERR_INVALID_ARG, "a match index is outside the keypoint set it refers to");
    // residualStats order: camera-major, match order, block order (velo.h:934-976)
    int nb = 0;
    for (int c = 0; c < C; c++) for (int i = 0; i < n_matches[c]; i++) {
        const VisMatchOut &m = mo[(size_t)c * MM + i];
        for (int b = 0; b < m.n; b++) { if (blocks && nb < block_capacity) blocks[nb] = m.b[b]; nb++; }
    }
    if (blocks && nb > block_capacity) return fail(ctx, VELO_ERR_CAPACITY, "block_capacity too small");
    if (n_blocks) *n_blocks = nb;
    if (neq) memcpy(neq, out, sizeof(out));
    return VELO_OK;
}

// ------------------------------------------------------------------------------------------------ device-resident frameToFrame (f1)
static int check_range(velo_gpu_ctx *ctx, int slot0, int count);
#define F2F_EVAL_CTAS_MAX 148
#define F2F_VIS_CTAS_MAX 32
struct F2FJob { int slot_M, set1, slot_S, set2; };

// buffers for n simultaneous solves (kept; grown when a larger batch arrives)
static int f2f_ensure(velo_gpu_ctx *ctx, int n) {
    if (n <= ctx->f2f_cap) return VELO_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    f2f_free(ctx);
    const DevBuffers &B = ctx->B;
    const size_t N = B.N, CM = (size_t)B.C * B.MM, runs = (size_t)launch_icp_runs_cap(B.N);
#define F2F_ALLOC(ptr, count) do { CK(cudaMalloc((void **)&(ptr), (count) * sizeof(*(ptr)))); CK(cudaMemsetAsync((ptr), 0, (count) * sizeof(*(ptr)), ctx->stream)); } while (0)
    F2F_ALLOC(ctx->d_f2f_lm, (size_t)n); F2F_ALLOC(ctx->d_f2f_icp_units, (size_t)n); F2F_ALLOC(ctx->d_f2f_vis_units, (size_t)n);
    F2F_ALLOC(ctx->d_f2f_frozen, (size_t)n * N); F2F_ALLOC(ctx->d_f2f_sel, (size_t)n * CM); F2F_ALLOC(ctx->d_f2f_poses, (size_t)n * 6);
    F2F_ALLOC(ctx->d_f2f_icp_partial, (size_t)n * runs * VELO_MAX_PASSES * 64); F2F_ALLOC(ctx->d_f2f_icp_out, (size_t)n * VELO_NEQ_STRIDE);
    F2F_ALLOC(ctx->d_f2f_eval_partial, (size_t)n * F2F_EVAL_CTAS_MAX * 64); F2F_ALLOC(ctx->d_f2f_eval_out, (size_t)n * VELO_NEQ_STRIDE);
    F2F_ALLOC(ctx->d_f2f_vis_partial, (size_t)n * F2F_VIS_CTAS_MAX * 64); F2F_ALLOC(ctx->d_f2f_vis_out, (size_t)n * VELO_NEQ_STRIDE);
    F2F_ALLOC(ctx->d_f2f_ndone, (size_t)4);
#undef F2F_ALLOC
    CK(cudaMallocHost((void **)&ctx->h_f2f_lm, (size_t)n * sizeof(LmState)));
    CK(cudaMallocHost((void **)&ctx->h_f2f_icp_units, (size_t)n * sizeof(IcpUnit)));
    CK(cudaMallocHost((void **)&ctx->h_f2f_vis_units, (size_t)n * sizeof(VisUnit)));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->f2f_cap = n;
    return VELO_OK;
}

// frameToFrame (velo.h:616-907) for n independent frame pairs at once.  Per f2f iteration the visual block list of every pair is
// frozen at its current pose (k_visual, free mode, type masks kept per pair); per ICP iteration the correspondences are frozen
// (k_icp_pass writes compact records per pair); then all pairs run their Levenberg-Marquardt solves side by side: evaluation
// kernels over (pair, CTA), one controller thread per pair, converged pairs drop out.  The host looks at a counter of finished
// pairs every 8 controller steps, and once per solve reads the n poses back: the pose constants of the NEXT correspondence pass
// (cos / sin / 1/theta) are formed with the host libm so that the narrowed float queries equal the reference's bit for bit (H8).
static int f2f_run(velo_gpu_ctx *ctx, int n, const F2FJob *jobs, bool visual, const int *d_lmv, const float4 *d_lmx, int enable_icp, int icp_skip,
                   double *transforms, velo_f2f_report *reports) {
    const int F2F = ctx->prm.f2f_iterations, ICP = enable_icp ? ctx->prm.icp_iterations : 1;
    if (F2F < 1 || ICP < 1 || F2F * ICP > VELO_MAX_SOLVES) return fail(ctx, VELO_ERR_CAPACITY, "f2f_iterations * icp_iterations exceeds VELO_MAX_SOLVES");
    int rc = f2f_ensure(ctx, n); if (rc) return rc;
    const DevBuffers &B = ctx->B;
    Launcher L = launcher(ctx);
    const VisTun tun = vis_tun(ctx);
    const int max_lm = 50, CM = B.C * B.MM;
    const int eval_ctas = std::max(1, std::min(F2F_EVAL_CTAS_MAX, (ctx->sm_count * 8 + n - 1) / n));
    const int vis_ctas = n == 1 ? F2F_VIS_CTAS_MAX : 4;
    const int icp_ctas = auto_ctas(ctx, n, n == 1 ? ctx->icp_partial_ctas : 32);
    if (reports) memset(reports, 0, (size_t)n * sizeof(velo_f2f_report));
    CK(cudaMemcpyAsync(ctx->d_f2f_poses, transforms, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    launch_lm_init(L, ctx->d_f2f_lm, n, ctx->d_f2f_poses, max_lm, ctx->d_f2f_ndone);
    CK(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    int solve = 0;
    for (int iter = 1; iter <= F2F; iter++) {
        if (visual) {   // freeze the visual block list of this f2f iteration at each pair's current pose (velo.h:622-792)
            CK(cudaStreamSynchronize(ctx->stream));     // the unit staging below is host memory an earlier copy may still read
            for (int u = 0; u < n; u++) {
                VisUnit *vu = &ctx->h_f2f_vis_units[u];
                memset(vu, 0, sizeof(*vu));
                vu->slot1 = jobs[u].slot_M; vu->set1 = jobs[u].set1; vu->slot2 = jobs[u].slot_S; vu->set2 = jobs[u].set2; vu->iter = iter;
            }
            CK(cudaMemcpyAsync(ctx->d_f2f_vis_units, ctx->h_f2f_vis_units, (size_t)n * sizeof(VisUnit), cudaMemcpyHostToDevice, ctx->stream));
            launch_visual(L, B, ctx->dcal, ctx->d_f2f_vis_units, n, tun, d_lmv, d_lmx, ctx->d_f2f_vis_partial, ctx->d_f2f_vis_out, nullptr, vis_ctas,
                          VisFixed{ nullptr, ctx->d_f2f_sel, ctx->d_f2f_lm, CM, 1 }, ctx->d_flags);
        }
        for (int ii = 0; ii < ICP; ii++, solve++) {
            if (enable_icp) {   // freeze the ICP correspondences at each pair's current pose (velo.h:806-894)
                CK(cudaStreamSynchronize(ctx->stream));
                for (int u = 0; u < n; u++) {
                    IcpUnit *iu = &ctx->h_f2f_icp_units[u];
                    init_icp_unit(ctx, iu, jobs[u].slot_M, jobs[u].slot_S, icp_skip);
                    add_icp_pass(ctx, iu, transforms + 6 * (size_t)u, iter);
                }
                CK(cudaMemcpyAsync(ctx->d_f2f_icp_units, ctx->h_f2f_icp_units, (size_t)n * sizeof(IcpUnit), cudaMemcpyHostToDevice, ctx->stream));
                launch_icp(L, B, ctx->dcal, ctx->d_f2f_icp_units, n, 1, icp_ctas, ctx->d_f2f_icp_partial, ctx->d_f2f_icp_out, 1, nullptr, 0, ctx->d_f2f_frozen, B.N);
            }
            launch_lm_init(L, ctx->d_f2f_lm, n, nullptr, max_lm, ctx->d_f2f_ndone);
            int n_done = 0;
            for (int ev = 0; ev <= max_lm && n_done < n;) {
                for (int k = 0; k < 8 && ev <= max_lm; k++, ev++) {   // a few controller steps per host look
                    if (visual)
                        launch_visual(L, B, ctx->dcal, ctx->d_f2f_vis_units, n, tun, d_lmv, d_lmx, ctx->d_f2f_vis_partial, ctx->d_f2f_vis_out, nullptr, vis_ctas,
                                      VisFixed{ ctx->d_f2f_sel, nullptr, ctx->d_f2f_lm, CM, 0 }, nullptr);
                    if (enable_icp)
                        launch_icp_eval(L, B, n, ctx->d_f2f_frozen, B.N, ctx->d_f2f_icp_out, ctx->d_f2f_icp_units, ctx->d_f2f_lm, ctx->prm.loss_thresh_3DPD,
                                        ctx->prm.weight_3DPD, ctx->d_f2f_eval_partial, eval_ctas, ctx->d_f2f_eval_out);
                    launch_lm_step(L, ctx->d_f2f_lm, n, enable_icp ? ctx->d_f2f_eval_out : nullptr, visual ? ctx->d_f2f_vis_out : nullptr, ctx->d_f2f_ndone);
                }
                CK(cudaMemcpyAsync(&n_done, ctx->d_f2f_ndone, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
            }
            CK(cudaMemcpyAsync(ctx->h_f2f_lm, ctx->d_f2f_lm, (size_t)n * sizeof(LmState), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            CK(cudaGetLastError());
            for (int u = 0; u < n; u++) {
                const LmState &hs = ctx->h_f2f_lm[u];
                memcpy(transforms + 6 * (size_t)u, hs.x, 6 * sizeof(double));
                if (reports) {
                    velo_f2f_report &rep = reports[u];
                    const int k = rep.n_solves++;
                    rep.lm_iterations[k] = hs.iter; rep.accepted_steps[k] = hs.accepted; rep.reason[k] = hs.reason; rep.n_blocks[k] = hs.n_blocks;
                    rep.initial_cost[k] = hs.init_cost; rep.final_cost[k] = hs.cost;
                    memcpy(rep.pose[k], hs.x, 6 * sizeof(double));
                }
            }
        }
    }
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->f2f_last_n = n;
    if (bad) return fail(ctx, VELO_ERR_INVALID_ARG, "a match index is outside the keypoint set it refers to");
    return VELO_OK;
}

extern "C" int velo_gpu_frame_to_frame(velo_gpu_ctx *ctx, int slot_M, int set1, int slot_S, int set2,
                                       const int *n_matches, const int *matches, const int *lm_valid, const float *lm_xyz,
                                       int enable_icp, int icp_skip, double transform[6], velo_f2f_report *report) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (check_slot(ctx, slot_M) || check_slot(ctx, slot_S)) return VELO_ERR_INVALID_ARG;
    if (!transform || icp_skip < 1 || set1 < 0 || set1 >= VELO_NUM_KP_SETS || set2 < 0 || set2 >= VELO_NUM_KP_SETS) return fail(ctx, VELO_ERR_INVALID_ARG, "bad argument");
    CK(cudaSetDevice(ctx->device));
    int st = slot_status(ctx, slot_M); if (st) return st;
    st = slot_status(ctx, slot_S); if (st) return st;
    const DevBuffers &B = ctx->B;
    const int C = B.C, MM = B.MM;
    const bool visual = n_matches != nullptr;
    if (visual) {       // install the match lists (same layout as velo_gpu_visual_residuals)
        int off = 0, tot = 0;
        for (int c = 0; c < C; c++) { if (n_matches[c] < 0 || n_matches[c] > MM) return fail(ctx, VELO_ERR_CAPACITY, "more matches than max_matches"); tot += n_matches[c]; }
        if (tot > 0 && !matches) return fail(ctx, VELO_ERR_INVALID_ARG, "null matches");
        for (int c = 0; c < C; c++) {
            if (n_matches[c] > 0) {
                CK(cudaMemcpyAsync(B.matches + 2 * (((size_t)slot_M * C + c) * MM), matches + 2 * (size_t)off, (size_t)n_matches[c] * 2 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
                if (lm_valid) {
                    CK(cudaMemcpyAsync(ctx->d_lm_valid + (size_t)c * MM, lm_valid + off, (size_t)n_matches[c] * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
                    CK(cudaMemcpyAsync(ctx->d_lm_xyz + (size_t)c * MM, lm_xyz + 4 * (size_t)off, (size_t)n_matches[c] * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
                }
            }
            off += n_matches[c];
        }
        CK(cudaMemcpyAsync(B.n_matches + (size_t)slot_M * C, n_matches, C * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));     // the caller's arrays may go away
    }
    const F2FJob job = { slot_M, set1, slot_S, set2 };
    return f2f_run(ctx, 1, &job, visual, (visual && lm_valid) ? ctx->d_lm_valid : nullptr, (visual && lm_valid) ? ctx->d_lm_xyz : nullptr,
                   enable_icp, icp_skip, transform, report);
}

// the same for the frame pairs (slot s, slot s-1) of a batch whose ingest / index / projection / association stages have run
// (velo_gpu_batch_run): frame1 = slot s with its tracked keypoints (set 1), frame2 = slot s-1 with its detected keypoints (set 0),
// matches as uploaded with the batch; no landmarks.
extern "C" int velo_gpu_batch_frame_to_frame(velo_gpu_ctx *ctx, int slot0, int count, int first_has_prev, int enable_visual, int enable_icp,
                                             double *transforms, velo_f2f_report *reports) {
    if (!ctx || !transforms) return VELO_ERR_INVALID_ARG;
    if (check_range(ctx, slot0, count)) return VELO_ERR_INVALID_ARG;
    if (!enable_visual && !enable_icp) return fail(ctx, VELO_ERR_INVALID_ARG, "nothing to solve: neither visual nor ICP terms");
    CK(cudaSetDevice(ctx->device));
    const int skip_first = (first_has_prev && slot0 > 0) ? 0 : 1;
    const int n = count - skip_first;
    if (n <= 0) return VELO_OK;
    std::vector<int> stv(count);
    CK(cudaMemcpyAsync(stv.data(), ctx->B.status + slot0, count * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < count; i++) if (stv[i] != 0) return fail(ctx, stv[i], "a scan of the batch has more rings than max_rings");
    std::vector<F2FJob> jobs(n);
    for (int u = 0; u < n; u++) { const int s = slot0 + skip_first + u; jobs[u] = F2FJob{ s, 1, s - 1, 0 }; }
    return f2f_run(ctx, n, jobs.data(), enable_visual != 0, nullptr, nullptr, enable_icp, ctx->prm.icp_skip,
                   transforms + 6 * (size_t)skip_first, reports ? reports + skip_first : nullptr);
}

extern "C" int velo_gpu_f2f_selection(velo_gpu_ctx *ctx, unsigned char *sel, int capacity) {
    if (!ctx || !sel) return VELO_ERR_INVALID_ARG;
    const size_t n = (size_t)ctx->B.C * ctx->B.MM;
    if ((size_t)capacity < n) return fail(ctx, VELO_ERR_CAPACITY, "capacity smaller than num_cams * max_matches");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->d_f2f_sel || ctx->f2f_last_n != 1) return fail(ctx, VELO_ERR_STATE, "no preceding velo_gpu_frame_to_frame call");
    CK(cudaMemcpyAsync(sel, ctx->d_f2f_sel, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VELO_OK;
}

// util::pose_mat2vec (utility.h:67-82): ceres::AngleAxisToRotationMatrix [recall: for theta^2 > DBL_EPSILON the Rodrigues form
// R = cos I + (1 - cos) a a^T + sin [a]x with a = w / theta, else the first-order I + [w]x] + translation
extern "C" int velo_pose_vec2mat(const double x[6], double T[16]) {
    if (!x || !T) return VELO_ERR_INVALID_ARG;
    double R[3][3];
    const double th2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    if (th2 > DBL_EPSILON) {
        const double th = sqrt(th2), wx = x[0] / th, wy = x[1] / th, wz = x[2] / th, c = cos(th), s = sin(th);
        R[0][0] = c + wx * wx * (1.0 - c);      R[1][0] = wz * s + wx * wy * (1.0 - c); R[2][0] = -wy * s + wx * wz * (1.0 - c);
        R[0][1] = wx * wy * (1.0 - c) - wz * s; R[1][1] = c + wy * wy * (1.0 - c);      R[2][1] = wx * s + wy * wz * (1.0 - c);
        R[0][2] = wy * s + wx * wz * (1.0 - c); R[1][2] = -wx * s + wy * wz * (1.0 - c); R[2][2] = c + wz * wz * (1.0 - c);
    } else {
        R[0][0] = 1.0; R[1][0] = x[2]; R[2][0] = -x[1];
        R[0][1] = -x[2]; R[1][1] = 1.0; R[2][1] = x[0];
        R[0][2] = x[1]; R[1][2] = -x[0]; R[2][2] = 1.0;
    }
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) T[4 * i + j] = R[i][j]; T[4 * i + 3] = x[3 + i]; }
    T[12] = T[13] = T[14] = 0.0; T[15] = 1.0;
    return VELO_OK;
}

// ------------------------------------------------------------------------------------------------ batched triangulation (f3)
extern "C" int velo_gpu_triangulate(velo_gpu_ctx *ctx, int n, const int *off3, const velo_tri_obs3 *obs3, const int *off2, const velo_tri_obs2 *obs2,
                                    const double *camera_poses, int n_frames, const float *init_xyz, const int *has_init, float *out_xyz, int *iterations) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (n < 0 || n_frames < 0 || (n > 0 && (!off3 || !off2 || !out_xyz)) || (has_init && !init_xyz)) return fail(ctx, VELO_ERR_INVALID_ARG, "bad triangulation arguments");
    if (n == 0) return VELO_OK;
    for (int l = 0; l < n; l++) if (off3[l + 1] < off3[l] || off2[l + 1] < off2[l]) return fail(ctx, VELO_ERR_INVALID_ARG, "offsets must be non-decreasing");
    const int n3 = off3[n] - off3[0], n2 = off2[n] - off2[0];
    if (off3[0] != 0 || off2[0] != 0 || (n3 > 0 && !obs3) || (n2 > 0 && !obs2) || ((n3 + n2) > 0 && !camera_poses)) return fail(ctx, VELO_ERR_INVALID_ARG, "bad observation arrays");
    CK(cudaSetDevice(ctx->device));
    // scratch: one arena kept by the context and grown on demand (no allocation on the steady-state path)
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_off = up((size_t)(n + 1) * sizeof(int)), b_o3 = up((size_t)std::max(n3, 1) * sizeof(velo_tri_obs3)), b_o2 = up((size_t)std::max(n2, 1) * sizeof(velo_tri_obs2)),
                 b_pose = up((size_t)std::max(n_frames, 1) * 6 * sizeof(double)), b_xyz = up((size_t)n * 3 * sizeof(float)), b_int = up((size_t)n * sizeof(int));
    const size_t need = 2 * b_off + b_o3 + b_o2 + b_pose + 2 * b_xyz + 2 * b_int;
    if (need > ctx->tri_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_tri) { cudaFree(ctx->d_tri); ctx->d_tri = nullptr; ctx->tri_cap = 0; }
        CK(cudaMalloc((void **)&ctx->d_tri, need + need / 2));
        ctx->tri_cap = need + need / 2;
    }
    char *a = ctx->d_tri;
    int *d_off3 = (int *)a; a += b_off; int *d_off2 = (int *)a; a += b_off;
    velo_tri_obs3 *d_o3 = (velo_tri_obs3 *)a; a += b_o3; velo_tri_obs2 *d_o2 = (velo_tri_obs2 *)a; a += b_o2;
    double *d_poses = (double *)a; a += b_pose;
    float *d_out = (float *)a; a += b_xyz; float *d_init = (float *)a; a += b_xyz;
    int *d_it = (int *)a; a += b_int; int *d_has = (int *)a;
    CK(cudaMemcpyAsync(d_off3, off3, (n + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_off2, off2, (n + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if (n3 > 0) CK(cudaMemcpyAsync(d_o3, obs3, (size_t)n3 * sizeof(velo_tri_obs3), cudaMemcpyHostToDevice, ctx->stream));
    if (n2 > 0) CK(cudaMemcpyAsync(d_o2, obs2, (size_t)n2 * sizeof(velo_tri_obs2), cudaMemcpyHostToDevice, ctx->stream));
    if (n_frames > 0) CK(cudaMemcpyAsync(d_poses, camera_poses, (size_t)n_frames * 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (has_init) { CK(cudaMemcpyAsync(d_has, has_init, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream)); CK(cudaMemcpyAsync(d_init, init_xyz, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream)); }
    launch_triangulate(launcher(ctx), n, d_off3, d_o3, d_off2, d_o2, d_poses, n_frames, ctx->dcal, ctx->prm.loss_thresh_3D2D, ctx->prm.weight_3D2D,
                       has_init ? d_init : nullptr, has_init ? d_has : nullptr, d_out, d_it);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_xyz, d_out, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (iterations) CK(cudaMemcpyAsync(iterations, d_it, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VELO_OK;
}

// ------------------------------------------------------------------------------------------------ Hamming matcher (f4)
extern "C" int velo_gpu_match_hamming(velo_gpu_ctx *ctx, const uint8_t *query, int n_query, const uint8_t *train, int n_train, int desc_bytes,
                                      double match_thresh, int *pairs, int *n_pairs, int *best_idx, int *best_dist) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (n_query < 0 || n_train < 0 || desc_bytes < 8 || desc_bytes > 64 || (desc_bytes % 8) || !n_pairs || (n_query > 0 && (!query || !pairs)) || (n_train > 0 && !train))
        return fail(ctx, VELO_ERR_INVALID_ARG, "bad descriptor arguments (desc_bytes must be a multiple of 8, at most 64)");
    CK(cudaSetDevice(ctx->device));
    *n_pairs = 0;
    if (n_query == 0) return VELO_OK;
    const int words = desc_bytes / 8;
    const size_t qb = (size_t)n_query * desc_bytes, tb = (size_t)n_train * desc_bytes;
    if (qb > ctx->ham_cap_q) {
        if (ctx->d_hq) { cudaFree(ctx->d_hq); cudaFree(ctx->d_hbest); ctx->d_hq = nullptr; ctx->d_hbest = nullptr; ctx->ham_cap_q = 0; }
        CK(cudaMalloc((void **)&ctx->d_hq, qb)); CK(cudaMalloc((void **)&ctx->d_hbest, (size_t)n_query * sizeof(unsigned long long)));
        ctx->ham_cap_q = qb;
    }
    if (tb > ctx->ham_cap_t) { if (ctx->d_ht) cudaFree(ctx->d_ht); CK(cudaMalloc((void **)&ctx->d_ht, tb ? tb : 8)); ctx->ham_cap_t = tb; }
    CK(cudaMemcpyAsync(ctx->d_hq, query, qb, cudaMemcpyHostToDevice, ctx->stream));
    if (tb) CK(cudaMemcpyAsync(ctx->d_ht, train, tb, cudaMemcpyHostToDevice, ctx->stream));
    launch_hamming(launcher(ctx), ctx->d_hq, n_query, ctx->d_ht, n_train, words, ctx->d_hbest, ctx->sm_count);
    CK(cudaGetLastError());
    std::vector<unsigned long long> key(n_query);
    CK(cudaMemcpyAsync(key.data(), ctx->d_hbest, (size_t)n_query * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<int> idx(n_query), dist(n_query);
    for (int i = 0; i < n_query; i++) {
        if (key[i] == ~0ull) { idx[i] = -1; dist[i] = 0x7fffffff; }               // empty train set
        else { idx[i] = (int)(key[i] & 0xffffffffull); dist[i] = (int)(key[i] >> 32); }
    }
    if (best_idx) memcpy(best_idx, idx.data(), n_query * sizeof(int));
    if (best_dist) memcpy(best_dist, dist.data(), n_query * sizeof(int));
    if (n_train == 0) return VELO_OK;                      // BFMatcher returns no matches for an empty train set
    // velo.h:536-548: DMatch::distance is a float; the comparison promotes to double
    double min_dist = 1e9;
    for (int i = 0; i < n_query; i++) if ((double)(float)dist[i] < min_dist) min_dist = (double)(float)dist[i];
    const double lim = (1.5 * min_dist > match_thresh) ? 1.5 * min_dist : match_thresh;
    int n = 0;
    for (int i = 0; i < n_query; i++) {
        if ((double)(float)dist[i] > lim) continue;
        pairs[2 * n] = i; pairs[2 * n + 1] = idx[i]; n++;
    }
    *n_pairs = n;
    return VELO_OK;
}

// ------------------------------------------------------------------------------------------------ batched path
static int check_range(velo_gpu_ctx *ctx, int slot0, int count) {
    if (slot0 < 0 || count < 1 || slot0 + count > ctx->B.S) return fail(ctx, VELO_ERR_INVALID_ARG, "slot range out of bounds");
    return VELO_OK;
}

// copy the inputs of batch entries [i0, i0+count) (slots slot0+i0 ...) to the device on `st`
// st_clear: the stream that voids the old projections of the slots.  A memset is a (tiny) KERNEL: on a copy stream it would queue for
// an SM slot behind the resident CTAs of the correspondence kernel (milliseconds each) and stall every copy issued after it — measured
// as 28 instead of 50 GB/s once the GPU was busy — so the pipelined call gives the chunk's compute stream here.
static int upload_range(velo_gpu_ctx *ctx, int slot0_all, int i0, int count, const velo_batch_inputs *src, cudaStream_t st, cudaStream_t st_clear) {
    const int slot0 = slot0_all + i0;
    velo_batch_inputs inl = *src, *in = &inl;
    {
        const size_t C_ = ctx->B.C, F_ = ctx->B.F, MM_ = ctx->B.MM, per_ = VELO_NUM_KP_SETS * C_;
        if (inl.scans) inl.scans += (size_t)i0 * ctx->prm.max_points * (inl.scan_stride_floats == 3 ? 3 : 4);
        if (inl.n_points) inl.n_points += i0;
        if (inl.kp) inl.kp += (size_t)i0 * per_ * F_ * 2;
        if (inl.n_kp) inl.n_kp += (size_t)i0 * per_;
        if (inl.matches) inl.matches += (size_t)i0 * C_ * MM_ * 2;
        if (inl.n_matches) inl.n_matches += (size_t)i0 * C_;
        if (inl.icp_poses) inl.icp_poses += (size_t)i0 * inl.n_passes * 6;
        if (inl.vis_poses) inl.vis_poses += (size_t)i0 * inl.n_vis_iters * 6;
    }
    const DevBuffers &B = ctx->B;
    const size_t C = B.C, F = B.F, MM = B.MM;
    if (in->scans && in->n_points) {
        if (in->scan_stride_floats != 0 && in->scan_stride_floats != 3 && in->scan_stride_floats != 4) return fail(ctx, VELO_ERR_INVALID_ARG, "scan_stride_floats must be 0, 3 or 4");
        const size_t rec = (in->scan_stride_floats == 3 ? 3 : 4) * sizeof(float);
        int nmax = 0;
        for (int i = 0; i < count; i++) {
            if (in->n_points[i] < 0 || in->n_points[i] > ctx->prm.max_points) return fail(ctx, VELO_ERR_CAPACITY, "scan has more points than max_points");
            nmax = std::max(nmax, in->n_points[i]);
        }
        // one strided copy for the chunk; only the records the longest scan of the chunk has, not max_points per scan
        if (nmax > 0) CK(cudaMemcpy2DAsync(B.raw + (size_t)slot0 * B.N, (size_t)B.N * sizeof(float4), in->scans, (size_t)ctx->prm.max_points * rec,
                                           (size_t)nmax * rec, count, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(B.n_points + slot0, in->n_points, count * sizeof(int), cudaMemcpyHostToDevice, st));
        for (int i = 0; i < count; i++) { ctx->h_npoints[slot0 + i] = in->n_points[i]; ctx->h_stride[slot0 + i] = (int)(rec / sizeof(float)); }
        CK(cudaMemcpyAsync(B.raw_stride + slot0, ctx->h_stride.data() + slot0, count * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(B.proj_count + (size_t)slot0 * B.C * B.R, 0, (size_t)count * B.C * B.R * sizeof(int), st_clear));   // projections of the old scans are void
    }
    if (in->kp && in->n_kp) {
        const size_t per = VELO_NUM_KP_SETS * C;
        for (size_t i = 0; i < count * per; i++) if (in->n_kp[i] < 0 || in->n_kp[i] > (int)F) return fail(ctx, VELO_ERR_CAPACITY, "more keypoints than max_features");
        CK(cudaMemcpyAsync(B.kp + (size_t)slot0 * per * F, in->kp, count * per * F * sizeof(float2), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(B.n_kp + (size_t)slot0 * per, in->n_kp, count * per * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    if (in->matches && in->n_matches) {
        for (size_t i = 0; i < count * C; i++) if (in->n_matches[i] < 0 || in->n_matches[i] > (int)MM) return fail(ctx, VELO_ERR_CAPACITY, "more matches than max_matches");
        CK(cudaMemcpyAsync(B.matches + (size_t)slot0 * C * MM * 2, in->matches, count * C * MM * 2 * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(B.n_matches + (size_t)slot0 * C, in->n_matches, count * C * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    if (in->icp_poses && in->pass_iter) {
        if (in->n_passes < 1 || in->n_passes > B.P) return fail(ctx, VELO_ERR_CAPACITY, "n_passes exceeds max_icp_passes");
        ctx->batch_passes = in->n_passes;
        for (int p = 0; p < in->n_passes; p++) if (in->pass_iter[p] < 1) return fail(ctx, VELO_ERR_INVALID_ARG, "pass_iter must be >= 1");
        for (int i = 0; i < count; i++) {
            const int slot = slot0 + i;
            IcpUnit *u = &ctx->h_icp_units[slot];
            init_icp_unit(ctx, u, slot == 0 ? -1 : slot, slot - 1, ctx->prm.icp_skip);
            for (int p = 0; p < in->n_passes; p++) add_icp_pass(ctx, u, in->icp_poses + 6 * ((size_t)i * in->n_passes + p), in->pass_iter[p]);
        }
        CK(cudaMemcpyAsync(ctx->d_icp_units + slot0, ctx->h_icp_units + slot0, (size_t)count * sizeof(IcpUnit), cudaMemcpyHostToDevice, st));
    }
    if (in->vis_poses) {
        const int V = ctx->prm.f2f_iterations;
        if (in->n_vis_iters < 1 || in->n_vis_iters > V) return fail(ctx, VELO_ERR_CAPACITY, "n_vis_iters exceeds f2f_iterations");
        ctx->batch_vis = in->n_vis_iters;
        for (int i = 0; i < count; i++) for (int it = 0; it < in->n_vis_iters; it++) {
            const int slot = slot0 + i;
            VisUnit *u = &ctx->h_vis_units[(size_t)slot * V + it];
            memset(u, 0, sizeof(*u));
            // frame1 = current slot, tracked set (1); frame2 = previous slot, detected set (0)  (main.cpp:388-405, velo.h:609-610)
            u->slot1 = slot; u->set1 = 1; u->slot2 = slot - 1; u->set2 = 0; u->iter = it + 1;
            memcpy(u->pose, in->vis_poses + 6 * ((size_t)i * in->n_vis_iters + it), 6 * sizeof(double));
        }
        CK(cudaMemcpyAsync(ctx->d_vis_units + (size_t)slot0 * V, ctx->h_vis_units + (size_t)slot0 * V, (size_t)count * V * sizeof(VisUnit), cudaMemcpyHostToDevice, st));
    }
    return VELO_OK;
}

extern "C" int velo_gpu_batch_upload(velo_gpu_ctx *ctx, int slot0, int count, const velo_batch_inputs *in) {
    if (!ctx || !in) return VELO_ERR_INVALID_ARG;
    if (check_range(ctx, slot0, count)) return VELO_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    return upload_range(ctx, slot0, 0, count, in, ctx->stream, ctx->stream);
}

// the selected stages for slots [slot0, slot0 + count) on L.stream; partial-sum areas are addressed by slot, so launches for
// disjoint slot ranges may run concurrently on different streams
static int run_stages(velo_gpu_ctx *ctx, const Launcher &L, int slot0, int count, int stages, int first_has_prev) {
    const DevBuffers &B = ctx->B;
    cudaStream_t st = L.stream;
    if (stages & VELO_STAGE_INGEST) launch_ingest(L, B, ctx->dcal, slot0, count);
    if (stages & VELO_STAGE_INDEX) launch_index(L, B, ctx->dcal, slot0, count);
    if (stages & VELO_STAGE_PROJECT) launch_project(L, B, ctx->dcal, slot0, count);
    if (stages & VELO_STAGE_ASSOC) launch_assoc(L, B, ctx->dcal, slot0, count, 0, VELO_NUM_KP_SETS, 0, B.C);
    const int skip_first = (first_has_prev && slot0 > 0) ? 0 : 1;
    const int pairs = count - skip_first, s_first = slot0 + skip_first;
    if ((stages & VELO_STAGE_ICP) && pairs > 0 && ctx->batch_passes > 0) {
        const int n_units = pairs;
        const int ctas = auto_ctas(ctx, n_units, 32);
        if (skip_first) CK(cudaMemsetAsync(ctx->d_icp_out + (size_t)slot0 * B.P * VELO_NEQ_STRIDE, 0, (size_t)B.P * VELO_NEQ_STRIDE * sizeof(double), st));
        launch_icp(L, B, ctx->dcal, ctx->d_icp_units + s_first, n_units, ctx->batch_passes, ctas,
                   ctx->d_icp_partial + (size_t)s_first * launch_icp_runs_cap(B.N) * VELO_MAX_PASSES * 64,
                   ctx->d_icp_out + (size_t)s_first * B.P * VELO_NEQ_STRIDE, B.P, nullptr, 0, nullptr, 0, ctx->search_stats);
    }
    if ((stages & VELO_STAGE_VISUAL) && pairs > 0 && ctx->batch_vis > 0) {
        const int V = ctx->prm.f2f_iterations;
        if (ctx->batch_vis != V) return fail(ctx, VELO_ERR_STATE, "batched visual stage needs n_vis_iters == f2f_iterations");
        const int n_units = pairs * V;
        if (skip_first) CK(cudaMemsetAsync(ctx->d_vis_out + (size_t)slot0 * V * VELO_NEQ_STRIDE, 0, (size_t)V * VELO_NEQ_STRIDE * sizeof(double), st));
        launch_visual(L, B, ctx->dcal, ctx->d_vis_units + (size_t)s_first * V, n_units, vis_tun(ctx), nullptr, nullptr,
                      ctx->d_vis_partial + (size_t)s_first * V * 4 * 64, ctx->d_vis_out + (size_t)s_first * V * VELO_NEQ_STRIDE, nullptr, 4,
                      VisFixed{ nullptr, nullptr, nullptr, 0, 0 }, ctx->d_flags);
    }
    CK(cudaGetLastError());
    return VELO_OK;
}

extern "C" int velo_gpu_batch_run(velo_gpu_ctx *ctx, int slot0, int count, int stages, int first_has_prev) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (check_range(ctx, slot0, count)) return VELO_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    if (stages & VELO_STAGE_VISUAL) CK(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    return run_stages(ctx, launcher(ctx), slot0, count, stages, first_has_prev);
}

extern "C" int velo_gpu_batch_download(velo_gpu_ctx *ctx, int slot0, int count, double *icp_neq, double *vis_neq, int *has_depth, int *n_hits) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (check_range(ctx, slot0, count)) return VELO_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    const DevBuffers &B = ctx->B;
    const size_t per = VELO_NUM_KP_SETS * (size_t)B.C;
    const int V = ctx->prm.f2f_iterations;
    if (icp_neq) CK(cudaMemcpyAsync(icp_neq, ctx->d_icp_out + (size_t)slot0 * B.P * VELO_NEQ_STRIDE, (size_t)count * B.P * VELO_NEQ_STRIDE * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (vis_neq) CK(cudaMemcpyAsync(vis_neq, ctx->d_vis_out + (size_t)slot0 * V * VELO_NEQ_STRIDE, (size_t)count * V * VELO_NEQ_STRIDE * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (has_depth) CK(cudaMemcpyAsync(has_depth, B.has_depth + (size_t)slot0 * per * B.F, (size_t)count * per * B.F * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (n_hits) CK(cudaMemcpyAsync(n_hits, B.n_hits + (size_t)slot0 * per, (size_t)count * per * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    int bad = 0;
    if (vis_neq) CK(cudaMemcpyAsync(&bad, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    if (bad) return fail(ctx, VELO_ERR_INVALID_ARG, "a match index is outside the keypoint set it refers to");
    return VELO_OK;
}

// keypoints_with_depth of the batch (velo.h:479 hands this cloud to the caller): [count][sets][cams][max_features] float4 {x, y, z, 1};
// the first n_hits[slot][set][cam] records of an image are valid, in keypoint order (has_depth[k] = index into them, velo.h:480)
extern "C" int velo_gpu_batch_download_kpwd(velo_gpu_ctx *ctx, int slot0, int count, float *kpwd) {
    if (!ctx || !kpwd) return VELO_ERR_INVALID_ARG;
    if (check_range(ctx, slot0, count)) return VELO_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    const DevBuffers &B = ctx->B;
    const size_t per = VELO_NUM_KP_SETS * (size_t)B.C * B.F;
    CK(cudaMemcpyAsync(kpwd, B.kpwd + (size_t)slot0 * per, (size_t)count * per * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VELO_OK;
}

// upload -> run -> download of a whole batch with the uploads of chunk c+1 overlapping the kernels of chunk c
// tuning aid (tools/): VELO_FE_UPLOAD_REPEAT=k copies every chunk k times, which puts one GPU into the copy-bound regime that eight
// GPUs sharing a host are in
static int fe_upload_repeat() { static const int r = [] { const char *e = getenv("VELO_FE_UPLOAD_REPEAT"); return e ? std::max(1, atoi(e)) : 1; }(); return r; }
static int batch_frontend_impl(velo_gpu_ctx *ctx, int slot0, int count, const velo_batch_inputs *in, int chunk,
                               double *icp_neq, double *vis_neq, int *has_depth, int *n_hits, float *kpwd);
extern "C" int velo_gpu_batch_frontend(velo_gpu_ctx *ctx, int slot0, int count, const velo_batch_inputs *in, int chunk,
                                       double *icp_neq, double *vis_neq, int *has_depth, int *n_hits) {
    return batch_frontend_impl(ctx, slot0, count, in, chunk, icp_neq, vis_neq, has_depth, n_hits, nullptr);
}
// the same call that also returns keypoints_with_depth (layout of velo_gpu_batch_download_kpwd); each chunk's cloud is copied back on
// its own stream as soon as the chunk's association stage is done, under the kernels of the later chunks
extern "C" int velo_gpu_batch_frontend_kpwd(velo_gpu_ctx *ctx, int slot0, int count, const velo_batch_inputs *in, int chunk,
                                            double *icp_neq, double *vis_neq, int *has_depth, int *n_hits, float *kpwd) {
    return batch_frontend_impl(ctx, slot0, count, in, chunk, icp_neq, vis_neq, has_depth, n_hits, kpwd);
}
static int batch_frontend_impl(velo_gpu_ctx *ctx, int slot0, int count, const velo_batch_inputs *in, int chunk,
                               double *icp_neq, double *vis_neq, int *has_depth, int *n_hits, float *kpwd) {
    if (!ctx || !in) return VELO_ERR_INVALID_ARG;
    if (check_range(ctx, slot0, count)) return VELO_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    // chunk boundaries: a fixed size if the caller gives one; otherwise a small first chunk (the only upload nothing can hide) growing
    // geometrically up to count/8.  A chunk's kernels can only start when ALL of it has arrived, so a chunk must not take longer to
    // upload than its predecessor takes to compute: the growth factor is 0.9 x (time of the previous call / time of its uploads),
    // clamped to [1.15, 2] — 2 when the copies are fast (one GPU alone on the host: 55 GB/s), ~1.2 when eight GPUs share it (23 GB/s).
    // (Measured with VELO_FE_UPLOAD_REPEAT / VELO_FE_TRACE: in the copy-bound regime the call ends 2.5 ms after the last copy, so
    // chunks that shrink again towards the end buy nothing; the copy rate under load is what limits eight GPUs on one host.)
    std::vector<int> cut(1, 0);
    if (chunk > 0) { for (int i = chunk; i < count; i += chunk) cut.push_back(i); }
    else {
        const int cap = std::max(32, count / 8);
        double n = std::max(16, count / 64);
        for (int i = (int)n; i < count; n = std::min(n * ctx->fe_growth, (double)cap), i += (int)n) cut.push_back(i);
    }
    if (chunk <= 0 && cut.size() > 1 && count - cut.back() < 16) cut.pop_back();      // no tiny last chunk
    cut.push_back(count);
    const int nchunks = (int)cut.size() - 1;
    if (!ctx->copy_stream) CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    while ((int)ctx->chunk_ev.size() < nchunks + 1) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->chunk_ev.push_back(e); }
    CK(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    if (!ctx->fe_up0) { CK(cudaEventCreate(&ctx->fe_up0)); CK(cudaEventCreate(&ctx->fe_up1)); CK(cudaEventCreate(&ctx->fe_c0)); CK(cudaEventCreate(&ctx->fe_c1)); }
    CK(cudaEventRecord(ctx->fe_c0, ctx->stream));
    // the copy stream must not overwrite slots that earlier work on the compute stream may still read
    CK(cudaEventRecord(ctx->chunk_ev[nchunks], ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ev[nchunks], 0));
    CK(cudaEventRecord(ctx->fe_up0, ctx->copy_stream));
    // chunks alternate between two compute streams, so the tail of one chunk's correspondence launch overlaps the next chunk's
    // kernels; a chunk's frame pairs read the previous chunk's last scan, hence the wait on its index / projection stages
    // (per-kernel profiling needs serial launches: one stream then)
    if (!ctx->stream2) CK(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    const bool two = !ctx->profile && nchunks > 1;
    while ((int)ctx->chunk_ev.size() < 2 * nchunks + 2) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->chunk_ev.push_back(e); }
    cudaEvent_t *ev_light = ctx->chunk_ev.data() + nchunks + 1;
    if (two) CK(cudaStreamWaitEvent(ctx->stream2, ctx->chunk_ev[nchunks], 0));
    const int light = VELO_STAGE_INGEST | VELO_STAGE_INDEX | VELO_STAGE_PROJECT | VELO_STAGE_ASSOC;
    static const bool trace = getenv("VELO_FE_TRACE") != nullptr;       // tuning aid: per-chunk completion times on stderr
    std::vector<cudaEvent_t> tr;
    if (trace) { tr.resize(3 * (size_t)nchunks); for (auto &e : tr) cudaEventCreate(&e); }
    for (int c = 0; c < nchunks; c++) {
        cudaStream_t st = (two && (c & 1)) ? ctx->stream2 : ctx->stream;
        const Launcher L = launcher_on(ctx, st);
        // host-side preparation of the chunk (pose packs: trigonometry + forward-mode dR per pass) and its copies are issued right
        // before its kernels, so preparing chunk c+1 overlaps the device work of chunk c
        int rc = upload_range(ctx, slot0, cut[c], cut[c + 1] - cut[c], in, ctx->copy_stream, st);
        if (rc) return rc;
        for (int rep = 1; rep < fe_upload_repeat(); rep++) upload_range(ctx, slot0, cut[c], cut[c + 1] - cut[c], in, ctx->copy_stream, st);
        CK(cudaEventRecord(ctx->chunk_ev[c], ctx->copy_stream));
        if (trace) cudaEventRecord(tr[3 * c], ctx->copy_stream);
        CK(cudaStreamWaitEvent(st, ctx->chunk_ev[c], 0));
        rc = run_stages(ctx, L, slot0 + cut[c], cut[c + 1] - cut[c], light, 1);
        if (rc) return rc;
        CK(cudaEventRecord(ev_light[c], st));
        if (kpwd) {
            if (!ctx->d2h_stream) CK(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
            const size_t per = VELO_NUM_KP_SETS * (size_t)ctx->B.C * ctx->B.F;
            CK(cudaStreamWaitEvent(ctx->d2h_stream, ev_light[c], 0));
            CK(cudaMemcpyAsync(kpwd + (size_t)cut[c] * per * 4, ctx->B.kpwd + (size_t)(slot0 + cut[c]) * per, (size_t)(cut[c + 1] - cut[c]) * per * sizeof(float4),
                               cudaMemcpyDeviceToHost, ctx->d2h_stream));
        }
        if (trace) cudaEventRecord(tr[3 * c + 1], st);
        if (c > 0) CK(cudaStreamWaitEvent(st, ev_light[c - 1], 0));
        rc = run_stages(ctx, L, slot0 + cut[c], cut[c + 1] - cut[c], VELO_STAGE_ICP | VELO_STAGE_VISUAL, 1);
        if (rc) return rc;
        if (trace) cudaEventRecord(tr[3 * c + 2], st);
    }
    CK(cudaEventRecord(ctx->fe_up1, ctx->copy_stream));
    if (two) { CK(cudaEventRecord(ev_light[nchunks], ctx->stream2)); CK(cudaStreamWaitEvent(ctx->stream, ev_light[nchunks], 0)); }
    ctx->launch_stream = ctx->stream;
    const int rc = velo_gpu_batch_download(ctx, slot0, count, icp_neq, vis_neq, has_depth, n_hits);
    if (kpwd) CK(cudaStreamSynchronize(ctx->d2h_stream));
    if (trace) {
        cudaDeviceSynchronize();
        fprintf(stderr, "[fe] growth %.2f, %d chunks: size copy_done light_done icp_done (ms)\n", ctx->fe_growth, nchunks);
        for (int c = 0; c < nchunks; c++) {
            float a = 0, b = 0, d = 0;
            cudaEventElapsedTime(&a, ctx->fe_c0, tr[3 * c]); cudaEventElapsedTime(&b, ctx->fe_c0, tr[3 * c + 1]); cudaEventElapsedTime(&d, ctx->fe_c0, tr[3 * c + 2]);
            fprintf(stderr, "[fe] %4d %8.2f %8.2f %8.2f\n", cut[c + 1] - cut[c], a, b, d);
        }
        for (auto &e : tr) cudaEventDestroy(e);
    }
    if (rc == VELO_OK && chunk <= 0) {        // how fast the copies were against the whole call: the next call's chunk growth
        CK(cudaEventRecord(ctx->fe_c1, ctx->stream)); CK(cudaEventSynchronize(ctx->fe_c1)); CK(cudaEventSynchronize(ctx->fe_up1));
        float up = 0.f, all = 0.f;
        CK(cudaEventElapsedTime(&up, ctx->fe_up0, ctx->fe_up1)); CK(cudaEventElapsedTime(&all, ctx->fe_c0, ctx->fe_c1));
        if (up > 0.f && all > 0.f) ctx->fe_growth = std::min(2.0f, std::max(1.15f, 0.9f * all / up));
    }
    return rc;
}

extern "C" int velo_gpu_batch_counts(velo_gpu_ctx *ctx, int slot0, int count, int *n_points, int *n_rings, int *proj_total, int *status) {
    if (!ctx) return VELO_ERR_INVALID_ARG;
    if (check_range(ctx, slot0, count)) return VELO_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    const DevBuffers &B = ctx->B;
    std::vector<int> nr(count), pc((size_t)count * B.C * B.R);
    CK(cudaMemcpyAsync(nr.data(), B.n_rings + slot0, count * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(pc.data(), B.proj_count + (size_t)slot0 * B.C * B.R, pc.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (n_points) CK(cudaMemcpyAsync(n_points, B.n_points + slot0, count * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (status) CK(cudaMemcpyAsync(status, B.status + slot0, count * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < count; i++) {
        if (n_rings) n_rings[i] = nr[i];
        if (proj_total) for (int c = 0; c < B.C; c++) {
            int t = 0;
            for (int s = 0; s < nr[i]; s++) t += pc[((size_t)i * B.C + c) * B.R + s];
            proj_total[(size_t)i * B.C + c] = t;
        }
    }
    return VELO_OK;
}
