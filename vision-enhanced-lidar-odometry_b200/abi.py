"""ctypes mirror of include/velo_gpu.h (POD structs and constants)."""
import ctypes as C

ABI_VERSION = 2
MAX_CAMS = 4
NUM_KP_SETS = 2
NEQ = 28
NEQ_STRIDE = 64
NUM_KERNELS = 12

RES_3D3D, RES_3D2D, RES_2D3D, RES_2D2D, RES_3DPD = 0, 1, 2, 3, 4
STAGE_INGEST, STAGE_INDEX, STAGE_PROJECT, STAGE_ASSOC, STAGE_ICP, STAGE_VISUAL, STAGE_ALL = 1, 2, 4, 8, 16, 32, 63

STATUS = {0: "VELO_OK", 1: "VELO_ERR_NO_DEVICE", 2: "VELO_ERR_CUDA", 3: "VELO_ERR_INVALID_ARG",
          4: "VELO_ERR_CAPACITY", 5: "VELO_ERR_STATE"}


class Params(C.Structure):
    _fields_ = [
        ("num_cams", C.c_int), ("icp_skip", C.c_int), ("f2f_iterations", C.c_int), ("icp_iterations", C.c_int),
        ("enable_2d2d", C.c_int), ("enable_3d2d", C.c_int), ("abs_truncates", C.c_int), ("reserved0", C.c_int),
        ("weight_3D2D", C.c_double), ("weight_2D2D", C.c_double), ("weight_3DPD", C.c_double),
        ("loss_thresh_3D2D", C.c_double), ("loss_thresh_2D2D", C.c_double), ("loss_thresh_3DPD", C.c_double),
        ("loss_thresh_3D3D", C.c_double), ("depth_assoc_thresh", C.c_double), ("outlier_reject", C.c_double),
        ("correspondence_thresh_icp", C.c_double), ("icp_norm_condition", C.c_double),
        ("max_slots", C.c_int), ("max_points", C.c_int), ("max_rings", C.c_int), ("max_features", C.c_int),
        ("max_matches", C.c_int), ("max_icp_passes", C.c_int), ("ctas_per_icp_unit", C.c_int), ("reserved1", C.c_int),
    ]


class Calib(C.Structure):
    _fields_ = [
        ("velo_to_cam", C.c_float * 16),
        ("cam_trans", (C.c_float * 4) * MAX_CAMS),
        ("cam_K", (C.c_float * 9) * MAX_CAMS),
        ("cam_Kinv", (C.c_float * 9) * MAX_CAMS),
        ("min_x", C.c_double * MAX_CAMS), ("max_x", C.c_double * MAX_CAMS),
        ("min_y", C.c_double * MAX_CAMS), ("max_y", C.c_double * MAX_CAMS),
        ("img_width", C.c_int), ("img_height", C.c_int),
    ]


class IcpCorr(C.Structure):
    _fields_ = [
        ("src_ring", C.c_int32), ("src_idx", C.c_int32), ("np_s_i", C.c_int32), ("np_i", C.c_int32),
        ("np_s_j", C.c_int32), ("np_j", C.c_int32), ("np_k", C.c_int32), ("kept", C.c_int32),
        ("normal", C.c_float * 3), ("v0", C.c_float * 3), ("residual", C.c_double), ("jacobian", C.c_double * 6),
    ]


class VisBlock(C.Structure):
    _fields_ = [
        ("cam", C.c_int32), ("match", C.c_int32), ("type", C.c_int32), ("n_res", C.c_int32),
        ("residual", C.c_double * 3), ("jacobian", C.c_double * 18),
    ]


MAX_SOLVES = 16


class F2FReport(C.Structure):
    _fields_ = [("n_solves", C.c_int), ("lm_iterations", C.c_int * MAX_SOLVES), ("accepted_steps", C.c_int * MAX_SOLVES),
                ("reason", C.c_int * MAX_SOLVES), ("n_blocks", C.c_int * MAX_SOLVES),
                ("initial_cost", C.c_double * MAX_SOLVES), ("final_cost", C.c_double * MAX_SOLVES), ("pose", (C.c_double * 6) * MAX_SOLVES)]


class BatchInputs(C.Structure):
    _fields_ = [
        ("scans", C.c_void_p), ("n_points", C.c_void_p),
        ("kp", C.c_void_p), ("n_kp", C.c_void_p),
        ("matches", C.c_void_p), ("n_matches", C.c_void_p),
        ("icp_poses", C.c_void_p), ("pass_iter", C.c_void_p), ("n_passes", C.c_int),
        ("vis_poses", C.c_void_p), ("n_vis_iters", C.c_int),
        ("scan_stride_floats", C.c_int),
    ]


import numpy as np

TRI_OBS3_DTYPE = np.dtype([("frame", np.int32), ("x", np.float32), ("y", np.float32), ("z", np.float32)])
TRI_OBS2_DTYPE = np.dtype([("frame", np.int32), ("cam", np.int32), ("x", np.float32), ("y", np.float32)])
ICP_CORR_DTYPE = np.dtype([
    ("src_ring", np.int32), ("src_idx", np.int32), ("np_s_i", np.int32), ("np_i", np.int32),
    ("np_s_j", np.int32), ("np_j", np.int32), ("np_k", np.int32), ("kept", np.int32),
    ("normal", np.float32, 3), ("v0", np.float32, 3), ("residual", np.float64), ("jacobian", np.float64, 6)], align=True)
VIS_BLOCK_DTYPE = np.dtype([
    ("cam", np.int32), ("match", np.int32), ("type", np.int32), ("n_res", np.int32),
    ("residual", np.float64, 3), ("jacobian", np.float64, 18)], align=True)
assert ICP_CORR_DTYPE.itemsize == C.sizeof(IcpCorr), (ICP_CORR_DTYPE.itemsize, C.sizeof(IcpCorr))
assert VIS_BLOCK_DTYPE.itemsize == C.sizeof(VisBlock)

# functions declared in include/velo_gpu.h; tests check that the library exports every one
EXPORTS = [
    "velo_gpu_abi_version", "velo_gpu_default_params", "velo_gpu_calib_from_kitti", "velo_pixel2canonical",
    "velo_canonical2pixel", "velo_kitti_load_calib", "velo_kitti_load_scan", "velo_kitti_format_pose", "velo_gpu_create", "velo_gpu_destroy", "velo_gpu_last_error", "velo_gpu_sync",
    "velo_gpu_device_name", "velo_gpu_host_alloc", "velo_gpu_host_free", "velo_gpu_timer_begin", "velo_gpu_timer_end",
    "velo_gpu_profile_enable", "velo_gpu_search_stats_enable", "velo_gpu_profile_reset", "velo_gpu_profile_read", "velo_gpu_kernel_name",
    "velo_gpu_scan_upload", "velo_gpu_scan_upload_rings", "velo_gpu_projection_upload", "velo_gpu_scan_info", "velo_gpu_scan_download", "velo_gpu_project",
    "velo_gpu_project_download", "velo_gpu_depth_assoc", "velo_gpu_assoc_upload", "velo_gpu_f2f_selection", "velo_pose_vec2mat", "velo_gpu_icp_pass", "velo_gpu_icp_passes", "velo_gpu_visual_residuals", "velo_gpu_frame_to_frame", "velo_gpu_batch_frame_to_frame", "velo_gpu_match_hamming", "velo_gpu_triangulate",
    "velo_gpu_batch_upload", "velo_gpu_batch_run", "velo_gpu_batch_download", "velo_gpu_batch_download_kpwd", "velo_gpu_batch_frontend", "velo_gpu_batch_frontend_kpwd", "velo_gpu_launch_count",
    "velo_gpu_batch_counts",
]
