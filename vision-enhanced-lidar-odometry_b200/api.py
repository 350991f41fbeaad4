"""Thin ctypes binding of libvelo_gpu.so (include/velo_gpu.h) — the harness the tests and bench.py use.

This module adds nothing to the product: every method is one C-ABI call on host numpy buffers.  If the CUDA
library cannot be built/loaded, or no B200 is present, it raises — there is no CPU path.
"""
import ctypes as C
import os

import numpy as np

from . import _build, abi

_P = C.c_void_p
_lib = None


class VeloError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{abi.STATUS.get(code, code)}: {msg}")
        self.code = code


def lib():
    """Load libvelo_gpu.so (building it with nvcc when sources are newer). Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("VELO_GPU_LIB") or _build.build_gpu()      # VELO_GPU_LIB: a prebuilt tuning variant (tools/build_variants.py)
    L = C.CDLL(path)
    L.velo_gpu_last_error.restype = C.c_char_p
    L.velo_gpu_last_error.argtypes = [_P]
    L.velo_gpu_kernel_name.restype = C.c_char_p
    L.velo_gpu_kernel_name.argtypes = [C.c_int]
    L.velo_gpu_default_params.argtypes = [_P]
    L.velo_gpu_calib_from_kitti.argtypes = [_P, _P, C.c_int, C.c_int, _P]
    L.velo_pixel2canonical.argtypes = [_P, C.c_int, _P, C.c_int, _P]
    L.velo_canonical2pixel.argtypes = [_P, C.c_int, _P, C.c_int, _P]
    L.velo_kitti_load_calib.argtypes = [C.c_char_p, _P, _P]
    L.velo_kitti_load_scan.argtypes = [C.c_char_p, _P, C.c_int, C.POINTER(C.c_int)]
    L.velo_kitti_format_pose.argtypes = [_P, C.c_char_p, C.c_int]
    L.velo_gpu_create.argtypes = [C.c_int, _P, _P, C.POINTER(_P)]
    for fn in ("velo_gpu_destroy", "velo_gpu_sync", "velo_gpu_timer_begin", "velo_gpu_profile_reset"):
        getattr(L, fn).argtypes = [_P]
    L.velo_gpu_device_name.argtypes = [_P, _P, C.c_int]
    L.velo_gpu_host_alloc.argtypes = [C.POINTER(_P), C.c_uint64]
    L.velo_gpu_host_free.argtypes = [_P]
    L.velo_gpu_timer_end.argtypes = [_P, C.POINTER(C.c_float)]
    L.velo_gpu_profile_enable.argtypes = [_P, C.c_int]
    L.velo_gpu_batch_download_kpwd.argtypes = [_P, C.c_int, C.c_int, _P]
    L.velo_gpu_search_stats_enable.argtypes = [_P, C.c_int]
    L.velo_gpu_profile_read.argtypes = [_P, _P, _P]
    L.velo_gpu_scan_upload.argtypes = [_P, C.c_int, _P, C.c_int]
    L.velo_gpu_scan_upload_rings.argtypes = [_P, C.c_int, _P, _P, C.c_int]
    L.velo_gpu_projection_upload.argtypes = [_P, C.c_int, C.c_int, _P, _P, _P]
    L.velo_gpu_scan_info.argtypes = [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.velo_gpu_scan_download.argtypes = [_P, C.c_int, _P, _P]
    L.velo_gpu_project.argtypes = [_P, C.c_int, C.c_int]
    L.velo_gpu_project_download.argtypes = [_P, C.c_int, C.c_int, _P, _P, _P, C.POINTER(C.c_int)]
    L.velo_gpu_depth_assoc.argtypes = [_P, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, _P, C.POINTER(C.c_int)]
    L.velo_gpu_assoc_upload.argtypes = [_P, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, _P, C.c_int]
    L.velo_gpu_f2f_selection.argtypes = [_P, _P, C.c_int]
    L.velo_pose_vec2mat.argtypes = [_P, _P]
    L.velo_gpu_icp_pass.argtypes = [_P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), _P]
    L.velo_gpu_icp_passes.argtypes = [_P, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, _P, C.c_int, C.POINTER(C.c_int), _P]
    L.velo_gpu_visual_residuals.argtypes = [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, C.c_int, _P, C.c_int, C.POINTER(C.c_int), _P]
    L.velo_gpu_frame_to_frame.argtypes = [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P]
    L.velo_gpu_batch_frame_to_frame.argtypes = [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]
    L.velo_gpu_match_hamming.argtypes = [_P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_double, _P, C.POINTER(C.c_int), _P, _P]
    L.velo_gpu_triangulate.argtypes = [_P, C.c_int, _P, _P, _P, _P, _P, C.c_int, _P, _P, _P, _P]
    L.velo_gpu_batch_upload.argtypes = [_P, C.c_int, C.c_int, _P]
    L.velo_gpu_batch_run.argtypes = [_P, C.c_int, C.c_int, C.c_int, C.c_int]
    L.velo_gpu_batch_download.argtypes = [_P, C.c_int, C.c_int, _P, _P, _P, _P]
    L.velo_gpu_batch_frontend.argtypes = [_P, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P, _P]
    L.velo_gpu_batch_frontend_kpwd.argtypes = [_P, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P, _P, _P]
    L.velo_gpu_batch_counts.argtypes = [_P, C.c_int, C.c_int, _P, _P, _P, _P]
    L.velo_gpu_launch_count.argtypes = [_P, C.POINTER(C.c_int64)]
    if L.velo_gpu_abi_version() != abi.ABI_VERSION:
        raise RuntimeError("libvelo_gpu.so ABI version mismatch")
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data


def default_params(**kw):
    p = abi.Params()
    lib().velo_gpu_default_params(C.addressof(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def calib_from_kitti(P, Tr, w, h):
    """loadCalibration (kitti.h:59-108)."""
    cal = abi.Calib()
    P = np.ascontiguousarray(P, np.float32)
    Tr = np.ascontiguousarray(Tr, np.float32)
    rc = lib().velo_gpu_calib_from_kitti(_ptr(P), _ptr(Tr), w, h, C.addressof(cal))
    if rc:
        raise VeloError(rc, "calib_from_kitti")
    return cal


def pixel2canonical(cal, cam, pix):
    pix = np.ascontiguousarray(pix, np.float32)
    out = np.zeros_like(pix)
    lib().velo_pixel2canonical(C.addressof(cal), cam, _ptr(pix), len(pix), _ptr(out))
    return out


def canonical2pixel(cal, cam, can):
    can = np.ascontiguousarray(can, np.float32)
    out = np.zeros_like(can)
    lib().velo_canonical2pixel(C.addressof(cal), cam, _ptr(can), len(can), _ptr(out))
    return out


def kitti_load_calib(path):
    """calib.txt (kitti.h:66-105) -> (P[48], Tr[12])"""
    P = np.zeros(48, np.float32); Tr = np.zeros(12, np.float32)
    rc = lib().velo_kitti_load_calib(str(path).encode(), _ptr(P), _ptr(Tr))
    if rc:
        raise VeloError(rc, f"cannot parse {path}")
    return P, Tr


def kitti_load_scan(path, max_points=200000):
    """velodyne .bin (kitti.h:121-152) -> float32 [n, 4]"""
    buf = np.zeros((max_points, 4), np.float32)
    n = C.c_int()
    rc = lib().velo_kitti_load_scan(str(path).encode(), _ptr(buf), max_points, C.byref(n))
    if rc:
        raise VeloError(rc, f"cannot read {path}")
    return buf[:n.value]


def kitti_format_pose(T):
    """one line of results/<seq>.txt (kitti.h:202-216) for a 4x4 pose"""
    T = np.ascontiguousarray(T, np.float64).reshape(16)
    buf = C.create_string_buffer(512)
    rc = lib().velo_kitti_format_pose(_ptr(T), buf, 512)
    if rc:
        raise VeloError(rc, "format_pose")
    return buf.value.decode()


def pose_vec2mat(transform):
    """util::pose_mat2vec (utility.h:67-82): 6-vector -> 4x4"""
    t = np.ascontiguousarray(transform, np.float64)
    T = np.zeros(16, np.float64)
    lib().velo_pose_vec2mat(_ptr(t), _ptr(T))
    return T.reshape(4, 4)


class PinnedPool:
    """numpy views over cudaHostAlloc'ed memory (velo_gpu_host_alloc)."""

    def __init__(self):
        self._ptrs = []

    def zeros(self, shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = _P()
        rc = lib().velo_gpu_host_alloc(C.byref(p), max(n, 1))
        if rc:
            raise VeloError(rc, "host_alloc")
        self._ptrs.append(p)
        buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        a = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        a[...] = 0
        return a

    def close(self):
        for p in self._ptrs:
            lib().velo_gpu_host_free(p)
        self._ptrs = []


class Context:
    """One GPU context (velo_gpu_create)."""

    def __init__(self, prm, cal, device=0):
        self.L = lib()
        self.prm, self.cal = prm, cal
        self.h = _P()
        rc = self.L.velo_gpu_create(device, C.addressof(prm), C.addressof(cal), C.byref(self.h))
        if rc:
            raise VeloError(rc, (self.L.velo_gpu_last_error(None) or b"").decode())

    def _ck(self, rc):
        if rc:
            raise VeloError(rc, (self.L.velo_gpu_last_error(self.h) or b"").decode())

    def close(self):
        if self.h:
            self.L.velo_gpu_destroy(self.h)
            self.h = _P()

    def sync(self):
        self._ck(self.L.velo_gpu_sync(self.h))

    def device_name(self):
        buf = C.create_string_buffer(256)
        self._ck(self.L.velo_gpu_device_name(self.h, buf, 256))
        return buf.value.decode()

    # ---- timing
    def timer_begin(self):
        self._ck(self.L.velo_gpu_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_float()
        self._ck(self.L.velo_gpu_timer_end(self.h, C.byref(ms)))
        return ms.value

    def profile(self, on):
        self._ck(self.L.velo_gpu_profile_enable(self.h, int(on)))

    def search_stats(self, on):
        self._ck(self.L.velo_gpu_search_stats_enable(self.h, int(on)))

    def profile_reset(self):
        self._ck(self.L.velo_gpu_profile_reset(self.h))

    def profile_read(self):
        ms = np.zeros(abi.NUM_KERNELS, np.float32)
        n = np.zeros(abi.NUM_KERNELS, np.int32)
        self._ck(self.L.velo_gpu_profile_read(self.h, _ptr(ms), _ptr(n)))
        return {self.L.velo_gpu_kernel_name(k).decode(): (float(ms[k]), int(n[k])) for k in range(abi.NUM_KERNELS) if n[k]}

    def launch_count(self):
        n = C.c_int64()
        self._ck(self.L.velo_gpu_launch_count(self.h, C.byref(n)))
        return n.value

    # ---- single-frame path
    def scan_upload(self, slot, xyzr):
        xyzr = np.ascontiguousarray(xyzr, np.float32).reshape(-1, 4)
        self._ck(self.L.velo_gpu_scan_upload(self.h, slot, _ptr(xyzr), len(xyzr)))

    def scan_upload_rings(self, slot, xyz1, ring_start):
        xyz1 = np.ascontiguousarray(xyz1, np.float32).reshape(-1, 4)
        ring_start = np.ascontiguousarray(ring_start, np.int32)
        self._ck(self.L.velo_gpu_scan_upload_rings(self.h, slot, _ptr(xyz1), _ptr(ring_start), len(ring_start) - 1))

    def projection_upload(self, slot, cam, ring_count, proj, valid):
        ring_count = np.ascontiguousarray(ring_count, np.int32)
        proj = np.ascontiguousarray(proj, np.float32); valid = np.ascontiguousarray(valid, np.float32)
        self._ck(self.L.velo_gpu_projection_upload(self.h, slot, cam, _ptr(ring_count), _ptr(proj), _ptr(valid)))

    def scan_info(self, slot):
        a, b = C.c_int(), C.c_int()
        self._ck(self.L.velo_gpu_scan_info(self.h, slot, C.byref(a), C.byref(b)))
        return a.value, b.value

    def scan_download(self, slot):
        n, nr = self.scan_info(slot)
        pts = np.zeros((max(n, 1), 4), np.float32)
        rs = np.zeros(nr + 1, np.int32)
        self._ck(self.L.velo_gpu_scan_download(self.h, slot, _ptr(pts), _ptr(rs)))
        return pts[:n], rs

    def project(self, slot, cam):
        self._ck(self.L.velo_gpu_project(self.h, slot, cam))

    def project_download(self, slot, cam):
        n, nr = self.scan_info(slot)
        rc = np.zeros(max(nr, 1), np.int32)
        proj = np.zeros((max(n, 1), 2), np.float32)
        valid = np.zeros((max(n, 1), 4), np.float32)
        tot = C.c_int()
        self._ck(self.L.velo_gpu_project_download(self.h, slot, cam, _ptr(rc), _ptr(proj), _ptr(valid), C.byref(tot)))
        return rc[:nr], proj[:tot.value], valid[:tot.value]

    def depth_assoc(self, slot, cam, kp, set_=0):
        kp = np.ascontiguousarray(kp, np.float32).reshape(-1, 2)
        F = len(kp)
        hd = np.zeros(max(F, 1), np.int32)
        kpwd = np.zeros((max(F, 1), 4), np.float32)
        nh = C.c_int()
        self._ck(self.L.velo_gpu_depth_assoc(self.h, slot, cam, set_, _ptr(kp), F, _ptr(hd), _ptr(kpwd), C.byref(nh)))
        return hd[:F], kpwd[:nh.value]

    def assoc_upload(self, slot, cam, set_, kp, has_depth, kpwd):
        kp = np.ascontiguousarray(kp, np.float32).reshape(-1, 2); hd = np.ascontiguousarray(has_depth, np.int32)
        kw = np.ascontiguousarray(kpwd, np.float32).reshape(-1, 4)
        self._ck(self.L.velo_gpu_assoc_upload(self.h, slot, cam, set_, _ptr(kp), len(kp), _ptr(hd), _ptr(kw) if len(kw) else None, len(kw)))

    def f2f_selection(self):
        sel = np.zeros(self.prm.num_cams * self.prm.max_matches, np.uint8)
        self._ck(self.L.velo_gpu_f2f_selection(self.h, _ptr(sel), len(sel)))
        return sel.reshape(self.prm.num_cams, self.prm.max_matches)

    def icp_pass(self, slot_M, slot_S, pose, it, skip, want_corr=True):
        pose = np.ascontiguousarray(pose, np.float64)
        cap = self.prm.max_points
        corr = np.zeros(cap, abi.ICP_CORR_DTYPE) if want_corr else None
        nq, nk = C.c_int(), C.c_int()
        neq = np.zeros(abi.NEQ_STRIDE, np.float64)
        self._ck(self.L.velo_gpu_icp_pass(self.h, slot_M, slot_S, _ptr(pose), it, skip, _ptr(corr), cap, C.byref(nq), C.byref(nk), _ptr(neq)))
        return (corr[:nq.value] if want_corr else None), neq, nk.value

    def icp_passes(self, slot_M, slot_S, poses, iters, skip, want_corr=True):
        """all ICP passes of a frame pair in one launch of the fused kernel; returns (corr[n_passes][nq] | None, neq[n_passes][64])"""
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 6)
        iters = np.ascontiguousarray(iters, np.int32)
        n = len(iters)
        cap = self.prm.max_points
        corr = np.zeros((n, cap), abi.ICP_CORR_DTYPE) if want_corr else None
        nq = C.c_int()
        neq = np.zeros((n, abi.NEQ_STRIDE), np.float64)
        self._ck(self.L.velo_gpu_icp_passes(self.h, slot_M, slot_S, _ptr(poses), _ptr(iters), n, skip, _ptr(corr), cap, C.byref(nq), _ptr(neq)))
        return (corr[:, :nq.value] if want_corr else None), neq

    def visual_residuals(self, slot1, set1, slot2, set2, n_matches, matches, pose, it, lm_valid=None, lm_xyz=None):
        """matches: concatenated per camera, [sum(n_matches)][2]"""
        n_matches = np.ascontiguousarray(n_matches, np.int32)
        matches = np.ascontiguousarray(matches, np.int32).reshape(-1, 2)
        pose = np.ascontiguousarray(pose, np.float64)
        lm_valid = None if lm_valid is None else np.ascontiguousarray(lm_valid, np.int32)
        lm_xyz = None if lm_xyz is None else np.ascontiguousarray(lm_xyz, np.float32)
        cap = 3 * int(n_matches.sum()) + 1
        blocks = np.zeros(cap, abi.VIS_BLOCK_DTYPE)
        nb = C.c_int()
        neq = np.zeros(abi.NEQ_STRIDE, np.float64)
        self._ck(self.L.velo_gpu_visual_residuals(self.h, slot1, set1, slot2, set2, _ptr(n_matches), _ptr(matches), _ptr(lm_valid), _ptr(lm_xyz),
                                                  _ptr(pose), it, _ptr(blocks), cap, C.byref(nb), _ptr(neq)))
        return blocks[:nb.value], neq

    def frame_to_frame(self, slot_M, set1, slot_S, set2, transform, n_matches=None, matches=None, enable_icp=1, icp_skip=None, lm_valid=None, lm_xyz=None):
        """velo.h:598-919 with the device-resident solve; returns (transform[6], report dict)"""
        t = np.ascontiguousarray(transform, np.float64).copy()
        nm = None if n_matches is None else np.ascontiguousarray(n_matches, np.int32)
        mt = None if matches is None else np.ascontiguousarray(matches, np.int32).reshape(-1, 2)
        lv = None if lm_valid is None else np.ascontiguousarray(lm_valid, np.int32)
        lx = None if lm_xyz is None else np.ascontiguousarray(lm_xyz, np.float32)
        rep = abi.F2FReport()
        self._ck(self.L.velo_gpu_frame_to_frame(self.h, slot_M, set1, slot_S, set2, _ptr(nm), _ptr(mt), _ptr(lv), _ptr(lx), enable_icp,
                                                self.prm.icp_skip if icp_skip is None else icp_skip, _ptr(t), C.addressof(rep)))
        n = rep.n_solves
        return t, {"n_solves": n, "lm_iterations": list(rep.lm_iterations)[:n], "accepted_steps": list(rep.accepted_steps)[:n], "reason": list(rep.reason)[:n],
                   "n_blocks": list(rep.n_blocks)[:n], "initial_cost": list(rep.initial_cost)[:n], "final_cost": list(rep.final_cost)[:n],
                   "pose": np.array([list(rep.pose[i]) for i in range(n)])}

    def match_hamming(self, query, train, match_thresh=29.0):
        """matchFeatures (velo.h:499-550): uint8 [n, desc_bytes] descriptors -> (pairs [m, 2], best_idx, best_dist)"""
        query = np.ascontiguousarray(query, np.uint8); train = np.ascontiguousarray(train, np.uint8)
        nq, nt = len(query), len(train)
        db = query.shape[1] if nq else (train.shape[1] if nt else 64)
        pairs = np.zeros((max(nq, 1), 2), np.int32); bi = np.zeros(max(nq, 1), np.int32); bd = np.zeros(max(nq, 1), np.int32)
        n = C.c_int()
        self._ck(self.L.velo_gpu_match_hamming(self.h, _ptr(query) if nq else None, nq, _ptr(train) if nt else None, nt, db, match_thresh,
                                               _ptr(pairs), C.byref(n), _ptr(bi), _ptr(bd)))
        return pairs[:n.value], bi[:nq], bd[:nq]

    def triangulate(self, off3, obs3, off2, obs2, poses, init_xyz=None, has_init=None):
        """triangulatePoint (velo.h:1027-1130) for len(off3)-1 landmarks; obs3/obs2 structured arrays (abi.TRI_OBS*_DTYPE)"""
        off3 = np.ascontiguousarray(off3, np.int32); off2 = np.ascontiguousarray(off2, np.int32)
        obs3 = np.ascontiguousarray(obs3, abi.TRI_OBS3_DTYPE); obs2 = np.ascontiguousarray(obs2, abi.TRI_OBS2_DTYPE)
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 6)
        n = len(off3) - 1
        out = np.zeros((max(n, 1), 3), np.float32); it = np.zeros(max(n, 1), np.int32)
        ini = None if init_xyz is None else np.ascontiguousarray(init_xyz, np.float32)
        has = None if has_init is None else np.ascontiguousarray(has_init, np.int32)
        self._ck(self.L.velo_gpu_triangulate(self.h, n, _ptr(off3), _ptr(obs3) if len(obs3) else None, _ptr(off2), _ptr(obs2) if len(obs2) else None,
                                             _ptr(poses) if len(poses) else None, len(poses), _ptr(ini), _ptr(has), _ptr(out), _ptr(it)))
        return out[:n], it[:n]

    # ---- batched path
    def batch_upload(self, slot0, batch):
        bi = abi.BatchInputs(_ptr(batch.scans), _ptr(batch.n_points), _ptr(batch.kp), _ptr(batch.n_kp), _ptr(batch.matches), _ptr(batch.n_matches),
                             _ptr(batch.icp_poses), _ptr(batch.pass_iter), batch.n_passes, _ptr(batch.vis_poses), batch.n_vis, batch.scans.shape[-1])
        self._ck(self.L.velo_gpu_batch_upload(self.h, slot0, batch.count, C.addressof(bi)))

    def batch_run(self, slot0, count, stages=abi.STAGE_ALL, first_has_prev=0):
        self._ck(self.L.velo_gpu_batch_run(self.h, slot0, count, stages, first_has_prev))

    def batch_download(self, slot0, count, icp_neq=None, vis_neq=None, has_depth=None, n_hits=None):
        self._ck(self.L.velo_gpu_batch_download(self.h, slot0, count, _ptr(icp_neq), _ptr(vis_neq), _ptr(has_depth), _ptr(n_hits)))

    def batch_download_kpwd(self, slot0, count, out=None):
        """[count][sets][cams][max_features][4] float32; out may be a (pinned) array of that shape"""
        if out is None:
            out = np.zeros((count, abi.NUM_KP_SETS, self.prm.num_cams, self.prm.max_features, 4), np.float32)
        self._ck(self.L.velo_gpu_batch_download_kpwd(self.h, slot0, count, _ptr(out)))
        return out

    def batch_frontend(self, slot0, batch, chunk=0, icp_neq=None, vis_neq=None, has_depth=None, n_hits=None, kpwd=None):
        bi = abi.BatchInputs(_ptr(batch.scans), _ptr(batch.n_points), _ptr(batch.kp), _ptr(batch.n_kp), _ptr(batch.matches), _ptr(batch.n_matches),
                             _ptr(batch.icp_poses), _ptr(batch.pass_iter), batch.n_passes, _ptr(batch.vis_poses), batch.n_vis, batch.scans.shape[-1])
        if kpwd is None:
            self._ck(self.L.velo_gpu_batch_frontend(self.h, slot0, batch.count, C.addressof(bi), chunk, _ptr(icp_neq), _ptr(vis_neq), _ptr(has_depth), _ptr(n_hits)))
        else:
            self._ck(self.L.velo_gpu_batch_frontend_kpwd(self.h, slot0, batch.count, C.addressof(bi), chunk, _ptr(icp_neq), _ptr(vis_neq), _ptr(has_depth),
                                                         _ptr(n_hits), _ptr(kpwd)))

    def batch_frame_to_frame(self, slot0, count, transforms, enable_visual=1, enable_icp=1, first_has_prev=0):
        """velo.h:616-907 for every frame pair of the batch at once; transforms [count][6] -> (transforms, [report dict per slot])"""
        t = np.ascontiguousarray(transforms, np.float64).reshape(count, 6).copy()
        reps = (abi.F2FReport * count)()
        self._ck(self.L.velo_gpu_batch_frame_to_frame(self.h, slot0, count, first_has_prev, enable_visual, enable_icp, _ptr(t), C.addressof(reps)))
        out = []
        for rep in reps:
            n = rep.n_solves
            out.append({"n_solves": n, "lm_iterations": list(rep.lm_iterations)[:n], "accepted_steps": list(rep.accepted_steps)[:n], "reason": list(rep.reason)[:n],
                        "n_blocks": list(rep.n_blocks)[:n], "initial_cost": list(rep.initial_cost)[:n], "final_cost": list(rep.final_cost)[:n],
                        "pose": np.array([list(rep.pose[i]) for i in range(n)])})
        return t, out

    def batch_counts(self, slot0, count):
        npnt = np.zeros(count, np.int32); nr = np.zeros(count, np.int32)
        pt = np.zeros((count, self.prm.num_cams), np.int32); st = np.zeros(count, np.int32)
        self._ck(self.L.velo_gpu_batch_counts(self.h, slot0, count, _ptr(npnt), _ptr(nr), _ptr(pt), _ptr(st)))
        return npnt, nr, pt, st
