"""ctypes wrappers for the CPU oracle (libvelo_oracle.so) and, when built, the reference-slice library
(oracle/_ref/libvelo_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import importlib
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
abi = importlib.import_module("vision-enhanced-lidar-odometry_b200.abi")

ORACLE_LIB = os.path.join(HERE, "libvelo_oracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libvelo_ref.so")


def build_oracle(force=False):
    src = os.path.join(HERE, "velo_oracle.cpp")
    hdr = os.path.join(ROOT, "include", "velo_gpu.h")
    fresh = os.path.isfile(ORACLE_LIB) and all(os.path.getmtime(ORACLE_LIB) >= os.path.getmtime(s) for s in (src, hdr))
    if fresh and not force:
        return ORACLE_LIB
    if shutil.which("g++") is None and os.path.isfile(ORACLE_LIB):
        return ORACLE_LIB
    # no -march, no FMA contraction: mirrors the reference build (CMakeLists.txt:30).  Built under a lock into a temporary name
    # (several test / bench processes may start at once).
    import fcntl
    with open(ORACLE_LIB + ".lock", "w") as lf:
        fcntl.flock(lf, fcntl.LOCK_EX)
        fresh = os.path.isfile(ORACLE_LIB) and all(os.path.getmtime(ORACLE_LIB) >= os.path.getmtime(s) for s in (src, hdr))
        if not fresh or force:
            tmp = f"{ORACLE_LIB}.tmp.{os.getpid()}"
            subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-pthread", "-w", "-o", tmp, src], check=True)
            os.replace(tmp, ORACLE_LIB)
    return ORACLE_LIB


def build_ref(force=False):
    sys.path.insert(0, HERE)
    import build_ref as br
    return br.build(force)


def default_params(**kw):
    """velo_gpu_params with the reference's tunables (kitti.h:3-35, main.cpp:42-49), filled here in plain Python so that the CPU
    reference arm never loads the CUDA library."""
    p = abi.Params()
    p.num_cams, p.icp_skip, p.f2f_iterations, p.icp_iterations = 2, 200, 2, 3
    p.enable_2d2d, p.enable_3d2d, p.abs_truncates = 1, 1, 0
    p.weight_3D2D, p.weight_2D2D, p.weight_3DPD = 10.0, 500.0, 1.0
    p.loss_thresh_3D2D, p.loss_thresh_2D2D, p.loss_thresh_3DPD, p.loss_thresh_3D3D = 0.01, 0.00002, 0.1, 0.04
    p.depth_assoc_thresh, p.outlier_reject, p.correspondence_thresh_icp, p.icp_norm_condition = 0.015, 5.0, 0.5, 1e-5
    p.max_slots, p.max_points, p.max_rings, p.max_features, p.max_matches, p.max_icp_passes = 4, 131072, 128, 3000, 3000, 6
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


_P = C.c_void_p


def _ptr(a):
    return None if a is None else a.ctypes.data


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.oracle_segment.argtypes = [_P, C.c_int, _P, _P, _P, C.c_int]
        L.oracle_project.argtypes = [_P, _P, C.c_int, _P, C.c_int, _P, _P, _P]
        L.oracle_depth_assoc.argtypes = [_P, _P, _P, C.c_int, _P, C.c_int, C.c_double, C.c_int, _P, _P]
        L.oracle_transform_points.argtypes = [_P, C.c_int, _P, _P]
        L.oracle_icp_pass.argtypes = [_P, _P, C.c_int, _P, _P, C.c_int, _P, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P]
        L.oracle_nn_selfcheck.argtypes = [_P, C.c_int, _P, C.c_int]
        L.oracle_eval_functor.argtypes = [C.c_int, _P, _P, _P, _P]
        L.oracle_loss.argtypes = [C.c_int, C.c_double, C.c_double, _P]
        L.oracle_visual.argtypes = [C.c_int, C.c_int, C.c_int] + [_P] * 10 + [_P, _P, _P, C.c_int, _P, C.c_int, _P]
        L.oracle_calib_from_kitti.argtypes = [_P, _P, C.c_int, C.c_int, _P]
        L.oracle_pixel2canonical.argtypes = [_P, C.c_int, _P, C.c_int, _P]
        L.oracle_canonical2pixel.argtypes = [_P, C.c_int, _P, C.c_int, _P]
        L.oracle_frame_to_frame.argtypes = [_P, _P, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int] + [_P] * 10 + [C.c_int, C.c_int, _P, _P]
        L.oracle_match_hamming.argtypes = [_P, C.c_int, _P, C.c_int, C.c_int, C.c_double, _P, _P, _P]
        L.oracle_triangulate.argtypes = [C.c_int, _P, _P, _P, _P, _P, C.c_int, _P, _P, _P, _P, _P, _P]
        L.oracle_bench_frames.restype = C.c_double
        L.oracle_bench_frames.argtypes = [C.c_int, C.c_int, _P, _P, _P, C.c_int, _P, _P]

    # ---- a1
    def calib_from_kitti(self, P, Tr, w, h):
        cal = abi.Calib()
        self.lib.oracle_calib_from_kitti(_ptr(np.ascontiguousarray(P, np.float32)), _ptr(np.ascontiguousarray(Tr, np.float32)), w, h, C.addressof(cal))
        return cal

    def pixel2canonical(self, cal, cam, pix):
        pix = np.ascontiguousarray(pix, np.float32)
        out = np.zeros_like(pix)
        self.lib.oracle_pixel2canonical(C.addressof(cal), cam, _ptr(pix), len(pix), _ptr(out))
        return out

    def canonical2pixel(self, cal, cam, can):
        can = np.ascontiguousarray(can, np.float32)
        out = np.zeros_like(can)
        self.lib.oracle_canonical2pixel(C.addressof(cal), cam, _ptr(can), len(can), _ptr(out))
        return out

    # ---- a3
    def segment(self, xyzr, cal, max_rings=4096):
        xyzr = np.ascontiguousarray(xyzr, np.float32)
        n = len(xyzr)
        out = np.zeros((max(n, 1), 4), np.float32)
        rs = np.zeros(max_rings + 1, np.int32)
        nr = self.lib.oracle_segment(_ptr(xyzr), n, C.addressof(cal), _ptr(out), _ptr(rs), max_rings)
        return out[:n], rs[:min(nr, max_rings) + 1].copy(), nr

    # ---- a5
    def project(self, pts, rs, cal, cam):
        pts = np.ascontiguousarray(pts, np.float32)
        rs = np.ascontiguousarray(rs, np.int32)
        nr, n = len(rs) - 1, len(pts)
        rc = np.zeros(max(nr, 1), np.int32)
        proj = np.zeros((max(n, 1), 2), np.float32)
        valid = np.zeros((max(n, 1), 4), np.float32)
        tot = self.lib.oracle_project(_ptr(pts), _ptr(rs), nr, C.addressof(cal), cam, _ptr(rc), _ptr(proj), _ptr(valid))
        return rc[:nr], proj[:tot], valid[:tot]

    # ---- a7
    def depth_assoc(self, valid, proj, rc, kp, thresh=0.015, abs_truncates=0):
        valid = np.ascontiguousarray(valid, np.float32)
        proj = np.ascontiguousarray(proj, np.float32)
        rc = np.ascontiguousarray(rc, np.int32)
        kp = np.ascontiguousarray(kp, np.float32)
        F = len(kp)
        hd = np.zeros(max(F, 1), np.int32)
        kpwd = np.zeros((max(F, 1), 4), np.float32)
        nh = self.lib.oracle_depth_assoc(_ptr(valid), _ptr(proj), _ptr(rc), len(rc), _ptr(kp), F, thresh, abs_truncates, _ptr(hd), _ptr(kpwd))
        return hd[:F], kpwd[:nh]

    def transform_points(self, pts, pose):
        pts = np.ascontiguousarray(pts, np.float32)
        out = np.zeros_like(pts)
        pose = np.ascontiguousarray(pose, np.float64)
        self.lib.oracle_transform_points(_ptr(pts), len(pts), _ptr(pose), _ptr(out))
        return out

    # ---- a10/a11/a13
    def icp_pass(self, ptsM, rsM, ptsS, rsS, pose, it, skip, prm, mode=1):
        ptsM = np.ascontiguousarray(ptsM, np.float32); rsM = np.ascontiguousarray(rsM, np.int32)
        ptsS = np.ascontiguousarray(ptsS, np.float32); rsS = np.ascontiguousarray(rsS, np.int32)
        pose = np.ascontiguousarray(pose, np.float64)
        nq = int(sum(-(-int(l) // skip) for l in np.diff(rsM)))
        corr = np.zeros(max(nq, 1), abi.ICP_CORR_DTYPE)
        neq = np.zeros(abi.NEQ_STRIDE, np.float64)
        kept = C.c_int()
        q = self.lib.oracle_icp_pass(_ptr(ptsM), _ptr(rsM), len(rsM) - 1, _ptr(ptsS), _ptr(rsS), len(rsS) - 1, _ptr(pose), it, skip,
                                     C.addressof(prm), mode, _ptr(corr), C.addressof(kept), _ptr(neq))
        assert q == nq, (q, nq)
        return corr[:nq], neq, kept.value

    def nn_selfcheck(self, pts, queries):
        pts = np.ascontiguousarray(pts, np.float32); queries = np.ascontiguousarray(queries, np.float32)
        return self.lib.oracle_nn_selfcheck(_ptr(pts), len(pts), _ptr(queries), len(queries))

    def eval_functor(self, typ, k, pose):
        k = np.ascontiguousarray(k, np.float64); pose = np.ascontiguousarray(pose, np.float64)
        r = np.zeros(3); J = np.zeros(18)
        nr = self.lib.oracle_eval_functor(typ, _ptr(k), _ptr(pose), _ptr(r), _ptr(J))
        return r[:nr].copy(), J[:6 * nr].reshape(nr, 6).copy()

    def loss(self, kind, a, s):
        rho = np.zeros(3)
        self.lib.oracle_loss(kind, a, s, _ptr(rho))
        return rho

    # ---- a12/a13
    def visual(self, kp1, kp2, hd1, hd2, kpwd1, kpwd2, n_matches, matches, cal, prm, pose, it, lm_valid=None, lm_xyz=None, _lib=None, _fn="oracle_visual"):
        """kp*: [C][F][2], hd*: [C][F], kpwd*: [C][F][4], matches: [C][MM][2], n_matches[C]"""
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        kp1, kp2, kpwd1, kpwd2 = f32(kp1), f32(kp2), f32(kpwd1), f32(kpwd2)
        hd1, hd2, matches, n_matches = i32(hd1), i32(hd2), i32(matches), i32(n_matches)
        ncam, F = kp1.shape[0], kp1.shape[1]
        MM = matches.shape[1]
        lm_valid = None if lm_valid is None else i32(lm_valid)
        lm_xyz = None if lm_xyz is None else f32(lm_xyz)
        cap = 3 * int(n_matches.sum()) + 1
        blocks = np.zeros(cap, abi.VIS_BLOCK_DTYPE)
        neq = np.zeros(abi.NEQ_STRIDE, np.float64)
        pose = np.ascontiguousarray(pose, np.float64)
        if _fn == "oracle_visual":
            nb = self.lib.oracle_visual(ncam, F, MM, _ptr(kp1), _ptr(kp2), _ptr(hd1), _ptr(hd2), _ptr(kpwd1), _ptr(kpwd2), _ptr(n_matches), _ptr(matches),
                                        _ptr(lm_valid), _ptr(lm_xyz), C.addressof(cal), C.addressof(prm), _ptr(pose), it, _ptr(blocks), cap, _ptr(neq))
        else:
            nb = _lib.ref_visual(ncam, F, MM, _ptr(kp1), _ptr(kp2), _ptr(hd1), _ptr(hd2), _ptr(kpwd1), _ptr(kpwd2), _ptr(n_matches), _ptr(matches),
                                 _ptr(lm_valid), _ptr(lm_xyz), C.addressof(cal), _ptr(pose), it, _ptr(blocks), cap, _ptr(neq))
        return blocks[:nb], neq

    # ---- f3: triangulatePoint
    def triangulate(self, off3, obs3, off2, obs2, poses, cal, prm, init_xyz=None, has_init=None, _ref=None, ncam=2):
        off3 = np.ascontiguousarray(off3, np.int32); off2 = np.ascontiguousarray(off2, np.int32)
        obs3 = np.ascontiguousarray(obs3, abi.TRI_OBS3_DTYPE); obs2 = np.ascontiguousarray(obs2, abi.TRI_OBS2_DTYPE)
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 6)
        n = len(off3) - 1
        out = np.zeros((max(n, 1), 3), np.float32); it = np.zeros(max(n, 1), np.int32)
        ini = None if init_xyz is None else np.ascontiguousarray(init_xyz, np.float32)
        has = None if has_init is None else np.ascontiguousarray(has_init, np.int32)
        if _ref is None:
            self.lib.oracle_triangulate(n, _ptr(off3), _ptr(obs3), _ptr(off2), _ptr(obs2), _ptr(poses), len(poses), C.addressof(cal), C.addressof(prm),
                                        _ptr(ini), _ptr(has), _ptr(out), _ptr(it))
        else:
            _ref.ref_triangulate(n, _ptr(off3), _ptr(obs3), _ptr(off2), _ptr(obs2), _ptr(poses), len(poses), C.addressof(cal), ncam, _ptr(ini), _ptr(has), _ptr(out))
        return out[:n], it[:n]

    # ---- f4: matchFeatures
    def match_hamming(self, query, train, match_thresh=29.0):
        query = np.ascontiguousarray(query, np.uint8); train = np.ascontiguousarray(train, np.uint8)
        nq, nt = len(query), len(train)
        db = query.shape[1] if nq else 64
        pairs = np.zeros((max(nq, 1), 2), np.int32); bi = np.zeros(max(nq, 1), np.int32); bd = np.zeros(max(nq, 1), np.int32)
        n = self.lib.oracle_match_hamming(_ptr(query), nq, _ptr(train), nt, db, match_thresh, _ptr(pairs), _ptr(bi), _ptr(bd))
        return pairs[:n], bi[:nq], bd[:nq]

    # ---- f1: frameToFrame with the frozen-block LM solve
    def frame_to_frame(self, ptsM, rsM, ptsS, rsS, cal, prm, transform, vis=None, enable_icp=1, icp_skip=1):
        """vis = (kp1, kp2, hd1, hd2, kpwd1, kpwd2, n_matches, matches[C][MM][2]) or None"""
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        ptsM, ptsS, rsM, rsS = f32(ptsM), f32(ptsS), i32(rsM), i32(rsS)
        t = np.ascontiguousarray(transform, np.float64).copy()
        rep = abi.F2FReport()
        if vis is not None:
            kp1, kp2, hd1, hd2, kw1, kw2, nm, mt = vis
            kp1, kp2, kw1, kw2, hd1, hd2, nm, mt = f32(kp1), f32(kp2), f32(kw1), f32(kw2), i32(hd1), i32(hd2), i32(nm), i32(mt)
            ncam, F, MM = kp1.shape[0], kp1.shape[1], mt.shape[1]
            args = [_ptr(kp1), _ptr(kp2), _ptr(hd1), _ptr(hd2), _ptr(kw1), _ptr(kw2), _ptr(nm), _ptr(mt)]
        else:
            ncam, F, MM = prm.num_cams, 1, 1
            args = [None] * 8
        self.lib.oracle_frame_to_frame(_ptr(ptsM), _ptr(rsM), len(rsM) - 1, _ptr(ptsS), _ptr(rsS), len(rsS) - 1, ncam, F, MM, *args,
                                       C.addressof(cal), C.addressof(prm), enable_icp, icp_skip, _ptr(t), C.addressof(rep))
        n = rep.n_solves
        return t, {"n_solves": n, "lm_iterations": list(rep.lm_iterations)[:n], "accepted_steps": list(rep.accepted_steps)[:n], "reason": list(rep.reason)[:n],
                   "n_blocks": list(rep.n_blocks)[:n], "initial_cost": list(rep.initial_cost)[:n], "final_cost": list(rep.final_cost)[:n],
                   "pose": np.array([list(rep.pose[i]) for i in range(n)])}

    # ---- timed CPU baseline
    def bench_frames(self, batch, prm, cal, threads, stages=abi.STAGE_ALL, want_out=False):
        assert batch.scans.shape[-1] == 4, "the CPU reference reads KITTI float4 records (kitti.h:142)"
        bi = abi.BatchInputs(_ptr(batch.scans), _ptr(batch.n_points), _ptr(batch.kp), _ptr(batch.n_kp), _ptr(batch.matches), _ptr(batch.n_matches),
                             _ptr(batch.icp_poses), _ptr(batch.pass_iter), batch.n_passes, _ptr(batch.vis_poses), batch.n_vis, 4)
        icp = np.zeros((batch.count, batch.n_passes, abi.NEQ_STRIDE)) if want_out else None
        vis = np.zeros((batch.count, batch.n_vis, abi.NEQ_STRIDE)) if want_out else None
        sec = self.lib.oracle_bench_frames(batch.count, threads, C.addressof(prm), C.addressof(cal), C.addressof(bi), stages, _ptr(icp), _ptr(vis))
        return sec, icp, vis


class Ref:
    """The reference's own source lines (oracle/_ref/libvelo_ref.so); None-like when not built."""

    def __init__(self):
        if not os.path.isfile(REF_LIB):
            build_ref()
        if not os.path.isfile(REF_LIB):
            raise FileNotFoundError(REF_LIB)
        self.lib = C.CDLL(REF_LIB)
        L = self.lib
        L.ref_segment.argtypes = [_P, C.c_int, _P, _P, _P, C.c_int]
        L.ref_project.argtypes = [_P, _P, C.c_int, _P, C.c_int, _P, _P, _P]
        L.ref_depth_assoc.argtypes = [_P, _P, _P, C.c_int, _P, C.c_int, _P, _P]
        L.ref_transform_points.argtypes = [_P, C.c_int, _P, _P]
        L.ref_eval_functor.argtypes = [C.c_int, _P, _P, _P, _P]
        L.ref_eval_functor_plain.argtypes = [C.c_int, _P, _P, _P]
        L.ref_icp_pass.argtypes = [_P, _P, C.c_int, _P, _P, C.c_int, _P, C.c_int, C.c_int, _P, C.c_int, _P]
        L.ref_visual.argtypes = [C.c_int, C.c_int, C.c_int] + [_P] * 10 + [_P, _P, C.c_int, _P, C.c_int, _P]
        L.ref_constants.argtypes = [_P]
        L.ref_match_hamming.argtypes = [_P, C.c_int, _P, C.c_int, C.c_int, _P]
        L.ref_triangulate.argtypes = [C.c_int, _P, _P, _P, _P, _P, C.c_int, _P, C.c_int, _P, _P, _P]

    def segment(self, xyzr, cal, max_rings=4096):
        xyzr = np.ascontiguousarray(xyzr, np.float32)
        n = len(xyzr)
        out = np.zeros((max(n, 1), 4), np.float32)
        rs = np.zeros(max_rings + 1, np.int32)
        nr = self.lib.ref_segment(_ptr(xyzr), n, C.addressof(cal), _ptr(out), _ptr(rs), max_rings)
        return out[:n], rs[:min(nr, max_rings) + 1].copy(), nr

    def project(self, pts, rs, cal, cam):
        pts = np.ascontiguousarray(pts, np.float32); rs = np.ascontiguousarray(rs, np.int32)
        nr, n = len(rs) - 1, len(pts)
        rc = np.zeros(max(nr, 1), np.int32); proj = np.zeros((max(n, 1), 2), np.float32); valid = np.zeros((max(n, 1), 4), np.float32)
        tot = self.lib.ref_project(_ptr(pts), _ptr(rs), nr, C.addressof(cal), cam, _ptr(rc), _ptr(proj), _ptr(valid))
        return rc[:nr], proj[:tot], valid[:tot]

    def depth_assoc(self, valid, proj, rc, kp):
        valid = np.ascontiguousarray(valid, np.float32); proj = np.ascontiguousarray(proj, np.float32)
        rc = np.ascontiguousarray(rc, np.int32); kp = np.ascontiguousarray(kp, np.float32)
        F = len(kp)
        hd = np.zeros(max(F, 1), np.int32); kpwd = np.zeros((max(F, 1), 4), np.float32)
        nh = self.lib.ref_depth_assoc(_ptr(valid), _ptr(proj), _ptr(rc), len(rc), _ptr(kp), F, _ptr(hd), _ptr(kpwd))
        return hd[:F], kpwd[:nh]

    def transform_points(self, pts, pose):
        pts = np.ascontiguousarray(pts, np.float32); out = np.zeros_like(pts); pose = np.ascontiguousarray(pose, np.float64)
        self.lib.ref_transform_points(_ptr(pts), len(pts), _ptr(pose), _ptr(out))
        return out

    def eval_functor(self, typ, k, pose):
        k = np.ascontiguousarray(k, np.float64); pose = np.ascontiguousarray(pose, np.float64)
        r = np.zeros(3); J = np.zeros(18)
        nr = self.lib.ref_eval_functor(typ, _ptr(k), _ptr(pose), _ptr(r), _ptr(J))
        return r[:nr].copy(), J[:6 * nr].reshape(nr, 6).copy()

    def eval_functor_plain(self, typ, k, pose):
        k = np.ascontiguousarray(k, np.float64); pose = np.ascontiguousarray(pose, np.float64)
        r = np.zeros(3)
        nr = self.lib.ref_eval_functor_plain(typ, _ptr(k), _ptr(pose), _ptr(r))
        return r[:nr].copy()

    def icp_pass(self, ptsM, rsM, ptsS, rsS, pose, it, skip):
        ptsM = np.ascontiguousarray(ptsM, np.float32); rsM = np.ascontiguousarray(rsM, np.int32)
        ptsS = np.ascontiguousarray(ptsS, np.float32); rsS = np.ascontiguousarray(rsS, np.int32)
        pose = np.ascontiguousarray(pose, np.float64)
        nq = int(sum(-(-int(l) // skip) for l in np.diff(rsM)))
        corr = np.zeros(max(nq, 1), abi.ICP_CORR_DTYPE); neq = np.zeros(abi.NEQ_STRIDE, np.float64)
        k = self.lib.ref_icp_pass(_ptr(ptsM), _ptr(rsM), len(rsM) - 1, _ptr(ptsS), _ptr(rsS), len(rsS) - 1, _ptr(pose), it, skip, _ptr(corr), nq, _ptr(neq))
        return corr[:k], neq

    def visual(self, oracle, *a, **kw):
        return Oracle.visual(oracle, *a, _lib=self.lib, _fn="ref_visual", **kw)

    def match_hamming(self, query, train):
        query = np.ascontiguousarray(query, np.uint8); train = np.ascontiguousarray(train, np.uint8)
        nq = len(query)
        pairs = np.zeros((max(nq, 1), 2), np.int32)
        n = self.lib.ref_match_hamming(_ptr(query), nq, _ptr(train), len(train), query.shape[1] if nq else 64, _ptr(pairs))
        return pairs[:n]

    def constants(self):
        out = np.zeros(32)
        n = self.lib.ref_constants(_ptr(out))
        return out[:n]
