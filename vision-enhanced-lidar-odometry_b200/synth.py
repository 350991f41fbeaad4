"""ctypes binding of host/velo_synth.c — seeded synthetic KITTI-shaped inputs."""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build_synth())
        _lib.velo_synth_scan.restype = C.c_int
        _lib.velo_synth_scan.argtypes = [C.c_uint64, C.c_int, C.c_void_p, C.c_int]
        _lib.velo_synth_calib.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib.velo_synth_features.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.velo_synth_pose.argtypes = [C.c_uint64, C.c_int, C.c_void_p]
        _lib.velo_synth_pose_guess.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_void_p]
        _lib.velo_synth_pose_guess_spread.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return _lib


SEED = 1234


def calib_raw(rig=0):
    """(P[48] f32, Tr[12] f32, width, height) — KITTI-like calib.txt contents (kitti.h:66-105)."""
    P = np.zeros(48, np.float32)
    Tr = np.zeros(12, np.float32)
    w, h = C.c_int(), C.c_int()
    lib().velo_synth_calib(rig, P.ctypes.data, Tr.ctypes.data, C.byref(w), C.byref(h))
    return P, Tr, w.value, h.value


def scan(frame, seed=SEED, out=None, max_points=140000):
    """KITTI .bin layout float4 {x,y,z,reflectance}; returns (array[n,4], n)."""
    buf = out if out is not None else np.zeros((max_points, 4), np.float32)
    n = lib().velo_synth_scan(seed, frame, buf.ctypes.data, buf.shape[0])
    return (buf if out is not None else buf[:n]), n


def features(frame, F, ncam=2, rig=0, seed=SEED, kpA=None, kpB=None, match=None):
    kpA = np.zeros((ncam, F, 2), np.float32) if kpA is None else kpA
    kpB = np.zeros((ncam, F, 2), np.float32) if kpB is None else kpB
    match = np.zeros((ncam, F), np.int32) if match is None else match
    lib().velo_synth_features(seed, frame, rig, ncam, F, kpA.ctypes.data, kpB.ctypes.data, match.ctypes.data)
    return kpA, kpB, match


def pose(frame, seed=SEED):
    out = np.zeros(6, np.float64)
    lib().velo_synth_pose(seed, frame, out.ctypes.data)
    return out


POSE_SPREADS = {"tight": 0, "spec": 1}


def pose_guess(frame, pass_idx, seed=SEED, spread="tight"):
    """supplied pose of ICP pass `pass_idx`.  "tight": truth +- 0.004 rad / 0.04 m, shrinking 1/(1+pass).  "spec" (SURVEY.md 8(d)):
    pass 0 = the reference's start (0,0,0,0,0,1) (main.cpp:170), later passes truth +- 0.02 rad / (0.05, 0.05, 0.2) m, shrinking."""
    out = np.zeros(6, np.float64)
    lib().velo_synth_pose_guess_spread(seed, frame, pass_idx, POSE_SPREADS[spread], out.ctypes.data)
    return out


class Batch:
    """Host-side batch in the layout of velo_batch_inputs (include/velo_gpu.h).

    `count` scans: frames frame0 .. frame0+count-1.  Frame pairs (t, t-1) exist for t >= 1 (slot 0 is the halo).
    Arrays may be views of pinned memory (pass `alloc`)."""

    def __init__(self, frame0, count, prm, rig=0, seed=SEED, alloc=None, threads=None, pose_spread="tight"):
        C_, F, NP, MM = prm.num_cams, prm.max_features, prm.max_points, prm.max_matches
        n_passes = prm.f2f_iterations * prm.icp_iterations
        n_vis = prm.f2f_iterations
        mk = alloc if alloc is not None else (lambda shape, dt: np.zeros(shape, dt))
        self.count, self.frame0 = count, frame0
        self.scans = mk((count, NP, 4), np.float32)
        self.n_points = mk((count,), np.int32)
        self.kp = mk((count, 2, C_, F, 2), np.float32)
        self.n_kp = mk((count, 2, C_), np.int32)
        self.matches = mk((count, C_, MM, 2), np.int32)
        self.n_matches = mk((count, C_), np.int32)
        self.icp_poses = mk((count, n_passes, 6), np.float64)
        self.pass_iter = mk((n_passes,), np.int32)
        self.vis_poses = mk((count, n_vis, 6), np.float64)
        self.n_passes, self.n_vis = n_passes, n_vis
        for p in range(n_passes):
            self.pass_iter[p] = p // prm.icp_iterations + 1          # velo.h:616,800
        L = lib()

        def gen(i):
            fr = frame0 + i
            self.n_points[i] = L.velo_synth_scan(seed, fr, self.scans[i].ctypes.data, NP)
            kpA, kpB, m = features(fr, F, C_, rig, seed)
            self.kp[i, 0], self.kp[i, 1] = kpA, kpB
            self.n_kp[i] = F
            for c in range(C_):
                idx = np.nonzero(m[c])[0][:MM].astype(np.int32)
                self.n_matches[i, c] = len(idx)
                self.matches[i, c, :len(idx), 0] = idx
                self.matches[i, c, :len(idx), 1] = idx
            for p in range(n_passes):
                self.icp_poses[i, p] = pose_guess(fr, p, seed, pose_spread)
            for it in range(n_vis):
                self.vis_poses[i, it] = pose_guess(fr, it * prm.icp_iterations, seed, pose_spread)

        import os
        nt = threads or min(32, os.cpu_count() or 1)
        with ThreadPoolExecutor(nt) as ex:
            list(ex.map(gen, range(count)))

    def xyz(self, alloc=None):
        """the same batch with packed {x,y,z} scan records (velo_batch_inputs.scan_stride_floats = 3)"""
        v = self.view(0, self.count)
        mk = alloc if alloc is not None else (lambda shape, dt: np.zeros(shape, dt))
        v.scans = mk(self.scans.shape[:-1] + (3,), np.float32)
        v.scans[...] = self.scans[..., :3]
        return v

    def view(self, first, count):
        """entries [first, first + count) as a batch of their own (arrays are views; entry `first` becomes the halo scan)"""
        v = object.__new__(Batch)
        v.count, v.frame0 = count, self.frame0 + first
        for k in ("scans", "n_points", "kp", "n_kp", "matches", "n_matches", "icp_poses", "vis_poses"):
            setattr(v, k, getattr(self, k)[first:first + count])
        v.pass_iter, v.n_passes, v.n_vis = self.pass_iter, self.n_passes, self.n_vis
        return v
