#!/usr/bin/env python3
"""Generate tests/golden/velo_golden.npz from the reference's OWN source lines (oracle/_ref/libvelo_ref.so, built by
oracle/build_ref.py from /root/reference).  Run in the build container only:  python tests/golden/make_golden.py

Inputs are small seeded synthetic cases; every output array is what the reference code returned for them.  The
fixture travels to the GPU box (where /root/reference does not exist) and pins both the oracle and the CUDA path.
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(HERE)]
import pyoracle  # noqa: E402

velo = importlib.import_module("vision-enhanced-lidar-odometry_b200")
abi, synth = velo.abi, velo.synth


def thin(frame, rings=range(28, 34), stride=2):
    raw, n = synth.scan(frame)
    x, y = raw[:, 0], raw[:, 1]
    flag = np.zeros(n, bool); flag[1:] = (x[1:] > 0) & ((y[1:] > 0) != (y[:-1] > 0))
    ring = np.cumsum(flag)
    return np.ascontiguousarray(raw[np.nonzero(np.isin(ring, list(rings)))[0][::stride]])


def main():
    ref = pyoracle.Ref()
    orc = pyoracle.Oracle()      # only used for calibration packing (inputs), not for any golden output
    P, Tr, w, h = synth.calib_raw(0)
    cal = orc.calib_from_kitti(P, Tr, w, h)
    prm = velo.api.default_params()
    g = {"P": P, "Tr": Tr, "wh": np.array([w, h], np.int32)}
    F = 300
    per_frame = {}
    for f in (60, 61):
        raw = thin(f)
        pts, rs, nr = ref.segment(raw, cal)
        g[f"raw{f}"], g[f"pts{f}"], g[f"rs{f}"] = raw, pts, rs
        kpA, kpB, m = synth.features(f, F)
        g[f"kpA{f}"], g[f"kpB{f}"], g[f"m{f}"] = kpA, kpB, m
        hd = np.zeros((2, 2, F), np.int32); kw = np.zeros((2, 2, F, 4), np.float32)
        for cam in (0, 1):
            rc, proj, valid = ref.project(pts, rs, cal, cam)
            g[f"rc{f}_{cam}"], g[f"proj{f}_{cam}"], g[f"valid{f}_{cam}"] = rc, proj, valid
            for s, kp in enumerate((kpA[cam], kpB[cam])):
                hh, kk = ref.depth_assoc(valid, proj, rc, kp)
                hd[s, cam] = hh; kw[s, cam, : len(kk)] = kk
        g[f"hd{f}"], g[f"kw{f}"] = hd, kw
        per_frame[f] = (pts, rs, kpA, kpB, m, hd, kw)
    ptsS, rsS = per_frame[60][:2]
    ptsM, rsM = per_frame[61][:2]
    for it, skip, pidx in ((1, 1, 0), (2, 1, 3), (1, 4, 1)):
        pose = synth.pose_guess(61, pidx)
        corr, neq = ref.icp_pass(ptsM, rsM, ptsS, rsS, pose, it, skip)
        g[f"icp_pose_{it}_{skip}"], g[f"icp_corr_{it}_{skip}"], g[f"icp_neq_{it}_{skip}"] = pose, corr, neq
    tp = np.array([0.013, -0.021, 0.008, 0.04, -0.03, 1.07])
    g["tp_pose"], g["tp_out"] = tp, ref.transform_points(ptsM[::11], tp)
    # visual blocks
    _, _, kpA60, _, _, hd60, kw60 = per_frame[60]
    _, _, _, kpB61, m61, hd61, kw61 = per_frame[61]
    matches = np.zeros((2, F, 2), np.int32); nm = np.zeros(2, np.int32)
    for cam in (0, 1):
        idx = np.nonzero(m61[cam])[0]
        nm[cam] = len(idx); matches[cam, : len(idx), 0] = idx; matches[cam, : len(idx), 1] = idx
    g["vis_matches"], g["vis_nm"] = matches, nm
    for it, pidx in ((1, 0), (2, 3)):
        pose = synth.pose_guess(61, pidx)
        blocks, neq = ref.visual(orc, kpB61, kpA60, hd61[1], hd60[0], kw61[1], kw60[0], nm, matches, cal, prm, pose, it)
        g[f"vis_pose_{it}"], g[f"vis_blocks_{it}"], g[f"vis_neq_{it}"] = pose, blocks, neq
    # functors: costfunctions.h verbatim
    rng = np.random.default_rng(77)
    ks, poses, rs_, Js = [], [], [], []
    sizes = {abi.RES_3DPD: 9, abi.RES_3D3D: 6, abi.RES_3D2D: 8, abi.RES_2D3D: 8, abi.RES_2D2D: 7}
    for typ, n in sizes.items():
        for _ in range(4):
            k = np.zeros(9); k[:n] = rng.normal(size=n)
            if typ in (abi.RES_3D2D, abi.RES_2D3D):
                k[2] = 6 + abs(k[2]) * 5
            pose = rng.normal(size=6) * [0.02, 0.02, 0.02, 0.1, 0.1, 1.0]
            r, J = ref.eval_functor(typ, k[:n], pose)
            rr = np.zeros(3); JJ = np.zeros((3, 6)); rr[: len(r)] = r; JJ[: len(r)] = J
            ks.append(np.concatenate([[typ, n], k])); poses.append(pose); rs_.append(rr); Js.append(JJ)
    g["fun_k"], g["fun_pose"], g["fun_r"], g["fun_J"] = np.array(ks), np.array(poses), np.array(rs_), np.array(Js)
    g["constants"] = ref.constants()
    out = os.path.join(HERE, "velo_golden.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
