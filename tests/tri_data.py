"""Synthetic landmark observations for the triangulation tests: L landmarks seen from n_frames camera poses,
with 3-D (lidar-associated) and 2-D (canonical image) observations in the reference's order (camera-major, frame ascending)."""
import importlib
import numpy as np


def make(seed, L=300, n_frames=6, ncam=2, cam_tx=(0.0, -0.5371657)):
    velo = importlib.import_module("vision-enhanced-lidar-odometry_b200")
    abi = velo.abi
    rng = np.random.default_rng(seed)
    poses = np.zeros((n_frames, 6))
    for f in range(n_frames):   # world -> camera f is p_f = R(-w) (p - t): the reference's convention (costfunctions.h:312-321)
        poses[f] = np.concatenate([rng.normal(0, 0.02, 3), [rng.normal(0, 0.1), rng.normal(0, 0.05), 1.0 * f + rng.normal(0, 0.1)]])

    def rot(w, p):
        th = np.linalg.norm(w)
        if th < 1e-12:
            return p + np.cross(w, p)
        u = w / th
        return p * np.cos(th) + np.cross(u, p) * np.sin(th) + u * np.dot(u, p) * (1 - np.cos(th))

    off3, off2, o3, o2, truth = [0], [0], [], [], []
    for l in range(L):
        P = np.array([rng.uniform(-8, 8), rng.uniform(-2, 1.5), rng.uniform(8, 40)])
        truth.append(P)
        frames = np.sort(rng.choice(n_frames, size=int(rng.integers(0, n_frames + 1)), replace=False))
        n3 = 0
        for f in frames:
            if rng.random() < 0.4:
                m = rot(-poses[f, :3], P - poses[f, 3:]) + rng.normal(0, 0.03, 3)
                o3.append((int(f), *m.astype(np.float32))); n3 += 1
        for cam in range(ncam):
            for f in frames:
                if rng.random() < 0.7:
                    m = rot(-poses[f, :3], P - poses[f, 3:]) + np.array([cam_tx[cam], 0, 0])
                    if m[2] > 1.0:
                        o2.append((int(f), cam, np.float32(m[0] / m[2] + rng.normal(0, 3e-4)), np.float32(m[1] / m[2] + rng.normal(0, 3e-4))))
        off3.append(len(o3)); off2.append(len(o2))
    obs3 = np.array(o3, dtype=abi.TRI_OBS3_DTYPE) if o3 else np.zeros(0, abi.TRI_OBS3_DTYPE)
    obs2 = np.array(o2, dtype=abi.TRI_OBS2_DTYPE) if o2 else np.zeros(0, abi.TRI_OBS2_DTYPE)
    return np.array(off3, np.int32), obs3, np.array(off2, np.int32), obs2, poses, np.array(truth)
