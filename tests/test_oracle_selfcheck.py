"""Oracle internal consistency: the per-ring kd-tree used by the timed CPU baseline returns exactly what brute force
(the truth, ties -> lower index) returns; generator invariants the benchmarks rely on."""
import numpy as np


def test_kdtree_equals_bruteforce(velo, oracle, calib, frames):
    _, pts, rs = frames[7]
    rng = np.random.default_rng(0)
    for s in (0, 17, 40, 63):
        ring = pts[rs[s]: rs[s + 1]]
        q = ring[rng.integers(0, len(ring), 300)].copy()
        q[:, :3] += rng.normal(size=(300, 3)).astype(np.float32) * 0.3
        far = (rng.normal(size=(100, 4)) * 30).astype(np.float32)
        assert oracle.nn_selfcheck(ring, np.concatenate([q, far, ring[:50]])) == 0


def test_kdtree_exact_ties(oracle):
    pts = np.zeros((64, 4), np.float32)
    pts[:, 0] = np.repeat(np.arange(8), 8)          # 8 copies of each of 8 positions: every query has 8-fold ties
    q = np.zeros((8, 4), np.float32); q[:, 0] = np.arange(8) + 0.25
    assert oracle.nn_selfcheck(pts, q) == 0


def test_icp_kdtree_mode_equals_bruteforce_mode(velo, oracle, calib, params):
    from conftest import small_scan
    rawM, rawS = small_scan(velo, 8), small_scan(velo, 7)
    ptsM, rsM, _ = oracle.segment(rawM, calib)
    ptsS, rsS, _ = oracle.segment(rawS, calib)
    pose = velo.synth.pose_guess(8, 0)
    a, na, ka = oracle.icp_pass(ptsM, rsM, ptsS, rsS, pose, 1, 3, params, 0)
    b, nb, kb = oracle.icp_pass(ptsM, rsM, ptsS, rsS, pose, 1, 3, params, 1)
    assert a.tobytes() == b.tobytes() and na.tobytes() == nb.tobytes() and ka == kb > 100


def test_generator_shape(velo, oracle, calib):
    """~120k points, 64 rings of ~1900, ~18k in-FOV survivors per camera, deterministic"""
    raw, n = velo.synth.scan(42)
    raw2, n2 = velo.synth.scan(42)
    assert n == n2 and raw.tobytes() == raw2.tobytes()
    assert 117000 < n < 124000
    pts, rs, nr = oracle.segment(raw, calib)
    assert nr == 64 and 1500 < np.diff(rs).min() and np.diff(rs).max() <= 2083
    rc, proj, valid = oracle.project(pts, rs, calib, 0)
    assert 14000 < rc.sum() < 22000
    kpA, kpB, m = velo.synth.features(42, 2000)
    hd, _ = oracle.depth_assoc(valid, proj, rc, kpA[0])
    assert 600 < (hd >= 0).sum() < 1800 and m.sum() > 2000
    p = velo.synth.pose(42)
    assert np.abs(p[:3]).max() < 0.05 and 0.7 < p[5] < 1.3


def test_oracle_frame_to_frame_recovers_motion(velo, oracle, calib, params):
    """SURVEY §8(f1) on the CPU side: frozen-block LM (the stand-in for ceres::Solve) inside the reference's 2 x 3 schedule
    recovers the synthetic ground-truth motion from the reference's initial guess (main.cpp:170) to noise level."""
    from conftest import small_scan
    rawM, rawS = small_scan(velo, 8, range(10, 54), 2), small_scan(velo, 7, range(10, 54), 2)
    ptsM, rsM, _ = oracle.segment(rawM, calib)
    ptsS, rsS, _ = oracle.segment(rawS, calib)
    truth = velo.synth.pose(8)
    x, rep = oracle.frame_to_frame(ptsM, rsM, ptsS, rsS, calib, params, np.array([0, 0, 0, 0, 0, 1.0]), None, 1, 5)
    assert rep["n_solves"] == 6 and all(r in (1, 2, 3) for r in rep["reason"])
    assert all(f <= i for f, i in zip(rep["final_cost"], rep["initial_cost"]))
    assert np.abs(x[:3] - truth[:3]).max() < 1e-3 and np.abs(x[3:] - truth[3:]).max() < 1e-2
