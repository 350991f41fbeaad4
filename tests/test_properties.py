"""Property tests (hypothesis): the oracle restatement against the reference's own source lines on random, ragged and
degenerate inputs — random ring counts (incl. 1-point and empty rings), keypoints on FOV edges, duplicated x, z ties,
exact-distance ties in the neighbour search.  CPU only; skipped when oracle/_ref is not available."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

SET = dict(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


def _rings(rng, n_rings, max_len, quant=None):
    """random ring-structured cam-0 cloud roughly in front of the camera; quant snaps coordinates to a grid to create ties"""
    lens = rng.integers(0, max_len + 1, n_rings)
    pts = []
    for s, L in enumerate(lens):
        az = np.sort(rng.uniform(-0.9, 0.9, L))
        r = rng.uniform(4, 30, L) if rng.random() < 0.5 else np.full(L, rng.uniform(4, 30))
        x = r * np.sin(az); z = r * np.cos(az); y = np.full(L, -0.5 + 0.12 * s) + rng.normal(0, 0.02, L)
        p = np.stack([x, y, z, np.ones(L)], 1)
        if quant:
            p[:, :3] = np.round(p[:, :3] / quant) * quant
        pts.append(p)
    rs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    return (np.concatenate(pts) if len(pts) else np.zeros((0, 4))).astype(np.float32), rs


@settings(**SET)
@given(seed=st.integers(0, 10**6), n_rings=st.integers(0, 9), max_len=st.integers(0, 40), quant=st.sampled_from([None, 0.5, 0.05]))
def test_project_property(oracle, ref, calib, seed, n_rings, max_len, quant):
    rng = np.random.default_rng(seed)
    pts, rs = _rings(rng, n_rings, max_len, quant)
    for cam in (0, 1):
        a = oracle.project(pts, rs, calib, cam)
        b = ref.project(pts, rs, calib, cam)
        assert np.array_equal(a[0], b[0]) and a[1].tobytes() == b[1].tobytes() and a[2].tobytes() == b[2].tobytes()


@settings(**SET)
@given(seed=st.integers(0, 10**6), n_rings=st.integers(1, 8), max_len=st.integers(0, 30), F=st.integers(0, 60), dup=st.booleans())
def test_depth_assoc_property(oracle, ref, calib, seed, n_rings, max_len, F, dup):
    rng = np.random.default_rng(seed)
    pts, rs = _rings(rng, n_rings, max_len, 0.05 if dup else None)      # duplicated x after quantisation (hazard H4)
    rc, proj, valid = oracle.project(pts, rs, calib, 0)
    kx = rng.uniform(calib.min_x[0], calib.max_x[0], F)
    ky = rng.uniform(calib.min_y[0], calib.max_y[0], F)
    if F > 3 and len(proj):                                            # keypoints exactly on projected points and FOV edges
        kx[0], ky[0] = proj[0]
        kx[1] = calib.min_x[0]; kx[2] = np.nextafter(np.float32(calib.max_x[0]), np.float32(0))
    kp = np.stack([kx, ky], 1).astype(np.float32)
    a = oracle.depth_assoc(valid, proj, rc, kp)
    b = ref.depth_assoc(valid, proj, rc, kp)
    assert np.array_equal(a[0], b[0]) and a[1].tobytes() == b[1].tobytes()


@settings(**SET)
@given(seed=st.integers(0, 10**6), n=st.integers(0, 300))
def test_segment_property(oracle, ref, calib, seed, n):
    rng = np.random.default_rng(seed)
    raw = (rng.normal(size=(n, 4)) * [10, 0.5, 1, 1]).astype(np.float32)   # small |y| => frequent seam crossings, tiny rings
    raw[rng.random(n) < 0.1, 1] = 0.0                                      # y == 0 exactly (sign test edge)
    a = oracle.segment(raw, calib)
    b = ref.segment(raw, calib)
    assert a[2] == b[2] and np.array_equal(a[1], b[1]) and a[0].tobytes() == b[0].tobytes()


@settings(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(seed=st.integers(0, 10**6), quant=st.sampled_from([None, 0.25]), it=st.sampled_from([1, 2]), skip=st.sampled_from([1, 3]))
def test_icp_property(velo, oracle, ref, params, seed, quant, it, skip):
    """random ring clouds incl. a grid-snapped variant where exact distance ties are common (ties -> lower ring, lower index)"""
    rng = np.random.default_rng(seed)
    ptsS, rsS = _rings(rng, 6, 60, quant)
    ptsM, rsM = _rings(rng, 3, 30, quant)
    pose = np.concatenate([rng.normal(0, 0.01, 3), rng.normal(0, 0.05, 3)]) if quant is None else np.zeros(6)
    cr, nr = ref.icp_pass(ptsM, rsM, ptsS, rsS, pose, it, skip)
    for mode in (0, 1):
        co, no, kept = oracle.icp_pass(ptsM, rsM, ptsS, rsS, pose, it, skip, params, mode)
        ck = co[co["kept"] == 1]
        assert len(ck) == len(cr) == kept
        for f in ("src_ring", "src_idx", "np_s_i", "np_i", "np_s_j", "np_j", "np_k"):
            assert np.array_equal(ck[f], cr[f]), f
        assert ck["normal"].tobytes() == cr["normal"].tobytes() and ck["residual"].tobytes() == cr["residual"].tobytes()
