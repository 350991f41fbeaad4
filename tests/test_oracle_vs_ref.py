"""The oracle restatement against the reference's OWN source lines (oracle/_ref/libvelo_ref.so, built by
oracle/build_ref.py from /root/reference).  Bit-exact for floats produced by f32 arithmetic, indices exact.
Skipped when the reference-slice library is not available."""
import numpy as np
import pytest
from conftest import small_scan


def test_constants_match_reference(ref, params):
    c = ref.constants()
    p = params
    got = [4, p.max_features, p.icp_skip, p.f2f_iterations, p.icp_iterations, p.weight_3D2D, p.weight_2D2D, p.weight_3DPD,
           p.loss_thresh_3D2D, p.loss_thresh_2D2D, p.loss_thresh_3DPD, p.loss_thresh_3D3D, p.depth_assoc_thresh,
           p.outlier_reject, p.correspondence_thresh_icp, p.icp_norm_condition]
    assert list(c) == [float(v) for v in got]


@pytest.mark.parametrize("frame", [3, 50])
def test_segment_bit_exact(velo, oracle, ref, calib, frame):
    raw, n = velo.synth.scan(frame)
    a, rsa, nra = oracle.segment(raw, calib)
    b, rsb, nrb = ref.segment(raw, calib)
    assert nra == nrb == 64
    assert np.array_equal(rsa, rsb)
    assert a.tobytes() == b.tobytes()


def test_segment_ragged(velo, oracle, ref, calib):
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 5, 97):
        raw = rng.normal(size=(n, 4)).astype(np.float32) * 10
        a, rsa, nra = oracle.segment(raw, calib)
        b, rsb, nrb = ref.segment(raw, calib)
        assert nra == nrb
        assert np.array_equal(rsa, rsb)
        assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("cam", [0, 1, 2])
def test_project_bit_exact(oracle, ref, calib, frames, cam):
    _, pts, rs = frames[7]
    rca, pa, va = oracle.project(pts, rs, calib, cam)
    rcb, pb, vb = ref.project(pts, rs, calib, cam)
    assert np.array_equal(rca, rcb)
    assert pa.tobytes() == pb.tobytes() and va.tobytes() == vb.tobytes()
    assert rca.sum() > 10000


def test_project_occlusion_cases(oracle, ref, calib):
    """hand-made ring exercising pop / skip / z-tie push (velo.h:351-365)"""
    z = np.array([10, 10, 5, 12, 5, 5, 20, 3, 3, 30], np.float32)
    cx = np.array([-0.2, -0.1, -0.15, -0.12, 0.0, -0.05, 0.1, 0.05, 0.02, 0.3], np.float32)
    pts = np.stack([cx * z, 0.05 * z, z, np.ones_like(z)], 1).astype(np.float32)
    rs = np.array([0, len(z)], np.int32)
    rca, pa, va = oracle.project(pts, rs, calib, 0)
    rcb, pb, vb = ref.project(pts, rs, calib, 0)
    assert np.array_equal(rca, rcb) and pa.tobytes() == pb.tobytes() and va.tobytes() == vb.tobytes()
    assert 0 < rca[0] < len(z)


def test_depth_assoc_bit_exact(velo, oracle, ref, calib, frames):
    _, pts, rs = frames[7]
    kpA, kpB, _ = velo.synth.features(7, 2000)
    for cam in (0, 1):
        rc, proj, valid = oracle.project(pts, rs, calib, cam)
        for kp in (kpA[cam], kpB[cam]):
            ha, ka = oracle.depth_assoc(valid, proj, rc, kp)
            hb, kb = ref.depth_assoc(valid, proj, rc, kp)
            assert np.array_equal(ha, hb)
            assert ka.tobytes() == kb.tobytes()
            assert (ha >= 0).sum() > 500


def test_depth_assoc_edge_rings(oracle, ref, calib):
    """rings with 0/1 points reset the bracket (velo.h:400-403); keypoints on FOV edges"""
    rc = np.array([3, 1, 0, 4, 4, 2], np.int32)
    rng = np.random.default_rng(1)
    proj, valid = [], []
    for s, c in enumerate(rc):
        xs = np.sort(rng.uniform(-0.01, 0.01, c)).astype(np.float32)
        for x in xs:
            proj.append((x, 0.02 * s - 0.05)); valid.append((x * 9, 1.0, 9.0 + s, 1.0))
    proj = np.array(proj, np.float32); valid = np.array(valid, np.float32)
    kp = np.stack([rng.uniform(-0.012, 0.012, 400), rng.uniform(-0.06, 0.08, 400)], 1).astype(np.float32)
    ha, ka = oracle.depth_assoc(valid, proj, rc, kp)
    hb, kb = ref.depth_assoc(valid, proj, rc, kp)
    assert np.array_equal(ha, hb) and ka.tobytes() == kb.tobytes()
    assert (ha >= 0).any() and (ha < 0).any()


def test_transform_point_bit_exact(oracle, ref, frames):
    _, pts, _ = frames[7]
    for pose in ([0.01, -0.02, 0.005, 0.03, -0.04, 1.1], [0, 0, 0, 0, 0, 1.0], [1e-9, 0, 1e-9, 0.1, 0.2, 0.3], [0.5, -1.0, 2.0, 1, 2, 3]):
        a = oracle.transform_points(pts[::37], pose)
        b = ref.transform_points(pts[::37], pose)
        assert a.tobytes() == b.tobytes()


def _rand_consts(rng, typ, abi):
    n = {abi.RES_3DPD: 9, abi.RES_3D3D: 6, abi.RES_3D2D: 8, abi.RES_2D3D: 8, abi.RES_2D2D: 7}[typ]
    k = rng.normal(size=n)
    if typ == abi.RES_3DPD:
        k[:3] *= 10; k[6:9] = k[:3] + rng.normal(size=3) * 0.2; k[3:6] /= np.linalg.norm(k[3:6])
    elif typ == abi.RES_3D3D:
        k[:3] *= 10; k[3:] = k[:3] + rng.normal(size=3) * 0.1
    elif typ in (abi.RES_3D2D, abi.RES_2D3D):
        k[:3] = [rng.normal() * 3, rng.normal(), 5 + abs(rng.normal()) * 10]; k[3:5] = k[:2] / k[2] + rng.normal(size=2) * 0.01; k[5:] = [-0.537, 0, 0]
    else:
        k[:4] *= 0.3; k[4:] = [-0.537, 0, 0]
    return k.astype(np.float32).astype(np.float64)


def test_functors_match_reference(velo, oracle, ref):
    """costfunctions.h:17-220 compiled verbatim vs the restated functors: residuals and autodiff Jacobians."""
    abi = velo.abi
    rng = np.random.default_rng(2)
    poses = [rng.normal(size=6) * [0.02, 0.02, 0.02, 0.1, 0.1, 1.0] for _ in range(6)] + [np.array([0, 0, 0, 0, 0, 1.0]), np.array([1e-9, -1e-9, 0, 0.1, 0, 1])]
    for typ in (abi.RES_3DPD, abi.RES_3D3D, abi.RES_3D2D, abi.RES_2D3D, abi.RES_2D2D):
        for pose in poses:
            k = _rand_consts(rng, typ, abi)
            ra, Ja = oracle.eval_functor(typ, k, pose)
            rb, Jb = ref.eval_functor(typ, k, pose)
            assert ra.tobytes() == rb.tobytes(), (typ, pose)
            assert Ja.tobytes() == Jb.tobytes(), (typ, pose)
            rp = ref.eval_functor_plain(typ, k, pose)
            np.testing.assert_allclose(rp, ra, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("it,skip", [(1, 1), (2, 1), (1, 7)])
def test_icp_pass_matches_reference(velo, oracle, ref, calib, params, it, skip):
    """velo.h:806-894 verbatim vs oracle (brute force AND kd-tree): indices exact, normals bit exact, r/J equal."""
    rawM, rawS = small_scan(velo, 8), small_scan(velo, 7)
    ptsM, rsM, _ = oracle.segment(rawM, calib)
    ptsS, rsS, _ = oracle.segment(rawS, calib)
    ptsM, rsM = ptsM[: rsM[6]], rsM[:7]          # 6 query rings keeps the brute-force shim fast
    pose = velo.synth.pose_guess(8, 0)
    cr, neq_r = ref.icp_pass(ptsM, rsM, ptsS, rsS, pose, it, skip)
    for mode in (0, 1):
        co, neq_o, kept = oracle.icp_pass(ptsM, rsM, ptsS, rsS, pose, it, skip, params, mode)
        ck = co[co["kept"] == 1]
        assert kept == len(cr) == len(ck) and kept > (50 if it == 1 else 5)
        for f in ("src_ring", "src_idx", "np_s_i", "np_i", "np_s_j", "np_j", "np_k"):
            assert np.array_equal(ck[f], cr[f]), f
        assert ck["normal"].tobytes() == cr["normal"].tobytes()
        assert ck["v0"].tobytes() == cr["v0"].tobytes()
        assert ck["residual"].tobytes() == cr["residual"].tobytes()
        assert ck["jacobian"].tobytes() == cr["jacobian"].tobytes()
        np.testing.assert_allclose(neq_o[:58], neq_r[:58], rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("it", [1, 2])
def test_visual_blocks_match_reference(velo, oracle, ref, calib, params, frames, it):
    abi = velo.abi
    F = 600
    data = {}
    for f in (7, 8):
        _, pts, rs = frames[f]
        kpA, kpB, m = velo.synth.features(f, F)
        hd = np.zeros((2, 2, F), np.int32); kw = np.zeros((2, 2, F, 4), np.float32)
        for cam in (0, 1):
            rc, proj, valid = oracle.project(pts, rs, calib, cam)
            for s, kp in enumerate((kpA[cam], kpB[cam])):
                h, k = oracle.depth_assoc(valid, proj, rc, kp)
                hd[s, cam] = h; kw[s, cam, : len(k)] = k
        data[f] = (kpA, kpB, m, hd, kw)
    kpA7, _, _, hd7, kw7 = data[7]
    _, kpB8, m8, hd8, kw8 = data[8]
    MM = F
    matches = np.zeros((2, MM, 2), np.int32); nm = np.zeros(2, np.int32)
    lm_valid = np.zeros((2, MM), np.int32); lm_xyz = np.zeros((2, MM, 4), np.float32)
    rng = np.random.default_rng(5)
    for cam in (0, 1):
        idx = np.nonzero(m8[cam])[0]
        nm[cam] = len(idx); matches[cam, : len(idx), 0] = idx; matches[cam, : len(idx), 1] = idx
        for j in range(0, len(idx), 9):          # a few triangulated landmarks override the lidar depth (velo.h:634-644)
            lm_valid[cam, j] = 1; lm_xyz[cam, j] = [rng.normal() * 3, rng.normal(), 8 + rng.uniform() * 10, 1]
    pose = velo.synth.pose_guess(8, 3) if it == 2 else velo.synth.pose_guess(8, 0)
    for lmv, lmx in ((None, None), (lm_valid, lm_xyz)):
        bo, neq_o = oracle.visual(kpB8, kpA7, hd8[1], hd7[0], kw8[1], kw7[0], nm, matches, calib, params, pose, it, lmv, lmx)
        br, neq_r = ref.visual(oracle, kpB8, kpA7, hd8[1], hd7[0], kw8[1], kw7[0], nm, matches, calib, params, pose, it, lmv, lmx)
        assert len(bo) == len(br) > 200
        for f in ("cam", "match", "type", "n_res"):
            assert np.array_equal(bo[f], br[f]), f
        assert set(np.unique(bo["type"])) >= {abi.RES_3D3D, abi.RES_3D2D, abi.RES_2D3D, abi.RES_2D2D}
        assert bo["residual"].tobytes() == br["residual"].tobytes()
        assert bo["jacobian"].tobytes() == br["jacobian"].tobytes()
        np.testing.assert_allclose(neq_o[:58], neq_r[:58], rtol=1e-12, atol=1e-300)


def _descriptors(rng, n, base=None, flips=0):
    d = rng.integers(0, 256, size=(n, 64), dtype=np.uint8) if base is None else base.copy()
    if base is not None:
        for i in range(n):
            for b in rng.integers(0, 512, flips):
                d[i, b // 8] ^= np.uint8(1 << (b % 8))
    return d


def test_match_hamming_matches_reference(oracle, ref):
    """velo.h:499-550 verbatim (BFMatcher stand-in: lowest index wins ties) vs the oracle: 64-byte FREAK-sized descriptors,
    a train set made of noisy copies (real matches), duplicates (ties) and clutter"""
    rng = np.random.default_rng(3)
    q = _descriptors(rng, 200)
    t = np.concatenate([_descriptors(rng, 120, q[:120], flips=12), q[5:9], q[5:9], _descriptors(rng, 150)])
    for tt in (t, t[:1], t[rng.permutation(len(t))]):
        po, bi, bd = oracle.match_hamming(q, tt)
        pr = ref.match_hamming(q, tt)
        assert np.array_equal(po, pr)
    assert len(po) > 50 and len(po) < 200


def test_triangulation_matches_reference(velo, oracle, ref, calib, params):
    """triangulatePoint (velo.h:1027-1130) compiled verbatim (ceres::Solve = the LM stand-in) vs the oracle,
    with and without initial guesses; landmarks with no, only 2-D, only 3-D and mixed observations"""
    import tri_data
    off3, obs3, off2, obs2, poses, truth = tri_data.make(1)
    a, it = oracle.triangulate(off3, obs3, off2, obs2, poses, calib, params)
    b, _ = oracle.triangulate(off3, obs3, off2, obs2, poses, calib, params, _ref=ref.lib)
    np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-6)
    n3, n2 = np.diff(off3), np.diff(off2)
    good = (n3 >= 1) & (n2 >= 3)
    assert good.sum() > 50 and np.abs(a[good] - truth[good]).max() < 0.25          # and it triangulates
    none = (n3 == 0) & (n2 == 0)
    assert none.any() and np.all(a[none] == [0, 0, 10])                            # untouched start (velo.h:1041)
    init = (truth + 0.3).astype(np.float32); has = (np.arange(len(truth)) % 2).astype(np.int32)
    a2, _ = oracle.triangulate(off3, obs3, off2, obs2, poses, calib, params, init, has)
    b2, _ = oracle.triangulate(off3, obs3, off2, obs2, poses, calib, params, init, has, _ref=ref.lib)
    np.testing.assert_allclose(a2, b2, rtol=1e-6, atol=1e-6)
