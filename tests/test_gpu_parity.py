"""-m gpu: every CUDA stage, called through the C ABI (libvelo_gpu.so), against the CPU oracle on identical inputs.
Bars (BASELINE.json north_star): indices / integer outputs bit-exact; f32 geometry bit-exact (same IEEE ops in the
same order); f64 residuals within 1e-5 relative; normal equations within 1e-4 relative."""
import numpy as np
import pytest
from conftest import small_scan

pytestmark = pytest.mark.gpu

RTOL_RES = 1e-5      # fp32/fp64 residual tolerance stated by north_star
RTOL_NEQ = 1e-4      # J^T J tolerance stated by north_star


@pytest.fixture(scope="module")
def ctx(velo, calib):
    prm = velo.api.default_params(max_slots=4, max_features=8192, max_matches=8192, num_cams=2)
    c = velo.api.Context(prm, calib)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx4(velo, oracle):
    """off-road rig: 4 cameras (BASELINE configs[3])"""
    P, Tr, w, h = velo.synth.calib_raw(1)
    cal = oracle.calib_from_kitti(P, Tr, w, h)
    prm = velo.api.default_params(max_slots=2, max_features=8192, max_matches=8192, num_cams=4)
    c = velo.api.Context(prm, cal)
    yield c
    c.close()


def test_calibration_matches_oracle(velo, oracle):
    for rig in (0, 1):
        P, Tr, w, h = velo.synth.calib_raw(rig)
        a = velo.api.calib_from_kitti(P, Tr, w, h)
        b = oracle.calib_from_kitti(P, Tr, w, h)
        assert bytes(a) == bytes(b)


@pytest.mark.parametrize("frame", [7, 123])
def test_ingest_bit_exact(velo, oracle, calib, ctx, frame):
    raw, n = velo.synth.scan(frame)
    ctx.scan_upload(0, raw)
    pts, rs = ctx.scan_download(0)
    opts, ors, onr = oracle.segment(raw, calib)
    assert ctx.scan_info(0) == (n, onr)
    assert np.array_equal(rs, ors)
    assert pts.tobytes() == opts.tobytes()


def test_ingest_ragged_and_empty(velo, oracle, calib, ctx):
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 31, 32, 33, 1000):
        raw = (rng.normal(size=(n, 4)) * 10).astype(np.float32)
        ctx.scan_upload(1, raw)
        opts, ors, onr = oracle.segment(raw, calib)
        if onr > ctx.prm.max_rings:          # random points cross the seam constantly: legitimately over capacity
            with pytest.raises(velo.api.VeloError):
                ctx.scan_download(1)
            continue
        pts, rs = ctx.scan_download(1)
        assert np.array_equal(rs, ors), n
        assert pts.tobytes() == opts.tobytes(), n


def test_ingest_too_many_rings_is_an_error(velo, ctx):
    n = 4000
    raw = np.zeros((n, 4), np.float32)
    raw[:, 0] = 5.0
    raw[:, 1] = np.where(np.arange(n) % 2 == 0, 1.0, -1.0)     # every point crosses the seam -> n rings
    ctx.scan_upload(1, raw)
    with pytest.raises(velo.api.VeloError) as e:
        ctx.scan_info(1)
    assert e.value.code == 4


@pytest.mark.parametrize("cam", [0, 1])
def test_project_bit_exact(velo, oracle, calib, ctx, cam):
    raw, n = velo.synth.scan(7)
    ctx.scan_upload(0, raw)
    ctx.project(0, cam)
    rc, proj, valid = ctx.project_download(0, cam)
    opts, ors, _ = oracle.segment(raw, calib)
    orc, oproj, ovalid = oracle.project(opts, ors, calib, cam)
    assert np.array_equal(rc, orc)
    assert proj.tobytes() == oproj.tobytes()
    assert valid.tobytes() == ovalid.tobytes()
    assert rc.sum() > 10000


def test_project_occlusion_stress(velo, oracle, calib, ctx):
    """noisy depth so that pops / skips / re-pushes happen constantly (velo.h:351-365)"""
    raw, n = velo.synth.scan(9)
    rng = np.random.default_rng(3)
    raw = raw.copy()
    raw[:, :3] *= (1.0 + rng.normal(size=(n, 1)).astype(np.float32) * 0.05)
    ctx.scan_upload(0, raw)
    opts, ors, _ = oracle.segment(raw, calib)
    for cam in (0, 1):
        ctx.project(0, cam)
        rc, proj, valid = ctx.project_download(0, cam)
        orc, oproj, ovalid = oracle.project(opts, ors, calib, cam)
        assert np.array_equal(rc, orc)
        assert proj.tobytes() == oproj.tobytes() and valid.tobytes() == ovalid.tobytes()


@pytest.mark.parametrize("F", [0, 1, 2000, 8000])
def test_depth_assoc_bit_exact(velo, oracle, calib, ctx, F):
    raw, n = velo.synth.scan(7)
    ctx.scan_upload(0, raw)
    opts, ors, _ = oracle.segment(raw, calib)
    kpA, kpB, _ = velo.synth.features(7, max(F, 1))
    for cam in (0, 1):
        ctx.project(0, cam)
        orc, oproj, ovalid = oracle.project(opts, ors, calib, cam)
        for s, kp in enumerate((kpA[cam][:F], kpB[cam][:F])):
            hd, kpwd = ctx.depth_assoc(0, cam, kp, s)
            ohd, okpwd = oracle.depth_assoc(ovalid, oproj, orc, kp)
            assert np.array_equal(hd, ohd)
            assert kpwd.tobytes() == okpwd.tobytes()


def test_depth_assoc_abs_truncates_variant(velo, oracle, calib):
    """hazard H1: the alternative binding of abs() (int abs(int)) is a runtime switch on both sides"""
    prm = velo.api.default_params(max_slots=1, abs_truncates=1)
    c = velo.api.Context(prm, calib)
    try:
        raw, n = velo.synth.scan(7)
        c.scan_upload(0, raw)
        c.project(0, 0)
        opts, ors, _ = oracle.segment(raw, calib)
        orc, oproj, ovalid = oracle.project(opts, ors, calib, 0)
        kp = velo.synth.features(7, 2000)[0][0]
        hd, kpwd = c.depth_assoc(0, 0, kp, 0)
        ohd, okpwd = oracle.depth_assoc(ovalid, oproj, orc, kp, abs_truncates=1)
        assert np.array_equal(hd, ohd) and kpwd.tobytes() == okpwd.tobytes()
        ohd0, _ = oracle.depth_assoc(ovalid, oproj, orc, kp, abs_truncates=0)
        assert (ohd >= 0).sum() >= (ohd0 >= 0).sum()
    finally:
        c.close()


def _check_icp(velo, oracle, params, ctx, ptsM, rsM, ptsS, rsS, pose, it, skip, mode):
    corr, neq, kept = ctx.icp_pass(1, 0, pose, it, skip)
    ocorr, oneq, okept = oracle.icp_pass(ptsM, rsM, ptsS, rsS, pose, it, skip, params, mode)
    assert len(corr) == len(ocorr)
    for f in ("src_ring", "src_idx", "kept", "np_s_i", "np_i", "np_s_j", "np_j", "np_k"):
        bad = np.nonzero(corr[f] != ocorr[f])[0]
        assert len(bad) == 0, (f, len(bad), corr[bad[:3]], ocorr[bad[:3]])
    assert kept == okept
    k = ocorr["kept"] == 1
    assert corr["normal"][k].tobytes() == ocorr["normal"][k].tobytes()
    assert corr["v0"][k].tobytes() == ocorr["v0"][k].tobytes()
    np.testing.assert_allclose(corr["residual"][k], ocorr["residual"][k], rtol=RTOL_RES, atol=1e-9)
    np.testing.assert_allclose(corr["jacobian"][k], ocorr["jacobian"][k], rtol=RTOL_RES, atol=1e-9)
    scale = np.abs(oneq[:56]).max()
    np.testing.assert_allclose(neq[:56], oneq[:56], rtol=RTOL_NEQ, atol=RTOL_NEQ * 1e-6 * max(scale, 1e-30))
    assert neq[56] == oneq[56] and neq[58] == oneq[58]
    return kept


@pytest.mark.parametrize("it,skip", [(1, 1), (2, 1), (1, 5), (1, 200)])
def test_icp_small_vs_bruteforce_truth(velo, oracle, calib, params, ctx, it, skip):
    """thinned scans: the oracle's brute-force NN (truth; ties -> lower index) defines the indices"""
    rawM, rawS = small_scan(velo, 8), small_scan(velo, 7)
    ctx.scan_upload(1, rawM); ctx.scan_upload(0, rawS)
    ptsM, rsM, _ = oracle.segment(rawM, calib)
    ptsS, rsS, _ = oracle.segment(rawS, calib)
    pose = velo.synth.pose_guess(8, 0)
    kept = _check_icp(velo, oracle, params, ctx, ptsM, rsM, ptsS, rsS, pose, it, skip, 0)
    if skip == 1:
        assert kept > 100


@pytest.mark.parametrize("it,pass_idx", [(1, 0), (2, 3)])
def test_icp_full_scan_indices_exact(velo, oracle, calib, params, ctx, it, pass_idx):
    """BASELINE configs[2]: 120k vs 120k, icp_skip = 1.  Oracle in kd-tree mode (itself checked against brute force)."""
    rawM, nM = velo.synth.scan(8)
    rawS, nS = velo.synth.scan(7)
    ctx.scan_upload(1, rawM); ctx.scan_upload(0, rawS)
    ptsM, rsM, _ = oracle.segment(rawM, calib)
    ptsS, rsS, _ = oracle.segment(rawS, calib)
    pose = velo.synth.pose_guess(8, pass_idx)
    kept = _check_icp(velo, oracle, params, ctx, ptsM, rsM, ptsS, rsS, pose, it, 1, 1)
    assert kept > (50000 if it == 1 else 5000)


def test_icp_degenerate_inputs(velo, oracle, calib, params, ctx):
    """duplicated points (exact distance ties), far-away pose (no correspondences), identity pose (small-angle branch)"""
    rawS = small_scan(velo, 7)
    rawM = rawS.copy()
    rawS = np.concatenate([rawS, rawS[-200:]])            # duplicates inside the last ring: ties broken by lower index
    ctx.scan_upload(1, rawM); ctx.scan_upload(0, rawS)
    ptsM, rsM, _ = oracle.segment(rawM, calib)
    ptsS, rsS, _ = oracle.segment(rawS, calib)
    for pose in ([0, 0, 0, 0, 0, 0], [1e-9, 0, -1e-9, 0.01, 0, 0.02], [0, 0, 0, 0, 0, 500.0], [0.3, 0.2, -0.4, 1, -2, 3]):
        _check_icp(velo, oracle, params, ctx, ptsM, rsM, ptsS, rsS, np.array(pose, np.float64), 1, 3, 0)


def _visual_inputs(velo, oracle, calib, ctx, F, ncam, frames=(7, 8)):
    out = {}
    for slot, f in enumerate(frames):
        raw, n = velo.synth.scan(f)
        ctx.scan_upload(slot, raw)
        pts, rs, _ = oracle.segment(raw, calib)
        kpA, kpB, m = velo.synth.features(f, F, ncam, 1 if ncam == 4 else 0)
        hd = np.zeros((2, ncam, F), np.int32); kw = np.zeros((2, ncam, F, 4), np.float32)
        for cam in range(ncam):
            ctx.project(slot, cam)
            rc, proj, valid = oracle.project(pts, rs, calib, cam)
            for s, kp in enumerate((kpA[cam], kpB[cam])):
                h, k = oracle.depth_assoc(valid, proj, rc, kp)
                gh, gk = ctx.depth_assoc(slot, cam, kp, s)
                assert np.array_equal(h, gh) and k.tobytes() == gk.tobytes()
                hd[s, cam] = h; kw[s, cam, : len(k)] = k
        out[f] = (kpA, kpB, m, hd, kw)
    return out


@pytest.mark.parametrize("it", [1, 2])
def test_visual_residuals_parity(velo, oracle, calib, params, ctx, it):
    F = 2000
    d = _visual_inputs(velo, oracle, calib, ctx, F, 2)
    kpA7, _, _, hd7, kw7 = d[7]
    _, kpB8, m8, hd8, kw8 = d[8]
    matches = np.zeros((2, F, 2), np.int32); nm = np.zeros(2, np.int32); cat = []
    lm_valid = np.zeros((2, F), np.int32); lm_xyz = np.zeros((2, F, 4), np.float32)
    rng = np.random.default_rng(5)
    for cam in (0, 1):
        idx = np.nonzero(m8[cam])[0]
        nm[cam] = len(idx); matches[cam, : len(idx), 0] = idx; matches[cam, : len(idx), 1] = idx
        cat.append(matches[cam, : len(idx)])
        for j in range(0, len(idx), 9):
            lm_valid[cam, j] = 1; lm_xyz[cam, j] = [rng.normal() * 3, rng.normal(), 8 + rng.uniform() * 10, 1]
    cat = np.concatenate(cat)
    pose = velo.synth.pose_guess(8, 3 if it == 2 else 0)
    for use_lm in (False, True):
        lmv = lm_valid if use_lm else None
        lmx = lm_xyz if use_lm else None
        ob, oneq = oracle.visual(kpB8, kpA7, hd8[1], hd7[0], kw8[1], kw7[0], nm, matches, calib, params, pose, it, lmv, lmx)
        glv = np.concatenate([lm_valid[c, : nm[c]] for c in (0, 1)]) if use_lm else None
        glx = np.concatenate([lm_xyz[c, : nm[c]] for c in (0, 1)]) if use_lm else None
        gb, gneq = ctx.visual_residuals(1, 1, 0, 0, nm, cat, pose, it, glv, glx)
        assert len(gb) == len(ob) > 1000
        for f in ("cam", "match", "type", "n_res"):
            assert np.array_equal(gb[f], ob[f]), f
        np.testing.assert_allclose(gb["residual"], ob["residual"], rtol=RTOL_RES, atol=1e-12)
        np.testing.assert_allclose(gb["jacobian"], ob["jacobian"], rtol=RTOL_RES, atol=1e-12)
        scale = np.abs(oneq[:56]).max()
        np.testing.assert_allclose(gneq[:56], oneq[:56], rtol=RTOL_NEQ, atol=RTOL_NEQ * 1e-6 * scale)
        assert gneq[56] == oneq[56] and gneq[57] == oneq[57]


def test_offroad_rig_four_cameras(velo, oracle, ctx4):
    """BASELINE configs[3]: 4 cameras, 8k features per image: projection + association parity on every camera"""
    cal = ctx4.cal
    raw, n = velo.synth.scan(11)
    ctx4.scan_upload(0, raw)
    pts, rs, _ = oracle.segment(raw, cal)
    kpA, kpB, _ = velo.synth.features(11, 8000, 4, 1)
    ctx4.project(0, 0)
    for cam in range(4):
        rc, proj, valid = ctx4.project_download(0, cam)
        orc, oproj, ovalid = oracle.project(pts, rs, cal, cam)
        assert np.array_equal(rc, orc) and proj.tobytes() == oproj.tobytes() and valid.tobytes() == ovalid.tobytes()
        hd, kpwd = ctx4.depth_assoc(0, cam, kpA[cam], 0)
        ohd, okpwd = oracle.depth_assoc(ovalid, oproj, orc, kpA[cam])
        assert np.array_equal(hd, ohd) and kpwd.tobytes() == okpwd.tobytes()
        assert (hd >= 0).sum() > 2000


def test_batched_path_equals_single_frame_path_and_oracle(velo, oracle, calib):
    """the throughput path (batch_upload / batch_run / batch_download) on 3 scans = 2 frame pairs, full schedule
    (2 f2f iterations x 3 ICP passes, icp_skip 20 to keep the CPU side quick) vs the oracle's timed-baseline driver."""
    prm = velo.api.default_params(max_slots=3, max_features=1000, max_matches=1000, icp_skip=20)
    c = velo.api.Context(prm, calib)
    try:
        b = velo.synth.Batch(20, 3, prm)
        c.batch_upload(0, b)
        c.batch_run(0, 3)
        icp = np.zeros((3, b.n_passes, velo.abi.NEQ_STRIDE)); vis = np.zeros((3, b.n_vis, velo.abi.NEQ_STRIDE))
        hd = np.zeros((3, 2, prm.num_cams, prm.max_features), np.int32); nh = np.zeros((3, 2, prm.num_cams), np.int32)
        c.batch_download(0, 3, icp, vis, hd, nh)
        sec, oicp, ovis = oracle.bench_frames(b, prm, calib, threads=2, want_out=True)
        assert np.all(icp[0] == 0) and np.all(vis[0] == 0)
        for t in (1, 2):
            for p in range(b.n_passes):
                sc = np.abs(oicp[t, p, :56]).max()
                np.testing.assert_allclose(icp[t, p, :56], oicp[t, p, :56], rtol=RTOL_NEQ, atol=RTOL_NEQ * 1e-6 * sc)
                assert icp[t, p, 56] == oicp[t, p, 56] and icp[t, p, 58] == oicp[t, p, 58]
            for it in range(b.n_vis):
                sc = np.abs(ovis[t, it, :56]).max()
                np.testing.assert_allclose(vis[t, it, :56], ovis[t, it, :56], rtol=RTOL_NEQ, atol=RTOL_NEQ * 1e-6 * sc)
                assert vis[t, it, 56] == ovis[t, it, 56]
        # association results of the batch equal the oracle's
        kw = c.batch_download_kpwd(0, 3)
        for t in range(3):
            pts, rs, _ = oracle.segment(b.scans[t, : b.n_points[t]], calib)
            for cam in range(prm.num_cams):
                rc, proj, valid = oracle.project(pts, rs, calib, cam)
                for s in range(2):
                    ohd, okw = oracle.depth_assoc(valid, proj, rc, b.kp[t, s, cam])
                    assert np.array_equal(hd[t, s, cam], ohd)
                    assert nh[t, s, cam] == (ohd >= 0).sum()
                    assert kw[t, s, cam, : len(okw)].tobytes() == okw.tobytes()      # keypoints_with_depth of the batch (velo.h:479)
        npnt, nr, ptot, st = c.batch_counts(0, 3)
        assert np.array_equal(npnt, b.n_points) and np.all(nr == 64) and np.all(st == 0) and np.all(ptot > 10000)
        # the one-call pipelined path (chunked upload overlapping compute) gives byte-identical results
        for chunk in (1, 2, 0):
            icp2 = np.zeros_like(icp); vis2 = np.zeros_like(vis); hd2 = np.zeros_like(hd); nh2 = np.zeros_like(nh)
            kw2 = np.zeros_like(kw)
            c.batch_frontend(0, b, chunk, icp2, vis2, hd2, nh2, kpwd=kw2 if chunk != 2 else None)
            assert icp2.tobytes() == icp.tobytes() and vis2.tobytes() == vis.tobytes() and hd2.tobytes() == hd.tobytes() and nh2.tobytes() == nh.tobytes()
            for t, sidx, cam in np.ndindex(3, 2, prm.num_cams):
                n = nh[t, sidx, cam]
                assert chunk == 2 or kw2[t, sidx, cam, :n].tobytes() == kw[t, sidx, cam, :n].tobytes()
    finally:
        c.close()


def test_normal_equations_do_not_depend_on_the_launch_shape(velo, calib):
    """sums are kept per run of 64 queries and added in run order, so the CTAs per frame pair (and hence which warp processed
    which run) must not change a single bit of the result"""
    outs = []
    for ctas in (0, 1, 5, 32):
        prm = velo.api.default_params(max_slots=3, max_features=200, max_matches=200, icp_skip=3)
        prm.ctas_per_icp_unit = ctas
        c = velo.api.Context(prm, calib)
        try:
            b = velo.synth.Batch(31, 3, prm)
            c.batch_upload(0, b)
            c.batch_run(0, 3)
            icp = np.zeros((3, b.n_passes, velo.abi.NEQ_STRIDE))
            c.batch_download(0, 3, icp, None, None, None)
            outs.append(icp)
        finally:
            c.close()
    assert outs[0][1:, :, 58].min() > 30000            # many runs per frame pair
    for o in outs[1:]:
        assert o.tobytes() == outs[0].tobytes()


def test_round_trip_properties_at_full_size(velo, calib):
    """size-independent properties at BASELINE size (no oracle): ingest is a permutation of a rigid transform,
    projection survivors are a subsequence inside the FOV, has_depth is a stable enumeration, JtJ is PSD."""
    prm = velo.api.default_params(max_slots=2, max_features=2000)
    c = velo.api.Context(prm, calib)
    try:
        rawM, nM = velo.synth.scan(301)
        rawS, nS = velo.synth.scan(300)
        c.scan_upload(1, rawM); c.scan_upload(0, rawS)
        pts, rs = c.scan_download(1)
        assert len(pts) == nM and rs[0] == 0 and rs[-1] == nM and np.all(np.diff(rs) > 0)
        # rigid transform preserves pairwise distances of the multiset: compare sorted norms about the velodyne origin
        T = np.array(list(calib.velo_to_cam), np.float64).reshape(4, 4)
        back = (pts[:, :3].astype(np.float64) - T[:3, 3]) @ T[:3, :3]
        assert np.allclose(np.sort(np.linalg.norm(back, axis=1)), np.sort(np.linalg.norm(rawM[:, :3].astype(np.float64), axis=1)), atol=2e-3)
        c.project(1, 0)
        rc, proj, valid = c.project_download(1, 0)
        assert np.all(proj[:, 0] >= calib.min_x[0]) and np.all(proj[:, 0] < calib.max_x[0])
        assert np.all(proj[:, 1] >= calib.min_y[0]) and np.all(proj[:, 1] < calib.max_y[0])
        kp = velo.synth.features(301, 2000)[0][0]
        hd, kpwd = c.depth_assoc(1, 0, kp, 0)
        hits = hd[hd >= 0]
        assert np.array_equal(hits, np.arange(len(hits))) and len(kpwd) == len(hits)
        corr, neq, kept = c.icp_pass(1, 0, velo.synth.pose_guess(301, 0), 1, 1)
        H = np.zeros((6, 6)); H[np.triu_indices(6)] = neq[:21]; H = H + H.T - np.diag(np.diag(H))
        assert np.linalg.eigvalsh(H).min() > -1e-9 * np.abs(H).max()
        k = corr["kept"] == 1
        assert kept == k.sum() and np.all(corr["np_s_i"][k] != corr["np_s_j"][k])
        assert np.allclose(np.linalg.norm(corr["normal"][k], axis=1), 1.0, atol=1e-5)
        # determinism: a second run is byte-identical
        corr2, neq2, _ = c.icp_pass(1, 0, velo.synth.pose_guess(301, 0), 1, 1)
        assert corr.tobytes() == corr2.tobytes() and neq.tobytes() == neq2.tobytes()
    finally:
        c.close()


def test_cuda_path_against_golden_fixture(velo, oracle):
    """tests/golden/velo_golden.npz holds outputs of the reference's own source lines: the CUDA path must reproduce them
    without any oracle in the loop (indices + f32 geometry bit-exact, f64 residuals 1e-5, normal equations 1e-4)."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "velo_golden.npz"))
    cal = velo.api.calib_from_kitti(gold["P"], gold["Tr"], int(gold["wh"][0]), int(gold["wh"][1]))
    F = gold["kpA60"].shape[1]
    prm = velo.api.default_params(max_slots=2, max_features=F, max_matches=F)
    c = velo.api.Context(prm, cal)
    try:
        for slot, f in ((0, 60), (1, 61)):
            c.scan_upload(slot, gold[f"raw{f}"])
            pts, rs = c.scan_download(slot)
            assert np.array_equal(rs, gold[f"rs{f}"]) and pts.tobytes() == gold[f"pts{f}"].tobytes()
            for cam in (0, 1):
                c.project(slot, cam)
                rc, proj, valid = c.project_download(slot, cam)
                assert np.array_equal(rc, gold[f"rc{f}_{cam}"])
                assert proj.tobytes() == gold[f"proj{f}_{cam}"].tobytes() and valid.tobytes() == gold[f"valid{f}_{cam}"].tobytes()
                for s, key in enumerate(("kpA", "kpB")):
                    hd, kw = c.depth_assoc(slot, cam, gold[f"{key}{f}"][cam], s)
                    assert np.array_equal(hd, gold[f"hd{f}"][s, cam]) and kw.tobytes() == gold[f"kw{f}"][s, cam, : len(kw)].tobytes()
        for it, skip in ((1, 1), (2, 1), (1, 4)):
            corr, neq, kept = c.icp_pass(1, 0, gold[f"icp_pose_{it}_{skip}"], it, skip)
            g = gold[f"icp_corr_{it}_{skip}"]
            ck = corr[corr["kept"] == 1]
            assert len(ck) == len(g) == kept
            for fld in ("src_ring", "src_idx", "np_s_i", "np_i", "np_s_j", "np_j", "np_k"):
                assert np.array_equal(ck[fld], g[fld]), fld
            assert ck["normal"].tobytes() == g["normal"].tobytes() and ck["v0"].tobytes() == g["v0"].tobytes()
            np.testing.assert_allclose(ck["residual"], g["residual"], rtol=RTOL_RES, atol=1e-9)
            np.testing.assert_allclose(ck["jacobian"], g["jacobian"], rtol=RTOL_RES, atol=1e-9)
            gn = gold[f"icp_neq_{it}_{skip}"]
            np.testing.assert_allclose(neq[:56], gn[:56], rtol=RTOL_NEQ, atol=RTOL_NEQ * 1e-6 * np.abs(gn[:56]).max())
        nm = gold["vis_nm"]
        cat = np.concatenate([gold["vis_matches"][cam, : nm[cam]] for cam in (0, 1)])
        for it in (1, 2):
            b, neq = c.visual_residuals(1, 1, 0, 0, nm, cat, gold[f"vis_pose_{it}"], it)
            g = gold[f"vis_blocks_{it}"]
            assert len(b) == len(g)
            for fld in ("cam", "match", "type", "n_res"):
                assert np.array_equal(b[fld], g[fld]), fld
            np.testing.assert_allclose(b["residual"], g["residual"], rtol=RTOL_RES, atol=1e-12)
            np.testing.assert_allclose(b["jacobian"], g["jacobian"], rtol=RTOL_RES, atol=1e-12)
            gn = gold[f"vis_neq_{it}"]
            np.testing.assert_allclose(neq[:56], gn[:56], rtol=RTOL_NEQ, atol=RTOL_NEQ * 1e-6 * np.abs(gn[:56]).max())
    finally:
        c.close()


def test_cpp_dropin_header(velo, oracle, calib, tmp_path):
    """include/velo_dropin.hpp (the reference's C++ signatures) compiled with g++ against libvelo_gpu.so and driven like main.cpp
    drives the reference, on a KITTI-layout tree written here (<root>/00/calib.txt, <root>/00/velodyne/%06d.bin; kitti.h:59-152):
    loadCalibration -> ScansLRU::get -> ScanData(dataset, frame) -> projectLidarToCamera -> featureDepthAssociation ->
    icpCorrespondences -> frameToFrame(velo.h:598-614 parameter list), then LRU eviction / slot reuse and recycled host addresses.
    Every output is compared with the oracle."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "vision-enhanced-lidar-odometry_b200")
    exe = tmp_path / "dropin_main"
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "cpp", "dropin_main.cpp"),
                    "-L", pkg, "-lvelo_gpu", f"-Wl,-rpath,{pkg}", "-o", str(exe)], check=True)
    P, Tr, w, h = velo.synth.calib_raw(0)
    seq = tmp_path / "00"
    (seq / "velodyne").mkdir(parents=True)
    with open(seq / "calib.txt", "w") as f:                                    # the layout kitti.h:66-105 parses
        for c in range(4):
            f.write(f"P{c}: " + " ".join("%.9g" % v for v in P[12 * c: 12 * c + 12]) + "\n")
        f.write("Tr: " + " ".join("%.9g" % v for v in Tr) + "\n")
    FR0, F, ncam = 100, 1200, 2
    raws = [small_scan(velo, FR0 + k, range(8, 56), 2) for k in range(5)]
    for k, raw in enumerate(raws):
        raw.tofile(seq / "velodyne" / ("%06d.bin" % k))
    kpA0 = velo.synth.features(FR0, F)[0]
    _, kpB1, m1 = velo.synth.features(FR0 + 1, F)
    kpA0.tofile(tmp_path / "kpA0.bin"); kpB1.tofile(tmp_path / "kpB1.bin")
    prm = velo.api.default_params(max_slots=6, icp_skip=4, max_features=1500, max_matches=1500)     # what dropin_main.cpp sets
    MM, Fp = prm.max_matches, prm.max_features
    matches = np.zeros((ncam, MM, 2), np.int32); nm = np.zeros(ncam, np.int32)
    for cam in range(ncam):
        idx = np.nonzero(m1[cam])[0]
        nm[cam] = len(idx); matches[cam, : len(idx), 0] = idx; matches[cam, : len(idx), 1] = idx
    np.concatenate([matches.ravel(), nm]).astype(np.int32).tofile(tmp_path / "matches.bin")
    pose = velo.synth.pose_guess(FR0 + 1, 0)
    pose.tofile(tmp_path / "pose.bin")
    r = subprocess.run([str(exe), str(tmp_path), "00", str(w), str(h)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    seg = [oracle.segment(raw, calib)[:2] for raw in raws]
    pts0, rs0 = seg[0]; pts1, rs1 = seg[1]
    hd = {}; kw = {}
    for fr, kp_all in ((0, kpA0), (1, kpB1)):
        hd[fr] = np.full((ncam, Fp), 0, np.int32); kw[fr] = np.zeros((ncam, Fp, 4), np.float32)
        for cam in range(ncam):
            rc, proj, valid = oracle.project(*seg[fr], calib, cam)
            h_, k_ = oracle.depth_assoc(valid, proj, rc, kp_all[cam])
            hd[fr][cam, :F] = h_; kw[fr][cam, : len(k_)] = k_
            if fr == 1:
                assert np.array_equal(np.fromfile(tmp_path / f"out_rc{cam}.bin", np.int32), rc)
                assert np.fromfile(tmp_path / f"out_proj{cam}.bin", np.float32).tobytes() == proj.tobytes()
                assert np.array_equal(np.fromfile(tmp_path / f"out_hd{cam}.bin", np.int32), h_)
                assert np.fromfile(tmp_path / f"out_kpwd{cam}.bin", np.float32).tobytes() == k_.tobytes()
                assert (h_ >= 0).sum() > 100
    corr = np.fromfile(tmp_path / "out_corr.bin", velo.abi.ICP_CORR_DTYPE)
    ocorr, oneq, okept = oracle.icp_pass(pts1, rs1, pts0, rs0, pose, 1, 5, prm, 1)
    for f in ("src_ring", "src_idx", "kept", "np_s_i", "np_i", "np_s_j", "np_j", "np_k"):
        assert np.array_equal(corr[f], ocorr[f]), f
    neq = np.fromfile(tmp_path / "out_neq.bin", np.float64)
    np.testing.assert_allclose(neq[:56], oneq[:56], rtol=RTOL_NEQ, atol=RTOL_NEQ * 1e-6 * np.abs(oneq[:56]).max())
    # ---- frameToFrame adapter vs the oracle's frame_to_frame
    pad = lambda a: np.concatenate([a, np.zeros((a.shape[0], Fp - a.shape[1]) + a.shape[2:], a.dtype)], 1)
    vis = (pad(kpB1), pad(kpA0), hd[1], hd[0], kw[1], kw[0], nm, matches)
    guess = np.array([0, 0, 0, 0, 0, 1.0])
    ox, orep = oracle.frame_to_frame(pts1, rs1, pts0, rs0, calib, prm, guess, vis, 1, prm.icp_skip)
    gx = np.fromfile(tmp_path / "out_f2f_transform.bin", np.float64)
    rep = velo.abi.F2FReport.from_buffer_copy((tmp_path / "out_f2f_report.bin").read_bytes())
    assert rep.n_solves == orep["n_solves"] == prm.f2f_iterations * prm.icp_iterations
    assert list(rep.n_blocks)[: rep.n_solves] == orep["n_blocks"] and list(rep.lm_iterations)[: rep.n_solves] == orep["lm_iterations"]
    np.testing.assert_allclose(gx, ox, rtol=0, atol=1e-8)
    truth = velo.synth.pose(FR0 + 1)
    assert np.abs(gx[:3] - truth[:3]).max() < 2e-3 and np.abs(gx[3:] - truth[3:]).max() < 2e-2
    T = np.fromfile(tmp_path / "out_f2f_T.bin", np.float64).reshape(4, 4)
    from scipy.spatial.transform import Rotation
    np.testing.assert_allclose(T[:3, :3], Rotation.from_rotvec(gx[:3]).as_matrix(), atol=1e-14)
    np.testing.assert_allclose(T[:3, 3], gx[3:], atol=0); assert np.array_equal(T[3], [0, 0, 0, 1])
    # good_matches / residual_type = the block list of the LAST f2f iteration, assembled at the pose the first iteration ended with
    ob, _ = oracle.visual(vis[0], vis[1], vis[2], vis[3], vis[4], vis[5], nm, matches, calib, prm, orep["pose"][prm.icp_iterations - 1], prm.f2f_iterations)
    for cam in range(ncam):
        gm = np.fromfile(tmp_path / f"out_f2f_good{cam}.bin", np.int32).reshape(-1, 3)
        ex = ob[ob["cam"] == cam]
        assert len(gm) == len(ex) > 200
        assert np.array_equal(gm[:, 0], matches[cam, ex["match"], 0]) and np.array_equal(gm[:, 1], matches[cam, ex["match"], 1])
        assert np.array_equal(gm[:, 2], ex["type"])                              # RESIDUAL_* (velo.h:3-8) numbering == VELO_RES_*
    assert np.isfinite(np.fromfile(tmp_path / "out_f2f_lm_transform.bin", np.float64)).all()
    # ---- ScansLRU eviction / slot reuse, recycled host addresses
    rc, proj, _ = oracle.project(*seg[0], calib, 0)
    assert np.array_equal(np.fromfile(tmp_path / "out_rc_lru0.bin", np.int32), rc) and np.fromfile(tmp_path / "out_proj_lru0.bin", np.float32).tobytes() == proj.tobytes()
    for k in (4, 2):
        rc, proj, _ = oracle.project(*seg[k], calib, 1)
        assert np.array_equal(np.fromfile(tmp_path / f"out_rc_foreign{k}.bin", np.int32), rc), k
        assert np.fromfile(tmp_path / f"out_proj_foreign{k}.bin", np.float32).tobytes() == proj.tobytes(), k


def test_random_ragged_ring_clouds(velo, oracle, calib, params, ctx):
    """random ring-structured clouds installed with scan_upload_rings: empty and 1-point rings, grid-snapped coordinates
    (exact distance ties, duplicated x, z ties), identity / random poses — projection, association and ICP vs the oracle."""
    from test_properties import _rings
    rng = np.random.default_rng(2024)
    for case in range(24):
        quant = [None, 0.25, 0.05][case % 3]
        ptsS, rsS = _rings(rng, int(rng.integers(1, 10)), int(rng.integers(0, 70)), quant)
        ptsM, rsM = _rings(rng, int(rng.integers(1, 6)), int(rng.integers(0, 40)), quant)
        ctx.scan_upload_rings(0, ptsS, rsS); ctx.scan_upload_rings(1, ptsM, rsM)
        gp, grs = ctx.scan_download(1)
        assert gp.tobytes() == ptsM.tobytes() and np.array_equal(grs, rsM)
        for cam in (0, 1):
            ctx.project(1, cam)
            rc, proj, valid = ctx.project_download(1, cam)
            orc, oproj, ovalid = oracle.project(ptsM, rsM, calib, cam)
            assert np.array_equal(rc, orc) and proj.tobytes() == oproj.tobytes() and valid.tobytes() == ovalid.tobytes(), case
            F = int(rng.integers(0, 80))
            kp = np.stack([rng.uniform(calib.min_x[cam], calib.max_x[cam], F), rng.uniform(calib.min_y[cam], calib.max_y[cam], F)], 1).astype(np.float32)
            hd, kw = ctx.depth_assoc(1, cam, kp, 0)
            ohd, okw = oracle.depth_assoc(ovalid, oproj, orc, kp)
            assert np.array_equal(hd, ohd) and kw.tobytes() == okw.tobytes(), case
        pose = np.zeros(6) if quant else np.concatenate([rng.normal(0, 0.01, 3), rng.normal(0, 0.05, 3)])
        for it, skip in ((1, 1), (2, 1), (1, 3)):
            corr, neq, kept = ctx.icp_pass(1, 0, pose, it, skip)
            ocorr, oneq, okept = oracle.icp_pass(ptsM, rsM, ptsS, rsS, pose, it, skip, params, 0)
            assert len(corr) == len(ocorr) and kept == okept, case
            for f in ("src_ring", "src_idx", "kept", "np_s_i", "np_i", "np_s_j", "np_j", "np_k"):
                assert np.array_equal(corr[f], ocorr[f]), (case, f)
            k = ocorr["kept"] == 1
            assert corr["normal"][k].tobytes() == ocorr["normal"][k].tobytes()
            np.testing.assert_allclose(corr["residual"][k], ocorr["residual"][k], rtol=RTOL_RES, atol=1e-9)
            if kept:
                np.testing.assert_allclose(neq[:56], oneq[:56], rtol=RTOL_NEQ, atol=RTOL_NEQ * 1e-6 * np.abs(oneq[:56]).max() + 1e-300)


def test_icp_more_than_64_rings(velo, oracle, calib, params, ctx):
    """100 target rings / 90 source rings (two 64-bit words of ring masks: the word loops of the cell seeds and of the search, rings
    63 | 64 neighbours across the word boundary), dense in elevation so that a dozen rings lie inside the threshold; the fused
    multi-pass launch (cell seeds in pass 0, the previous pair as the bound afterwards) against the oracle's brute force."""
    from test_properties import _rings
    rng = np.random.default_rng(77)
    for quant in (None, 0.05):
        ptsS, rsS = _rings(rng, 100, 60, quant)
        ptsM, rsM = _rings(rng, 90, 40, quant)
        ctx.scan_upload_rings(0, ptsS, rsS); ctx.scan_upload_rings(1, ptsM, rsM)
        iters = [1, 1, 2]
        poses = np.stack([np.concatenate([rng.normal(0, 0.01, 3), rng.normal(0, 0.05, 3)]) for _ in iters])
        corr, neq = ctx.icp_passes(1, 0, poses, iters, 1)
        for p, it in enumerate(iters):
            ocorr, oneq, okept = oracle.icp_pass(ptsM, rsM, ptsS, rsS, poses[p], it, 1, params, 0)
            assert corr.shape[1] == len(ocorr)
            for f in ("src_ring", "src_idx", "kept", "np_s_i", "np_i", "np_s_j", "np_j", "np_k"):
                assert np.array_equal(corr[p][f], ocorr[f]), (quant, p, f)
            assert neq[p, 56] == okept and neq[p, 58] == len(ocorr)
            if okept:
                np.testing.assert_allclose(neq[p, :56], oneq[:56], rtol=RTOL_NEQ, atol=RTOL_NEQ * 1e-6 * np.abs(oneq[:56]).max() + 1e-300)
        assert (ocorr["np_s_i"] >= 64).any() and (ocorr["np_s_i"] < 64).any()


def test_depth_assoc_projection_larger_than_shared_memory(velo, oracle, calib, ctx):
    """more in-FOV points than the association kernel stages in shared memory (24 576): the global-memory path"""
    rng = np.random.default_rng(9)
    n_rings, L = 48, 900
    pts = []
    for s in range(n_rings):
        az = np.sort(rng.uniform(-0.65, 0.65, L)); r = 12.0 + rng.normal(0, 0.05, L)
        pts.append(np.stack([r * np.sin(az), np.full(L, -1.9 + 0.08 * s) + rng.normal(0, 0.01, L), r * np.cos(az), np.ones(L)], 1))
    pts = np.concatenate(pts).astype(np.float32)
    rs = (np.arange(n_rings + 1) * L).astype(np.int32)
    ctx.scan_upload_rings(0, pts, rs)
    ctx.project(0, 0)
    rc, proj, valid = ctx.project_download(0, 0)
    orc, oproj, ovalid = oracle.project(pts, rs, calib, 0)
    assert rc.sum() > 24576 and np.array_equal(rc, orc) and proj.tobytes() == oproj.tobytes()
    kp = np.stack([rng.uniform(-0.6, 0.6, 3000), rng.uniform(-0.16, 0.16, 3000)], 1).astype(np.float32)
    hd, kw = ctx.depth_assoc(0, 0, kp, 0)
    ohd, okw = oracle.depth_assoc(ovalid, oproj, orc, kp)
    assert np.array_equal(hd, ohd) and kw.tobytes() == okw.tobytes() and (hd >= 0).sum() > 1000


def test_frame_to_frame_device_solve(velo, oracle, calib, ctx):
    """velo_gpu_frame_to_frame (frozen blocks + device-resident LM in place of ceres::Solve) against the oracle's
    identical restatement, ICP only and ICP + visual terms, from the reference's initial guess (main.cpp:170).  Tolerance =
    solver tolerance (function_tolerance 1e-6): the two implementations take the same LM path, sums differ at 1e-13."""
    prm = ctx.prm
    rawM, rawS = small_scan(velo, 8, range(8, 56), 2), small_scan(velo, 7, range(8, 56), 2)
    ctx.scan_upload(1, rawM); ctx.scan_upload(0, rawS)
    ptsM, rsM, _ = oracle.segment(rawM, calib)
    ptsS, rsS, _ = oracle.segment(rawS, calib)
    truth = velo.synth.pose(8)
    guess = np.array([0, 0, 0, 0, 0, 1.0])
    skip = 4
    # ---- ICP only
    gx, grep = ctx.frame_to_frame(1, 1, 0, 0, guess, enable_icp=1, icp_skip=skip)
    ox, orep = oracle.frame_to_frame(ptsM, rsM, ptsS, rsS, calib, prm, guess, None, 1, skip)
    assert grep["n_solves"] == orep["n_solves"] == prm.f2f_iterations * prm.icp_iterations
    assert grep["n_blocks"] == orep["n_blocks"] and grep["reason"] == orep["reason"]
    assert grep["lm_iterations"] == orep["lm_iterations"] and grep["accepted_steps"] == orep["accepted_steps"]
    np.testing.assert_allclose(grep["pose"], orep["pose"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(grep["final_cost"], orep["final_cost"], rtol=1e-8)
    np.testing.assert_allclose(gx, ox, rtol=0, atol=1e-8)
    assert np.abs(gx[:3] - truth[:3]).max() < 1e-3 and np.abs(gx[3:] - truth[3:]).max() < 1e-2     # and it recovers the true motion
    assert grep["final_cost"][0] < grep["initial_cost"][0]
    # ---- ICP + visual terms (full frameToFrame schedule)
    F = 800
    data = {}
    for slot, f, raw in ((0, 7, rawS), (1, 8, rawM)):
        pts, rs, _ = oracle.segment(raw, calib)
        kpA, kpB, m = velo.synth.features(f, F)
        hd = np.zeros((2, 2, F), np.int32); kw = np.zeros((2, 2, F, 4), np.float32)
        for cam in (0, 1):
            ctx.project(slot, cam)
            rc, proj, valid = oracle.project(pts, rs, calib, cam)
            for s, kp in enumerate((kpA[cam], kpB[cam])):
                h, k = oracle.depth_assoc(valid, proj, rc, kp)
                ctx.depth_assoc(slot, cam, kp, s)
                hd[s, cam] = h; kw[s, cam, : len(k)] = k
        data[f] = (kpA, kpB, m, hd, kw)
    kpA7, _, _, hd7, kw7 = data[7]
    _, kpB8, m8, hd8, kw8 = data[8]
    MM = ctx.prm.max_matches
    matches = np.zeros((2, MM, 2), np.int32); nm = np.zeros(2, np.int32); cat = []
    for cam in (0, 1):
        idx = np.nonzero(m8[cam])[0]
        nm[cam] = len(idx); matches[cam, : len(idx), 0] = idx; matches[cam, : len(idx), 1] = idx
        cat.append(matches[cam, : len(idx)])
    Fp = ctx.prm.max_features
    pad = lambda a, shape: np.concatenate([a, np.zeros((a.shape[0], shape - a.shape[1]) + a.shape[2:], a.dtype)], 1)
    vis = (pad(kpB8, Fp), pad(kpA7, Fp), pad(hd8[1], Fp), pad(hd7[0], Fp), pad(kw8[1], Fp), pad(kw7[0], Fp), nm, matches)
    gx, grep = ctx.frame_to_frame(1, 1, 0, 0, guess, n_matches=nm, matches=np.concatenate(cat), enable_icp=1, icp_skip=skip)
    ox, orep = oracle.frame_to_frame(ptsM, rsM, ptsS, rsS, calib, prm, guess, vis, 1, skip)
    assert grep["n_blocks"] == orep["n_blocks"] and grep["reason"] == orep["reason"] and grep["lm_iterations"] == orep["lm_iterations"]
    np.testing.assert_allclose(grep["pose"], orep["pose"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(gx, ox, rtol=0, atol=1e-8)
    assert grep["n_blocks"][0] > orep["n_blocks"][0] - 1 and max(grep["n_blocks"]) > 8000
    # ---- visual terms only (the shipped configuration: ENABLE_ICP commented out, main.cpp:43)
    gx, grep = ctx.frame_to_frame(1, 1, 0, 0, guess, n_matches=nm, matches=np.concatenate(cat), enable_icp=0, icp_skip=skip)
    ox, orep = oracle.frame_to_frame(ptsM, rsM, ptsS, rsS, calib, prm, guess, vis, 0, skip)
    assert grep["n_solves"] == orep["n_solves"] == prm.f2f_iterations and grep["n_blocks"] == orep["n_blocks"]
    np.testing.assert_allclose(gx, ox, rtol=0, atol=1e-7)
    assert np.abs(gx[3:] - truth[3:]).max() < 0.05


def test_match_hamming_bit_exact(velo, oracle, ctx):
    """matchFeatures (velo.h:499-550) — brute-force Hamming 1-NN + min-distance filter, integer exact, ties -> lower index"""
    from test_oracle_vs_ref import _descriptors
    rng = np.random.default_rng(11)
    q = _descriptors(rng, 3000)
    t = np.concatenate([_descriptors(rng, 1800, q[:1800], flips=14), q[7:12], q[7:12], _descriptors(rng, 1200)])
    t = t[rng.permutation(len(t))]
    for qq, tt in ((q, t), (q[:1], t), (q, t[:1]), (q[:0], t), (q, t[:0]), (q[:129], t[:257])):
        gp, gi, gd = ctx.match_hamming(qq, tt)
        op, oi, od = oracle.match_hamming(qq, tt)
        assert np.array_equal(gp, op)
        if len(tt):
            assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    gp, gi, gd = ctx.match_hamming(q, t)
    assert 1500 < len(gp) < 3000
    gp16, _, _ = ctx.match_hamming(q[:, :16].copy(), t[:, :16].copy())          # shorter descriptors (desc_bytes = 16)
    op16, _, _ = oracle.match_hamming(q[:, :16].copy(), t[:, :16].copy())
    assert np.array_equal(gp16, op16)


def test_batched_triangulation(velo, oracle, calib, params, ctx):
    """velo_gpu_triangulate (one thread per landmark, LM on 3 parameters) vs the oracle's restatement of
    triangulatePoint (velo.h:1027-1130); "to solver tolerance" (function_tolerance 1e-6) on the float outputs."""
    import tri_data
    off3, obs3, off2, obs2, poses, truth = tri_data.make(5, L=4000, n_frames=8)
    g, git = ctx.triangulate(off3, obs3, off2, obs2, poses)
    o, oit = oracle.triangulate(off3, obs3, off2, obs2, poses, calib, params)
    np.testing.assert_allclose(g, o, rtol=2e-6, atol=2e-6)
    assert (git == oit).mean() > 0.999
    n3, n2 = np.diff(off3), np.diff(off2)
    good = (n3 >= 1) & (n2 >= 3)
    assert good.sum() > 500 and np.abs(g[good] - truth[good]).max() < 0.3
    init = (truth + 0.3).astype(np.float32); has = (np.arange(len(truth)) % 2).astype(np.int32)
    g2, _ = ctx.triangulate(off3, obs3, off2, obs2, poses, init, has)
    o2, _ = oracle.triangulate(off3, obs3, off2, obs2, poses, calib, params, init, has)
    np.testing.assert_allclose(g2, o2, rtol=2e-6, atol=2e-6)
    e, _ = ctx.triangulate(np.zeros(1, np.int32), obs3[:0], np.zeros(1, np.int32), obs2[:0], poses)
    assert len(e) == 0


# ------------------------------------------------------------------------------------------------ round 2: the benchmarked code path
@pytest.mark.parametrize("spread", ["tight", "spec"])
def test_fused_multipass_kernel_every_pass_vs_oracle(velo, oracle, calib, params, ctx, spread):
    """The code path bench.py times: ONE launch of k_icp_pass for the 2 x 3 passes of a frame pair (velo.h:616,800), pass p+1
    seeded by the correspondences of pass p, icp_skip = 1 on full scans (120k x 120k).  The records of EVERY pass are compared
    with the oracle at that pass's pose: indices bit-exact, normals / plane points bit-exact, residual + Jacobian 1e-5,
    normal equations 1e-4.  "spec" = SURVEY 8(d) poses: pass 0 from the reference's start (0,0,0,0,0,1) (main.cpp:170)."""
    rawM, _ = velo.synth.scan(8)
    rawS, _ = velo.synth.scan(7)
    ctx.scan_upload(1, rawM); ctx.scan_upload(0, rawS)
    ptsM, rsM, _ = oracle.segment(rawM, calib)
    ptsS, rsS, _ = oracle.segment(rawS, calib)
    iters = [p // params.icp_iterations + 1 for p in range(params.f2f_iterations * params.icp_iterations)]
    poses = np.stack([velo.synth.pose_guess(8, p, spread=spread) for p in range(len(iters))])
    corr, neq = ctx.icp_passes(1, 0, poses, iters, 1)
    assert corr.shape[0] == len(iters) == 6
    kept = []
    for p, it in enumerate(iters):
        ocorr, oneq, okept = oracle.icp_pass(ptsM, rsM, ptsS, rsS, poses[p], it, 1, params, 1)
        assert corr.shape[1] == len(ocorr)
        for f in ("src_ring", "src_idx", "kept", "np_s_i", "np_i", "np_s_j", "np_j", "np_k"):
            bad = np.nonzero(corr[p][f] != ocorr[f])[0]
            assert len(bad) == 0, (p, f, len(bad), corr[p][bad[:3]], ocorr[bad[:3]])
        k = ocorr["kept"] == 1
        assert corr[p]["normal"][k].tobytes() == ocorr["normal"][k].tobytes() and corr[p]["v0"][k].tobytes() == ocorr["v0"][k].tobytes()
        np.testing.assert_allclose(corr[p]["residual"][k], ocorr["residual"][k], rtol=RTOL_RES, atol=1e-9)
        np.testing.assert_allclose(corr[p]["jacobian"][k], ocorr["jacobian"][k], rtol=RTOL_RES, atol=1e-9)
        sc = np.abs(oneq[:56]).max()
        np.testing.assert_allclose(neq[p, :56], oneq[:56], rtol=RTOL_NEQ, atol=RTOL_NEQ * 1e-6 * sc)
        assert neq[p, 56] == oneq[56] == okept and neq[p, 58] == oneq[58]
        kept.append(okept)
        # and the single-pass entry point gives the same bits as the fused launch
        if p in (0, 4):
            c1, n1, _ = ctx.icp_pass(1, 0, poses[p], it, 1)
            assert c1.tobytes() == corr[p].tobytes() and n1[:59].tobytes() == neq[p, :59].tobytes()
    assert kept[0] > (50000 if spread == "tight" else 10000) and kept[-1] > 5000


def test_bench_configuration_three_full_scans_vs_oracle(velo, oracle, calib):
    """bench.py's configuration on 3 full scans (2 frame pairs): icp_skip = 1, 6 passes, 2000 features, batch_run and the one-call
    batch_frontend vs the oracle's timed-baseline driver: block / query counts equal, H / g / cost within 1e-4."""
    import os
    prm = velo.api.default_params(max_slots=3, max_points=131072, max_rings=64, max_features=2000, max_matches=2000, icp_skip=1)
    c = velo.api.Context(prm, calib)
    try:
        b = velo.synth.Batch(1000, 3, prm)
        icp = np.zeros((3, b.n_passes, velo.abi.NEQ_STRIDE)); vis = np.zeros((3, b.n_vis, velo.abi.NEQ_STRIDE))
        hd = np.zeros((3, 2, prm.num_cams, prm.max_features), np.int32); nh = np.zeros((3, 2, prm.num_cams), np.int32)
        c.batch_frontend(0, b, 0, icp, vis, hd, nh)
        c.batch_upload(0, b); c.batch_run(0, 3)
        icp2 = np.zeros_like(icp); vis2 = np.zeros_like(vis)
        c.batch_download(0, 3, icp2, vis2, None, None)
        assert icp2[:, :, :59].tobytes() == icp[:, :, :59].tobytes() and vis2.tobytes() == vis.tobytes()
        sec, oicp, ovis = oracle.bench_frames(b, prm, calib, threads=min(2, os.cpu_count() or 1), want_out=True)
        for t in (1, 2):
            for p in range(b.n_passes):
                sc = np.abs(oicp[t, p, :56]).max()
                np.testing.assert_allclose(icp[t, p, :56], oicp[t, p, :56], rtol=RTOL_NEQ, atol=RTOL_NEQ * 1e-6 * sc)
                assert icp[t, p, 56] == oicp[t, p, 56] and icp[t, p, 58] == oicp[t, p, 58] == b.n_points[t]
            for it in range(b.n_vis):
                sc = np.abs(ovis[t, it, :56]).max()
                np.testing.assert_allclose(vis[t, it, :56], ovis[t, it, :56], rtol=RTOL_NEQ, atol=RTOL_NEQ * 1e-6 * sc)
                assert vis[t, it, 56] == ovis[t, it, 56] and vis[t, it, 57] == ovis[t, it, 57]
        assert icp[1:, 0, 56].min() > 50000
    finally:
        c.close()


def test_offroad_rig_visual_and_batched_path(velo, oracle, ctx4):
    """BASELINE configs[3] end to end: 4 cameras x 8000 features.  (a) visual residual blocks (order, types, r, J) and normal
    equations at C = 4 vs the oracle, iter 1 and 2; (b) the batched path at C = 4 (full front end per frame) vs the oracle's driver."""
    cal = ctx4.cal
    F = 8000
    d = _visual_inputs(velo, oracle, cal, ctx4, F, 4, frames=(11, 12))
    kpA0, _, _, hd0, kw0 = d[11]
    _, kpB1, m1, hd1, kw1 = d[12]
    matches = np.zeros((4, F, 2), np.int32); nm = np.zeros(4, np.int32); cat = []
    for cam in range(4):
        idx = np.nonzero(m1[cam])[0]
        nm[cam] = len(idx); matches[cam, : len(idx), 0] = idx; matches[cam, : len(idx), 1] = idx
        cat.append(matches[cam, : len(idx)])
    cat = np.concatenate(cat)
    prm4 = ctx4.prm
    for it in (1, 2):
        pose = velo.synth.pose_guess(12, 3 if it == 2 else 0)
        ob, oneq = oracle.visual(kpB1, kpA0, hd1[1], hd0[0], kw1[1], kw0[0], nm, matches, cal, prm4, pose, it)
        gb, gneq = ctx4.visual_residuals(1, 1, 0, 0, nm, cat, pose, it)
        assert len(gb) == len(ob) > 4000 and set(np.unique(ob["cam"])) == {0, 1, 2, 3}
        for f in ("cam", "match", "type", "n_res"):
            assert np.array_equal(gb[f], ob[f]), f
        np.testing.assert_allclose(gb["residual"], ob["residual"], rtol=RTOL_RES, atol=1e-12)
        np.testing.assert_allclose(gb["jacobian"], ob["jacobian"], rtol=RTOL_RES, atol=1e-12)
        np.testing.assert_allclose(gneq[:56], oneq[:56], rtol=RTOL_NEQ, atol=RTOL_NEQ * 1e-6 * np.abs(oneq[:56]).max())
        assert gneq[56] == oneq[56] and gneq[57] == oneq[57]
    # (b) batched path, 4 cameras
    prm = velo.api.default_params(max_slots=3, max_features=F, max_matches=F, num_cams=4, icp_skip=20, max_rings=64)
    c = velo.api.Context(prm, cal)
    try:
        b = velo.synth.Batch(40, 3, prm, rig=1)
        icp = np.zeros((3, b.n_passes, velo.abi.NEQ_STRIDE)); vis = np.zeros((3, b.n_vis, velo.abi.NEQ_STRIDE))
        hd = np.zeros((3, 2, 4, F), np.int32); nh = np.zeros((3, 2, 4), np.int32)
        c.batch_frontend(0, b, 0, icp, vis, hd, nh)
        sec, oicp, ovis = oracle.bench_frames(b, prm, cal, threads=2, want_out=True)
        for t in (1, 2):
            for p in range(b.n_passes):
                np.testing.assert_allclose(icp[t, p, :56], oicp[t, p, :56], rtol=RTOL_NEQ, atol=RTOL_NEQ * 1e-6 * np.abs(oicp[t, p, :56]).max())
                assert icp[t, p, 56] == oicp[t, p, 56] and icp[t, p, 58] == oicp[t, p, 58]
            for it in range(b.n_vis):
                np.testing.assert_allclose(vis[t, it, :56], ovis[t, it, :56], rtol=RTOL_NEQ, atol=RTOL_NEQ * 1e-6 * np.abs(ovis[t, it, :56]).max())
                assert vis[t, it, 56] == ovis[t, it, 56] > 4000
        pts, rs, _ = oracle.segment(b.scans[2, : b.n_points[2]], cal)
        for cam in range(4):
            rc, proj, valid = oracle.project(pts, rs, cal, cam)
            for s in range(2):
                ohd, _ = oracle.depth_assoc(valid, proj, rc, b.kp[2, s, cam])
                assert np.array_equal(hd[2, s, cam], ohd) and nh[2, s, cam] == (ohd >= 0).sum()
    finally:
        c.close()


def test_bad_inputs_are_errors_not_garbage(velo, calib, ctx):
    """caller-supplied match indices outside their keypoint sets -> VELO_ERR_INVALID_ARG; a depth association on a slot whose scan
    was replaced without projecting it again finds nothing (the old projection is void)"""
    raw, _ = velo.synth.scan(7)
    ctx.scan_upload(0, raw); ctx.scan_upload(1, raw)
    kp = velo.synth.features(7, 500)[0][0]
    for slot in (0, 1):
        for cam in (0, 1):
            ctx.project(slot, cam)
            for s in (0, 1):
                ctx.depth_assoc(slot, cam, kp, s)
    nm = np.array([2, 0], np.int32)
    ok = np.array([[0, 1], [2, 3]], np.int32)
    ctx.visual_residuals(1, 1, 0, 0, nm, ok, np.zeros(6), 1)
    for bad in ([[0, 500], [2, 3]], [[-1, 1], [2, 3]], [[0, 1], [1 << 20, 3]]):
        with pytest.raises(velo.api.VeloError) as e:
            ctx.visual_residuals(1, 1, 0, 0, nm, np.array(bad, np.int32), np.zeros(6), 1)
        assert e.value.code == 3
    hd, _ = ctx.depth_assoc(0, 0, kp, 0)
    assert (hd >= 0).sum() > 50
    ctx.scan_upload(0, raw)                      # new scan in the slot, not projected yet
    hd, kw = ctx.depth_assoc(0, 0, kp, 0)
    assert (hd >= 0).sum() == 0 and len(kw) == 0


def test_xyz_scan_records_equal_kitti_records(velo, calib):
    """velo_batch_inputs.scan_stride_floats = 3 (packed xyz, a quarter less to upload; loadPoints drops the reflectance anyway,
    kitti.h:145-148): every output of the batched front end is byte-identical to the KITTI float4 upload."""
    prm = velo.api.default_params(max_slots=3, max_features=600, max_matches=600, icp_skip=7, max_rings=64)
    c = velo.api.Context(prm, calib)
    try:
        b = velo.synth.Batch(50, 3, prm)
        outs = []
        for bb in (b, b.xyz()):
            icp = np.zeros((3, b.n_passes, velo.abi.NEQ_STRIDE)); vis = np.zeros((3, b.n_vis, velo.abi.NEQ_STRIDE))
            hd = np.zeros((3, 2, prm.num_cams, prm.max_features), np.int32); nh = np.zeros((3, 2, prm.num_cams), np.int32)
            c.batch_frontend(0, bb, 0, icp, vis, hd, nh)
            pts, rs = c.scan_download(2)
            outs.append((icp.tobytes(), vis.tobytes(), hd.tobytes(), nh.tobytes(), pts.tobytes(), rs.tobytes()))
            assert icp[1:, :, 56].min() > 1000 and vis[1:, :, 56].min() > 100
        assert outs[0] == outs[1]
    finally:
        c.close()


def test_batched_frame_to_frame_vs_oracle(velo, oracle, calib):
    """velo_gpu_batch_frame_to_frame: the live frameToFrame loop (velo.h:616-907: every correspondence pass at the pose its pair has
    reached) for all frame pairs of a batch side by side, one LM controller per pair.  Each pair vs the oracle's frame_to_frame from
    the reference's start (0,0,0,0,0,1): same LM path (blocks, iterations, accepted steps, termination reasons), poses to 1e-8; and
    vs the single-pair entry point."""
    F = 800
    prm = velo.api.default_params(max_slots=4, max_features=F, max_matches=F, icp_skip=4, max_rings=64)
    c = velo.api.Context(prm, calib)
    try:
        b = velo.synth.Batch(60, 4, prm)
        c.batch_upload(0, b)
        A = velo.abi
        c.batch_run(0, 4, A.STAGE_INGEST | A.STAGE_INDEX | A.STAGE_PROJECT | A.STAGE_ASSOC)
        guess = np.array([0, 0, 0, 0, 0, 1.0])
        t, reps = c.batch_frame_to_frame(0, 4, np.tile(guess, (4, 1)))
        assert np.array_equal(t[0], guess) and reps[0]["n_solves"] == 0           # the halo slot has no previous scan
        seg = [oracle.segment(b.scans[s, : b.n_points[s]], calib)[:2] for s in range(4)]
        hd = np.zeros((4, 2, 2, F), np.int32); kw = np.zeros((4, 2, 2, F, 4), np.float32)
        for s in range(4):
            for cam in (0, 1):
                rc, proj, valid = oracle.project(*seg[s], calib, cam)
                for st in (0, 1):
                    h, k = oracle.depth_assoc(valid, proj, rc, b.kp[s, st, cam])
                    hd[s, st, cam] = h; kw[s, st, cam, : len(k)] = k
        for s in (1, 2, 3):
            vis = (b.kp[s, 1], b.kp[s - 1, 0], hd[s, 1], hd[s - 1, 0], kw[s, 1], kw[s - 1, 0], b.n_matches[s], b.matches[s])
            ox, orep = oracle.frame_to_frame(*seg[s], *seg[s - 1], calib, prm, guess, vis, 1, prm.icp_skip)
            g = reps[s]
            assert g["n_solves"] == orep["n_solves"] == 6
            assert g["n_blocks"] == orep["n_blocks"] and g["reason"] == orep["reason"], (s, g["n_blocks"], orep["n_blocks"])
            assert g["lm_iterations"] == orep["lm_iterations"] and g["accepted_steps"] == orep["accepted_steps"]
            np.testing.assert_allclose(g["pose"], orep["pose"], rtol=0, atol=1e-8)
            np.testing.assert_allclose(t[s], ox, rtol=0, atol=1e-8)
            truth = velo.synth.pose(60 + s)
            assert np.abs(t[s][:3] - truth[:3]).max() < 1e-3 and np.abs(t[s][3:] - truth[3:]).max() < 2e-2
            cat = np.concatenate([b.matches[s, cam, : b.n_matches[s, cam]] for cam in (0, 1)])
            sx, srep = c.frame_to_frame(s, 1, s - 1, 0, guess, n_matches=b.n_matches[s], matches=cat, enable_icp=1, icp_skip=prm.icp_skip)
            assert srep["lm_iterations"] == g["lm_iterations"] and srep["n_blocks"] == g["n_blocks"]
            np.testing.assert_allclose(sx, t[s], rtol=0, atol=1e-9)
        # ICP terms only / visual terms only
        t2, reps2 = c.batch_frame_to_frame(0, 4, np.tile(guess, (4, 1)), enable_visual=0)
        ox, orep = oracle.frame_to_frame(*seg[2], *seg[1], calib, prm, guess, None, 1, prm.icp_skip)
        assert reps2[2]["n_blocks"] == orep["n_blocks"] and reps2[2]["lm_iterations"] == orep["lm_iterations"]
        np.testing.assert_allclose(t2[2], ox, rtol=0, atol=1e-8)
        t3, reps3 = c.batch_frame_to_frame(0, 4, np.tile(guess, (4, 1)), enable_icp=0)
        vis = (b.kp[2, 1], b.kp[1, 0], hd[2, 1], hd[1, 0], kw[2, 1], kw[1, 0], b.n_matches[2], b.matches[2])
        ox, orep = oracle.frame_to_frame(*seg[2], *seg[1], calib, prm, guess, vis, 0, prm.icp_skip)
        assert reps3[2]["n_solves"] == orep["n_solves"] == prm.f2f_iterations and reps3[2]["n_blocks"] == orep["n_blocks"]
        np.testing.assert_allclose(t3[2], ox, rtol=0, atol=1e-7)
    finally:
        c.close()
