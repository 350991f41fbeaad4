"""The CPU oracle against the committed golden vectors (tests/golden/velo_golden.npz), which were produced by the
reference's own source lines (tests/golden/make_golden.py -> oracle/_ref).  Runs anywhere (no GPU, no /root/reference)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "velo_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def gcal(oracle, gold):
    return oracle.calib_from_kitti(gold["P"], gold["Tr"], int(gold["wh"][0]), int(gold["wh"][1]))


def test_golden_constants(gold, params):
    p = params
    got = [4, p.max_features, p.icp_skip, p.f2f_iterations, p.icp_iterations, p.weight_3D2D, p.weight_2D2D, p.weight_3DPD,
           p.loss_thresh_3D2D, p.loss_thresh_2D2D, p.loss_thresh_3DPD, p.loss_thresh_3D3D, p.depth_assoc_thresh,
           p.outlier_reject, p.correspondence_thresh_icp, p.icp_norm_condition]
    assert list(gold["constants"]) == [float(v) for v in got]


@pytest.mark.parametrize("f", [60, 61])
def test_golden_segment_project_assoc(oracle, gold, gcal, f):
    pts, rs, nr = oracle.segment(gold[f"raw{f}"], gcal)
    assert np.array_equal(rs, gold[f"rs{f}"]) and pts.tobytes() == gold[f"pts{f}"].tobytes()
    for cam in (0, 1):
        rc, proj, valid = oracle.project(pts, rs, gcal, cam)
        assert np.array_equal(rc, gold[f"rc{f}_{cam}"])
        assert proj.tobytes() == gold[f"proj{f}_{cam}"].tobytes() and valid.tobytes() == gold[f"valid{f}_{cam}"].tobytes()
        for s, key in enumerate(("kpA", "kpB")):
            hd, kw = oracle.depth_assoc(valid, proj, rc, gold[f"{key}{f}"][cam])
            assert np.array_equal(hd, gold[f"hd{f}"][s, cam])
            assert kw.tobytes() == gold[f"kw{f}"][s, cam, : len(kw)].tobytes()
            assert (hd >= 0).sum() > 5


def test_golden_transform_point(oracle, gold):
    out = oracle.transform_points(gold["pts61"][::11], gold["tp_pose"])
    assert out.tobytes() == gold["tp_out"].tobytes()


@pytest.mark.parametrize("it,skip", [(1, 1), (2, 1), (1, 4)])
@pytest.mark.parametrize("mode", [0, 1])
def test_golden_icp(oracle, gold, params, it, skip, mode):
    corr, neq, kept = oracle.icp_pass(gold["pts61"], gold["rs61"], gold["pts60"], gold["rs60"], gold[f"icp_pose_{it}_{skip}"], it, skip, params, mode)
    g = gold[f"icp_corr_{it}_{skip}"]
    ck = corr[corr["kept"] == 1]
    assert len(ck) == len(g) == kept and kept > 3
    for f in ("src_ring", "src_idx", "np_s_i", "np_i", "np_s_j", "np_j", "np_k"):
        assert np.array_equal(ck[f], g[f]), f
    for f in ("normal", "v0", "residual", "jacobian"):
        assert ck[f].tobytes() == g[f].tobytes(), f
    np.testing.assert_allclose(neq[:58], gold[f"icp_neq_{it}_{skip}"][:58], rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("it", [1, 2])
def test_golden_visual(oracle, gold, gcal, params, it):
    b, neq = oracle.visual(gold["kpB61"], gold["kpA60"], gold["hd61"][1], gold["hd60"][0], gold["kw61"][1], gold["kw60"][0],
                           gold["vis_nm"], gold["vis_matches"], gcal, params, gold[f"vis_pose_{it}"], it)
    g = gold[f"vis_blocks_{it}"]
    assert len(b) == len(g) > 50
    for f in ("cam", "match", "type", "n_res"):
        assert np.array_equal(b[f], g[f]), f
    assert b["residual"].tobytes() == g["residual"].tobytes() and b["jacobian"].tobytes() == g["jacobian"].tobytes()
    np.testing.assert_allclose(neq[:58], gold[f"vis_neq_{it}"][:58], rtol=1e-12, atol=1e-300)


def test_golden_functors(oracle, gold):
    for k, pose, r, J in zip(gold["fun_k"], gold["fun_pose"], gold["fun_r"], gold["fun_J"]):
        typ, n = int(k[0]), int(k[1])
        ro, Jo = oracle.eval_functor(typ, k[2:2 + n], pose)
        assert ro.tobytes() == r[: len(ro)].tobytes() and Jo.tobytes() == J[: len(ro)].tobytes()
