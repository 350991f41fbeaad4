"""-m "not gpu": the C-ABI library loads and exports every symbol include/velo_gpu.h declares, its host-side functions
agree with the oracle, and without a GPU it fails loudly (no CPU fallback).  No compute call is made."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_and_binding_agree(velo):
    hdr = open(os.path.join(ROOT, "include", "velo_gpu.h")).read()
    declared = set(re.findall(r"\b(velo_[a-z0-9_]+)\s*\(", hdr)) - {"velo_status"}
    assert declared == set(velo.abi.EXPORTS), declared ^ set(velo.abi.EXPORTS)
    lib = velo.api.lib()
    for name in velo.abi.EXPORTS:
        assert hasattr(lib, name), name
    assert lib.velo_gpu_abi_version() == velo.abi.ABI_VERSION


def test_struct_sizes_match_c(velo, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "velo_gpu.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",sizeof(velo_gpu_params),sizeof(velo_gpu_calib),sizeof(velo_icp_corr),sizeof(velo_vis_block),sizeof(velo_batch_inputs));return 0;}\n')
    exe = tmp_path / "sz"
    import subprocess
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    a = velo.abi
    assert sizes == [C.sizeof(a.Params), C.sizeof(a.Calib), C.sizeof(a.IcpCorr), C.sizeof(a.VisBlock), C.sizeof(a.BatchInputs)]


def test_default_params_are_kitti_h(velo):
    p = velo.api.default_params()
    assert (p.num_cams, p.icp_skip, p.f2f_iterations, p.icp_iterations, p.max_features) == (2, 200, 2, 3, 3000)   # kitti.h:3-10
    assert (p.weight_3D2D, p.weight_2D2D, p.weight_3DPD) == (10, 500, 1)
    assert (p.loss_thresh_3D2D, p.loss_thresh_2D2D, p.loss_thresh_3DPD, p.loss_thresh_3D3D) == (0.01, 0.00002, 0.1, 0.04)
    assert (p.depth_assoc_thresh, p.outlier_reject, p.correspondence_thresh_icp, p.icp_norm_condition) == (0.015, 5.0, 0.5, 1e-5)
    assert p.enable_2d2d == 1 and p.enable_3d2d == 1 and p.abs_truncates == 0                                        # main.cpp:44-45


@pytest.mark.parametrize("rig", [0, 1])
def test_calibration_equals_oracle(velo, oracle, rig):
    P, Tr, w, h = velo.synth.calib_raw(rig)
    a = velo.api.calib_from_kitti(P, Tr, w, h)
    b = oracle.calib_from_kitti(P, Tr, w, h)
    assert bytes(a) == bytes(b)
    assert abs(a.cam_trans[1][0] - (-0.537 if rig else -386.1448 / 718.856)) < 1e-4     # kitti.h:74-78
    assert a.min_x[0] < 0 < a.max_x[0] and a.min_y[0] < 0 < a.max_y[0]


def test_pixel_canonical_round_trip(velo, oracle, calib):
    rng = np.random.default_rng(0)
    pix = np.stack([rng.uniform(0, 1241, 500), rng.uniform(0, 376, 500)], 1).astype(np.float32)
    can = velo.api.pixel2canonical(calib, 0, pix)
    assert can.tobytes() == oracle.pixel2canonical(calib, 0, pix).tobytes()
    back = velo.api.canonical2pixel(calib, 0, can)
    assert back.tobytes() == oracle.canonical2pixel(calib, 0, can).tobytes()
    assert np.abs(back - pix).max() < 1e-2


def test_create_fails_loudly_without_gpu(velo, calib):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(velo.api.VeloError) as e:
        velo.api.Context(velo.api.default_params(), calib)
    assert e.value.code == 1 and "no CPU fallback" in str(e.value)


def test_invalid_params_rejected(velo, calib):
    for kw in ({"num_cams": 5}, {"max_rings": 1000}, {"max_points": 1 << 21}, {"icp_skip": 0}):
        with pytest.raises(velo.api.VeloError) as e:
            velo.api.Context(velo.api.default_params(**kw), calib)
        assert e.value.code == 3


def test_product_does_not_import_oracle():
    """the product path must never route through oracle/: no file of the package or include/ mentions it"""
    pkg = os.path.join(ROOT, "vision-enhanced-lidar-odometry_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".hpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "pyoracle" not in txt and "velo_oracle" not in txt and "libvelo_ref" not in txt, os.path.join(dp, f)


def test_kitti_wire_formats(velo, tmp_path):
    """calib.txt / velodyne .bin / pose line formats of kitti.h:59-152,202-216"""
    P, Tr, w, h = velo.synth.calib_raw(0)
    lines = [f"P{c}: " + " ".join(repr(float(x)) for x in P[12 * c: 12 * c + 12]) for c in range(4)]
    lines.append("Tr: " + " ".join(repr(float(x)) for x in Tr))
    (tmp_path / "calib.txt").write_text("\n".join(lines) + "\n")
    P2, Tr2 = velo.api.kitti_load_calib(tmp_path / "calib.txt")
    assert P2.tobytes() == P.tobytes() and Tr2.tobytes() == Tr.tobytes()
    raw, n = velo.synth.scan(3)
    raw[:5000].tofile(tmp_path / "000003.bin")
    back = velo.api.kitti_load_scan(tmp_path / "000003.bin")
    assert back.tobytes() == raw[:5000].tobytes()
    with pytest.raises(velo.api.VeloError):
        velo.api.kitti_load_scan(tmp_path / "missing.bin")
    T = np.eye(4); T[0, 3] = 1.5; T[2, 3] = -0.25; T[1, 1] = 0.999999123
    assert velo.api.kitti_format_pose(T) == "1 0 0 1.5 0 0.999999 0 0 0 0 1 -0.25 "


def test_query_atan2_error_bound():
    """The pruning windows of the correspondence search are padded by 1e-5 rad for atan2_q (csrc/velo_common.cuh); its error,
    emulated in float32 with the kernel's operation order and the coefficients read from the source, must stay below 2.5e-6."""
    import importlib.util, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("check_atan2", os.path.join(root, "tools", "check_atan2.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    src = open(os.path.join(root, "vision-enhanced-lidar-odometry_b200", "csrc", "velo_common.cuh")).read()
    body = src[src.index("float atan2_q("):src.index("float r = p * a;")]
    coef = [float(x) for x in re.findall(r"([-+]? ?\d\.\d{10,})f", body.replace("- 0", "-0").replace("+ 0", "+0"))]
    assert len(coef) == 6
    assert np.allclose(sorted(abs(c) for c in coef), sorted(abs(float(c)) for c in mod.c), rtol=0, atol=1e-7)
    assert mod.max_error(300_000) < 2.5e-6


def test_no_fused_packed_f32_in_product_sass():
    """Hazard H14: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with explicit rounding and -fmad=false
    (the scalar forms are respected).  The candidate distance uses FADD2 / FMUL2 only for operations that do not feed a packed
    add; a fused packed op anywhere in the product library would silently break bit-exact indices."""
    import shutil, subprocess
    so = os.path.join(ROOT, "vision-enhanced-lidar-odometry_b200", "libvelo_gpu.so")
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(so) or not os.path.exists(cuobjdump):
        pytest.skip("library or cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True, check=True).stdout
    assert "FMUL2" in sass and "FADD2" in sass          # the packed path is really there
    assert "FFMA2" not in sass


def test_concurrent_builds_do_not_corrupt_the_library(velo):
    """eight ranks of one torchrun import the package at once: a (re)build must be serialised and published atomically (a stale
    header once made all ranks rebuild libvelo_gpu.so into the same file: "file too short")"""
    import subprocess
    import sys
    code = ("import importlib, ctypes; b = importlib.import_module('vision-enhanced-lidar-odometry_b200._build'); "
            "p = b.build_synth(force=True); ctypes.CDLL(p).velo_synth_calib")
    procs = [subprocess.Popen([sys.executable, "-c", code], cwd=ROOT, stderr=subprocess.PIPE) for _ in range(6)]
    for p in procs:
        _, err = p.communicate(timeout=120)
        assert p.returncode == 0, err.decode()[-500:]


def test_pose_vec2mat_is_rodrigues(velo):
    """util::pose_mat2vec (utility.h:67-82): AngleAxisToRotationMatrix + translation, incl. the first-order branch near zero"""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(4)
    for t in [rng.normal(0, 0.5, 6) for _ in range(20)] + [np.zeros(6), np.array([1e-9, -2e-9, 1e-9, 1, 2, 3])]:
        T = velo.api.pose_vec2mat(t)
        np.testing.assert_allclose(T[:3, :3], Rotation.from_rotvec(t[:3]).as_matrix(), atol=1e-14)
        assert np.array_equal(T[:3, 3], t[3:]) and np.array_equal(T[3], [0, 0, 0, 1])
