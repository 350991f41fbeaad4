"""-m "not gpu": csrc/velo_functors.h (the residual functors the CUDA kernels evaluate: rotation applied as a linear map with a
per-pose derivative pack instead of per-block autodiff) compiled for the host and compared with the oracle's dual-number
evaluation of costfunctions.h — residuals and Jacobians of all five functors, general / tiny / zero rotations."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_K = {0: 6, 1: 8, 2: 8, 3: 7, 4: 9}      # VELO_RES_3D3D, 3D2D, 2D3D, 2D2D, 3DPD: constructor doubles


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = tmp_path_factory.mktemp("functors") / "libfunctors.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "vision-enhanced-lidar-odometry_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "functors_shim.cpp"), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    lib.functor_eval.argtypes = [C.c_int] + [C.c_void_p] * 4
    return lib


def _inputs(rng, typ):
    k = np.zeros(N_K[typ])
    if typ == 0:
        k[:] = np.concatenate([rng.normal(0, 8, 3), rng.normal(0, 8, 3)])
    elif typ in (1, 2):
        k[:] = np.concatenate([rng.normal(0, 8, 2), [abs(rng.normal(12, 5)) + 2], rng.uniform(-0.8, 0.8, 2), [-0.537, 0.0, 0.0]])
    elif typ == 3:
        k[:] = np.concatenate([rng.uniform(-0.8, 0.8, 4), [-0.537, 0.01, -0.02]])
    else:
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        k[:] = np.concatenate([rng.normal(0, 10, 3), n, rng.normal(0, 10, 3)])
    return k


@pytest.mark.parametrize("typ", [0, 1, 2, 3, 4])
def test_linear_map_functors_equal_autodiff(shim, oracle, typ):
    rng = np.random.default_rng(100 + typ)
    poses = [np.concatenate([rng.normal(0, 0.02, 3), rng.normal(0, 0.1, 2), [1 + rng.normal(0, 0.2)]]) for _ in range(40)]
    poses += [np.array([0.0, 0, 0, 0, 0, 1]), np.array([1e-9, -1e-9, 2e-9, 0.01, 0, 1]), np.array([0.7, -1.1, 0.4, 2, -3, 5]), np.array([3e-8, 0, 0, 0, 0, 0.5])]
    worst = 0.0
    for pose in poses:
        for _ in range(5):
            k = _inputs(rng, typ)
            r = np.zeros(3); J = np.zeros(18)
            n = shim.functor_eval(typ, k.ctypes.data, pose.ctypes.data, r.ctypes.data, J.ctypes.data)
            orr, oJ = oracle.eval_functor(typ, k, pose)
            assert n == len(orr)
            sr = max(np.abs(orr).max(), 1e-3); sJ = max(np.abs(oJ).max(), 1e-3)
            worst = max(worst, np.abs(r[:n] - orr).max() / sr, np.abs(J[:6 * n].reshape(n, 6) - oJ).max() / sJ)
    assert worst < 1e-12, worst
