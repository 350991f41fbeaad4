import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def velo():
    return importlib.import_module("vision-enhanced-lidar-odometry_b200")


@pytest.fixture(scope="session")
def pyoracle():
    import pyoracle as po
    return po


@pytest.fixture(scope="session")
def oracle(pyoracle):
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def ref(pyoracle):
    """The reference's own source lines, compiled (oracle/_ref).  Built here when /root/reference exists."""
    try:
        return pyoracle.Ref()
    except (FileNotFoundError, OSError) as e:
        pytest.skip(f"oracle/_ref/libvelo_ref.so not available: {e}")


@pytest.fixture(scope="session")
def calib(velo, oracle):
    P, Tr, w, h = velo.synth.calib_raw(0)
    return oracle.calib_from_kitti(P, Tr, w, h)


@pytest.fixture(scope="session")
def params(velo):
    return velo.api.default_params()


@pytest.fixture(scope="session")
def frames(velo, oracle, calib):
    """Two consecutive synthetic frames, segmented by the oracle: dict frame -> (raw, pts, ring_start)."""
    out = {}
    for f in (7, 8):
        raw, n = velo.synth.scan(f)
        pts, rs, nr = oracle.segment(raw, calib)
        out[f] = (raw, pts, rs)
    return out


def small_scan(velo, frame, keep_rings=range(20, 44), az_stride=4):
    """A thinned synthetic scan (still KITTI ordered) for tests that need brute-force NN."""
    raw, n = velo.synth.scan(frame)
    x, y = raw[:, 0], raw[:, 1]
    flag = np.zeros(n, bool)
    flag[1:] = (x[1:] > 0) & ((y[1:] > 0) != (y[:-1] > 0))
    ring = np.cumsum(flag)
    keep = np.isin(ring, list(keep_rings))
    idx = np.nonzero(keep)[0][::az_stride]
    # keep ring seams intact: always keep first and last point of each kept ring
    return np.ascontiguousarray(raw[np.sort(idx)])
