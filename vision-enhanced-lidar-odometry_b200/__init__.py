"""velo-b200: B200-native (sm_100a) front end of VELO (lichunshang/vision-enhanced-lidar-odometry).

The product is the C-ABI shared library built from csrc/ (include/velo_gpu.h) plus the C++ drop-in adapters in
include/velo_dropin.hpp.  This Python package only builds and binds it for the tests and bench.py:
    abi    — ctypes mirror of include/velo_gpu.h
    api    — Context: one method per C-ABI call
    synth  — seeded synthetic KITTI-shaped inputs (host/velo_synth.c)
Import it with importlib.import_module("vision-enhanced-lidar-odometry_b200") (the directory name is not an identifier).
"""
from . import _build, abi  # noqa: F401

__all__ = ["_build", "abi", "api", "synth"]


def __getattr__(name):
    if name in ("api", "synth"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
