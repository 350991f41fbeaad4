"""In-tree builds: nvcc (sm_100a) for csrc/*.cu -> libvelo_gpu.so, gcc for host/velo_synth.c -> libvelo_synth.so.

Both are plain shared libraries with a C ABI (include/velo_gpu.h); no torch headers are involved.
"""
import contextlib
import fcntl
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
GPU_LIB = os.path.join(PKG, "libvelo_gpu.so")
SYNTH_LIB = os.path.join(HOST, "libvelo_synth.so")

# -fmad=false: the reference is an FMA-free x86-64 build (CMakeLists.txt:30); every index-determining
# float expression must round like it (SURVEY.md hazard H2).
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xcompiler", "-ffp-contract=off",   # host code decides calibration / pose constants bit for bit
    "-shared", "-cudart", "static",
]


@contextlib.contextmanager
def _build_lock(out):
    """One builder at a time per output (eight ranks of one torchrun import this package at once); the others wait and then find
    the output fresh."""
    with open(out + ".lock", "w") as f:
        fcntl.flock(f, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(f, fcntl.LOCK_UN)


def _run_to(cmd, out):
    """run a compiler command whose last two arguments are `-o out`, writing to a temporary name and renaming (a reader never sees
    a half-written library)"""
    tmp = f"{out}.tmp.{os.getpid()}"
    subprocess.run(cmd[:-1] + [tmp], check=True)
    os.replace(tmp, out)


def _newer(out, srcs):
    return os.path.isfile(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs)


def find_nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.isfile(c):
            return c
    return None


def gpu_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def gpu_deps():
    return gpu_sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [
        os.path.join(ROOT, "include", "velo_gpu.h")]


def build_gpu(force=False, verbose=False):
    if not force and _newer(GPU_LIB, gpu_deps()):
        return GPU_LIB
    nvcc = find_nvcc()
    if nvcc is None:
        if os.path.isfile(GPU_LIB):
            return GPU_LIB  # prebuilt library shipped with the snapshot
        raise RuntimeError("nvcc not found and libvelo_gpu.so is not built")
    with _build_lock(GPU_LIB):
        if not force and _newer(GPU_LIB, gpu_deps()):
            return GPU_LIB      # another process built it while this one waited
        extra = os.environ.get("VELO_NVCC_EXTRA", "").split()      # tuning experiments only (tools/)
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + [
            "-I", os.path.join(ROOT, "include"), "-I", CSRC] + gpu_sources() + ["-o", GPU_LIB]
        _run_to(cmd, GPU_LIB)
    return GPU_LIB


def build_synth(force=False):
    src = os.path.join(HOST, "velo_synth.c")
    if not force and _newer(SYNTH_LIB, [src]):
        return SYNTH_LIB
    if shutil.which("gcc") is None and os.path.isfile(SYNTH_LIB):
        return SYNTH_LIB
    with _build_lock(SYNTH_LIB):
        if not force and _newer(SYNTH_LIB, [src]):
            return SYNTH_LIB
        _run_to(["gcc", "-O2", "-fPIC", "-shared", "-Wall", src, "-lm", "-o", SYNTH_LIB], SYNTH_LIB)
    return SYNTH_LIB


def build_all(force=False, verbose=False):
    build_synth(force)
    build_gpu(force, verbose)
