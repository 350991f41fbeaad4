#!/usr/bin/env python3
"""Build tuning variants of libvelo_gpu.so HERE (nvcc cross-compiles without a GPU) so that the GPU box only runs them:
objects of the unchanged translation units are compiled once, the listed .cu files are recompiled per variant.
usage: tools/build_variants.py tag1="-DFLAG ..." tag2="..."        -> exp_libs/libvelo_<tag>.so   (exp_libs/ is git-ignored)
       VARIANT_SRCS=velo_icp.cu,velo_kernels.cu selects the files that see the flags (default velo_icp.cu)."""
import importlib, os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b = importlib.import_module("vision-enhanced-lidar-odometry_b200._build")
OBJ = "/tmp/velo_obj"; OUT = os.path.join(ROOT, "exp_libs"); os.makedirs(OBJ, exist_ok=True); os.makedirs(OUT, exist_ok=True)
flags = [f for f in b.NVCC_FLAGS if f not in ("-shared",)]
inc = ["-I", os.path.join(ROOT, "include"), "-I", b.CSRC]
var_srcs = os.environ.get("VARIANT_SRCS", "velo_icp.cu").split(",")
ALL = os.environ.get("VARIANT_ALL") is not None      # flags that change a header constant: every file sees them

def cc(src, extra, out):
    subprocess.run([b.find_nvcc()] + flags + extra + inc + ["-c", src, "-o", out], check=True, stderr=subprocess.DEVNULL if not os.environ.get("V") else None)
    return out

def base_obj(src):
    o = os.path.join(OBJ, os.path.basename(src) + ".o")
    deps = b.gpu_deps()
    if not (os.path.isfile(o) and all(os.path.getmtime(o) >= os.path.getmtime(d) for d in deps)): cc(src, [], o)
    return o

def variant(arg):
    tag, _, fl = arg.partition("=")
    objs = []
    for s in b.gpu_sources():
        if ALL or os.path.basename(s) in var_srcs: objs.append(cc(s, fl.split(), os.path.join(OBJ, f"{tag}.{os.path.basename(s)}.o")))
        else: objs.append(base_obj(s))
    out = os.path.join(OUT, f"libvelo_{tag}.so")
    subprocess.run([b.find_nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", out], check=True)
    return out

with ThreadPoolExecutor(8) as ex:
    list(ex.map(base_obj, [s for s in b.gpu_sources() if ALL is False and os.path.basename(s) not in var_srcs]))
    for o in ex.map(variant, sys.argv[1:]): print(o)
