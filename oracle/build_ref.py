#!/usr/bin/env python3
"""Build oracle/_ref/libvelo_ref.so from the reference's own source lines.

TEST INFRASTRUCTURE ONLY.  Runs only where /root/reference exists (the build container).  The hot-path
line ranges listed in SLICES are cut from the reference headers into a *temporary* directory, compiled
verbatim together with oracle/ref_glue.cpp against the stand-in headers in oracle/ref_shim/, and the
temporary directory is removed: the only output is oracle/_ref/libvelo_ref.so (git-ignored, shipped to
the GPU box by gpurun like any other built .so).  No reference source is copied into the repository.

Each slice carries an anchor string that must occur on its first line, so a changed reference fails loudly.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VELO_REFERENCE_DIR", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libvelo_ref.so")

# name: (file, first line, last line, anchor on first line, anchor on last line)
SLICES = {
    "ref_utility_a.inc": ("utility.h", 1, 56, "const double INF", "}"),
    "ref_utility_b.inc": ("utility.h", 97, 103, "static void transform_point", "}"),
    "ref_kitti_consts.inc": ("kitti.h", 3, 35, "const int num_cams", "loop_close_thresh"),
    "ref_kitti_segment.inc": ("kitti.h", 154, 185, "void segmentPoints(", "}"),
    "ref_velo_enum.inc": ("velo.h", 3, 8, "enum ResidualType", "};"),
    "ref_velo_project.inc": ("velo.h", 329, 375, "void projectLidarToCamera(", "}"),
    "ref_velo_assoc.inc": ("velo.h", 377, 497, "std::vector<int> featureDepthAssociation(", "}"),
    "ref_velo_match.inc": ("velo.h", 499, 550, "void matchFeatures(", "}"),
    "ref_velo_triangulate.inc": ("velo.h", 1027, 1130, "void triangulatePoint(", "}"),
    "ref_velo_visual.inc": ("velo.h", 622, 792, "for(int cam = 0; cam<num_cams; cam++) {", "}"),
    "ref_velo_icp_a.inc": ("velo.h", 806, 874, "for(int sm = 0; sm < scans_M.size() * enable_icp; sm++) {", "N /= N.norm();"),
    "ref_velo_icp_b.inc": ("velo.h", 875, 894, "ceres::CostFunction* cost_function =", "}"),
}


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "velo.h"))


def build(force: bool = False) -> bool:
    if not available():
        return os.path.isfile(OUT)
    srcs = [os.path.join(HERE, "ref_glue.cpp"), os.path.join(HERE, "ref_shim", "velo_ref_shim.hpp"), __file__]
    if (not force and os.path.isfile(OUT)
            and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs)):
        return True
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="velo_ref_")
    try:
        for name, (fn, a, b, anchor_a, anchor_b) in SLICES.items():
            with open(os.path.join(REF, fn)) as f:
                lines = f.readlines()
            if anchor_a not in lines[a - 1] or anchor_b not in lines[b - 1]:
                raise RuntimeError(f"{fn}:{a}-{b} does not match its anchors; reference changed?")
            with open(os.path.join(tmp, name), "w") as f:
                f.writelines(lines[a - 1:b])
        cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w",
               "-I", tmp, "-I", REF, "-I", os.path.join(HERE, "ref_shim"), "-I", os.path.join(HERE, "..", "include"),
               os.path.join(HERE, "ref_glue.cpp"), "-o", OUT]
        subprocess.run(cmd, check=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("libvelo_ref.so:", "built" if ok else "unavailable (no /root/reference)")
