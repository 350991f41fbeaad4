"""bench.py's reference arm on the host (no GPU): the line the driver parses has the contract's keys, and the arm runs without the
CUDA library (it is built from oracle/ + host/velo_synth.c only)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    env = dict(os.environ, VELO_GPU_LIB="/nonexistent/libvelo_gpu.so")      # loading the product library would fail loudly
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--icp-skip", "50",
                          "--features", "300"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("frames/s") and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "frame pairs" in cb["sample"]
    for k in ("workload", "frames_per_step", "points_per_scan", "features_per_image", "icp_passes", "icp_skip"):
        assert k in d["config"], k
