"""Timings of the SURVEY §8(f) rows (f1 device-resident frameToFrame, f3 batched triangulation, f4 Hamming matcher) through the
C ABI with host buffers (H2D/D2H included), next to the CPU oracle on one host core.  Prints one JSON line per row."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import pyoracle
import tri_data
from test_oracle_vs_ref import _descriptors
velo = importlib.import_module("vision-enhanced-lidar-odometry_b200")
api, synth = velo.api, velo.synth
orc = pyoracle.Oracle()
P, Tr, w, h = synth.calib_raw(0)
cal = api.calib_from_kitti(P, Tr, w, h)
prm = api.default_params(max_slots=2, max_features=2000, max_matches=2000)
ctx = api.Context(prm, cal)


def best(fn, n=5):
    fn(); ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return min(ts)


# ---- f1: full frameToFrame (2 x 3 solves) on a full-size frame pair
rawS, _ = synth.scan(700); rawM, _ = synth.scan(701)
ctx.scan_upload(0, rawS); ctx.scan_upload(1, rawM)
ptsS, rsS, _ = orc.segment(rawS, cal); ptsM, rsM, _ = orc.segment(rawM, cal)
guess = np.array([0, 0, 0, 0, 0, 1.0]); truth = synth.pose(701)
for skip in (1, 20, 200):
    g = best(lambda: ctx.frame_to_frame(1, 1, 0, 0, guess, enable_icp=1, icp_skip=skip), 3)
    gx, rep = ctx.frame_to_frame(1, 1, 0, 0, guess, enable_icp=1, icp_skip=skip)
    t = time.perf_counter(); ox, orep = orc.frame_to_frame(ptsM, rsM, ptsS, rsS, cal, prm, guess, None, 1, skip); c = time.perf_counter() - t
    print(json.dumps({"row": "f1 frame_to_frame (ICP terms, 6 solves)", "icp_skip": skip, "gpu_ms": round(g * 1e3, 2), "cpu_1core_ms": round(c * 1e3, 1),
                      "lm_iterations": rep["lm_iterations"], "blocks_first_solve": rep["n_blocks"][0],
                      "pose_err_vs_truth": [float(np.abs(gx[:3] - truth[:3]).max()), float(np.abs(gx[3:] - truth[3:]).max())],
                      "max_abs_diff_vs_oracle": float(np.abs(gx - ox).max())}))
# ---- f1 batched: the live frameToFrame loop for B frame pairs side by side (one LM controller per pair), full scans, icp_skip = 1
Bn = int(os.environ.get("F2F_BATCH", "200"))
prm_b = api.default_params(max_slots=Bn + 1, max_points=131072, max_rings=64, max_features=2000, max_matches=2000, icp_skip=1)
cb = api.Context(prm_b, cal)
bb = synth.Batch(2000, Bn + 1, prm_b)
cb.batch_upload(0, bb)
A = velo.abi
cb.batch_run(0, Bn + 1, A.STAGE_INGEST | A.STAGE_INDEX | A.STAGE_PROJECT | A.STAGE_ASSOC)
g0 = np.tile(guess, (Bn + 1, 1))
for name, kw in (("ICP + visual terms", {}), ("ICP terms only", {"enable_visual": 0}), ("visual terms only (the shipped configuration, main.cpp:43)", {"enable_icp": 0})):
    g = best(lambda: cb.batch_frame_to_frame(0, Bn + 1, g0, **kw), 2)
    tb, rb = cb.batch_frame_to_frame(0, Bn + 1, g0, **kw)
    err = np.array([[np.abs(tb[s][:3] - synth.pose(2000 + s)[:3]).max(), np.abs(tb[s][3:] - synth.pose(2000 + s)[3:]).max()] for s in range(1, Bn + 1)])
    print(json.dumps({"row": "f1 batched frame_to_frame, " + name, "pairs": Bn, "icp_skip": 1, "gpu_ms": round(g * 1e3, 1), "pairs_per_s": round(Bn / g, 1),
                      "lm_iterations_mean_per_solve": [round(float(np.mean([r["lm_iterations"][k] for r in rb[1:]])), 2) for k in range(rb[1]["n_solves"])],
                      "pose_err_vs_truth_median": [float(np.median(err[:, 0])), float(np.median(err[:, 1]))], "pose_err_vs_truth_max": [float(err[:, 0].max()), float(err[:, 1].max())]}))
cb.close()
# ---- f4: matchFeatures, 3000 x 3000 FREAK-sized descriptors
rng = np.random.default_rng(0)
q = _descriptors(rng, 3000); t_ = np.concatenate([_descriptors(rng, 2000, q[:2000], flips=14), _descriptors(rng, 1000)])
g = best(lambda: ctx.match_hamming(q, t_))
t = time.perf_counter(); op, _, _ = orc.match_hamming(q, t_); c = time.perf_counter() - t
gp, _, _ = ctx.match_hamming(q, t_)
print(json.dumps({"row": "f4 match_hamming 3000x3000x64B", "gpu_ms": round(g * 1e3, 3), "cpu_1core_ms": round(c * 1e3, 1), "pairs": len(gp), "identical": bool(np.array_equal(gp, op))}))
# ---- f3: triangulatePoint for 20 000 landmarks
off3, obs3, off2, obs2, poses, truth3 = tri_data.make(3, L=20000, n_frames=10)
g = best(lambda: ctx.triangulate(off3, obs3, off2, obs2, poses))
t = time.perf_counter(); o, _ = orc.triangulate(off3, obs3, off2, obs2, poses, cal, prm); c = time.perf_counter() - t
gg, _ = ctx.triangulate(off3, obs3, off2, obs2, poses)
print(json.dumps({"row": "f3 triangulate 20000 landmarks", "observations": int(len(obs3) + len(obs2)), "gpu_ms": round(g * 1e3, 3), "cpu_1core_ms": round(c * 1e3, 1),
                  "max_abs_diff_vs_oracle": float(np.abs(gg - o).max())}))
ctx.close()
