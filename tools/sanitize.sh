#!/bin/bash
# compute-sanitizer passes over the smoke test + a small batched run (memcheck, racecheck, initcheck)
cd "$(dirname "$0")/.."
cat > /tmp/san_small.py <<'PY'
import importlib, numpy as np, sys
sys.path.insert(0, '.')
import __graft_entry__ as g
g.smoke()
v = importlib.import_module('vision-enhanced-lidar-odometry_b200')
api, syn = v.api, v.synth
P, Tr, w, h = syn.calib_raw(0); cal = api.calib_from_kitti(P, Tr, w, h)
prm = api.default_params(max_slots=3, max_features=500, max_matches=500, icp_skip=40, max_rings=64)
c = api.Context(prm, cal)
b = syn.Batch(20, 3, prm)
icp = np.zeros((3, b.n_passes, 64)); vis = np.zeros((3, b.n_vis, 64)); hd = np.zeros((3, 2, 2, 500), np.int32); nh = np.zeros((3, 2, 2), np.int32)
c.batch_frontend(0, b, 1, icp, vis, hd, nh)
print('batch ok', icp[1:, :, 56].ravel())
c.close()
PY
for tool in memcheck racecheck initcheck; do
  echo "== $tool"; compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_small.py 2>&1 | grep -E "ERROR SUMMARY|smoke ok|batch ok|Invalid|Race|Uninit|hazard" | head -12
done
