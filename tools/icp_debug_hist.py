"""Per-query search statistics of k_icp_pass (debug build with -DVELO_ICP_DEBUG into /tmp; needs nvcc + a GPU)."""
import ctypes as C, importlib, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
v = importlib.import_module('vision-enhanced-lidar-odometry_b200')
b = v._build
out = '/tmp/libvelo_gpu_dbg.so'
subprocess.run([b.find_nvcc()] + b.NVCC_FLAGS + ['-DVELO_ICP_DEBUG', '-I', os.path.join(ROOT, 'include'), '-I', b.CSRC] + b.gpu_sources() + ['-o', out], check=True)
b.GPU_LIB = out
b.build_gpu = lambda *a, **k: out
api, syn = v.api, v.synth
P, Tr, w, h = syn.calib_raw(0); cal = api.calib_from_kitti(P, Tr, w, h)
c = api.Context(api.default_params(max_slots=2), cal)
a, _ = syn.scan(1000); bb, _ = syn.scan(1001)
c.scan_upload(0, a); c.scan_upload(1, bb)
pts, rs = c.scan_download(1)
for it, p in ((1, 0), (2, 3)):
    corr, neq, kept = c.icp_pass(1, 0, syn.pose_guess(1001, p), it, 1)
    seed, exh = corr['jacobian'][:, 3], corr['jacobian'][:, 4]
    rings = corr['jacobian'][:, 5] % 1000; mask = corr['jacobian'][:, 5] // 1000
    k = corr['kept']
    rng = np.linalg.norm(pts[:, :3], axis=1)
    print(f'iter {it}: mean exh {exh.mean():.1f}  percentiles 50/90/99/max', np.percentile(exh, [50, 90, 99, 100]))
    for name, sel in (('kept', k == 1), ('rejected(<2 rings)', k == 0)):
        print(f'   {name}: n={sel.sum()} exh mean {exh[sel].mean():.1f} rings {rings[sel].mean():.2f} mask {mask[sel].mean():.2f} share of all cand {exh[sel].sum() / exh.sum():.2f}')
    for lo, hi in ((0, 6), (6, 10), (10, 15), (15, 25), (25, 100)):
        sel = (rng >= lo) & (rng < hi)
        print(f'   range {lo}-{hi} m: n={sel.sum()} exh mean {exh[sel].mean():.1f} rings {rings[sel].mean():.2f} d_j median {np.sqrt(np.median((corr["residual"][sel]) ** 2)):.3f}')
    ring_id = corr['src_ring']
    print('   by source ring (exh mean):', [round(float(exh[ring_id == r].mean()), 0) for r in range(0, 64, 4)])
    if it == 1:
        near = rng < 6
        print('   near percentiles 10/50/75/90/99:', np.percentile(exh[near], [10, 50, 75, 90, 99]))
        heavy = np.nonzero(exh > 800)[0]
        print('   heavy n', len(heavy), 'by src ring', np.bincount(ring_id[heavy], minlength=64)[40:])
        for i in heavy[:: max(1, len(heavy) // 12)][:12]:
            r = corr[i]
            dj = -1.0
            print('    q', i, 'ring', r['src_ring'], 'idx', r['src_idx'], 'kept', r['kept'], 'si', r['np_s_i'], 'sj', r['np_s_j'], 'seed', seed[i], 'exh', exh[i], 'rings', rings[i], 'mask', mask[i], 'range %.2f' % rng[i], 'pt', pts[i, :3])
        mid = np.nonzero((exh > 100) & (exh < 200) & near)[0]
        for i in mid[:: max(1, len(mid) // 8)][:8]:
            r = corr[i]
            print('    mid q', i, 'ring', r['src_ring'], 'kept', r['kept'], 'si', r['np_s_i'], 'sj', r['np_s_j'], 'seed', seed[i], 'exh', exh[i], 'rings', rings[i], 'mask', mask[i], 'range %.2f' % rng[i])
