#!/bin/bash
cd "$(dirname "$0")/.."
for c in 8 16 32 1000000; do
  VELO_NVCC_EXTRA="-DICP_SCAN_CHUNK=$c" python -c "
import importlib; b=importlib.import_module('vision-enhanced-lidar-odometry_b200._build'); b.build_gpu(force=True)" 2>&1 | grep -i " error"
  python bench.py --frames ${FRAMES:-200} --steps 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('chunk $c:', d['value'], 'frames/s  icp ms', k['icp_pass']['ms_per_launch'])"
done
