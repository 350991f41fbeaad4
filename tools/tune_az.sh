#!/bin/bash
cd "$(dirname "$0")/.."
for az in 512 1024 2048; do
  VELO_NVCC_EXTRA="-DVELO_AZ_BINS=$az" python -c "
import importlib; b=importlib.import_module('vision-enhanced-lidar-odometry_b200._build'); b.build_gpu(force=True)" 2>&1 | grep -i " error"
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "icp_small or icp_degenerate or golden" 2>&1 | tail -1
  python bench.py --frames ${FRAMES:-200} --steps 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('AZ $az:', d['value'], 'frames/s  icp ms', k['icp_pass']['ms_per_launch'], 'index', k['index_build']['ms_per_launch'], k['index_masks']['ms_per_launch'], d['icp_search']['per_pass_candidates_per_query'])"
done
