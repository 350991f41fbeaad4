#!/bin/bash
# build variants of the ICP kernel (launch bounds / CTA size) on the GPU box and bench each briefly
cd "$(dirname "$0")/.."
for cfg in ${CFGS:-128,6 256,3 256,2 192,4}; do
  t=${cfg%,*}; b=${cfg#*,}
  VELO_NVCC_EXTRA="-DICP_THREADS=$t -DICP_MIN_BLOCKS=$b $EXTRA" python -c "
import importlib; b=importlib.import_module('vision-enhanced-lidar-odometry_b200._build'); b.build_gpu(force=True)" 2>&1 | grep -i error
  python bench.py --frames ${FRAMES:-100} --steps 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('threads $t minblocks $b:', d['value'], 'frames/s  icp ms', d['kernels']['icp_pass']['ms_per_launch'])"
done
