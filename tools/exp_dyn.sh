#!/bin/bash
# run length (chunks per dynamically scheduled run) of k_icp_pass at the bench size
cd "$(dirname "$0")/.."
IFS=';'
for ex in ${EXPS:--DICP_RUN_CHUNKS=4;-DICP_RUN_CHUNKS=8}; do
  unset IFS
  VELO_NVCC_EXTRA="$ex" python -c "
import importlib; b=importlib.import_module('vision-enhanced-lidar-odometry_b200._build'); b.build_gpu(force=True)" 2>&1 | grep -i " error"
  for c in ${CTAS:-9}; do
  python bench.py --frames ${FRAMES:-1000} --steps 2 --warmup 3 --no-cpu --icp-ctas $c 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('[$ex] ctas $c:', d['value'], 'frames/s  icp ms', k['icp_pass']['ms_per_launch'], 'reduce ms', k['neq_reduce']['ms_per_launch'], 'x', k['neq_reduce']['launches'])"
  done
done
