#!/bin/bash
# projection kernel shape sweep: warps (= rings) per CTA
cd "$(dirname "$0")/.."
IFS=';'
for ex in ${EXPS:--DPROJ_WARPS=4;-DPROJ_WARPS=4 -DPROJ_STAGES=4;-DPROJ_WARPS=8;-DPROJ_WARPS=2 -DPROJ_STAGES=12}; do
  unset IFS
  VELO_NVCC_EXTRA="$ex" python -c "
import importlib; b=importlib.import_module('vision-enhanced-lidar-odometry_b200._build'); b.build_gpu(force=True)" 2>&1 | grep -i " error"
  python bench.py --frames ${FRAMES:-400} --steps 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']['project_occlude']; print('[$ex]', 'project ms', k['ms_per_launch'], k['GBps'])"
done
