#!/bin/bash
# timing experiments: switch parts of k_icp_pass off (results may then be wrong or slower) to see what each part costs
cd "$(dirname "$0")/.."
IFS=';'
for ex in ${EXPS:-;-DEXP_NO_ACCUM;-DEXP_NO_ACCUM -DEXP_NO_EPILOGUE;-DEXP_NO_PHASE2}; do
  unset IFS
  VELO_NVCC_EXTRA="$ex" python -c "
import importlib; b=importlib.import_module('vision-enhanced-lidar-odometry_b200._build'); b.build_gpu(force=True)" 2>&1 | grep -i " error"
  python bench.py --frames ${FRAMES:-100} --steps 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$ex]', 'icp ms', d['kernels']['icp_pass']['ms_per_launch'], d['icp_search']['per_pass_candidates_per_query'], d['icp_search']['per_pass_rings_scanned_per_query'])"
done
