#!/bin/bash
# build variants of libvelo_gpu.so on the GPU box (EXPS: ';'-separated nvcc flag sets) and bench each briefly; the default build is restored last
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
IFS=";" read -ra LIST <<< "${EXPS-;}"
for ex in "${LIST[@]}" ""; do
  VELO_NVCC_EXTRA="$ex" python -c "
import importlib; b=importlib.import_module('vision-enhanced-lidar-odometry_b200._build'); b.build_gpu(force=True)" 2>&1 | grep -iE " error|ptxas fatal"
  [ -z "$ex" ] && [ -n "$DONE_DEFAULT" ] && break
  [ -z "$ex" ] && DONE_DEFAULT=1
  timeout ${BENCH_TIMEOUT:-150} python bench.py --frames ${FRAMES:-200} --steps ${STEPS:-3} --no-cpu --no-parity ${BENCH_ARGS} 2>>$O/exp_err.log | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); k=d['kernels']; s=d['icp_search']
    print('[$ex]', 'value', d['value'], 'e2e', d['e2e']['value'], '| icp ms', k['icp_pass']['ms_per_launch'], 'assoc', k['assoc_search']['ms_per_launch'], 'visual', k['visual_residuals']['ms_per_launch'], 'index', k['index_build']['ms_per_launch'], '| cand', s['per_pass_candidates_per_query'], 'rings', s['per_pass_rings_scanned_per_query'])
except Exception as e: print('[$ex] FAILED', e)
" | tee -a $O/exp_variants.log
  if tail -1 $O/exp_variants.log | grep -q FAILED; then echo "stopping: a variant failed or hung (the GPU may be wedged)"; break; fi
done
