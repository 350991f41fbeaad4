#!/bin/bash
# on the GPU box: bench every prebuilt variant library exp_libs/libvelo_<tag>.so (tools/build_variants.py) briefly.  TAGS="a b c" restricts / orders.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
LIST=${TAGS:-$(ls exp_libs/libvelo_*.so | sed 's/.*libvelo_//; s/\.so$//')}
for t in $LIST; do
  VELO_GPU_LIB=$PWD/exp_libs/libvelo_$t.so timeout ${BENCH_TIMEOUT:-150} python bench.py --frames ${FRAMES:-200} --steps ${STEPS:-3} --no-cpu ${BENCH_ARGS---no-parity} 2>>$O/exp_err.log | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); k=d['kernels']; s=d['icp_search']
    print('[$t]', 'value', d['value'], 'e2e', d['e2e']['value'], '| icp ms', k['icp_pass']['ms_per_launch'], 'assoc', k['assoc_search']['ms_per_launch'], 'visual', k['visual_residuals']['ms_per_launch'], 'index', k['index_build']['ms_per_launch'], 'parity', d.get('parity_checked'), '| cand', s['per_pass_candidates_per_query'], 'rings', s['per_pass_rings_scanned_per_query'])
except Exception as e: print('[$t] FAILED', e)
" | tee -a $O/exp_variants.log
  if tail -1 $O/exp_variants.log | grep -q FAILED; then echo "stopping: a variant failed or hung (the GPU may be wedged)"; break; fi
done
