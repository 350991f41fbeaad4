#!/bin/bash
# copy-bound regime of velo_gpu_batch_frontend on ONE GPU (every chunk copied k times);
# VELO_FE_TRACE prints per-chunk completion times of the last calls
cd "$(dirname "$0")/.."
for cfg in "1 0" "3 0"; do set -- $cfg
  env VELO_FE_UPLOAD_REPEAT=$1 VELO_FE_TRACE=1 timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu --no-parity 2>gpurun_out/fe_trace_$1_$2.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('repeat $1 no_tail $2: device', d['ms_per_step'], 'ms  e2e', d['e2e']['ms_per_step'], 'ms')" | tee -a gpurun_out/exp_frontend.log
done
