#!/bin/bash
# CTAs per frame pair of k_icp_pass at the bench size
cd "$(dirname "$0")/.."
for c in ${CTAS:-9 32}; do
  python bench.py --frames ${FRAMES:-1000} --steps 2 --warmup 3 --no-cpu --icp-ctas $c 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ctas $c:', d['value'], 'frames/s  icp ms', d['kernels']['icp_pass']['ms_per_launch'])"
done
