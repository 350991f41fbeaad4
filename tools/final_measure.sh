#!/bin/bash
# round-end measurement refresh on the GPU box: both bench arms, ncu launch list, ncu captures (reports are reduced to CSV on the
# box: gpurun only brings back 64 MiB).  tools/make_profiles.py turns gpurun_out/final_* into the committed profiles/ files.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/final_pytest.txt; fi
timeout 600 python bench.py 2> $O/final_bench.err | tail -1 > $O/final_bench.json
timeout 300 python bench.py --pose-spread spec --no-cpu 2>> $O/final_bench.err | tail -1 > $O/final_bench_spec.json
timeout 300 python bench.py --scan-format xyz --no-cpu --no-parity 2>> $O/final_bench.err | tail -1 > $O/final_bench_xyz.json
timeout 300 python bench.py --rig 1 --features 8000 --frames 200 --no-cpu 2>> $O/final_bench.err | tail -1 > $O/final_bench_offroad.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2> $O/final_ref.err | tail -1 > $O/final_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-parity > /dev/null 2>&1
if [ -z "$SKIP_ICP" ]; then timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_icp_pass -c 1 -o $O/final_icp python bench.py --frames 1000 --steps 1 --warmup 0 --no-cpu --no-parity > /dev/null 2>&1; fi
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section WarpStateStats --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 40 -o /tmp/final_all python bench.py --frames 200 --steps 1 --warmup 0 --no-cpu --no-parity > /dev/null 2>&1
ncu -i /tmp/final_all.ncu-rep --page raw --csv > $O/final_all_raw.csv 2>/dev/null
timeout 300 python tools/bench_next_rows.py > $O/final_next_rows.jsonl 2>/dev/null
cat $O/final_pytest.txt 2>/dev/null; head -c 300 $O/final_bench.json; echo; ls -la $O | tail -12; du -sh $O
