#!/bin/bash
# c4 kernel list (decoder forward / backward), refreshed ncu launch list of the bench command, default bench line
mkdir -p gpurun_out
timeout 600 python tools/prof_c4.py 64 > gpurun_out/e1_prof_c4.log 2>&1; echo "prof_c4 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_r2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/prof_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_r2.csv
timeout 900 python bench.py > gpurun_out/e1_bench.json 2> gpurun_out/e1_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/e1_bench.json
