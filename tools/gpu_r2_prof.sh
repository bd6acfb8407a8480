#!/bin/bash
# round-2 profiling call: ncu launch list of the bench command, full captures of the GEMM variants and the attention kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_r2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/prof_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_r2.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 12 -f -o gpurun_out/r2_gemm_full \
  python tools/prof_fused.py ncu 64 > gpurun_out/prof_gemm.log 2>&1
echo "gemm full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_space_tc|attn_time_v2" -c 4 -f -o gpurun_out/r2_attn_full \
  python tools/prof_kernels.py attn 16 2 > gpurun_out/prof_attn.log 2>&1
echo "attn full rc=$?"
ls -la gpurun_out/*.ncu-rep
