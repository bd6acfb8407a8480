#!/bin/bash
# round-2 call 2: state check after restore (full GPU suite), stand-alone timing + ncu full capture of the fused-LN GEMMs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c2_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c2_gpu_suite.log 2>&1
echo "suite rc=$?"; tail -5 gpurun_out/c2_gpu_suite.log
timeout 300 python tools/prof_fused.py time 64 10 > gpurun_out/c2_fused_time.log 2>&1
echo "time rc=$?"; cat gpurun_out/c2_fused_time.log | tail -12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 12 -f -o gpurun_out/r2_fused_gemm \
  python tools/prof_fused.py ncu 64 > gpurun_out/c2_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/c2_ncu.log
ls -la gpurun_out | head -30
