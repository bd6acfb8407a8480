#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_ln.py tests/test_gpu_kernels.py -q -x > gpurun_out/c4_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/c4_tests.log
HH_B200_LIB=tools/ab/libhh_b200_trace.so timeout 300 python tools/prof_fused.py trace 64 > gpurun_out/c4_trace.log 2>&1
echo "trace rc=$?"; cat gpurun_out/c4_trace.log | tail -12
timeout 300 python tools/prof_fused.py time 64 10 > gpurun_out/c4_time.log 2>&1
echo "time rc=$?"; head -9 gpurun_out/c4_time.log
