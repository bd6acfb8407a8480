#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_fused_ln.py tests/test_gpu_kernels.py -q -x -k "gemm" > gpurun_out/c10_tests.log 2>&1
rc=$?; echo "gemm tests rc=$rc"; tail -3 gpurun_out/c10_tests.log
if [ $rc -ne 0 ]; then echo ABORT; exit 1; fi
HH_B200_LIB=tools/ab/libhh_b200_trace.so timeout 120 python tools/prof_fused.py trace 64 > gpurun_out/c10_trace.log 2>&1; cat gpurun_out/c10_trace.log | tail -9 | cut -c1-260
bash tools/ab_bench.sh tools/ab/libhh_b200_roleslo.so 2 2>&1 | tee gpurun_out/c10_ab.log
