#!/bin/bash
mkdir -p gpurun_out
HH_B200_LIB=tools/ab/libhh_b200_trace.so timeout 300 python tools/prof_fused.py trace 64 > gpurun_out/c3_trace.log 2>&1
echo "trace rc=$?"; cat gpurun_out/c3_trace.log | tail -12
