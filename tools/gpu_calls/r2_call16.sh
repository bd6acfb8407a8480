#!/bin/bash
mkdir -p gpurun_out
for pace in 0 300 600 1000 1500 2200; do
echo "--- pace $pace ns"
HH_GEMM_RES_PACE_NS=$pace HH_B200_LIB=tools/ab/libhh_b200_trace.so PROF_ONLY=fc2_res_wb,fc2_res_nowb timeout 120 python tools/prof_fused.py trace 64 2>&1 | tail -2 | cut -c1-170
done
