#!/bin/bash
mkdir -p gpurun_out
HH_B200_LIB=tools/ab/libhh_b200_trace.so PROF_ONLY=proj_plain,proj_res_nowb,proj_res_wb,fc2_plain,fc2_res_wb,fc2_res_nowb timeout 120 python tools/prof_fused.py trace 64 > gpurun_out/c15_trace.log 2>&1; cat gpurun_out/c15_trace.log | tail -6 | cut -c1-200
echo "--- no residual loads"
HH_B200_LIB=tools/ab/libhh_b200_tracenoload.so PROF_ONLY=proj_res_nowb,proj_res_wb,fc2_res_wb,fc2_res_nowb timeout 120 python tools/prof_fused.py trace 64 > gpurun_out/c15_trace_noload.log 2>&1; cat gpurun_out/c15_trace_noload.log | tail -4 | cut -c1-200
