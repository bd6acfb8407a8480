#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_fused_ln.py -q -x > gpurun_out/c19_tests.log 2>&1
rc=$?; echo "fused tests rc=$rc"; tail -2 gpurun_out/c19_tests.log
if [ $rc -ne 0 ]; then echo ABORT; exit 1; fi
HH_B200_LIB=tools/ab/libhh_b200_trace.so PROF_ONLY=fc2_plain,fc2_res_wb,fc2_res_nowb timeout 120 python tools/prof_fused.py trace 64 2>&1 | tail -3 | cut -c1-170
HH_GEMM_NO_LONGK=1 HH_B200_LIB=tools/ab/libhh_b200_trace.so PROF_ONLY=fc2_res_wb,fc2_res_nowb timeout 120 python tools/prof_fused.py trace 64 2>&1 | tail -2 | cut -c1-170
for v in "" 1 "" 1; do
HH_GEMM_NO_LONGK=$v timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['config']['kernel_ms_per_step']
print('no_longk=${v:-0}', round(d['value'],1), 'clips/s', {n:round(v['ms_per_step'],2) for n,v in k.items() if n in ('gemm_qkv','gemm_proj','gemm_fc1','gemm_fc2')}, d['clocks']['sm_mhz'])"
done
