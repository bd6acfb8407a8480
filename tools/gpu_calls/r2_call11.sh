#!/bin/bash
mkdir -p gpurun_out
HH_B200_LIB=tools/ab/libhh_b200_hints.so timeout 120 python -m pytest tests/test_gpu_fused_ln.py -q -x > gpurun_out/c11_tests.log 2>&1
rc=$?; echo "fused tests (hints) rc=$rc"; tail -2 gpurun_out/c11_tests.log
if [ $rc -ne 0 ]; then echo ABORT; exit 1; fi
HH_B200_LIB=tools/ab/libhh_b200_tracehints.so PROF_ONLY=proj_res_nowb,proj_res_wb,fc2_res_wb timeout 120 python tools/prof_fused.py trace 64 > gpurun_out/c11_trace.log 2>&1; cat gpurun_out/c11_trace.log | tail -3 | cut -c1-260
for lib in "" tools/ab/libhh_b200_hints.so "" tools/ab/libhh_b200_hints.so; do
HH_B200_LIB=$lib timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['config']['kernel_ms_per_step']
print('${lib:-default}', round(d['value'],1), 'clips/s', {n:round(v['ms_per_step'],2) for n,v in k.items() if n in ('gemm_qkv','gemm_proj','gemm_fc1','gemm_fc2','attn_time','attn_space')}, d['clocks']['sm_mhz'])"
done
for c in c1 c3 c4; do
timeout 400 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c11_bench_$c.json 2> gpurun_out/c11_bench_$c.err
echo "bench $c rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/c11_bench_$c.json')); print('$c', d['metric'], round(d['value'],1), round(d['ms_per_step'],2), 'e2e', d.get('e2e',{}).get('value'), 'roof', round(d['roofline']['frac'],3), d['config'].get('parts_ms'))"
done
