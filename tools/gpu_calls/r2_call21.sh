#!/bin/bash
for lib in "" tools/ab/libhh_b200_attn1t.so "" tools/ab/libhh_b200_attn1t.so; do
HH_B200_LIB=$lib timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['config']['kernel_ms_per_step']
print('${lib:-kept}', round(d['value'],1), 'clips/s', {n:round(v['ms_per_step'],2) for n,v in k.items() if n in ('gemm_qkv','gemm_proj','attn_time','attn_space')}, d['clocks']['sm_mhz'])"
done
