#!/bin/bash
# A/B on one box: alternate bench.py between the in-tree library and a variant (tools/build_variant.sh).
#   tools/ab_bench.sh tools/ab/libhh_b200_NAME.so [rounds]
V=$1; R=${2:-2}
for i in $(seq $R); do
  for lib in "" "$V"; do
    HH_B200_LIB=$lib python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['config']['kernel_ms_per_step']
print('${lib:-default}', round(d['value'],1), 'clips/s', {n:round(v['ms_per_step'],2) for n,v in k.items() if n in ('gemm_qkv','gemm_proj','gemm_fc1','gemm_fc2','layernorm','attn_time','attn_space')}, d['clocks']['sm_mhz'])"
  done
done
