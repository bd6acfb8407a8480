#!/bin/bash
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_kernels.py -q -x -k "test_divided_attention" > gpurun_out/c14_attn_quick.log 2>&1
rc=$?; echo "attn quick rc=$rc"; tail -2 gpurun_out/c14_attn_quick.log
if [ $rc -ne 0 ]; then echo "ABORT"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c14_suite.log 2>&1
echo "suite rc=$?"; tail -3 gpurun_out/c14_suite.log
for i in 1 2; do
timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['config']['kernel_ms_per_step']
print(round(d['value'],1), 'clips/s', {n:round(v['ms_per_step'],2) for n,v in k.items() if n in ('gemm_qkv','gemm_proj','gemm_fc1','gemm_fc2','attn_time','attn_space')}, d['clocks']['sm_mhz'])"
done
