#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" > gpurun_out/c8_attn_tests.log 2>&1
echo "attn tests rc=$?"; tail -5 gpurun_out/c8_attn_tests.log
HH_ATTN_TRACE=1 timeout 120 python tools/attn_trace.py > gpurun_out/c8_attn_trace.log 2>&1; grep "attn trace" gpurun_out/c8_attn_trace.log | grep -E "mma task [1-5]|softmax.h. task [1-5]|producer" | head -40
timeout 300 python tools/time_kernels.py > gpurun_out/c8_time_kernels.log 2>&1; tail -12 gpurun_out/c8_time_kernels.log
timeout 600 python -m pytest tests/test_gpu_parity_l14.py tests/test_gpu_model.py -q -x > gpurun_out/c8_model_tests.log 2>&1
echo "model tests rc=$?"; tail -5 gpurun_out/c8_model_tests.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c8_bench.json')); k=d['config']['kernel_ms_per_step']
print(round(d['value'],1), round(d['ms_per_step'],2), {n:round(v['ms_per_step'],2) for n,v in k.items()}, d['clocks']['sm_mhz'])
PY
