#!/bin/bash
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_kernels.py -q -x -k "test_divided_attention" > gpurun_out/c12_attn_quick.log 2>&1
rc=$?; echo "attn quick rc=$rc"; tail -3 gpurun_out/c12_attn_quick.log
if [ $rc -ne 0 ]; then echo "ABORT: attention kernel broken"; exit 1; fi
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" > gpurun_out/c12_attn_tests.log 2>&1
echo "attn tests rc=$?"; tail -2 gpurun_out/c12_attn_tests.log
HH_ATTN_TRACE=1 timeout 60 python tools/attn_trace.py > gpurun_out/c12_attn_trace.log 2>&1; grep "attn trace" gpurun_out/c12_attn_trace.log | grep -E "producer task [2-5]|mma task [2-5]|helper task [2-5]|softmax.h. task [2-5]" | head -24
timeout 120 python tools/time_kernels.py > gpurun_out/c12_time_kernels.log 2>&1; tail -1 gpurun_out/c12_time_kernels.log
