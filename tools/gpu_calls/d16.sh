#!/bin/bash
for lib in "" $(ls tools/ab/libhh_b200_*.so 2>/dev/null); do
  echo "== ${lib:-default}"
  HH_B200_LIB=$lib python tools/time_kernels.py 64 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('attn_space ms', d['attn_space']['ms'])"
  HH_B200_LIB=$lib HH_ATTN_TRACE=1 python tools/attn_trace.py 2>&1 | grep -E "GHz|softmax.h[01] task [34]|mma task 3|helper task [34]"
  HH_B200_LIB=$lib python -m pytest tests/test_gpu_kernels.py -q -x -k "divided_attention or deterministic_under" 2>&1 | tail -2
done
