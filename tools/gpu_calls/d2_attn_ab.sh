#!/bin/bash
# attention A/B: stand-alone kernel time + CTA-0 timeline for the in-tree library and every tools/ab variant
for lib in "" $(ls tools/ab/libhh_b200_*.so 2>/dev/null); do
  echo "== ${lib:-default}"
  HH_B200_LIB=$lib python tools/time_kernels.py 64 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('attn_space ms', d['attn_space']['ms'])"
  HH_B200_LIB=$lib HH_ATTN_TRACE=1 python tools/attn_trace.py 2>&1 | grep -E "GHz|softmax.h[01] task [34]|helper task 3|mma task 3" 
done
