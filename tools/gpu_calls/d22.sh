#!/bin/bash
mkdir -p gpurun_out
for lib in "" tools/ab/libhh_b200_noturns.so; do
  echo "== ${lib:-default}"
  HH_B200_LIB=$lib python tools/time_kernels.py 64 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('attn_space ms', d['attn_space']['ms'])"
  HH_B200_LIB=$lib HH_ATTN_TRACE=1 python tools/attn_trace.py 2>&1 | grep -E "GHz|softmax.h[01] task [34]|mma task 3|helper task 3"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_space_tc" -c 2 -f -o gpurun_out/r2_attn_space_final \
  python tools/prof_kernels.py attn 16 1 > gpurun_out/d22_prof.log 2>&1
echo "ncu rc=$?"
