#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_space_tc" -c 2 -f -o gpurun_out/d20_attn_z \
  python tools/prof_kernels.py attn 16 1 > gpurun_out/d20_prof.log 2>&1
echo "rc=$?"; ls -la gpurun_out/d20_attn_z.ncu-rep
