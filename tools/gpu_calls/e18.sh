#!/bin/bash
# ncu --set full of the new cross-attention backward kernel and of the final forward lin3 launch (box head)
mkdir -p gpurun_out
timeout 120 ncu --set full --clock-control none --import-source on -k regex:cross_bwd_mma --launch-skip 6 -c 1 -f -o gpurun_out/e18_cross_bwd \
  python tools/prof_train_step.py 64 > gpurun_out/e18_a.log 2>&1; echo "cross rc=$?"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:lin3 --launch-skip 44 -c 1 -f -o gpurun_out/e18_lin3_box0 \
  python tools/prof_train_step.py 64 > gpurun_out/e18_b.log 2>&1; echo "lin3 rc=$?"
ls -la gpurun_out/*.ncu-rep
