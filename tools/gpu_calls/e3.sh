#!/bin/bash
# full ncu capture of five pipelined-linear launches of one decoder training forward + backward (indices from the launch list)
mkdir -p gpurun_out
cap() {  # name, skip
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:lin3 --launch-skip $2 -c 1 -f -o gpurun_out/e3_$1 \
    python tools/prof_train_step.py 64 > gpurun_out/e3_b.log 2>&1; echo "$1 rc=$?"
}
cap fwd_small 2
cap fwd_ffn2 7
cap fwd_box0 44
cap wgrad_box 49
cap dgrad_box 50
ls -la gpurun_out
