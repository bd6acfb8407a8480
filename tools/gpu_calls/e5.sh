#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "linear or decoder or backward or train or dropout or cross" 2>&1 | tail -3 | tee gpurun_out/e5_tests.log
timeout 600 python tools/prof_c4.py 64 > gpurun_out/e5_prof_c4.log 2>&1; echo "prof_c4 rc=$?"; grep -E "^parts|^==" gpurun_out/e5_prof_c4.log; sed -n '/== backward/,$p' gpurun_out/e5_prof_c4.log | head -14
cap() {  # name, skip
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:lin3 --launch-skip $2 -c 1 -f -o gpurun_out/e5_$1 \
    python tools/prof_train_step.py 64 > gpurun_out/e5_b.log 2>&1; echo "$1 rc=$?"
}
cap fwd_small 2
cap fwd_box0 44
