#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "linear or decoder or backward or train or dropout or cross" 2>&1 | tail -3 | tee gpurun_out/e6_tests.log
timeout 600 python tools/prof_c4.py 64 > gpurun_out/e6_prof_c4.log 2>&1; echo "prof_c4 rc=$?"; grep -E "^parts|^==|lin3" gpurun_out/e6_prof_c4.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/e6_launches_train.csv \
  python tools/prof_train_step.py 64 > gpurun_out/e6_a.log 2>&1; echo "list rc=$?"
