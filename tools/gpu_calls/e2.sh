#!/bin/bash
# pipelined linear family: parity tests, timing against the first-generation kernels, c4 kernel list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "linear or decoder or backward or train or dropout" 2>&1 | tail -8 | tee gpurun_out/e2_tests.log
python tools/time_lin3.py 2>&1 | tee gpurun_out/e2_time_new.log
HH_LIN_LEGACY=1 python tools/time_lin3.py 2>&1 | tee gpurun_out/e2_time_legacy.log
timeout 600 python tools/prof_c4.py 64 > gpurun_out/e2_prof_c4.log 2>&1; echo "prof_c4 rc=$?"; grep -E "^parts|^==|linear|lin3" gpurun_out/e2_prof_c4.log
