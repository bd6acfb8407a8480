#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cross or decoder or backward or train or dropout" 2>&1 | tail -15 | tee gpurun_out/e8_tests.log
timeout 600 python tools/prof_c4.py 64 > gpurun_out/e8_prof_c4.log 2>&1; echo "prof_c4 rc=$?"; grep -E "^parts|^==" gpurun_out/e8_prof_c4.log; sed -n '/== backward/,$p' gpurun_out/e8_prof_c4.log | head -12
