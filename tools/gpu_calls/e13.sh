#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 600 python -m pytest tests -m gpu -x -q -k "backward or train or dropout or cross or decoder" 2>&1 | tail -1; done | tee gpurun_out/e13_tests.log
timeout 300 python tools/prof_c4.py 64 > gpurun_out/e13_prof_c4.log 2>&1; grep -E "^parts|^==" gpurun_out/e13_prof_c4.log; sed -n '/== backward/,$p' gpurun_out/e13_prof_c4.log | head -14
