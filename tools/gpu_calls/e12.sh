#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 600 python -m pytest tests -m gpu -x -q -k "backward or train or dropout" 2>&1 | tail -1; done | tee gpurun_out/e12_tests.log
for v in 0 1; do echo "== HH_BWD_SINGLE_STREAM=$v"; HH_BWD_SINGLE_STREAM=$v timeout 300 python tools/prof_c4.py 64 2>&1 | grep -E "^parts|== backward"; done | tee gpurun_out/e12_ab.log
