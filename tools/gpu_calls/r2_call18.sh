#!/bin/bash
for i in 1 2; do
python tools/time_small_m.py 2>&1 | tail -1
HH_GEMM_NO_WAVE_RULE=1 python tools/time_small_m.py 2>&1 | tail -1
done
