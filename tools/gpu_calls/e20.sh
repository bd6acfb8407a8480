#!/bin/bash
# wgrad row-split heuristic A/B (rows per CTA of the split)
mkdir -p gpurun_out
for v in 128 64 256; do echo "== HH_LIN3_WGRAD_ROWS=$v"; HH_LIN3_WGRAD_ROWS=$v timeout 100 python tools/prof_c4.py 64 2>&1 | grep -E "^parts|lin3_kernel<2>"; done | tee gpurun_out/e20_ab.log
