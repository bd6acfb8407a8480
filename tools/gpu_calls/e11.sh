#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/e11_suite.log
HH_LIN_LEGACY=1 HH_CROSS_BWD_SIMT=1 timeout 600 python -m pytest tests -m gpu -x -q -k "linear or cross or backward or decoder or train" 2>&1 | tail -2 | tee gpurun_out/e11_alt.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
