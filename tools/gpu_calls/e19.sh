#!/bin/bash
# CUDA decoder against the new reference fixtures at the c1 / c4 decoder geometries
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "reference_golden" 2>&1 | tail -3 | tee gpurun_out/e19_tests.log
