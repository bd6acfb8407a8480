#!/bin/bash
# round-2 call 1: fused-LayerNorm GEMM epilogues -- kernel tests, full GPU suite, A/B bench on one box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_fused_ln.py -x -q > gpurun_out/c1_fused_tests.log 2>&1
echo "fused tests rc=$?"
tail -15 gpurun_out/c1_fused_tests.log
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fused_ln.py > gpurun_out/c1_gpu_suite.log 2>&1
echo "suite rc=$?"
tail -8 gpurun_out/c1_gpu_suite.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c1_bench_fused.json 2> gpurun_out/c1_bench_fused.err
echo "bench fused rc=$?"; cat gpurun_out/c1_bench_fused.json | head -c 3000
HH_LN_UNFUSED=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c1_bench_unfused.json 2> gpurun_out/c1_bench_unfused.err
echo "bench unfused rc=$?"; cat gpurun_out/c1_bench_unfused.json | head -c 3000
