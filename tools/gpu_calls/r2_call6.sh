#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity_l14.py "tests/test_gpu_model.py::test_encoder_against_reference_golden" -q -x -s > gpurun_out/c6_parity.log 2>&1
echo "parity rc=$?"; grep -E "EgoMCQ|clip [0-9]+:|passed|failed|Error|error" gpurun_out/c6_parity.log | tail -15
HH_B200_LIB=tools/ab/libhh_b200_trace4.so PROF_ONLY=proj_plain,fc2_plain,qkv_plain,fc1_plain timeout 300 python tools/prof_fused.py trace 64 > gpurun_out/c6_trace4.log 2>&1
echo "trace4 rc=$?"; tail -4 gpurun_out/c6_trace4.log
timeout 900 python bench.py > gpurun_out/c6_bench_default.json 2> gpurun_out/c6_bench_default.err
echo "bench rc=$?"; tail -3 gpurun_out/c6_bench_default.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/c6_bench_default.json'))
    print('value',d['value'],'e2e',d.get('e2e',{}).get('value'),'cpu',d.get('cpu_baseline',{}).get('value'))
    print('eager',json.dumps(d.get('gpu_eager_baseline'))[:600])
    print('extra',json.dumps(d.get('extra'))[:2500])
except Exception as e: print('ERR',e)
PY
