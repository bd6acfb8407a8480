#!/bin/bash
# full GPU suite + default bench line after the attention rework
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/d15_suite.log
timeout 600 python bench.py > gpurun_out/d15_bench.json 2> gpurun_out/d15_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/d15_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/d15_bench.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1))
for k,v in d['config']['kernel_ms_per_step'].items(): print(' ',k,{a:round(b,3) for a,b in v.items()})
print('clocks', d['clocks'])
PY
