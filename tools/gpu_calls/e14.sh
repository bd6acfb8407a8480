#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "backward or train or dropout or decoder or model or state" 2>&1 | tail -2 | tee gpurun_out/e14_tests.log
timeout 300 python tools/prof_c4.py 64 > gpurun_out/e14_prof_c4.log 2>&1; grep -E "^parts|^==" gpurun_out/e14_prof_c4.log
timeout 600 python bench.py --config c4 --no-cpu-baseline > gpurun_out/e14_bench_c4.json 2> gpurun_out/e14_bench_c4.err; echo "c4 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/e14_bench_c4.json')); print('c4 value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1)); print(d['config'].get('parts') or d.get('extra'))"
