#!/bin/bash
# full GPU suite, alternative-path runs of the new kernels, default bench line, c1 / c4 lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/e10_suite.log
HH_LIN_LEGACY=1 HH_CROSS_BWD_SIMT=1 timeout 600 python -m pytest tests -m gpu -x -q -k "linear or cross or backward or decoder or train" 2>&1 | tail -2 | tee gpurun_out/e10_alt.log
timeout 900 python bench.py > gpurun_out/e10_bench.json 2> gpurun_out/e10_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --config c4 --no-cpu-baseline > gpurun_out/e10_bench_c4.json 2> gpurun_out/e10_bench_c4.err; echo "c4 rc=$?"
timeout 600 python bench.py --config c1 --no-cpu-baseline > gpurun_out/e10_bench_c1.json 2> gpurun_out/e10_bench_c1.err; echo "c1 rc=$?"
python - <<'PY'
import json
for f in ("e10_bench","e10_bench_c4","e10_bench_c1"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f,'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1))
        if 'extra' in d: print('  c4',d['extra'].get('c4_train_step'),'\n  q',d['extra'].get('egomcq_question'))
    except Exception as e: print(f,'ERR',e)
PY
