#!/bin/bash
# final single-GPU evidence: full suite, smoke, default bench line, c1 / c3 / c4 lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/e15_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/e15_smoke.log
timeout 900 python bench.py > gpurun_out/e15_bench.json 2> gpurun_out/e15_bench.err; echo "bench rc=$?"
for c in c1 c3 c4; do timeout 600 python bench.py --config $c --no-cpu-baseline > gpurun_out/e15_bench_$c.json 2> gpurun_out/e15_bench_$c.err; echo "$c rc=$?"; done
python - <<'PY'
import json
for f in ("e15_bench","e15_bench_c1","e15_bench_c3","e15_bench_c4"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f,'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),d['clocks']['sm_mhz'])
        if 'extra' in d: print('  c4',d['extra'].get('c4_train_step'),'\n  q',d['extra'].get('egomcq_question_ms'))
    except Exception as e: print(f,'ERR',e)
PY
