#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-400
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'u8',round(d['e2e_u8']['value'],1),'cpu',d['cpu_baseline']['value'])
print('roofline',{k:(round(v,3) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k in('achieved','frac','share_of_step','traffic')})
for k,v in d['config']['kernel_ms_per_step'].items(): print(' ',k,{a:round(b,3) for a,b in v.items()})
print('path frac', d['config']['path_frac_of_sustained_peak'], d['config']['path_frac_of_burst_peak'])
print('eager', d['gpu_eager_baseline'].get('speedup_vs_bf16_autocast'))
print('extra', {k:v for k,v in d['extra'].items() if k.endswith('_ms')})
print('question', d['extra']['egomcq_question'])
PY
