#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/c5_bench_fused_$i.json 2> gpurun_out/c5_bench_fused_$i.err
echo "bench fused rc=$?"
HH_LN_UNFUSED=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/c5_bench_unfused_$i.json 2> gpurun_out/c5_bench_unfused_$i.err
echo "bench unfused rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c5_bench_*.json')):
    try:
        d=json.load(open(f)); k=d['config']['kernel_ms_per_step']
        print(f.split('/')[-1], round(d['value'],1), round(d['ms_per_step'],2), {n:round(v['ms_per_step'],2) for n,v in k.items() if n.startswith('gemm') or n in('layernorm','attn_time','attn_space')}, d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 gpurun_out/c5_bench_fused_1.err
