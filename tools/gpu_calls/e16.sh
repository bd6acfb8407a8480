#!/bin/bash
# 2-GPU sanity after the decoder rework: collective tests, c2 and c4 bench lines at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2 | tee gpurun_out/e16_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/e16_bench_n2.json 2> gpurun_out/e16_bench_n2.err; echo "n2 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c4 --steps 10 --warmup 3 > gpurun_out/e16_bench_c4_n2.json 2> gpurun_out/e16_bench_c4_n2.err; echo "c4 n2 rc=$?"
python - <<'PY'
import json
for f in ("e16_bench_n2","e16_bench_c4_n2"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f,'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),d.get('rank_ms_per_step'),d.get('collective'))
    except Exception as e: print(f,'ERR',e)
PY
