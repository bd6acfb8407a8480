#!/bin/bash
# c4 under torchrun at N = 2 after the fix of the rank-0-only profiled step (short, bounded)
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c4 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/e17_bench_c4_n2.json 2> gpurun_out/e17_bench_c4_n2.err; echo "c4 n2 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/e17_bench_c4_n2.json')); print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',d['e2e'] and round(d['e2e']['value'],1), d.get('rank_ms_per_step'))"
