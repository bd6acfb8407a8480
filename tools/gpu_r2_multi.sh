#!/bin/bash
# multi-GPU call: the collective checked on real GPUs + the N-rank bench line
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python -m pytest tests/test_gpu_multi.py -q -x -s > gpurun_out/multi_test_n$N.log 2>&1
echo "multi test rc=$?"; tail -4 gpurun_out/multi_test_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'rank spread',d['rank_ms_per_step'],'collective',d.get('collective'))
print('cpu_baseline' in d, d['clocks'])
PY
