#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/c7_suite.log 2>&1
echo "suite rc=$?"; grep -E "EgoMCQ L/14|clip [0-9]+:|passed|failed|Error" gpurun_out/c7_suite.log | tail -15
python - <<'PY' > gpurun_out/c7_sim.log 2>&1
import torch, sys
sys.path.insert(0,'.')
import bench
print(bench.extra_c3_sim(torch.device('cuda',0), bench.load_peaks()))
from helping_hand_for_egocentric_videos_b200 import ops
a=torch.randn(9728,256,device='cuda'); b=torch.randn(9728,256,device='cuda')
s=ops.sim_matrix(a,b)
ref=torch.nn.functional.normalize(a.double(),dim=-1)@torch.nn.functional.normalize(b.double(),dim=-1).t()
print('max err vs fp64', (s.double()-ref).abs().max().item())
s2=ops.sim_matrix(a,a); ref2=torch.nn.functional.normalize(a.double(),dim=-1); ref2=ref2@ref2.t()
print('self-sim max err', (s2.double()-ref2).abs().max().item())
PY
cat gpurun_out/c7_sim.log | tail -5
HH_ATTN_TRACE=1 python tools/attn_trace.py > gpurun_out/c7_attn_trace.log 2>&1; grep "attn trace" gpurun_out/c7_attn_trace.log | head -60
