"""CUDA-event timing of the satellite kernels at the headline shapes (64 clips, 16 frames, L/14): temporal / spatial
attention and LayerNorm, against the HBM roofline (algorithmic bytes of SURVEY 8d / MEASURED_PEAKS.json).
  python tools/time_kernels.py [clips] [reps] [T]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from helping_hand_for_egocentric_videos_b200 import ops, _lib as L  # noqa: E402

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
T = int(sys.argv[3]) if len(sys.argv) > 3 else 16
n, H = 256, 16
hbm = 6464.3
try:
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
M = clips * (1 + T * n)
torch.manual_seed(0)
lib = L.load()


def timed(fn):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2]


qkv = (torch.randn(M, 3 * H * 64, device="cuda") * 0.5).bfloat16()
o = torch.empty(M, H * 64, device="cuda", dtype=torch.bfloat16)
res = {}
for kind, name in ((1, "attn_time"), (0, "attn_space")):
    ms = timed(lambda: L.check(lib.hh_attention(L.ptr(qkv), L.ptr(o), clips, T, n, H, kind, L.stream_ptr()), name))
    byt = M * H * 64 * 2 * 4
    res[name] = {"ms": round(ms, 4), "GBps": round(byt / ms / 1e6, 1), "frac_hbm": round(byt / ms / 1e6 / hbm, 3)}
x = torch.randn(M, 1024, device="cuda")
w1 = torch.ones(1024, device="cuda")
ms = timed(lambda: ops.layernorm(x, w1, w1, 1e-6, want_f32=False, want_bf16=True))
byt = M * 1024 * 6
res["ln_f32_to_bf16"] = {"ms": round(ms, 4), "GBps": round(byt / ms / 1e6, 1), "frac_hbm": round(byt / ms / 1e6 / hbm, 3)}
print(json.dumps({"clips": clips, "T": T, "hbm_peak_GBps": hbm, **res}))
