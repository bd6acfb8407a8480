"""Launches each hot kernel a few times at the headline shapes (for ncu / launch lists). Not a benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helping_hand_for_egocentric_videos_b200 import ops, _lib as L  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
clips = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
M = clips * 4097
torch.manual_seed(0)
if which in ("gemm", "all"):
    for (N, K, epi) in [(3072, 1024, 0), (1024, 1024, 0), (4096, 1024, 1), (1024, 4096, 0)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") / 32).bfloat16()
        bias = torch.randn(N, device="cuda")
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for _ in range(reps):
            ops.gemm_bf16(a, w, bias, epilogue=epi, out=out)
if which in ("attn", "all"):
    B, T, n, H = clips, 16, 256, 16
    qkv = (torch.randn(B * (1 + T * n), 3 * H * 64, device="cuda") * 0.5).bfloat16()
    o = torch.empty(B * (1 + T * n), H * 64, device="cuda", dtype=torch.bfloat16)
    lib = L.load()
    for _ in range(reps):
        for kind in (0, 1, 2):
            L.check(lib.hh_attention(L.ptr(qkv), L.ptr(o), B, T, n, H, kind, L.stream_ptr()), "attn")
if which in ("ln", "all"):
    x = torch.randn(M, 1024, device="cuda")
    w1 = torch.ones(1024, device="cuda")
    for _ in range(reps):
        ops.layernorm(x, w1, w1, 1e-6, want_f32=False, want_bf16=True)
torch.cuda.synchronize()
print("done", which)
