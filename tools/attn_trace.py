"""CTA-0 timeline of the spatial attention kernel: HH_ATTN_TRACE=1 python tools/attn_trace.py  (stamps on stderr; summary in
profiles/r2_attention_timeline.md).  16 clips x 16 frames x 16 heads, n = 256."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helping_hand_for_egocentric_videos_b200 import _lib as L  # noqa: E402

B, T, n, H = 16, 16, 256, 16
qkv = (torch.randn(B * (1 + T * n), 3 * H * 64, device="cuda") * 0.5).bfloat16()
o = torch.empty(B * (1 + T * n), H * 64, device="cuda", dtype=torch.bfloat16)
L.check(L.load().hh_attention(L.ptr(qkv), L.ptr(o), B, T, n, H, 0, L.stream_ptr()), "hh_attention")
torch.cuda.synchronize()
