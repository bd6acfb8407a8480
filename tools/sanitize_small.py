"""Tiny invocations of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helping_hand_for_egocentric_videos_b200 import ops, synthetic  # noqa: E402
from helping_hand_for_egocentric_videos_b200.model import LaviLa, tfm_decoder, metric  # noqa: E402

torch.manual_seed(0)
for (M, N, K, epi) in [(300, 512, 128, 0), (129, 96, 72, 3), (9601, 256, 64, 1), (260, 31, 64, 3)]:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = torch.randn(N, K, device="cuda").bfloat16()
    ops.gemm_bf16(a, w, torch.randn(N, device="cuda"), epilogue=epi)
for (B, T, n, H) in [(1, 3, 16, 2), (1, 2, 50, 3), (1, 16, 256, 1)]:
    qkv = torch.randn(B * (1 + T * n), 3 * H * 64, device="cuda").bfloat16()
    ops.attention(qkv, B, T, n, H)
vis = LaviLa.SpaceTimeTransformer(img_size=56, patch_size=14, embed_dim=128, depth=1, num_heads=2, num_frames=3,
                                  time_init='zeros', ln_pre=True, act_layer=LaviLa.QuickGELU, num_classes=0)
tr = tfm_decoder.Cross_Attention(d_model=128, nhead=2, num_decoder_layers=2, dim_feedforward=256, normalize_before=True,
                                 return_intermediate_dec=True)
dec = tfm_decoder.ObjDecoder(tr, num_classes=30, num_queries=5, aux_loss=True, pred_traj=True, feature_dim=128,
                             num_frames=3, patches_per_frame=16)
synthetic.randomize_(vis, 1)
synthetic.randomize_(dec, 2)
vis, dec = vis.cuda(), dec.cuda().eval()
_, fmap = vis.forward_features(torch.randn(2, 3, 3, 56, 56, device="cuda"))
out, hs, _, _ = dec(fmap[:, 1:].unflatten(1, (3, 16)))
e = dec.obj_proj(hs[-1])[:, -1]
s = metric.sim_matrix(torch.randn(3, 256, device="cuda"), e)
ops.row_argmax(s)
ops.box_match_cost(torch.rand(10, 4, device="cuda"), torch.rand(3, 4, device="cuda"))
torch.cuda.synchronize()
print("sanitize_small done")
