"""One decoder training forward + backward at the c4 per-GPU share (for ncu launch lists)."""
import sys
import torch
sys.path.insert(0, ".")
from helping_hand_for_egocentric_videos_b200 import synthetic
from helping_hand_for_egocentric_videos_b200.model import tfm_decoder as D

B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 4
tr = D.Cross_Attention(normalize_before=True, return_intermediate_dec=True)
model = D.ObjDecoder(transformer=tr, num_classes=22047, num_queries=13, aux_loss=True, pred_traj=True, feature_dim=1024,
                     num_frames=T, patches_per_frame=256)
synthetic.randomize_(model, 1)
model = model.cuda().eval()
grid = torch.randn(B, T, 256, 1024, device="cuda")
for it in range(2):
    mo, hs, _, _ = model(grid)
    loss = hs[-1].square().mean() + mo["pred_boxes"].square().mean()
    loss.backward()
torch.cuda.synchronize()
