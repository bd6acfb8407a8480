"""PyTorch-eager GPU baseline beside the kernels (SURVEY.md section 8d 'GPU baseline to beat'): the reference's algorithm as
plain torch ops (the oracle restatement -- /root/reference itself does not exist on the GPU box) on the same B200, fp32 and
under bf16 autocast, L/14 + decoder nq=12, 16 frames.  Measurement utility only; nothing in the package imports it.
    python tools/torch_eager_gpu_baseline.py [clips]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from oracle import hh_oracle as O  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    T, nq = 16, 12
    torch.cuda.set_device(0)
    vsd = O.synth_state_dict(O.encoder_param_shapes(1024, 24, 14, 256, T), 0)
    dsd = O.synth_state_dict(O.decoder_param_shapes(512, nq + 1, 256, T, 1024, 22048, layers=6, ffn=2048, pred_traj=False), 1)
    vsd = {k: v.cuda() for k, v in vsd.items()}
    dsd = {k: v.cuda() for k, v in dsd.items()}
    video = torch.randn(B, T, 3, 224, 224, device="cuda")
    text = torch.randn(13, 256, device="cuda")

    def step():
        _, fmap = O.encoder_forward(video, vsd, 16)
        _, hs, _, _ = O.decoder_forward(fmap[:, 1:].unflatten(1, (T, 256)).float(), dsd, heads=8, pred_traj=False)
        vid = O.obj_proj(hs[-1], dsd)[:, -1]
        return O.sim_matrix(text, vid.float()).argmax(-1)

    res = {}
    for name, ctx in (("fp32 (TF32 off)", torch.autocast("cuda", enabled=False)),
                      ("bf16 autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        with torch.no_grad(), ctx:
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                step()
            e1.record()
            torch.cuda.synchronize()
        res[name] = {"ms_per_step": e0.elapsed_time(e1) / 3, "clips_per_s": B / (e0.elapsed_time(e1) / 3) * 1e3}
    print(json.dumps({"baseline": "torch eager ops of the oracle restatement on one B200", "clips": B, "frames": T, **res}))


if __name__ == "__main__":
    main()
