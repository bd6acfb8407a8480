"""Timings of the SURVEY section 8f rows (text tower, training step pieces, matcher, retrieval metrics) on one GPU, with
the CPU reference (oracle / numpy / scipy) beside each.  Writes one JSON object per row to stdout.
    python tools/bench_next_rows.py > gpurun_out/next_rows.jsonl"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from helping_hand_for_egocentric_videos_b200 import ops, synthetic  # noqa: E402
from helping_hand_for_egocentric_videos_b200.model import LaviLa, box_utils  # noqa: E402
from helping_hand_for_egocentric_videos_b200.utils import mAP, nDCG  # noqa: E402
from oracle import golden_cases as gc  # noqa: E402
from oracle import hh_oracle as O  # noqa: E402


def gpu_ms(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    torch.cuda.set_device(0)
    g = torch.Generator().manual_seed(0)
    # ---- 1. text tower: LARGE geometry (768 x 12 heads x 12 layers), 1024 captions per call
    clip = LaviLa.CLIP_OPENAI_TIMESFORMER_LARGE(num_frames=4)
    synthetic.randomize_(clip, 0)
    clip = clip.cuda().eval()
    G = 1024
    tokens = gc.make_tokens(G, 49408, g).cuda()
    ms = gpu_ms(lambda: clip.encode_text(tokens))
    fl = clip.text_flops_per_sequence()
    sd = {k: v.detach().cpu() for k, v in clip.state_dict().items() if not k.startswith("visual.")}
    t0 = time.perf_counter()
    with torch.no_grad():
        O.text_forward(tokens[:64].cpu(), sd, 12)
    cpu = (time.perf_counter() - t0) / 64
    print(json.dumps({"row": "f-1 text tower", "captions_per_call": G, "ms_per_call": ms, "captions_per_s": G / ms * 1e3,
                      "tflops": fl * G / ms / 1e9, "cpu_oracle_captions_per_s": 1 / cpu, "cpu_cores": torch.get_num_threads()}))
    del clip
    # ---- 2. matcher: 256 images x 10 queries, 0-4 targets (c4 share)
    bs, Q = 256, 10

    def rb(k):
        return torch.cat([0.2 + 0.6 * torch.rand(k, 2, generator=g), 0.02 + 0.35 * torch.rand(k, 2, generator=g)], -1)
    pb = rb(bs * Q).view(bs, Q, 4)
    sizes = torch.randint(0, 5, (bs,), generator=g).tolist()
    tb = [rb(k) for k in sizes]
    m = box_utils.build_matcher(None)
    outs = {"pred_logits": torch.zeros(bs, Q, 8).cuda(), "pred_boxes": pb.cuda()}
    tg = [{"boxes": b.cuda(), "labels": torch.zeros(len(b)).cuda()} for b in tb]
    ms = gpu_ms(lambda: m(outs, tg, exclude_class=True))
    t0 = time.perf_counter()
    for _ in range(5):
        O.hungarian_match(pb, tb)
    cpu = (time.perf_counter() - t0) / 5 * 1e3
    print(json.dumps({"row": "a17 HungarianMatcher (256 images x 10 queries)", "ms_gpu_incl_index_d2h": ms,
                      "ms_cpu_oracle_scipy": cpu}))
    # ---- 3. retrieval metrics at EPIC-MIR size
    N = 9668
    sim, rel = gc.synth_retrieval(N, N, 1)
    t0 = time.perf_counter()
    v1 = mAP.calculate_mAP(sim, rel)
    v2 = nDCG.calculate_nDCG(sim, rel)
    torch.cuda.synchronize()
    gpu_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    w1 = O.calculate_mAP(sim, rel)[0]
    w2 = O.calculate_nDCG(sim, rel)[0]
    cpu_s = time.perf_counter() - t0
    sd_, rd_ = torch.from_numpy(sim).cuda(), torch.from_numpy(rel).cuda()
    kms = gpu_ms(lambda: ops.retrieval_rows(sd_, rd_, 0), iters=2, warm=1)
    print(json.dumps({"row": "f-4 mAP + nDCG, 9668 x 9668 float64", "s_gpu_incl_h2d_of_1.5GB": gpu_s, "s_cpu_numpy": cpu_s,
                      "ap_kernel_ms_resident": kms, "equal_mAP": bool(v1 == w1), "equal_nDCG": bool(v2 == w2)}))


if __name__ == "__main__":
    main()
