"""Probe: does running two half-batches on two streams (LayerNorm / small kernels of one under the GEMMs of the other)
beat one full batch on one stream?  Encoder only, L/14, 16 frames.  python tools/overlap_probe.py [clips] [iters]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
from helping_hand_for_egocentric_videos_b200 import synthetic  # noqa: E402

vis0, _ = bench.build_modules(16, 12)
vis1, _ = bench.build_modules(16, 12)
vis0, vis1 = vis0.cuda(), vis1.cuda()
x = synthetic.synthetic_clips(clips, 16, 224, seed=3, device="cuda")
xa, xb = x[: clips // 2].contiguous(), x[clips // 2:].contiguous()
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()


def one_stream():
    vis0.forward_features(x)


def two_streams():
    with torch.cuda.stream(s0):
        vis0.forward_features(xa)
    with torch.cuda.stream(s1):
        vis1.forward_features(xb)


for name, fn in (("one stream, %d clips" % clips, one_stream), ("two streams, 2 x %d clips" % (clips // 2), two_streams),
                 ("one stream, %d clips" % clips, one_stream), ("two streams, 2 x %d clips" % (clips // 2), two_streams)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    print("%-28s %.1f ms per %d clips = %.1f clips/s" % (name, dt * 1e3, clips, clips / dt), flush=True)
