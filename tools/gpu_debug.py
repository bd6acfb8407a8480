"""Developer diagnostics for a GPU box: prints error patterns and timings kernel by kernel (not a test)."""
import math
import sys
import time
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helping_hand_for_egocentric_videos_b200 import ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
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


def gemm_check(M, N, K, epi=3):
    g = torch.Generator().manual_seed(0)
    a = torch.randn(M, K, generator=g).cuda().bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda().bfloat16()
    ref = a.float() @ w.float().t()
    try:
        out = ops.gemm_bf16(a, w, None, epilogue=epi).float()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print("GEMM %dx%dx%d epi%d EXC %s" % (M, N, K, epi, e))
        return False
    err = (out - ref).abs()
    ok = err.max().item() < 2e-2
    print("GEMM %dx%dx%d epi%d max err %.3e ref max %.2f %s" % (M, N, K, epi, err.max().item(), ref.abs().max().item(),
                                                               "OK" if ok else "MISMATCH"))
    if not ok:
        bad = err > 2e-2
        print("  bad frac %.4f; bad rows(first 16 of 8-row groups): %s" % (
            bad.float().mean().item(), bad.any(1).view(-1)[:128].int().tolist()))
        print("  bad cols (first 128): %s" % bad.any(0)[:128].int().tolist())
        print("  out[0,:8]", out[0, :8].tolist(), "\n  ref[0,:8]", ref[0, :8].tolist())
        print("  out[1,:8]", out[1, :8].tolist(), "\n  ref[1,:8]", ref[1, :8].tolist())
        print("  out[8,:8]", out[8, :8].tolist(), "\n  ref[8,:8]", ref[8, :8].tolist())
    return ok


def main():
    print(torch.cuda.get_device_name(0))
    ok = gemm_check(128, 128, 64)
    ok &= gemm_check(128, 256, 64)
    ok &= gemm_check(128, 256, 128)
    ok &= gemm_check(256, 512, 1024)
    ok &= gemm_check(300, 3072, 1024, epi=0)
    if not ok:
        print("GEMM broken; stopping")
        return
    # timings on the four encoder contractions at M = 16 clips x 4097 tokens
    M = 16 * 4097
    for (N, K, epi, nm) in [(3072, 1024, 0, "qkv"), (1024, 1024, 2, "proj"), (4096, 1024, 1, "fc1"), (1024, 4096, 2, "fc2")]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") / 32).bfloat16()
        bias = torch.randn(N, device="cuda")
        res = torch.randn(M, N, device="cuda") if epi == 2 else None
        out = torch.empty(M, N, device="cuda", dtype=torch.float32 if epi >= 2 else torch.bfloat16)
        ms = timeit(lambda: ops.gemm_bf16(a, w, bias, epilogue=epi, residual=res, out=out))
        ms_t = timeit(lambda: torch.matmul(a, w.t()))
        print("GEMM %-4s M=%d N=%d K=%d: %.3f ms  %.1f TFLOP/s   (torch.matmul bf16: %.3f ms %.1f TF/s)" % (
            nm, M, N, K, ms, 2.0 * M * N * K / ms / 1e9, ms_t, 2.0 * M * N * K / ms_t / 1e9))
    # attention timings
    B, T, n, H = 16, 16, 256, 16
    N_ = 1 + T * n
    qkv = (torch.randn(B * N_, 3 * H * 64, device="cuda") * 0.5).bfloat16()
    o = torch.empty(B * N_, H * 64, device="cuda", dtype=torch.bfloat16)
    from helping_hand_for_egocentric_videos_b200 import _lib as L
    lib = L.load()
    for kind, nm in ((0, "space"), (1, "time"), (2, "cls")):
        ms = timeit(lambda: L.check(lib.hh_attention(L.ptr(qkv), L.ptr(o), B, T, n, H, kind, L.stream_ptr()), "attn"))
        gb = (qkv.numel() + o.numel()) * 2 / 1e9
        print("attention %-5s B=%d: %.3f ms  (%.0f GB/s on q,k,v in + o out)" % (nm, B, ms, gb / ms * 1e3))
    x = torch.randn(B * N_, 1024, device="cuda")
    w1 = torch.ones(1024, device="cuda")
    ms = timeit(lambda: ops.layernorm(x, w1, w1, 1e-6, want_f32=False, want_bf16=True))
    print("layernorm M=%d: %.3f ms (%.0f GB/s)" % (B * N_, ms, x.numel() * 6 / 1e6 / ms))


if __name__ == "__main__":
    main()
