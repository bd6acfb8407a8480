"""Steady-state power / clock / throughput of each hot kernel class (NVML sampling while the kernel loops).
Energy per unit of work is what bounds a power-capped step: E_step = sum(P_k * t_k)."""
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helping_hand_for_egocentric_videos_b200 import ops, _lib as L  # noqa: E402

import pynvml

pynvml.nvmlInit()
H = pynvml.nvmlDeviceGetHandleByIndex(0)


def run(name, fn, units, unit_name, secs=2.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    samples = []
    stop = [False]

    def sampler():
        while not stop[0]:
            samples.append((pynvml.nvmlDeviceGetPowerUsage(H) / 1000.0,
                            pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM)))
            time.sleep(0.02)
    th = threading.Thread(target=sampler)
    th.start()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 0
    while time.perf_counter() - t0 < secs:
        for _ in range(10):
            fn()
        iters += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    stop[0] = True
    th.join()
    ms = e0.elapsed_time(e1) / iters
    half = samples[len(samples) // 2:]
    pw = sum(s[0] for s in half) / len(half)
    clk = sum(s[1] for s in half) / len(half)
    print("%-26s %8.3f ms/iter  %9.1f %s  power %6.0f W  sm %4.0f MHz  -> %8.3f J/iter" % (
        name, ms, units / ms / 1e9 if unit_name == "TFLOP/s" else units / ms / 1e6, unit_name, pw, clk, pw * ms / 1e3),
        flush=True)


def main():
    clips = 64
    M = clips * 4097
    for (N, K, epi, nm) in [(3072, 1024, 0, "gemm qkv"), (1024, 1024, 0, "gemm proj"), (4096, 1024, 1, "gemm fc1"),
                            (1024, 4096, 0, "gemm fc2")]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") / 32).bfloat16()
        bias = torch.randn(N, device="cuda")
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        run(nm, lambda: ops.gemm_bf16(a, w, bias, epilogue=epi, out=out), 2.0 * M * N * K, "TFLOP/s")
        if nm == "gemm qkv":
            run("torch.matmul (cuBLAS) qkv", lambda: torch.matmul(a, w.t(), out=out), 2.0 * M * N * K, "TFLOP/s")
        del a, w, out
    B, T, n, Hh = clips, 16, 256, 16
    N_ = 1 + T * n
    qkv = (torch.randn(B * N_, 3 * Hh * 64, device="cuda") * 0.5).bfloat16()
    o = torch.empty(B * N_, Hh * 64, device="cuda", dtype=torch.bfloat16)
    lib = L.load()
    gb = (qkv.numel() + o.numel()) * 2
    for kind, nm in ((0, "attn space"), (1, "attn time")):
        run(nm, lambda: L.check(lib.hh_attention(L.ptr(qkv), L.ptr(o), B, T, n, Hh, kind, L.stream_ptr()), "attn"),
            gb, "GB/s")
    del qkv, o
    x = torch.randn(M, 1024, device="cuda")
    w1 = torch.ones(1024, device="cuda")
    run("layernorm (r4 w2)", lambda: ops.layernorm(x, w1, w1, 1e-6, want_f32=False, want_bf16=True), x.numel() * 6, "GB/s")
    y = torch.empty_like(x)
    run("torch copy fp32", lambda: y.copy_(x), x.numel() * 8, "GB/s")
    print("idle power %.0f W" % (pynvml.nvmlDeviceGetPowerUsage(H) / 1000.0))


if __name__ == "__main__":
    main()
