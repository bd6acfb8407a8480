"""Stand-alone timing (CUDA events) and ncu target for the fused-LayerNorm GEMM epilogues next to their plain forms, at
the headline shapes (M = clips x 4097 rows).  Not a benchmark.

  python tools/prof_fused.py time [clips] [reps]    # per-variant ms, TFLOP/s, effective GB/s of the epilogue traffic
  ncu ... python tools/prof_fused.py ncu [clips]    # one launch of every variant
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helping_hand_for_egocentric_videos_b200 import ops  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "time"
clips = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
M = clips * 4097
D, Hd = 1024, 4096
torch.manual_seed(0)
dev = "cuda"


def mk(N, K):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / 32).bfloat16()
    b = torch.randn(N, device=dev)
    return a, w, b


variants = []


def add(name, fn, flops, extra_bytes):
    variants.append((name, fn, flops, extra_bytes))


a1, w_qkv, b_qkv = mk(3 * D, D)
_, w_proj, b_proj = mk(D, D)
_, w_fc1, b_fc1 = mk(Hd, D)
h, w_fc2, b_fc2 = mk(D, Hd)
x = torch.randn(M, D, device=dev)
gamma = torch.ones(D, device=dev)
beta = torch.zeros(D, device=dev)
out_d = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
out_3d = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
out_h = torch.empty(M, Hd, device=dev, dtype=torch.bfloat16)
wq_f, cs_q, bq_f = ops.fold_layernorm_weight(w_qkv.float(), gamma, beta, b_qkv)
w1_f, cs_1, b1_f = ops.fold_layernorm_weight(w_fc1.float(), gamma, beta, b_fc1)
_, stats = ops.gemm_bf16_res_stats(a1, w_proj, b_proj, x, writeback=False)

add("proj_plain", lambda: ops.gemm_bf16(a1, w_proj, b_proj, out=out_d), 2.0 * M * D * D, 0)
add("proj_res_nowb", lambda: ops.gemm_bf16_res_stats(a1, w_proj, b_proj, x, writeback=False), 2.0 * M * D * D, 4.0 * M * D)
add("proj_res_wb", lambda: ops.gemm_bf16_res_stats(a1, w_proj, b_proj, x, writeback=True), 2.0 * M * D * D, 8.0 * M * D)
add("fc2_plain", lambda: ops.gemm_bf16(h, w_fc2, b_fc2, out=out_d), 2.0 * M * D * Hd, 0)
add("fc2_res_wb", lambda: ops.gemm_bf16_res_stats(h, w_fc2, b_fc2, x, writeback=True), 2.0 * M * D * Hd, 8.0 * M * D)
add("fc2_res_nowb", lambda: ops.gemm_bf16_res_stats(h, w_fc2, b_fc2, x, writeback=False), 2.0 * M * D * Hd, 4.0 * M * D)
add("qkv_plain", lambda: ops.gemm_bf16(a1, w_qkv, b_qkv, out=out_3d), 2.0 * M * 3 * D * D, 0)
add("qkv_ln", lambda: ops.gemm_bf16_ln(a1, wq_f, bq_f, cs_q, stats, 1e-6), 2.0 * M * 3 * D * D, 0)
add("fc1_plain", lambda: ops.gemm_bf16(a1, w_fc1, b_fc1, epilogue=1, out=out_h), 2.0 * M * Hd * D, 0)
add("fc1_ln", lambda: ops.gemm_bf16_ln(a1, w1_f, b1_f, cs_1, stats, 1e-6, qgelu=True), 2.0 * M * Hd * D, 0)

only = os.environ.get("PROF_ONLY")
if only:
    variants = [v for v in variants if v[0] in only.split(",")]

if mode == "trace":
    # HH_B200_LIB=tools/ab/libhh_b200_trace.so (built with -DHH_GEMM_TRACE): where every role of the GEMM waits
    import ctypes
    from helping_hand_for_egocentric_videos_b200 import _lib
    dll = _lib.load()
    buf = (ctypes.c_ulonglong * (148 * 16))()
    names = ["kernel", "prod:empty", "mma:acc_empty", "mma:full", "epi:acc_full", "epi:res_full", "epi:store_read",
             "epi:prologue", "tiles", "epi1:acc_full", "epi1:res_full"]
    for name, fn, _, _ in variants:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        assert dll.hh_debug_gemm_trace(buf, 148 * 16) == 0
        rows = [[buf[c * 16 + i] for i in range(11)] for c in range(148)]
        tot = sum(r[0] for r in rows) / 148.0
        line = "%-14s kernel %8.0f cyc, tiles/CTA %5.1f |" % (name, tot, sum(r[8] for r in rows) / 148.0)
        for i in range(1, 11):
            if i == 8:
                continue
            line += " %s %4.1f%%" % (names[i], 100.0 * sum(r[i] for r in rows) / 148.0 / tot)
        print(line, flush=True)
    sys.exit(0)

if mode == "ncu":
    for name, fn, _, _ in variants:
        fn()
    torch.cuda.synchronize()
    print("done")
    sys.exit(0)

res = {}
for name, fn, flops, extra in variants:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res[name] = {"ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1), "extra_GBps": round(extra / ms / 1e6, 1)}
    print(name, res[name], flush=True)
print(json.dumps(res))
