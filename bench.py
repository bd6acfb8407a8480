#!/usr/bin/env python
"""Headline benchmark: clips/sec of the video-side forward path (TimeSformer-L/14 encoder -> object-aware decoder ->
obj_proj -> similarity scoring) at 16 frames x 224^2, nq = 12, batch 64 clips per GPU (BASELINE.json configs[2]).

  python bench.py --gpus 1 --steps K --warmup W                      # this repo (hand-written sm_100a path)
  torchrun ... bench.py --gpus N ...                                 # one rank per GPU, weak scaling, 1 all-gather
  python bench.py --impl reference ...                               # reference algorithm on the host CPU cores

One JSON line is printed by rank 0 (contract: task description / DESIGN.md section "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "clips/sec (16f x 224^2, nq=12)"
UNIT = "clips/s"


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return dict(burst=d.get("bf16_tflops", 1590.0), sustained=d.get("bf16_tflops_sustained", 1400.0),
                    hbm=d.get("hbm_gbs", 6650.0), source="MEASURED_PEAKS.json")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
            except ValueError:
                continue
            for nm, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------- model construction
def build_modules(frames, nq, seed=0):
    from helping_hand_for_egocentric_videos_b200.model import LaviLa, tfm_decoder
    from helping_hand_for_egocentric_videos_b200 import synthetic
    vis = LaviLa.SpaceTimeTransformer(img_size=224, patch_size=14, embed_dim=1024, depth=24, num_heads=16,
                                      num_frames=frames, time_init='zeros', attention_style='frozen-in-time',
                                      ln_pre=True, act_layer=LaviLa.QuickGELU, num_classes=0)
    tr = tfm_decoder.Cross_Attention(normalize_before=True, return_intermediate_dec=True)
    # run/test_epic.py:150-153 construction (the 16-frame path): pred_traj=False, num_queries = nq + 1
    dec = tfm_decoder.ObjDecoder(tr, num_classes=22047, num_queries=nq + 1, aux_loss=True, pred_traj=False,
                                 feature_dim=1024, num_frames=frames, patches_per_frame=256)
    synthetic.randomize_(vis, seed)
    synthetic.randomize_(dec, seed + 1)
    return vis.eval(), dec.eval()


def cpu_forward_clips(vsd, dsd, frames, nclips, threads):
    """Reference algorithm (oracle port) on the host: encoder + decoder + obj_proj + sim, fp32."""
    from oracle import hh_oracle as O
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(99)
    video = torch.randn(nclips, frames, 3, 224, 224, generator=g)
    text = torch.randn(13, 256, generator=g)
    t0 = time.perf_counter()
    with torch.no_grad():
        _, fmap = O.encoder_forward(video, vsd, 16)
        _, hs, _, _ = O.decoder_forward(fmap[:, 1:].unflatten(1, (frames, 256)), dsd, heads=8, pred_traj=False)
        vid = O.obj_proj(hs[-1], dsd)[:, -1]
        O.sim_matrix(text, vid).argmax(-1)
    return time.perf_counter() - t0


def run_reference(args, rank, emit):
    """--impl reference: the reference's algorithm (oracle port: the reference itself is Python that needs
    /root/reference, which does not exist on the GPU box) on all host cores; bounded sample of 4 clips per step."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vis, dec = build_modules(args.frames, args.nq)
    vsd = {k: v.detach() for k, v in vis.state_dict().items()}
    dsd = {k: v.detach() for k, v in dec.state_dict().items()}
    sample = 4                                    # ~10 s of host work per step on a 16-core box
    for _ in range(min(args.warmup, 1)):
        cpu_forward_clips(vsd, dsd, args.frames, sample, cores)
    steps = max(1, min(args.steps, 3))
    dts = [cpu_forward_clips(vsd, dsd, args.frames, sample, cores) for _ in range(steps)]
    dt = sum(dts) / len(dts)
    val = sample / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "TimeSformer-L/14 + tfm_decoder nq=%d, %d frames 224^2, batch %d clips/GPU, "
                                   "EgoMCQ-style scoring (BASELINE.json configs[2])" % (args.nq, args.frames, args.batch),
                       "frames": args.frames, "nq": args.nq, "clips_per_step": sample,
                       "sample": "bounded sample of the workload: %d clips per step on the host cores" % sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d clip(s) per step, %d steps, oracle/hh_oracle.py fp32 torch CPU" % (sample, steps)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU per step")
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--nq", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    # stdout carries exactly one JSON line: while the benchmark runs, file descriptor 1 points at stderr, so that
    # anything a C library prints there (NCCL's "NCCL version ..." banner when the box sets NCCL_DEBUG) cannot end up in
    # front of it; the descriptor is restored for the final print
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
    if args.impl == "reference":
        run_reference(args, rank, emit)
        return

    import torch.distributed as dist
    from helping_hand_for_egocentric_videos_b200 import ops, parallel, synthetic
    from helping_hand_for_egocentric_videos_b200.model import metric

    assert torch.cuda.is_available(), "bench.py (b200 arm) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(3, args.warmup)
    B, T, nq = args.batch, args.frames, args.nq

    vis_cpu, dec_cpu = build_modules(T, nq)
    cpu_sd = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_sd = ({k: v.detach().clone() for k, v in vis_cpu.state_dict().items()},
                  {k: v.detach().clone() for k, v in dec_cpu.state_dict().items()})
    vis, dec = vis_cpu.to(dev), dec_cpu.to(dev)

    host_video = synthetic.synthetic_clips(B, T, 224, seed=1234 + rank, pinned=True)   # 616 MB at B=64: > L2 (126 MB)
    video = host_video.to(dev, non_blocking=True)
    text = torch.randn(nq + 1, 256, generator=torch.Generator().manual_seed(7)).to(dev)  # caption embeddings (text tower
    #                                                                                      is outside the measured path)

    def step(v):
        _, fmap = vis.forward_features(v)
        grid = fmap[:, 1:].unflatten(1, (T, 256))
        _, hs, _, _ = dec(grid)
        vid = dec.obj_proj(hs[-1])[:, -1]
        if world > 1:
            (vid,) = parallel.all_gather_packed([vid])          # the one collective: embeddings over NVLink
        sim = metric.sim_matrix(text, vid)
        return ops.row_argmax(sim)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(warmup):
        step(video)
    flops_clip = vis.flops_per_clip() + dec.flops_per_clip(T)
    launches_step = vis.last_launches() + dec.last_launches() + 4 + (1 if world > 1 else 0)

    # ---- device-resident throughput (value), with per-kernel CUDA-event timing and clock sampling
    vis.set_profile(True)
    dec.set_profile(True)
    vis.profile(), dec.profile()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(lambda: step(video), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    prof = {}
    prof.update(vis.profile())
    prof.update(dec.profile())
    vis.set_profile(False)
    dec.set_profile(False)
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- end to end through the public modules with host buffers: pinned H2D of the clips + D2H of the choices
    e2e = None
    if not args.no_e2e:
        # Input pipeline: every step's clips cross PCIe from pinned host memory into one of two device buffers on a
        # copy stream, issued while the previous step computes (standard double buffering); the step waits for
        # its own copy, and its result is read back to the host before the next step starts.
        copy_stream = torch.cuda.Stream(device=dev)

        def measure_e2e(host, run_step):
            bufs = [torch.empty_like(host, device=dev), torch.empty_like(host, device=dev)]
            ready = [torch.cuda.Event(), torch.cuda.Event()]
            free = [torch.cuda.Event(), torch.cuda.Event()]
            for ev in free:
                ev.record()
            state = {"i": 0}

            def issue_copy(i):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(free[i % 2])
                    bufs[i % 2].copy_(host, non_blocking=True)
                    ready[i % 2].record(copy_stream)

            issue_copy(0)

            def e2e_step():
                i = state["i"]
                torch.cuda.current_stream().wait_event(ready[i % 2])
                issue_copy(i + 1)
                res = run_step(bufs[i % 2])
                free[i % 2].record()
                state["i"] = i + 1
                return res.cpu()
            for _ in range(2):
                e2e_step()
            ms = timed(e2e_step, args.steps) / args.steps
            torch.cuda.synchronize()
            return {"value": world * B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                    "h2d_bytes_per_step": host.numel() * host.element_size(), "d2h_bytes_per_step": (nq + 1) * 8}

        e2e = measure_e2e(host_video, step)
        # same step fed with raw uint8 frames [B,T,224,224,3] (what the video decoder produces): the loader's
        # /255 + mean/std normalisation is fused into the patch loader (hh_encoder_forward_u8), a quarter of the H2D bytes
        g8 = torch.Generator().manual_seed(4321 + rank)
        host_u8 = torch.randint(0, 256, (B, T, 224, 224, 3), generator=g8, dtype=torch.uint8).pin_memory()
        mean = [108.3272985 / 255, 116.7460125 / 255, 104.09373615000001 / 255]
        std = [68.5005327 / 255, 66.6321579 / 255, 70.32316305 / 255]

        def step_u8(frames):
            _, fmap = vis.forward_features_u8(frames, mean, std)
            grid = fmap[:, 1:].unflatten(1, (T, 256))
            _, hs, _, _ = dec(grid)
            vid = dec.obj_proj(hs[-1])[:, -1]
            if world > 1:
                (vid,) = parallel.all_gather_packed([vid])
            return ops.row_argmax(metric.sim_matrix(text, vid))
        e2e_u8 = measure_e2e(host_u8, step_u8)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    # dominant kernel = the tcgen05 GEMM (qkv / proj / fc1 / fc2 launches of the encoder)
    M = B * (1 + T * 256)
    gemm_flops = {"gemm_qkv": 2.0 * M * 3072 * 1024, "gemm_proj": 2.0 * M * 1024 * 1024,
                  "gemm_fc1": 2.0 * M * 4096 * 1024, "gemm_fc2": 2.0 * M * 1024 * 4096}
    g_ms = sum(prof[k][0] for k in gemm_flops if k in prof)
    g_fl = sum(gemm_flops[k] * prof[k][1] for k in gemm_flops if k in prof)
    gemm_tflops = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    per_class = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in prof.items()}
    for k in gemm_flops:
        if k in prof and prof[k][0] > 0:
            per_class[k]["tflops"] = gemm_flops[k] * prof[k][1] / (prof[k][0] * 1e-3) / 1e12
    # DRAM traffic of the dominant kernel per launch: from the committed `ncu --set full` capture at these very shapes
    # (profiles/r1_gemm_traffic.json), averaged over the launches of one step like `achieved`.
    traffic, traffic_detail = None, None
    tpath = os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")
    if os.path.isfile(tpath) and B == 64 and T == 16:
        with open(tpath) as f:
            tk = json.load(f)["kernels"]
        per_step = {"qkv": 48, "proj": 48, "fc1": 24, "fc2": 24}
        traffic = sum(tk[k]["dram_bytes_per_launch"] * c for k, c in per_step.items()) / 144   # bytes per launch
        traffic_detail = {"avg_dram_bytes_per_launch": traffic,
                          "avg_algorithmic_bytes_per_launch": sum(tk[k]["algorithmic_bytes"] * c for k, c in per_step.items()) / 144,
                          "source": "profiles/r1_gemm_ncu_full.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
    path_tflops = value / world * flops_clip / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "TimeSformer-L/14 + tfm_decoder nq=%d, %d frames 224^2, batch %d clips/GPU, "
                               "EgoMCQ-style scoring (BASELINE.json configs[2])" % (nq, T, B),
                   "clips_per_gpu": B, "frames": T, "nq": nq, "l2": "inputs larger than L2 (%.0f MB clips/step)" % (
                       host_video.numel() * 4 / 1e6), "parallelism": "dp%d" % world,
                   "flops_per_clip": flops_clip, "path_tflops_per_gpu": path_tflops,
                   "path_frac_of_sustained_peak": path_tflops / peaks["sustained"],
                   "kernel_ms_per_step": per_class},
        "roofline": {"bound": "tensor", "kernel": "gemm_kernel (tcgen05, encoder qkv/proj/fc1/fc2 launches)",
                     "achieved": gemm_tflops, "peak": peaks["sustained"], "unit": "TFLOP/s",
                     "frac": gemm_tflops / peaks["sustained"], "peak_source": peaks["source"] + " sustained bf16",
                     "share_of_step": g_ms / args.steps / ms_step if ms_step > 0 else None, "traffic": traffic,
                     "traffic_detail": traffic_detail},
        "clocks": clocks, "gpu_launches": launches_step * args.steps,
    }
    if e2e:
        line["e2e"] = e2e
        line["e2e_u8"] = e2e_u8
    if cpu_sd is not None:
        cores = os.cpu_count() or 1
        cpu_forward_clips(cpu_sd[0], cpu_sd[1], T, 1, cores)                 # warm the host thread pool / allocator
        dt = cpu_forward_clips(cpu_sd[0], cpu_sd[1], T, 4, cores)
        line["cpu_baseline"] = {"value": 4.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "4 clips (same L/14, %d-frame, nq=%d path) through oracle/hh_oracle.py, fp32 "
                                          "torch CPU, %.1f s" % (T, nq, dt)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
