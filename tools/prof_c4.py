"""Kernel list of the c4 training step's decoder forward and backward (torch.profiler / CUPTI, one step each): which kernels
the time goes to, and how much of the wall time between the CUDA events is kernel time at all.
  python tools/prof_c4.py [B]"""
import collections
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
import bench  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    ts = bench.TrainShare(B, "cuda:0")
    for _ in range(3):
        ts.step()
    torch.cuda.synchronize()
    ts.step(timed_parts=True)
    print("parts (CUDA events):", {k: round(v, 2) for k, v in ts.marks.items()})
    for part in ("decoder", "losses", "backward"):
        prof = profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA])

        class Ctx:
            def __enter__(self):
                torch.cuda.synchronize()
                prof.__enter__()

            def __exit__(self, *a):
                torch.cuda.synchronize()
                prof.__exit__(*a)

        ts.step(wrap={part: Ctx})
        agg = collections.OrderedDict()
        t_lo, t_hi = None, None
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA and e.name and not e.name.startswith("Memset") and \
                    not e.name.startswith("Memcpy"):
                a = agg.setdefault(e.name[:90], [0, 0.0])
                a[0] += 1
                a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
                lo, hi = e.time_range.start, e.time_range.end
                t_lo = lo if t_lo is None else min(t_lo, lo)
                t_hi = hi if t_hi is None else max(t_hi, hi)
        tot = sum(v for _, v in agg.values())
        n = sum(c for c, _ in agg.values())
        print("== %s: %d kernels, summed kernel time %.2f ms, first-start to last-end %.2f ms" % (part, n, tot / 1e3, (t_hi - t_lo) / 1e3))
        for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            print("  %-90s %4d %9.1f us" % (k, c, v))


if __name__ == "__main__":
    main()
