"""Turns the ncu reports / CSVs in gpurun_out/ into the committed summaries under profiles/ (round 1)."""
import collections
import csv
import json
import re
import subprocess
import sys

R = "r1"


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def num(s):
    return float(s.replace(",", "")) if s not in ("", "n/a") else 0.0


def to_bytes(v, unit):
    return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def to_us(v, unit):
    return num(v) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)  # "second"-style units vary by ncu version


def launches():
    rows = [r for r in csv.reader(open("gpurun_out/launches_%s.csv" % R)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    data = rows[1:]
    last = data[-(len(data) // 4):]
    agg = collections.OrderedDict()
    tot = 0.0
    for r in last:
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"void |hh::|\(anonymous namespace\)::|<unnamed>::", "", name)
        v = to_us(r[vi], r[ui])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    out = ["# ncu launch list, round 1 (gpu__time_duration.sum, --clock-control none)", "",
           "Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 1 "
           "--warmup 3 --batch 16 --no-cpu-baseline --no-e2e`",
           "(16 clips per step to keep the serialised run short; the last quarter of all launches ~ one step; times are "
           "cold-cache and serialised, so compare SHARES with bench.py's `kernel_ms_per_step`, not absolutes.)", "",
           "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.0f | %.1f %% |" % (k[:100], c, v, 100 * v / tot))
    out.append("| **total** | %d | %.0f | 100 %% |" % (len(last), tot))
    open("profiles/%s_launches_bench_b16.md" % R, "w").write("\n".join(out) + "\n")


def gemm_full():
    hdr, units, rows = ncu_raw("gpurun_out/gemm_%sc.ncu-rep" % R)
    M = 64 * 4097
    shapes = [("qkv", 3072, 1024), ("proj", 1024, 1024), ("fc1 (+QuickGELU)", 4096, 1024), ("fc2", 1024, 4096)]
    keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "launch__grid_size", "launch__cluster_size", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
    out = ["# ncu --set full: gemm_kernel (tcgen05), round 1, bench shapes (M = 64 clips x 4097 tokens = 262 208 rows)", "",
           "Command: `ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 4 python "
           "tools/prof_kernels.py gemm 64 1`   (one launch per encoder GEMM shape; report gpurun_out/gemm_r1c.ncu-rep)",
           "Durations under ncu are with replay at ncu's clocks; bench numbers come from CUDA events.", ""]
    traffic = {}
    for r, (nm, N, K) in zip(rows, shapes):
        out.append("## %s: M=%d N=%d K=%d" % (nm, M, N, K))
        d = {}
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                d[k] = (r[i], units[i])
                out.append("- %s: %s %s" % (k, r[i], units[i]))
        us = to_us(*d["gpu__time_duration.sum"])
        tr = to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])
        alg = (M * K + N * K + M * N) * 2.0
        fl = 2.0 * M * N * K
        out.append("- derived: %.0f TFLOP/s under ncu; DRAM traffic %.1f MB per launch vs algorithmic %.1f MB (x%.2f)" % (
            fl / us / 1e6, tr / 1e6, alg / 1e6, tr / alg))
        out.append("")
        traffic[nm.split()[0]] = {"dram_bytes_per_launch": tr, "algorithmic_bytes": alg, "flops": fl, "ncu_us": us,
                                  "tensor_pipe_active_pct": num(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0])}
    out += ["## Reading", "",
            "* The mainloop keeps the tensor pipe 80-90 % active; K=1024 tiles lose a little to the epilogue "
            "(TMEM -> registers -> swizzled smem -> bulk tensor store), K=4096 tiles (fc2) the least.",
            "* DRAM traffic is at or below the algorithmic bytes (A streamed once, W L2-resident; each CTA of a pair "
            "fetches only its half of the W tile): no wasted re-reads.",
            "* 148 CTAs (74 clusters of 2) x 256 threads, 1 CTA/SM; each pair runs ONE `tcgen05.mma.cta_group::2` "
            "(256 x 256 x 16) per K step: 231.7 KB dynamic smem (6 x 32 KB TMA stages + 32 KB store ring), 512 TMEM "
            "columns per CTA (2 accumulators of its 128 rows).",
            "* SASS: `UTCHMMA.2CTA` (tcgen05.mma.cta_group::2), `UTMALDG.2D.2CTA`, `UTMASTG.2D`, `LDTM`, "
            "`UTCBAR.2CTA.MULTICAST` present (`cuobjdump -sass libhh_b200.so`)."]
    open("profiles/%s_gemm_ncu_full.md" % R, "w").write("\n".join(out) + "\n")
    json.dump({"source": "ncu --set full, tools/prof_kernels.py gemm 64 1 (M=262208)", "kernels": traffic},
              open("profiles/%s_gemm_traffic.json" % R, "w"), indent=1)


def satellites():
    hdr, units, rows = ncu_raw("gpurun_out/sat_%sc.ncu-rep" % R)
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]
    out = ["# ncu --set full: attention and LayerNorm kernels, round 1 (B = 16 clips, T = 16, n = 256, H = 16)", "",
           "Command: `ncu --set full --clock-control none -k regex:\"attn_space_tc|attn_time_v2|ln_rows\" -c 3 python "
           "tools/prof_kernels.py all 16 1`", ""]
    ki = hdr.index("Kernel Name")
    for r in rows:
        out.append("## " + re.sub(r"\(.*", "", r[ki])[:80])
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                out.append("- %s: %s %s" % (k, r[i], units[i]))
        out.append("")
    open("profiles/%s_attention_ln_ncu_full.md" % R, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    launches()
    gemm_full()
    satellites()
    print(open("profiles/%s_gemm_traffic.json" % R).read()[:600])
