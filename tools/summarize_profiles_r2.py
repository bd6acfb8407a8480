"""Turns the round-2 ncu reports / CSVs in gpurun_out/ (tools/gpu_r2_prof.sh) into the committed summaries under profiles/."""
import collections
import csv
import json
import re
import subprocess

R = "r2"


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def num(s):
    return float(s.replace(",", "")) if s not in ("", "n/a") else 0.0


def to_bytes(v, unit):
    return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def to_us(v, unit):
    return num(v) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def short(name):
    name = re.sub(r"\(CUtensorMap_st.*|\(hh::.*|\(const .*|\(float.*|\(void.*|\(int.*|\(unsigned.*|\(__nv.*|\(at::.*", "", name)
    return re.sub(r"void |hh::|\(anonymous namespace\)::|<unnamed>::|\(int\)|\(bool\)", "", name)


def launches():
    rows = [r for r in csv.reader(open("gpurun_out/launches_%s.csv" % R)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    data = rows[1:]
    last = data[-(len(data) // 5):]          # 3 warm-up + 2 timed steps: the last fifth ~ one step
    agg = collections.OrderedDict()
    tot = 0.0
    for r in last:
        name = short(r[ki])
        v = to_us(r[vi], r[ui])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    out = ["# ncu launch list of the bench command, round 2 (gpu__time_duration.sum, --clock-control none)", "",
           "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv python bench.py --steps 2 "
           "--warmup 3 --no-cpu-baseline --no-e2e --no-extras`  (64 clips x 16 frames per step, LayerNorm folded into the GEMMs)",
           "(%d launches in all; the last fifth ~ one step; times are cold-cache and serialised, so compare SHARES with "
           "bench.py's `kernel_ms_per_step`, not absolutes.)" % len(data), "",
           "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.0f | %.1f %% |" % (k[:110], c, v, 100 * v / tot))
    out.append("| **total** | %d | %.0f | 100 %% |" % (len(last), tot))
    gem = sum(v for k, (c, v) in agg.items() if k.startswith("gemm_kernel"))
    out += ["", "`gemm_kernel` (all instantiations): %.1f %% of the serialised step." % (100 * gem / tot)]
    open("profiles/%s_launches_bench.md" % R, "w").write("\n".join(out) + "\n")


def gemm_full():
    hdr, units, rows = ncu_raw("gpurun_out/%s_gemm_full.ncu-rep" % R)
    M = 64 * 4097
    D, Hd = 1024, 4096
    names = [("(warm-up) proj res nowb", D, D, 4), ("proj plain", D, D, 0), ("proj res nowb", D, D, 4), ("proj res wb", D, D, 8),
             ("fc2 plain", D, Hd, 0), ("fc2 res wb", D, Hd, 8), ("qkv plain", 3 * D, D, 0), ("qkv ln", 3 * D, D, 0),
             ("fc1 plain (QuickGELU)", Hd, D, 0), ("fc1 ln (QuickGELU)", Hd, D, 0)]
    keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "launch__grid_size", "launch__cluster_size", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
    out = ["# ncu --set full: gemm_kernel (tcgen05) with the plain and the fused-LayerNorm epilogues, round 2 (M = 262 208 rows)", "",
           "Command: `ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 12 python "
           "tools/prof_fused.py ncu 64`  (one launch per variant; report gpurun_out/r2_gemm_full.ncu-rep).",
           "`res` = residual + LayerNorm-statistics producer epilogue (`wb`: fp32 sum written back in place), `ln` = LayerNorm "
           "folded into the consumer epilogue.  Durations under ncu are at ncu's clocks; bench numbers come from CUDA events.", ""]
    info = {}
    for r, (nm, N, K, extra) in zip(rows, names):
        out.append("## %s: M=%d N=%d K=%d" % (nm, M, N, K))
        d = {}
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                d[k] = (r[i], units[i])
                out.append("- %s: %s %s" % (k, r[i], units[i]))
        us = to_us(*d["gpu__time_duration.sum"])
        tr = to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])
        alg = (M * K + N * K + M * N) * 2.0 + extra * 1.0 * M * N
        fl = 2.0 * M * N * K
        out.append("- derived: %.0f TFLOP/s under ncu; DRAM traffic %.1f MB per launch vs algorithmic %.1f MB (x%.2f)" % (
            fl / us / 1e6, tr / 1e6, alg / 1e6, tr / alg))
        out.append("")
        info[nm] = {"dram_bytes_per_launch": tr, "algorithmic_bytes": alg, "flops": fl, "ncu_us": us,
                    "tensor_pipe_active_pct": num(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0])}
    out += ["## Reading", "",
            "* DRAM traffic stays at the algorithmic bytes for every variant (A streamed once, W L2-resident, the fp32 residual "
            "read once and -- with write-back -- written once): the fused epilogues add exactly the residual stream's bytes.",
            "* The folded-LayerNorm consumers (`ln`) run within 2 - 3 % of the plain epilogues; the residual producers (`res`) are "
            "14 - 54 % longer although their epilogue has slack -- the MMA thread waits for operands (`profiles/r2_gemm_epilogue_trace.md`)."]
    open("profiles/%s_gemm_ncu_full.md" % R, "w").write("\n".join(out) + "\n")

    def avg(a, b):
        return {k: 0.5 * (info[a][k] + info[b][k]) for k in info[a]}
    traffic = {"qkv": info["qkv ln"], "proj": avg("proj res nowb", "proj res wb"), "fc1": info["fc1 ln (QuickGELU)"],
               "fc2": info["fc2 res wb"]}
    json.dump({"source": "profiles/r2_gemm_ncu_full.md (ncu --set full, tools/prof_fused.py ncu 64, M=262208, fused-LayerNorm "
                         "epilogues as the step runs them: qkv / fc1 = ln consumers, proj = mean of res nowb / wb, fc2 = res wb; "
                         "dram__bytes_read.sum + dram__bytes_write.sum)",
               "kernels": traffic, "all_variants": info}, open("profiles/%s_gemm_traffic.json" % R, "w"), indent=1)


def attention():
    hdr, units, rows = ncu_raw("gpurun_out/%s_attn_full.ncu-rep" % R)
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]
    out = ["# ncu --set full: attention kernels, round 2 (B = 16 clips, T = 16, n = 256, H = 16)", "",
           "Command: `ncu --set full --clock-control none --import-source on -k regex:\"attn_space_tc|attn_time_v2\" -c 4 python "
           "tools/prof_kernels.py attn 16 2`  (report gpurun_out/r2_attn_full.ncu-rep)", ""]
    ki = hdr.index("Kernel Name")
    for r in rows:
        out.append("## " + short(r[ki])[:80])
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                out.append("- %s: %s %s" % (k, r[i], units[i]))
        out.append("")
    # the spatial kernel as the round ends (one thread per query row, per-half MMA threads, turn-taking exp pass)
    hdr, units, rows = ncu_raw("gpurun_out/%s_attn_space_final.ncu-rep" % R)
    keys += ["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
             "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second"]
    out += ["# Final spatial-attention kernel of the round (same problem: B = 16 clips, T = 16, n = 256, H = 16)", "",
            "Command: `ncu --set full --clock-control none --import-source on -k regex:\"attn_space_tc\" -c 2 python "
            "tools/prof_kernels.py attn 16 1`  (tools/gpu_calls/d22.sh; report gpurun_out/r2_attn_space_final.ncu-rep)", ""]
    ki = hdr.index("Kernel Name")
    for r in rows:
        out.append("## " + short(r[ki])[:80])
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                out.append("- %s: %s %s" % (k, r[i], units[i]))
        out.append("")
    out += ["Reading: 167.8 -> 131.9 us under ncu (bench: 23.6 -> 17.1 ms per 64-clip step); tensor pipe 27.6 -> 35.5 % active, "
            "XU (MUFU ex2) 47.6 -> 57.5 %, instructions 88.6 M -> 59.3 M (no cross-thread max / sum exchange, no ragged-chunk code), "
            "DRAM 37 -> 47 % of peak with unchanged bytes (512 MB per launch = the algorithmic q/k/v read + output write). "
            "The exp pass (one ex2 per logit: 66 k per 256-query task = 4.1 k cycles of the SM's 16-lane MUFU) is the floor; the "
            "kernel sits at 1.75 x that floor.", ""]
    open("profiles/%s_attention_ncu_full.md" % R, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    launches()
    gemm_full()
    attention()
    print(open("profiles/%s_gemm_traffic.json" % R).read()[:900])
