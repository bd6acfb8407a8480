"""Per-source-line warp-stall samples of one kernel in an ncu report (--set full --import-source on), using the line
table of the shipped cubin: ncu's CSV source page lists SASS only, `nvdisasm -g` gives the same instruction sequence
with `//## File ..., line N` markers, the two are zipped in order.

  python tools/ncu_lines.py REPORT.ncu-rep LAUNCH_INDEX CUBIN 'gemm_kernelILi256ELi6ELb1ELb1ELb1' [top]
"""
import csv
import re
import subprocess
import sys


def main():
    rep, skip, cubin, fn = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(skip), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][1][:160])
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    inst = [r for r in rows[2:] if len(r) >= len(hdr) and r[idx["# Samples"]].isdigit()]
    inst = inst[:len(inst) // 2] if len(inst) % 2 == 0 and inst[0][idx["Address"]] == inst[len(inst) // 2][idx["Address"]] else inst
    sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    lines, cur, on = [], None, False
    for ln in sass:
        if ln.startswith(".text."):
            on = fn in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]*)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    print("instructions: ncu %d, cubin %d" % (len(inst), len(lines)))
    n = min(len(inst), len(lines))
    agg, total = {}, 0
    for r, loc in zip(inst[:n], lines[:n]):
        s = int(r[idx["# Samples"]])
        total += s
        a = agg.setdefault(loc, {"s": 0, "ex": 0, "why": {}})
        a["s"] += s
        a["ex"] += int(r[idx["Instructions Executed"]] or 0)
        for h in stalls:
            v = int(r[idx[h]] or 0)
            if v:
                a["why"][h[6:]] = a["why"].get(h[6:], 0) + v
    print("total samples", total)
    src = {}
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1]["s"])[:top]:
        text = ""
        if loc:
            if loc[0] not in src:
                try:
                    src[loc[0]] = open("helping_hand_for_egocentric_videos_b200/csrc/" + loc[0]).read().splitlines()
                except OSError:
                    src[loc[0]] = []
            if 0 < loc[1] <= len(src[loc[0]]):
                text = src[loc[0]][loc[1] - 1].strip()[:70]
        why = sorted(a["why"].items(), key=lambda kv: -kv[1])[:3]
        print("%6d %5.1f%%  %-22s ex %-9d %-40s | %s" % (a["s"], 100.0 * a["s"] / total, "%s:%d" % loc if loc else "?", a["ex"],
                                                       why, text))


if __name__ == "__main__":
    main()
