"""Per-kernel SASS opcode counts of the shipped library (the mnemonics that prove tcgen05 / TMEM / TMA are in use:
B200_PROFILING.md).  python tools/sass_counts.py > profiles/r2_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "helping_hand_for_egocentric_videos_b200", "libhh_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "MUFU", "LDSM",
       "FFMA2", "FADD2", "FMNMX3", "UTCATOMSWS", "MEMBAR"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(CUtensorMap_st.*|\(hh::.*|\(const .*|\(float.*|\(void.*|\(int.*|\(unsigned.*|\(__nv.*", "", name)
        cur = re.sub(r"void |hh::|\(anonymous namespace\)::|<unnamed>::", "", name)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if not m:
        continue
    op = m.group(1)
    counts[cur]["_total"] += 1
    for o in OPS:
        if op == o or op.startswith(o + "."):
            counts[cur][o] += 1
    if op.startswith("UTCHMMA") and ".2CTA" in op:
        counts[cur]["UTCHMMA.2CTA"] += 1
print("# SASS opcode counts per kernel (`cuobjdump -sass libhh_b200.so`, sm_100a) -- round 2\n")
print("Tensor-core path mnemonics: `UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), `LDTM` / `STTM` = tcgen05.ld / st (TMEM), "
      "`UTMALDG` / `UTMASTG` = cp.async.bulk.tensor load / store (TMA), `UBLKCP` = cp.async.bulk, `UTCBAR` = tcgen05.commit, "
      "`SYNCS` = mbarrier ops, `HMMA` = legacy mma.sync, `MUFU` = SFU (ex2 / tanh / rsqrt).\n")
cols = [o for o in OPS]
print("| kernel | instr | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for k, c in counts.items():
    if not any(c[o] for o in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "HMMA", "UBLKCP")):
        continue
    print("| `%s` | %d | %s |" % (k[:90], c["_total"], " | ".join(str(c[o]) if c[o] else "" for o in cols)))
