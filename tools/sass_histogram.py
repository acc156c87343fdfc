#!/usr/bin/env python
"""cuobjdump -sass of libea_b200.so -> per-kernel opcode histogram of the Blackwell-specific instructions
(B200_PROFILING.md "What proves a Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
UTMALDG/UTMASTG = TMA tensor copies, HMMA would be the legacy mma.sync path (must be 0).
usage: python tools/sass_histogram.py [lib] > profiles/rNN_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "easyanimate_b200/libea_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU.EX2",
        "F2FP", "FFMA2", "FADD2", "USETMAXREG", "LDG", "STG", "ST.E", "LD.E"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k.endswith(".2CTA") and op.startswith("UTCHMMA") and ".2CTA" in op):
                per[cur][k] += 1
tot = collections.Counter()
print(f"# {lib}: SASS opcode counts per kernel (static instruction counts, cuobjdump -sass; sm_100a)")
print("# " + " ".join(f"{k:>12}" for k in ["instrs"] + KEYS))
for fn, c in per.items():
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(.*", "", name)[:100]
    if not any(c[k] for k in ("UTCHMMA", "LDTM", "UTMALDG", "HMMA")) and "--all" not in sys.argv:
        for k in c:
            tot[k] += c[k]
        continue
    print(f"{name}\n  " + " ".join(f"{c[k]:>12}" for k in ["_total"] + KEYS))
    for k in c:
        tot[k] += c[k]
print("TOTAL (all kernels incl. the elementwise ones not listed)\n  " + " ".join(f"{tot[k]:>12}" for k in ["_total"] + KEYS))
