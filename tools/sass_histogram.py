#!/usr/bin/env python
"""Per-kernel histogram of the SASS opcodes that prove which hardware path a kernel uses (B200_PROFILING.md): tcgen05 (UTCHMMA,
UTCBAR, LDTM / STTM = tcgen05.ld / st), TMA (UTMALDG / UTMASTG / UTMAREDG / UTMAPF), legacy warp MMA (HMMA), mbarrier (SYNCS),
cp.async (LDGSTS), ldmatrix (LDSM), atomics (RED / ATOM / ATOMS).

    python tools/sass_histogram.py [tim_b200/libtim_b200.so] > profiles/r02_sass_opcodes.txt
"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "tim_b200/libtim_b200.so"
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "HMMA", "LDSM", "LDGSTS", "SYNCS",
        "RED", "ATOM", "ATOMS", "FFMA", "MUFU"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip()
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        base = op.split(".")[0]
        per[cur][base] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            per[cur]["UTCHMMA.2CTA"] += 1
        per[cur]["_total"] += 1
print(f"# SASS opcode histogram of {LIB} (cuobjdump -sass); columns: " + " ".join(KEYS) + " | total instructions")
tot = collections.Counter()
for fn, c in per.items():
    name = demangle(fn)
    name = re.sub(r"tim::\(anonymous namespace\)::", "", name)
    name = name.split("(")[0][:110]
    print(f"{name:112s} " + " ".join(f"{c.get(k, 0):5d}" for k in KEYS) + f" | {c['_total']}")
    tot.update(c)
print(f"{'TOTAL':112s} " + " ".join(f"{tot.get(k, 0):5d}" for k in KEYS) + f" | {tot['_total']}")
