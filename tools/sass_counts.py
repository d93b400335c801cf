#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md), from
`cuobjdump -sass` of the built library.  Writes a table to stdout:  python tools/sass_counts.py > profiles/rNN_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "eo_diffusion_b200", "libeo_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "MUFU.EX2",
        "MUFU.TANH", "FFMA2", "FADD2", "HMMA", "total"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
counts, order, cur = {}, [], None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"^void ", "", name).split("(")[0]
        cur = name
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        c = counts[cur]
        c["total"] += 1
        for k in KEYS:
            if k != "total" and (op == k or op.startswith(k + ".")):
                c[k] += 1
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instruction counts per kernel (static SASS)")
print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG = TMA load/store,")
print("# UTMAPF = TMA prefetch, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be 0)")
w = max(len(n) for n in order)
print(f"{'kernel':{w}s} " + " ".join(f"{k:>13s}" for k in KEYS))
for n in order:
    c = counts[n]
    if not (c["UTCHMMA"] or c["LDTM"] or c["UTMALDG"] or c["UTMASTG"]):
        continue
    print(f"{n:{w}s} " + " ".join(f"{c[k]:13d}" for k in KEYS))
print("# kernels without tensor-core / TMA instructions (bandwidth and fp32-parity kernels):")
print("# " + ", ".join(n for n in order if not (counts[n]["UTCHMMA"] or counts[n]["LDTM"] or counts[n]["UTMALDG"] or counts[n]["UTMASTG"])))
