"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel family.
usage: python profiles/summarize_launches.py gpurun_out/launches_rNN.csv [index of the step to summarise]"""
import collections
import csv
import io
import re
import sys

path = sys.argv[1]
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
rows = [r for r in rows if "k_pack" not in r["Kernel Name"]]
# one sampler step = from one k_sinusoid (first kernel of the UNet forward) to the next
starts = [i for i, r in enumerate(rows) if "k_sinusoid" in r["Kernel Name"]]
if len(starts) >= 2:
    which = min(per_step, len(starts) - 2)
    rows = rows[starts[which]:starts[which + 1]]
fam = collections.OrderedDict()
for r in rows:
    name = re.sub(r"^void\s+", "", r["Kernel Name"])
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"eo::\(anonymous namespace\)::|eo::", "", name)
    e = fam.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += float(r["Metric Value"]) / 1e6
tot = sum(v[1] for v in fam.values())
print(f"# {path}: {len(rows)} launches, {tot:.3f} ms total (ncu-serialised, cold-cache: compare shares)")
print(f"{'kernel':44s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:44s} {v[0]:8d} {v[1]:10.3f} {100 * v[1] / tot:6.1f}%")
