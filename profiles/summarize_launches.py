"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`)
per kernel family over ONE sampler step, and optionally write the DRAM traffic per launch of one family as
profiles/<tag>_traffic.json (read by bench.py's roofline.traffic).
usage: python profiles/summarize_launches.py gpurun_out/launches_rNN.csv [step index] [--traffic FAMILY WORKLOAD BATCH OUT.json]"""
import collections
import csv
import io
import json
import re
import sys

args = sys.argv[1:]
traffic = None
if "--traffic" in args:
    i = args.index("--traffic")
    traffic = args[i + 1:i + 5]
    args = args[:i]
path = args[0]
per_step = int(args[1]) if len(args) > 1 else 0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
launches = collections.OrderedDict()
for r in csv.DictReader(io.StringIO("".join(lines))):
    e = launches.setdefault(r["ID"], {"name": r["Kernel Name"]})
    e[r["Metric Name"]] = float(r["Metric Value"])
rows = [e for e in launches.values() if "k_pack" not in e["name"]]
# one sampler step = from the first kernel of one UNet forward to the next: k_gather_rows when the sampler hoisted the
# timestep tables (round 2), k_sinusoid otherwise
marker = "k_gather_rows" if any("k_gather_rows" in r["name"] for r in rows) else "k_sinusoid"
starts = [i for i, r in enumerate(rows) if marker in r["name"]]
if len(starts) >= 2:
    which = min(per_step, len(starts) - 2)
    rows = rows[starts[which]:starts[which + 1]]
fam = collections.OrderedDict()
for r in rows:
    name = re.sub(r"^void\s+", "", r["name"])
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"eo::\(anonymous namespace\)::|eo::|unnamed>::", "", name)
    e = fam.setdefault(name, [0, 0.0, 0.0])
    e[0] += 1
    e[1] += r["gpu__time_duration.sum"] / 1e6
    e[2] += r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)
tot = sum(v[1] for v in fam.values())
print(f"# {path}: {len(rows)} launches of one sampler step, {tot:.3f} ms total (ncu-serialised, cold-cache: compare shares)")
print(f"{'kernel':36s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'DRAM GB':>9s} {'GB/launch':>10s}")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:36s} {v[0]:8d} {v[1]:10.3f} {100 * v[1] / tot:6.1f}% {v[2] / 1e9:9.3f} {v[2] / 1e9 / v[0]:10.4f}")
if traffic:
    family, workload, batch, out_path = traffic
    hits = [(k, v) for k, v in fam.items() if k.startswith(family)]
    n = sum(v[0] for _, v in hits)
    b = sum(v[2] for _, v in hits)
    res = {"kernel": family, "workload": workload, "batch": int(batch), "launches": n,
           "dram_bytes_per_launch": b / max(n, 1), "source": path.split("/")[-1],
           "how": "dram__bytes_read.sum + dram__bytes_write.sum of every launch of this family in one sampler step, averaged"}
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print("wrote", out_path, res)
