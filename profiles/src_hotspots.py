"""Top SASS instructions by stall samples from `ncu --page source --csv` (needs -lineinfo +
--import-source on).  usage: python profiles/src_hotspots.py prof.ncu-rep [N] [launch index]"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
# several launches: each has its own "Kernel Name" + header rows; keep the launch asked for
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
lo = starts[launch] + 2
hi = starts[launch + 1] if launch + 1 < len(starts) else len(rows)
print(rows[starts[launch]][1][:100])
body = [r for r in rows[lo:hi] if len(r) == len(hdr)]
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
totinst = sum(int(r[col["Instructions Executed"]] or 0) for r in body)
print(f"total samples {tot}, warp instructions executed {totinst}")
agg = {}
for s in stalls:
    agg[s] = sum(int(r[col[s]] or 0) for r in body)
print("stall mix:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
print(f"{'samples':>8s} {'%':>6s} {'executed':>12s}  top-stall          SASS")
for r in sorted(body, key=lambda r: -int(r[col["# Samples"]] or 0))[:n]:
    smp = int(r[col["# Samples"]] or 0)
    top = max(stalls, key=lambda s: int(r[col[s]] or 0))
    print(f"{smp:8d} {100 * smp / max(tot, 1):6.2f} {r[col['Instructions Executed']]:>12s}  {top[6:]:16s}  {r[col['Source']][:110]}")
