"""DRAM traffic per launch of one kernel from an `ncu --set full` report -> profiles/<tag>_traffic.json
(read by bench.py's roofline.traffic).  Run where ncu is installed; no GPU needed.
usage: python profiles/extract_traffic.py report.ncu-rep <kernel regex> <workload> <batch> <out.json>"""
import csv
import json
import re
import subprocess
import sys

rep, pat, workload, batch, out_path = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(v.replace(",", "")) * mult


tot, n, dur = 0.0, 0, 0.0
for r in rows[2:]:
    if not re.search(pat, r[col["Kernel Name"]]):
        continue
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    tot += rd + wr
    n += 1
kernel = {"k_conv_tc3": "k_conv_tc3", "k_attn_tc": "k_attn_tc"}.get(pat, pat)
res = {"kernel": kernel, "workload": workload, "batch": batch, "launches": n,
       "dram_bytes_per_launch": tot / max(n, 1), "source": rep.split("/")[-1]}
with open(out_path, "w") as f:
    json.dump(res, f, indent=1)
print(res)
