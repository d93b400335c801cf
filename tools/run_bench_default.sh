#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02i}
start=$(date +%s)
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo rc $? wall $(( $(date +%s) - start )) s
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["e2e_streamed"]["ms_per_step"], d["clocks"])
print({k:(v.get("ms_per_step"),v.get("frac")) for k,v in d["secondary"].items()})
PY
