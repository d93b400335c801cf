#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3
for w in c1 c2; do
  timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_pdl_$w.json 2> gpurun_out/bench_pdl_$w.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_pdl_$w.json").read().strip().splitlines()[-1])
print("$w", "ms/step %.3f" % d["ms_per_step"], "e2e %.3f" % d["e2e"]["ms_per_step"], "streamed %.3f" % d["e2e_streamed"]["ms_per_step"])
PY
done
for r in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary --short-e2e > gpurun_out/bench_pdl_c3_$r.json 2> gpurun_out/bench_pdl_c3_$r.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_pdl_c3_$r.json").read().strip().splitlines()[-1])
print("c3", "ms/step %.3f" % d["ms_per_step"], "e2e %.3f" % d["e2e"]["ms_per_step"], "clk", d["clocks"]["sm_mhz"], "conv %.1f" % d["roofline"]["achieved"])
PY
done
