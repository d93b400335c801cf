#!/bin/bash
# correctness of the small-shape tile choice + c1 / c2 lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_unet.py tests/test_gpu_round2.py -q -p no:cacheprovider -x 2>&1 | tail -3
for w in c1 c2; do
  timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu --breakdown gpurun_out/bd_small_$w.json > gpurun_out/bench_small_$w.json 2> gpurun_out/bench_small_$w.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_small_$w.json").read().strip().splitlines()[-1])
print("$w", "ms/step %.3f" % d["ms_per_step"], "e2e %.3f" % d["e2e"]["ms_per_step"], "streamed %.3f" % d["e2e_streamed"]["ms_per_step"], "launches/step", d["gpu_launches"] / d["steps"])
PY
  grep -A9 "per-kernel-family" gpurun_out/bench_small_$w.err
done
