#!/bin/bash
# A/B of library variants on the launch-bound workloads c1 / c2: tools/ab_small.sh TAG variant1 variant2 ...
# ("base" = libeo_b200.so, anything else libeo_b200_<variant>.so); two interleaved rounds, device-timed steps.
TAG=$1; shift
mkdir -p gpurun_out
for round in 1 2; do
  for wl in c1 c2; do
    for v in "$@"; do
      lib=eo_diffusion_b200/libeo_b200.so
      [ "$v" != "base" ] && lib=eo_diffusion_b200/libeo_b200_$v.so
      EO_B200_LIB=$PWD/$lib timeout 300 python bench.py --workload $wl --steps 50 --warmup 10 --no-cpu --no-secondary --short-e2e \
        > gpurun_out/ab_${TAG}_${wl}_${v}_$round.json 2> gpurun_out/ab_${TAG}_${wl}_${v}_$round.err
      python - <<PY
import json
d=json.loads(open("gpurun_out/ab_${TAG}_${wl}_${v}_$round.json").read().strip().splitlines()[-1])
print("$wl $v", $round, "ms/step %.4f" % d["ms_per_step"], "e2e %.4f" % d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], "clk", d["clocks"]["sm_mhz"])
PY
    done
  done
done
