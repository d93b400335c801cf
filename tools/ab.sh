#!/bin/bash
# A/B of library variants on ONE box: tools/ab.sh TAG variant1 variant2 ...   ("base" = libeo_b200.so)
# Each variant: bench.py default workload, device-timed steps + per-op breakdown; two interleaved rounds.
TAG=$1; shift
mkdir -p gpurun_out
for round in 1 2; do
  for v in "$@"; do
    lib=eo_diffusion_b200/libeo_b200.so
    [ "$v" != "base" ] && lib=eo_diffusion_b200/libeo_b200_$v.so
    EO_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary --short-e2e \
      --breakdown gpurun_out/bd_${TAG}_${v}_$round.json > gpurun_out/ab_${TAG}_${v}_$round.json 2> gpurun_out/ab_${TAG}_${v}_$round.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/ab_${TAG}_${v}_$round.json").read().strip().splitlines()[-1])
print("$v", $round, "ms/step %.3f" % d["ms_per_step"], "conv %.1f TF" % d["roofline"]["achieved"], "attn %.3f ms" % d["attention"]["ms"], "clk", d["clocks"]["sm_mhz"])
PY
  done
done
