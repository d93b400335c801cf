#!/bin/bash
# One gpurun call: GPU test suite, then the default bench line with the per-op breakdown.  Usage: tools/gpu_run.sh TAG
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/bd_$TAG.json > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"
tail -c 1500 gpurun_out/bench_$TAG.json
