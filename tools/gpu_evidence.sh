#!/bin/bash
# Evidence pass on a B200 (run under gpurun): GPU test suite with the parity numbers printed, default bench line,
# ncu launch list (+ DRAM bytes) of a few steps, one `--set full` capture of the conv and attention kernels.
# usage: bash tools/gpu_evidence.sh <tag>
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > $out/pytest_${tag}.log 2>&1; echo "pytest exit $?" >> $out/pytest_${tag}.log
grep -E "parity|passed|failed" $out/pytest_${tag}.log | tail -40
python bench.py --breakdown $out/bd_${tag}.json > $out/bench_${tag}.json 2> $out/bench_${tag}.err || { tail -20 $out/bench_${tag}.err; exit 1; }
tail -c 600 $out/bench_${tag}.json
for wl in c1 c2; do   # per-op breakdown of the launch-bound workloads
  python bench.py --workload $wl --steps 50 --warmup 10 --no-cpu --no-secondary --short-e2e --breakdown $out/bd_${wl}_${tag}.json \
    > $out/bench_${wl}_${tag}.json 2> $out/bench_${wl}_${tag}.err
done
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary --short-e2e"
$CMD > $out/plain_${tag}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file $out/launches_${tag}.csv $CMD > $out/ncu_launches_${tag}.log 2>&1
$CMD > $out/plain_${tag}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_conv_tc3 -s 83 -c 12 -f -o $out/prof_conv3_${tag} \
    $CMD > $out/ncu_conv3_${tag}.log 2>&1
$CMD > $out/plain_${tag}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_attn_tc6 -s 11 -c 1 -f -o $out/prof_attn_${tag} \
    $CMD > $out/ncu_attn_${tag}.log 2>&1
ls -la $out | tail -8
