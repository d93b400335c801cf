#!/bin/bash
# Evidence pass on a B200 (run under gpurun): default bench line, ncu launch list, DRAM traffic of every
# launch of one step, one `--set full` capture of the conv and attention kernels, attention phase trace.
# usage: bash tools/gpu_evidence.sh <tag>
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
python bench.py > $out/bench_${tag}.json 2> $out/bench_${tag}.err || { tail -20 $out/bench_${tag}.err; exit 1; }
cat $out/bench_${tag}.json
python tools/attn_trace.py > $out/attn_trace_${tag}.log 2>&1; cat $out/attn_trace_${tag}.log
CMD="python bench.py --steps 1 --warmup 3 --no-cpu"
$CMD > $out/plain_${tag}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file $out/launches_${tag}.csv $CMD > $out/ncu_launches_${tag}.log 2>&1
$CMD > $out/plain_${tag}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_conv_tc3 -s 81 -c 7 -f -o $out/prof_conv3_${tag} \
    $CMD > $out/ncu_conv3_${tag}.log 2>&1
$CMD > $out/plain_${tag}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_attn_tc -s 11 -c 1 -f -o $out/prof_attn_${tag} \
    $CMD > $out/ncu_attn_${tag}.log 2>&1
ls -la $out | tail -12
