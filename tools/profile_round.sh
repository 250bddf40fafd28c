#!/bin/bash
# Evidence for profiles/: bench line, ncu launch list of the same command, one --set full capture of the top kernels.
# usage (GPU box): bash tools/profile_round.sh <tag>
set -u
tag=${1:-rXX}
out=gpurun_out
python bench.py --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err || exit 1
cmd="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$cmd > $out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/launches_$tag.csv $cmd > $out/ncu_launches_$tag.log 2>&1
$cmd > $out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16_kernel|attention_pp_kernel|attention_kernel|rollout' -s 150 -c 12 -o $out/prof_$tag $cmd > $out/ncu_full_$tag.log 2>&1
ls -la $out | tail -8
