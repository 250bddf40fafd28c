#!/bin/bash
# ncu --set full captures of the small-launch kernels (B = 1 forward) and of the key-blocked attention kernels
# (ViT-B/16 at 384 px, B = 16); each command runs once without ncu first.  usage (GPU box): bash tools/profile_small.sh <tag>
set -u
tag=${1:-rXX}
out=gpurun_out
export STEPS=2 BLOCKS=1 WARM=2
MODEL=vit_b_16 BATCH=1 python tools/time_forward.py > $out/plain_small_$tag.log 2>&1 &&
MODEL=vit_b_16 BATCH=1 ncu --set full --clock-control none --import-source on \
  -k regex:'rollout_cluster_kernel|attention_pp_kernel|avg_parts_sum_kernel|gemm_bf16_kernel' -s 240 -c 10 -o $out/prof_small_$tag \
  python tools/time_forward.py > $out/ncu_small_$tag.log 2>&1
MODEL=vit_b_16_384 BATCH=16 python tools/time_forward.py > $out/plain_long_$tag.log 2>&1 &&
MODEL=vit_b_16_384 BATCH=16 ncu --set full --clock-control none --import-source on \
  -k regex:'attention_long' -s 24 -c 2 -o $out/prof_long_$tag python tools/time_forward.py > $out/ncu_long_$tag.log 2>&1
ls -la $out | tail -6
