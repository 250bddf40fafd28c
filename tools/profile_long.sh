#!/bin/bash
# ncu capture of the key-blocked attention kernels inside a 384-px forward
out=gpurun_out
cmd="python bench.py --model vit_b_16_384 --batch 64 --steps 2 --warmup 3 --no-cpu-baseline"
$cmd > $out/plain_long.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'attention_long' -s 24 -c 4 -o $out/prof_long $cmd > $out/ncu_long.log 2>&1
ls -la $out/prof_long.ncu-rep
