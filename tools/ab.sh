#!/bin/bash
# Same-box A/B: alternate the baseline worktree (ab_old/) and the working tree, N rounds; prints ms per step.
N=${1:-3}
for i in $(seq $N); do
  for side in ab_old .; do
    (cd $side && python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']
print('$side', round(d['value']), '%.3f ms' % d['ms_per_step'], {n:k[n]['ms'] for n in ('gemm_qkv','attention','gemm_out_proj','gemm_fc1_gelu','gemm_fc2') if n in k})")
  done
done
