#!/bin/bash
# BASELINE configs 3 and 5 (and 2) on N GPUs of one box:  gpurun --gpus N -- 'bash tools/scale_configs.sh N'
# One JSON line per (config, N) in gpurun_out/scale_<model>_b<batch>_<N>gpu.json.
N=${1:-1}
mkdir -p gpurun_out
for spec in "vit_s_16 1024" "vit_h_16_384 16" "vit_b_16 256"; do
  set -- $spec
  out=gpurun_out/scale_$1_b$2_${N}gpu.json
  if [ "$N" = "1" ]; then
    timeout 300 python bench.py --gpus 1 --model $1 --batch $2 --no-cpu-baseline --long-seconds 1.5 > $out 2> ${out%.json}.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --model $1 --batch $2 --no-cpu-baseline --long-seconds 1.5 > $out 2> ${out%.json}.err
  fi
  python - $out <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%s N=%d: %.0f img/s  %.3f ms/step  e2e %.0f  sustained %.0f  per-rank %s  gather=%s" % (
        d["config"]["workload"][:28], d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"],
        d.get("sustained", {}).get("value", 0), d.get("per_rank", {}).get("ms_per_step"), d["config"].get("gather")))
except Exception as ex:
    print("FAILED", sys.argv[1], ex)
P
done
