# same-box A/B of an env knob on the 384-px models: bash tools/ab384.sh KNOB=VALUE
show='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], "%.0f img/s %.3f ms" % (d["value"], d["ms_per_step"]), {k:v["ms"] for k,v in d["kernels"].items() if v["ms"]>0.5})'
for i in 1 2 3; do
  env $1 python bench.py --model vit_b_16_384 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$show" "knob"
  python bench.py --model vit_b_16_384 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$show" "tree"
done
