# same-box A/B/C: baseline library (tools/libv_base.so) vs the working tree, optionally with an env knob
# usage: bash tools/ab_lib.sh [rounds] [KNOB=VALUE]
rounds=${1:-4}; knob=${2:-X_UNUSED=1}
show='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], "%.0f img/s %.3f ms  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k:v["ms"] for k,v in d["kernels"].items() if v["ms"]>0.5})'
for i in $(seq $rounds); do
  VITB200_LIB=tools/libv_base.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "$show" base
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "$show" "new "
  env $knob python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "$show" "knob"
done
