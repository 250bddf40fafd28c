# same-box A/B: the in-tree library vs another build of it (default tools/libv_alt.so)
# usage: bash tools/ab_lib.sh [rounds] [other.so]
rounds=${1:-4}; alt=${2:-tools/libv_alt.so}
show='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], "%.0f img/s %.3f ms  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k:v["ms"] for k,v in d["kernels"].items() if v["ms"]>0.5})'
for i in $(seq $rounds); do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "$show" "tree"
  VITB200_LIB=$alt python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "$show" "alt "
done
