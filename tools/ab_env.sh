# same-box A/B of an environment knob: bash tools/ab_env.sh VAR=VALUE [rounds]
knob=$1; rounds=${2:-4}
for i in $(seq $rounds); do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('base   %.0f img/s %.3f ms  e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']), {k:v['ms'] for k,v in d['kernels'].items() if v['ms']>0.5})"
  env $knob python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('knob   %.0f img/s %.3f ms  e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']), {k:v['ms'] for k,v in d['kernels'].items() if v['ms']>0.5})"
done
