for spec in "vit_s_16 256" "vit_s_16 1024" "vit_s_16 4096" "vit_l_16 128" "vit_b_16_384 64" "vit_h_16_384 16" "vit_b_16 1" "vit_b_16 8" "vit_b_16 64"; do
  set -- $spec
  python bench.py --model $1 --batch $2 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1 B=$2: %.0f img/s  %.3f ms/step  e2e %.0f img/s  tensor %.1f%% burst / %.1f%% sustained  kernels: %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], 100*d['step_tensor']['frac_of_burst_peak'], 100*d['step_tensor']['frac_of_sustained_peak'], {k:v['ms'] for k,v in d['kernels'].items() if v['ms']>0.05}))"
done
