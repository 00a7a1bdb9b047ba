for m in tab fused step action; do
  timeout 250 python bench.py --model $m --n 1e8 --steps 10 --cpu-seconds 3 > gpurun_out/final4_$m.json 2> gpurun_out/final4_$m.err
done
python - <<'PY'
import json
for f in ['tab','fused','step','action']:
    try:
        d=json.loads(open(f'gpurun_out/final4_{f}.json').read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, round(d['value']/1e9,3), 'GQP/s', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), 'gathered', r.get('gathered') and round(r['gathered']['frac_of_hbm_peak'],3), 'e2e', d.get('e2e') and round(d['e2e']['value']/1e6,1))
    except Exception as e:
        print(f, 'ERR', e, open(f'gpurun_out/final4_{f}.err').read()[-500:])
PY
