#!/bin/bash
# device-side consumers: GPU suite, smoke, bench legs step / action, launch list
python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for m in step action; do
  python bench.py --model $m --n 1e8 --steps 10 --cpu-seconds 3 > gpurun_out/r1f_$m.json 2> gpurun_out/r1f_$m.err
done
python bench.py --model step --n 1e8 --steps 10 --cpu-seconds 0 --fused-exact > gpurun_out/r1f_step_exact.json 2> gpurun_out/r1f_step_exact.err
python bench.py --gpus 1 --steps 10 --cpu-seconds 3 > gpurun_out/r1f_vm.json 2> gpurun_out/r1f_vm.err
python - <<'PY'
import json
for f in ['step','action','step_exact','vm']:
    try:
        d=json.loads(open(f'gpurun_out/r1f_{f}.json').read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f"{f:10s} {d['value']/1e9:9.3f} GQP/s {str(d.get('ms_per_step'))[:7]:>8s} ms frac {str(r.get('frac'))[:5]} e2e {d.get('e2e') and round(d['e2e']['value']/1e6,1)} ({d.get('e2e') and d['e2e'].get('ms_per_step')}) cpu {d.get('cpu_baseline') and round(d['cpu_baseline']['value']/1e6,2)} launches {d.get('gpu_launches')} dc {d.get('e2e_device_consumers') and {k:v for k,v in d['e2e_device_consumers'].items() if k!='api'}}")
    except Exception as e:
        print(f, 'ERR', e, open(f'gpurun_out/r1f_{f}.err').read()[-600:])
PY
