#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python scratch/time_mat.py 2>&1 | tail -2
timeout 300 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/final3_reference.json 2> gpurun_out/final3_reference.err
timeout 600 python bench.py --gpus 1 > gpurun_out/final3_vm.json 2> gpurun_out/final3_vm.err
python - <<'PY'
import json
for f in ['reference','vm']:
    try:
        d=json.loads(open(f'gpurun_out/final3_{f}.json').read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, round(d['value']/1e9,3), 'GQP/s', d.get('ms_per_step'), 'frac', r.get('frac'), 'e2e', d.get('e2e') and round(d['e2e']['value']/1e6,1), 'cpu', d.get('cpu_baseline') and round(d['cpu_baseline']['value']/1e6,2), 'launches', d.get('gpu_launches'), d.get('clocks'), 'dc', d.get('e2e_device_consumers') and {k:round(v['value']/1e6,1) for k,v in d['e2e_device_consumers'].items() if isinstance(v,dict)})
    except Exception as e:
        print(f, 'ERR', e, open(f'gpurun_out/final3_{f}.err').read()[-600:])
PY
