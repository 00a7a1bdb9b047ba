#!/bin/bash
EO_FORM_STEP_DIRECT=1 timeout 300 python -m pytest tests/test_forms_gpu.py -m gpu -q -k "vm_residual or cylinder" 2>&1 | tail -2
for v in 1 0; do
EO_FORM_STEP_DIRECT=$v timeout 200 python bench.py --model step --n 1e8 --steps 10 --cpu-seconds 0 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('step direct=$v', d['config']['qp_per_gpu'], round(d['ms_per_step'],3), 'ms frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']/1e9,3))
    elif 'rror' in l: print(l.strip()[:300])
"
done
