#!/bin/bash
timeout 300 python -m pytest tests/test_forms_gpu.py tests/test_cylinder_gpu.py -m gpu -q 2>&1 | tail -3
for n in 2e7 1e8; do
timeout 200 python bench.py --model action --n $n --steps 10 --cpu-seconds 0 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('action', d['config']['qp_per_gpu'], round(d['ms_per_step'],3), 'ms frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']/1e9,3))
    elif 'rror' in l: print(l.strip()[:300])
"
done
