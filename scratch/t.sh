timeout 300 python -m pytest tests/test_forms_gpu.py tests/test_cylinder_gpu.py -m gpu -q 2>&1 | tail -3
for v in 1 0; do
for x in "" "--fused-exact"; do
EO_FORM_STEP2=$v timeout 200 python bench.py --model step --n 1e8 --steps 10 --cpu-seconds 0 --e2e-n 0 $x 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('step two_phase=$v $x', d['config']['qp_per_gpu'], round(d['ms_per_step'],3), 'ms frac', round(d['roofline']['frac'],3))
    elif 'rror' in l: print(l.strip()[:300])
"
done
done
