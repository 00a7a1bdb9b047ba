timeout 300 python -m pytest tests/test_tabulation_gpu.py tests/test_fullsize_gpu.py -m gpu -q 2>&1 | tail -2
for x in "" "--fused-exact"; do
timeout 200 python bench.py --model fused --n 1e8 --steps 10 --cpu-seconds 0 --e2e-n 0 $x 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('fused $x', d['config']['qp_per_gpu'], round(d['ms_per_step'],3), 'ms frac', round(d['roofline']['frac'],3))
    elif 'rror' in l: print(l.strip()[:300])
"
done
