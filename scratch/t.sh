timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for m in action step; do
timeout 200 python bench.py --model $m --n 1e8 --steps 10 --cpu-seconds 0 --e2e-n 0 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$m', d['config']['qp_per_gpu'], round(d['ms_per_step'],3), 'ms frac', round(d['roofline']['frac'],3))
    elif 'rror' in l: print(l.strip()[:300])
"
done
timeout 200 python scratch/time_vec.py 2>&1 | tail -2
