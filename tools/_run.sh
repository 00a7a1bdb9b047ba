python -m pytest tests/test_forms_gpu.py tests/test_switches_gpu.py -x -q -m gpu 2>&1 | tail -2
run() { python bench.py --model $M --steps 10 --warmup 3 --cpu-seconds 0 --e2e-n 0 $X 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$M', '$X', d['ms_per_step'], d['roofline']['frac'], d['value'])"; }
M=step6; for X in "--n 1e8" "--mesh-order morton --n 1e8"; do run; EO_STEP_CELL=0 run; done
