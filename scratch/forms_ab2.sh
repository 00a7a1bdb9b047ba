#!/bin/bash
B="--n 1e8 --steps 10 --cpu-seconds 0 --e2e-n 0"
run() { name=$1; model=$2; shift 2
  env "$@" python bench.py --model $model $B > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_$name.json').read().strip().splitlines()[-1])
    print("$name", round(d['value']/1e9,3), 'GQP/s', round(d['ms_per_step'],3), 'ms frac', round(d['roofline']['frac'],3))
except Exception as e:
    print("$name ERR", e, open('gpurun_out/ab_$name.err').read()[-400:])
PY
}
timeout 300 python -m pytest tests/test_forms_gpu.py tests/test_cylinder_gpu.py -m gpu -q 2>&1 | tail -3
run action_tma2 action EO_FORM_ACTION_TMA2=1
run action_tma1 action EO_FORM_ACTION_TMA2=0
