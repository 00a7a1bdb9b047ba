#!/bin/bash
# A/B of the form-kernel variants (env switches read by form.cu)
B="--n 1e8 --steps 10 --cpu-seconds 0"
run() { # name model env...
  name=$1; model=$2; shift 2
  env "$@" python bench.py --model $model $B > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_$name.json').read().strip().splitlines()[-1])
    print("$name", round(d['value']/1e9,3), 'GQP/s', round(d['ms_per_step'],3), 'ms frac', round(d['roofline']['frac'],3), 'e2e', d['e2e'] and round(d['e2e']['value']/1e6,1))
except Exception as e:
    print("$name ERR", e, open('gpurun_out/ab_$name.err').read()[-400:])
PY
}
run action_point_nopf action EO_FORM_ACTION_CELL=0 EO_FORM_PREFETCH=0
run action_point_pf action EO_FORM_ACTION_CELL=0 EO_FORM_PREFETCH=1
run action_cell action EO_FORM_ACTION_CELL=1
run step_nopf step EO_FORM_PREFETCH=0
run step_pf step EO_FORM_PREFETCH=1
python -m pytest tests/test_forms_gpu.py -m gpu -q 2>&1 | tail -2
EO_FORM_ACTION_CELL=1 python -m pytest tests/test_forms_gpu.py -m gpu -q 2>&1 | tail -2
