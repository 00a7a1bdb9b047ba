#!/bin/bash
B="--n 1e8 --steps 10 --cpu-seconds 0 --e2e-n 0"
run() { name=$1; model=$2; shift 2
  extra=""; if [ "${@: -2:1}" = "--" ]; then extra="${@: -1}"; set -- "${@:1:$#-2}"; fi; env "$@" python bench.py --model $model $B $extra > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_$name.json').read().strip().splitlines()[-1])
    print("$name", round(d['value']/1e9,3), 'GQP/s', round(d['ms_per_step'],3), 'ms frac', round(d['roofline']['frac'],3))
except Exception as e:
    print("$name ERR", e, open('gpurun_out/ab_$name.err').read()[-400:])
PY
}
timeout 300 python -m pytest tests/test_forms_gpu.py -m gpu -q 2>&1 | tail -3
EO_FORM_STEP_TMA=0 EO_FORM_ACTION_TMA=0 timeout 300 python -m pytest tests/test_forms_gpu.py -m gpu -q 2>&1 | tail -3
run step_tma step EO_FORM_STEP_TMA=1
run step_tma_exact step EO_FORM_STEP_TMA=1 -- --fused-exact
run step_point step EO_FORM_STEP_TMA=0
run action_tma action EO_FORM_ACTION_TMA=1
