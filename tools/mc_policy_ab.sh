#!/bin/bash
# tools/mc_policy_ab.sh - A/B of the pass-2 scheduling policy (EO_MC_POLICY=full,min_active,split_first,sleep_ns)
for pol in ${POLICIES:-"28,99,0,200" "28,99,1,200" "30,8,0,200" "30,8,1,200" "32,8,1,200" "32,6,1,100" "30,10,1,100" "31,9,1,50"}; do
  EO_MC_POLICY=$pol python bench.py --model mc --n ${N:-2e7} --steps 5 --cpu-seconds 0 --e2e-n 0 > gpurun_out/mc_pol.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/mc_pol.json").read().strip().splitlines()[-1])
print("policy $pol kernel_ms %.3f frac %.3f" % (d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
PY
done
