#!/bin/bash
# tools/mc_ab.sh - A/B of the Mohr-Coulomb pass-2 CTA shapes (EO_MC_CONFIG) at 2e7 and 1e8 points; run under gpurun
for c in ${CONFIGS:-0 1 2}; do for n in ${SIZES:-2e7 1e8}; do
  EO_MC_CONFIG=$c python bench.py --model mc --n $n --steps 5 --cpu-seconds 0 --e2e-n 0 > gpurun_out/mc_ab_${c}_${n}.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/mc_ab_${c}_${n}.json").read().strip().splitlines()[-1])
print("config $c n $n kernel_ms %.3f frac %.3f" % (d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
PY
done; done
