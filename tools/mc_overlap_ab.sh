#!/bin/bash
# tools/mc_overlap_ab.sh - A/B of the Mohr-Coulomb overlapped scheme (EO_MC_OVERLAP = min_n,chunk,first_mult; "0" = two-launch
# scheme) at 2e7 and 1e8 points; run under gpurun
for c in ${CONFIGS:-0 3000000,0,8}; do for n in ${SIZES:-2e7 1e8}; do
  tag=$(echo $c | tr ',' '_')
  EO_MC_OVERLAP=$c timeout 300 python bench.py --model mc --n $n --steps 5 --cpu-seconds 0 --e2e-n 0 > gpurun_out/mc_ov_${tag}_${n}.json 2>gpurun_out/mc_ov_${tag}_${n}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/mc_ov_${tag}_${n}.json").read().strip().splitlines()[-1])
    print("overlap $c n $n kernel_ms %.3f frac %.3f launches %s" % (d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["gpu_launches"]))
except Exception as e:
    print("overlap $c n $n FAILED", e)
PY
done; done
