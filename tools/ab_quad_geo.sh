# A/B of the tangent store form (EO_QUAD_STORE) and the per-cell geometry cache (EO_GEOM_CACHE) in the fused and
# residual-step kernels; run under gpurun.  Prints model, switches, ms per step, roofline fraction.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_tabulation_gpu.py tests/test_forms_gpu.py tests/test_cylinder_gpu.py -x -q -m gpu 2>&1 | tail -4
for M in ${MODELS:-fused step}; do
 for Q in 0 1; do for G in 0 1; do
  EO_QUAD_STORE=$Q EO_GEOM_CACHE=$G python bench.py --model $M --n 1e8 --steps 10 --warmup 3 --cpu-seconds 0 --e2e-n 0 ${EXTRA:-} > gpurun_out/ab_${M}_q${Q}g${G}.json 2> gpurun_out/ab_${M}_q${Q}g${G}.err
  python - <<P
import json
d=json.loads(open("gpurun_out/ab_${M}_q${Q}g${G}.json").read().strip().splitlines()[-1])
print("$M quad=$Q geo=$G", d.get("ms_per_step"), d["roofline"]["frac"])
P
 done; done
done
