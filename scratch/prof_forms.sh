#!/bin/bash
# ncu capture of the device-side consumer kernels (one GPU): --set full for the residual step and the tangent action,
# launch lists of the same bench commands, then the plain bench lines.
mkdir -p gpurun_out
B="--cpu-seconds 0 --e2e-n 0"
python -m pytest tests/test_forms_gpu.py tests/test_cylinder_gpu.py -m gpu -q 2>&1 | tail -3
for m in step action; do
  K=form_vm_step_kernel; [ $m = action ] && K=form_action_tma_kernel
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_r1f_$m \
    python bench.py --model $m --n 1e8 --steps 2 --warmup 3 $B > gpurun_out/ncu_r1f_$m.log 2>&1
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1f_$m.csv \
    python bench.py --model $m --n 1e8 --steps 5 --warmup 3 $B > gpurun_out/launches_r1f_$m.log 2>&1
done
for m in step action; do
  python bench.py --model $m --n 1e8 --steps 10 --cpu-seconds 3 > gpurun_out/r1f_$m.json 2> gpurun_out/r1f_$m.err
done
python examples/thick_walled_cylinder.py 20 64 2>&1 | tail -8
tail -c 400 gpurun_out/r1f_step.json gpurun_out/r1f_action.json
