#!/bin/bash
# Round-1 profile capture: launch list + one --set full capture per kernel (run under gpurun, one GPU).
set -x
mkdir -p gpurun_out
B="--cpu-seconds 0 --e2e-n 0"
for m in mc tab fused isihara heat; do
  N=1e8; [ $m = mc ] && N=2e7; [ $m = isihara ] && N=2e7
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_$m.csv \
    python bench.py --model $m --n $N --steps 5 --warmup 3 $B > gpurun_out/launches_r1_$m.log 2>&1
done
for m in mc tab fused isihara; do
  N=1e8; [ $m = mc ] && N=2e7; [ $m = isihara ] && N=2e7
  K=${m}_kernel; [ $m = fused ] && K=tab_vm_kernel
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_r1b_$m \
    python bench.py --model $m --n $N --steps 2 --warmup 3 $B > gpurun_out/ncu_r1b_$m.log 2>&1
done
# plain bench lines (no profiler)
for m in mc tab fused isihara heat; do
  N=1e8; [ $m = isihara ] && N=2e7
  timeout 600 python bench.py --model $m --n $N --steps 10 --warmup 3 > gpurun_out/bench_r1b_$m.json 2> gpurun_out/bench_r1b_$m.err
done
timeout 600 python bench.py > gpurun_out/bench_r1b_vm.json 2> gpurun_out/bench_r1b_vm.err
tail -c 600 gpurun_out/bench_r1b_*.json
