#!/bin/bash
python scratch/mc_dbg.py 2e7 0,3 2>&1 | grep scheme
python bench.py --model mc --n 1e8 --steps 10 --cpu-seconds 0 --e2e-n 0 > gpurun_out/bench_r1d_mc.json 2>gpurun_out/bench_r1d_mc.err; tail -c 400 gpurun_out/bench_r1d_mc.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1d_mc.csv python bench.py --model mc --n 2e7 --steps 5 --warmup 3 --cpu-seconds 0 --e2e-n 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mc_kernel|mc_trial" -s 40 -c 2 -f -o gpurun_out/prof_r1d_mc python bench.py --model mc --n 2e7 --steps 2 --warmup 3 --cpu-seconds 0 --e2e-n 0 > gpurun_out/ncu_r1d_mc.log 2>&1
grep -v "^==" gpurun_out/launches_r1d_mc.csv | tail -12
