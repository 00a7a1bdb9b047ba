#!/bin/bash
# Size sweep (1e6..1e8 QP per GPU; 1e9 needs >= 2 GPUs for von Mises) - device-resident legs only.
mkdir -p gpurun_out
: > gpurun_out/sweep_r1.jsonl
for m in vm mc heat jitvm; do
  for n in 1e6 1e7 1e8; do
    timeout 300 python bench.py --model $m --n $n --steps 10 --warmup 3 --cpu-seconds 0 --e2e-n 0 >> gpurun_out/sweep_r1.jsonl 2>> gpurun_out/sweep_r1.err
  done
done
# heat at 1e9 points fits one GPU (88 GB)
timeout 300 python bench.py --model heat --n 1e9 --steps 5 --warmup 3 --cpu-seconds 0 --e2e-n 0 >> gpurun_out/sweep_r1.jsonl 2>> gpurun_out/sweep_r1.err
# both arms exactly as the driver runs them
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/bench_r1c_reference.json 2> gpurun_out/bench_r1c_reference.err
timeout 600 python bench.py --gpus 1 > gpurun_out/bench_r1c_vm.json 2> gpurun_out/bench_r1c_vm.err
timeout 600 python bench.py --model mc --n 1e8 --steps 10 > gpurun_out/bench_r1c_mc.json 2> gpurun_out/bench_r1c_mc.err
python - <<'PY'
import json
for l in open('gpurun_out/sweep_r1.jsonl'):
    d=json.loads(l); print(d['config']['workload'][:28], d['config']['qp_per_gpu'], '%.3f ms'%d['ms_per_step'], '%.2f GQP/s'%(d['value']/1e9), d['roofline'].get('frac'))
for f in ['reference','vm','mc']:
    d=json.loads(open(f'gpurun_out/bench_r1c_{f}.json').read().strip().splitlines()[-1]); print(f, d['value']/1e6, 'MQP/s', d.get('e2e'), d.get('cpu_baseline',{}) and d['cpu_baseline'].get('cores'))
PY
