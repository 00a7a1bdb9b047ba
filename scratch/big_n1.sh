#!/bin/bash
# 5e8 quadrature points on ONE GPU (von Mises 120 GB, Mohr-Coulomb 126 GB of the 180 GB)
timeout 600 python bench.py --n 5e8 --steps 5 --warmup 3 --cpu-seconds 0 --no-device-consumers > gpurun_out/bench_5e8_vm_n1.json 2> gpurun_out/bench_5e8_vm_n1.err
timeout 700 python bench.py --model mc --n 5e8 --steps 3 --warmup 3 --cpu-seconds 0 --e2e-n 0 > gpurun_out/bench_5e8_mc_n1.json 2> gpurun_out/bench_5e8_mc_n1.err
python - <<PY
import json
for f in ['vm','mc']:
    try:
        d=json.loads(open('gpurun_out/bench_5e8_%s_n1.json'%f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['config']['qp_per_gpu'], '%.3f GQP/s'%(d['value']/1e9), d.get('ms_per_step'), 'frac', d['roofline'].get('frac'), 'e2e', d.get('e2e') and d['e2e']['value']/1e6)
    except Exception as e:
        print(f, 'ERR', e); print(open('gpurun_out/bench_5e8_%s_n1.err'%f).read()[-800:])
PY
nvidia-smi --query-gpu=memory.total --format=csv
