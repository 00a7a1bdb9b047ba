#!/bin/bash
# 1e9 quadrature points in total: 5e8 per GPU on 2 GPUs (von Mises 120 GB, Mohr-Coulomb 126 GB per GPU)
N=2
P=29711
run() { timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N "$@"; P=$((P+1)); }
run --qp-per-gpu 5e8 --steps 5 --warmup 3 --cpu-seconds 0 --no-device-consumers > gpurun_out/bench_1e9_vm_n$N.json 2> gpurun_out/bench_1e9_vm_n$N.err
run --model mc --qp-per-gpu 5e8 --steps 3 --warmup 3 --cpu-seconds 0 --e2e-n 0 > gpurun_out/bench_1e9_mc_n$N.json 2> gpurun_out/bench_1e9_mc_n$N.err
python - <<PY
import json
for f in ['vm','mc']:
    try:
        d=json.loads(open('gpurun_out/bench_1e9_%s_n$N.json'%f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['config']['qp_per_gpu'], '%.3f GQP/s'%(d['value']/1e9), d.get('ms_per_step'), 'frac', d['roofline'].get('frac'), 'e2e', d.get('e2e') and d['e2e']['value']/1e6)
    except Exception as e:
        print(f, 'ERR', e); print(open('gpurun_out/bench_1e9_%s_n$N.err'%f).read()[-800:])
PY
