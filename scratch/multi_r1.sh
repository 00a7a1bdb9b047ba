#!/bin/bash
N=${1:-2}
P=29511
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N "$@"; P=$((P+1)); }
run --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1e_reference_n$N.json 2> gpurun_out/bench_r1e_reference_n$N.err
run --steps 10 --warmup 3 > gpurun_out/bench_r1e_vm_n$N.json 2> gpurun_out/bench_r1e_vm_n$N.err
run --model mc --steps 5 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_r1e_mc_n$N.json 2> gpurun_out/bench_r1e_mc_n$N.err
python - <<PY
import json
for f in ['reference','vm','mc']:
    try:
        d=json.loads(open('gpurun_out/bench_r1e_%s_n$N.json'%f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], '%.3f GQP/s'%(d['value']/1e9), d.get('ms_per_step'), 'e2e', d.get('e2e') and d['e2e']['value']/1e6)
    except Exception as e:
        print(f, 'ERR', e); print(open('gpurun_out/bench_r1e_%s_n$N.err'%f).read()[-800:])
PY
