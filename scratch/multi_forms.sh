#!/bin/bash
N=${1:-2}
P=29611
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N "$@"; P=$((P+1)); }
run --steps 10 --warmup 3 > gpurun_out/bench_r1f_vm_n$N.json 2> gpurun_out/bench_r1f_vm_n$N.err
run --model step --steps 10 --warmup 3 > gpurun_out/bench_r1f_step_n$N.json 2> gpurun_out/bench_r1f_step_n$N.err
run --model action --steps 10 --warmup 3 > gpurun_out/bench_r1f_action_n$N.json 2> gpurun_out/bench_r1f_action_n$N.err
python - <<PY
import json
for f in ['vm','step','action']:
    try:
        d=json.loads(open('gpurun_out/bench_r1f_%s_n$N.json'%f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], '%.3f GQP/s'%(d['value']/1e9), d.get('ms_per_step'), 'e2e', d.get('e2e') and d['e2e']['value']/1e6, 'dc', d.get('e2e_device_consumers') and {k:(v['value']/1e6 if isinstance(v,dict) else None) for k,v in d['e2e_device_consumers'].items() if isinstance(v,dict)})
    except Exception as e:
        print(f, 'ERR', e); print(open('gpurun_out/bench_r1f_%s_n$N.err'%f).read()[-800:])
PY
