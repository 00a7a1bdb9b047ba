#!/bin/bash
# Round-1 closing run: full GPU suite, smoke, both arms of the headline bench, every secondary bench line.
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err
python bench.py --gpus 1 > gpurun_out/final_vm.json 2> gpurun_out/final_vm.err
for m in mc heat tab fused isihara jitvm jitfused; do
  N=1e8; [ $m = isihara ] && N=2e7
  python bench.py --model $m --n $N --steps 10 > gpurun_out/final_$m.json 2> gpurun_out/final_$m.err
done
python bench.py --model jitvm3d --n 5e7 --steps 10 --cpu-seconds 0 --e2e-n 0 > gpurun_out/final_jitvm3d.json 2> gpurun_out/final_jitvm3d.err
python - <<'PY'
import json
for f in ['reference','vm','mc','heat','tab','fused','isihara','jitvm','jitfused','jitvm3d']:
    try:
        d=json.loads(open(f'gpurun_out/final_{f}.json').read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f"{f:9s} {d['value']/1e9:9.3f} GQP/s  {str(d.get('ms_per_step'))[:7]:>8s} ms  frac {str(r.get('frac'))[:5]}  e2e {d.get('e2e') and round(d['e2e']['value']/1e6,1)}  cpu {d.get('cpu_baseline') and round(d['cpu_baseline']['value']/1e6,2)}  launches {d.get('gpu_launches')}  clocks {d.get('clocks',{}).get('sm_mhz')} {d.get('clocks',{}).get('reasons')}")
    except Exception as e:
        print(f, 'ERR', e, open(f'gpurun_out/final_{f}.err').read()[-300:])
PY
