#!/bin/bash
for st in 1; do echo "== EO_JIT_STAGED=$st"; EO_JIT_STAGED=$st python scratch/jit_bw.py 2>&1 | tail -13; done
echo "== auto"; python scratch/jit_bw.py 2>&1 | tail -13
python -m pytest tests/test_jit_gpu.py tests/test_assign_gpu.py -x -q 2>&1 | tail -3
for st in 0 1; do EO_JIT_STAGED=$st python bench.py --model jitvm --steps 10 --cpu-seconds 0 --e2e-n 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('jitvm staged=$st', d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])"; done
