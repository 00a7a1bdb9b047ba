#!/bin/bash
echo "== direct"; EO_JIT_STAGED=0 python scratch/jit_bw.py 2>&1 | tail -13
echo "== auto";   python scratch/jit_bw.py 2>&1 | tail -13
ncu --set full --clock-control none -k regex:eo_jit_entry -s 6 -c 1 -f -o gpurun_out/prof_r1g_jit_staged python scratch/jit_bw.py 9 > gpurun_out/ncu_r1g_jit_staged.log 2>&1
EO_JIT_STAGED=0 ncu --set full --clock-control none -k regex:eo_jit_entry -s 6 -c 1 -f -o gpurun_out/prof_r1g_jit_direct python scratch/jit_bw.py 9 > gpurun_out/ncu_r1g_jit_direct.log 2>&1
