SKIP=4 bash tools/prof_kernel.sh step form_vm_step_cell 1e8 r2l_step
SKIP=4 bash tools/prof_kernel.sh action form_action_tma 1e8 r2l_action
SKIP=4 bash tools/prof_kernel.sh action6 form_action_vm6 1e8 r2l_action6
rm -f gpurun_out/r2l_*.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2l_default_cmd_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r2l_default_cmd.log 2>&1
