set -u
bash tools/prof_kernel.sh fused tab_vm_kernel 1e8 r2e_fused
bash tools/prof_kernel.sh step form_vm_step_kernel 1e8 r2e_step
bash tools/prof_kernel.sh action form_action_tma_kernel 1e8 r2e_action
bash tools/prof_kernel.sh tab tab_kernel 1e8 r2e_tab
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2e_default_cmd_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r2e_default_cmd.log 2>&1
rm -f gpurun_out/r2e_tab.ncu-rep
ls -la gpurun_out | tail -20
