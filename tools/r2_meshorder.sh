#!/bin/bash
# tools/r2_meshorder.sh - roofline fraction and DRAM traffic of the cell-loop kernels under different mesh numberings
# (structured row-major / Z-order curve / random permutation); run under gpurun.  Output: gpurun_out/r2_meshorder.jsonl
# (bench lines) and gpurun_out/r2_meshorder_dram.csv (ncu dram bytes of one launch per case).
N=${N:-2e7}
: > gpurun_out/r2_meshorder.jsonl
: > gpurun_out/r2_meshorder_dram.csv
for order in ${ORDERS:-structured morton shuffled}; do for m in tab fused step action; do
  python bench.py --model $m --n $N --mesh-order $order --steps 5 --cpu-seconds 0 --e2e-n 0 2>/dev/null | tail -1 >> gpurun_out/r2_meshorder.jsonl
  case $m in tab) K=tab_kernel;; fused) K=tab_vm_kernel;; step) K=form_vm_step_kernel;; action) K=form_action_tma_kernel;; esac
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_op_read.sum \
      --clock-control none -k "regex:^$K\$" -s 4 -c 1 --csv python bench.py --model $m --n $N --mesh-order $order --steps 2 --cpu-seconds 0 --e2e-n 0 2>/dev/null \
      | grep -E "dram__|gpu__time|l1tex__t_sectors|lts__t_sectors" | sed "s/^/$order,$m,/" >> gpurun_out/r2_meshorder_dram.csv
done; done
python - <<'PY'
import json
for ln in open("gpurun_out/r2_meshorder.jsonl"):
    d=json.loads(ln); c=d["config"]; r=d["roofline"]
    print(c["mesh_order"], c["workload"][:28], "n", c["qp_per_gpu"], "kernel_ms %.3f frac %.3f" % (r["kernel_ms"], r["frac"]))
PY
