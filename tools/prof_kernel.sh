#!/bin/bash
# tools/prof_kernel.sh MODEL KERNEL_REGEX N TAG [extra bench args] - one `ncu --set full` capture of a kernel of
# `bench.py --model MODEL` plus the launch list of the same command (B200_PROFILING.md recipe); run under gpurun.
# Outputs in gpurun_out/: TAG.ncu-rep, TAG_raw.csv (--page raw), TAG_launches.csv
set -u
MODEL=$1; KREGEX=$2; N=$3; TAG=$4; shift 4
ARGS="--model $MODEL --n $N --warmup 3 --cpu-seconds 0 --e2e-n 0 $*"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py $ARGS --steps 3 > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" -s ${SKIP:-4} -c 1 -f -o gpurun_out/${TAG} \
    python bench.py $ARGS --steps 2 > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
tail -2 gpurun_out/${TAG}_full.log
