#!/bin/bash
# tools/sanitize.sh - compute-sanitizer passes (memcheck, initcheck, racecheck, synccheck) over __graft_entry__.smoke(),
# i.e. one small evaluation of every kernel on the hot path; run under gpurun.  Logs: gpurun_out/sanitize_<tool>.log
set -u
mkdir -p gpurun_out
for TOOL in memcheck initcheck racecheck synccheck; do
  timeout ${SAN_TIMEOUT:-500} compute-sanitizer --tool $TOOL --error-exitcode 86 --print-limit 20 \
      python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_$TOOL.log 2>&1
  echo "$TOOL rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$TOOL.log | tail -1)"
done
