#!/bin/bash
# tools/pipe_ab.sh - shapes of the host-vector chunk pipeline (cell chunks x vector pieces); EO_FORM_PIPE=0: serial path
for cfg in ${CFGS:-"0 16 128" "1 4 32" "1 8 64" "1 16 128" "1 32 256" "1 16 64" "1 32 128"}; do set -- $cfg
  EO_FORM_PIPE=$1 EO_FORM_PIPE_CHUNKS=$2 EO_FORM_PIPE_PIECES=$3 python bench.py --models "" --steps 3 --cpu-seconds 0 ${EXTRA:-} 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); dc=d['e2e_device_consumers']
print('pipe=$1 chunks=$2 pieces=$3 residual %.3f ms %.2f GQP/s | action %.3f ms %.2f GQP/s' % (dc['residual']['ms_per_step'], dc['residual']['value']/1e9, dc['tangent_action']['ms_per_step'], dc['tangent_action']['value']/1e9))"
done
