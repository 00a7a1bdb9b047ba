#!/bin/bash
for n in 2e7 5e7 1e8; do
  echo "== n=$n"
  EO_FORM_ACTION_TMA2=1 timeout 100 python bench.py --model action --n $n --steps 3 --warmup 1 --cpu-seconds 0 2>&1 | grep -v "^$" | cut -c1-300 | tail -4
done
