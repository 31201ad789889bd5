#!/bin/bash
mkdir -p gpurun_out
for st in 1 2 4; do
echo "=== bench streams $st"; timeout 900 python bench.py --steps 150 --no-cpu-baseline --no-e2e --no-features --streams $st 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['share_of_step'], d['roofline']['other'], d['clocks'])"
done
echo "=== e2e streams 2"; timeout 900 python bench.py --steps 100 --no-cpu-baseline --no-features --streams 2 2>&1 | tail -1 | cut -c1-1500
