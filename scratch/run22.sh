#!/bin/bash
echo "=== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "=== kernel table default"; timeout 600 python scratch/kernel_table.py 32 2>&1 | grep -E "conv"
echo "=== bench"; timeout 900 python bench.py --no-cpu-baseline --no-features 2>&1 | tail -1 > gpurun_out/bench_r22.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r22.json').read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['other'])
PY
