#!/bin/bash
echo "=== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"; timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_r24.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r24.json').read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['roofline']['achieved'], d['gpu_launches'], d['secondary'])
PY
