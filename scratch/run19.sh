#!/bin/bash
echo "=== kernel tests (c64)"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "c64" 2>&1 | tail -3
echo "=== kernel table"; timeout 600 python scratch/kernel_table.py 32 2>&1 | grep -E "conv1_"
echo "=== nst tests"; timeout 900 python -m pytest tests/test_gpu_nst.py -m gpu -q -x 2>&1 | tail -3
echo "=== bench"; timeout 900 python bench.py --no-cpu-baseline --no-features 2>&1 | tail -1 > gpurun_out/bench_r19.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r19.json').read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['evals'], d['clocks'])
PY
