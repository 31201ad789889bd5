#!/bin/bash
# A/B on one box: max-pool backward through index bytes (default) vs through the pre-pool activations
mkdir -p gpurun_out/r02
F="--steps 20 --no-cpu-baseline --no-gpu-library --no-e2e --no-features"
for rep in 1 2; do
  for o in 0 1; do
    python bench.py $F --opt pool_idx=$o 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('pool_idx=$o rep $rep: %.1f image-steps/s  %.2f ms/step  conv %.1f TF/s share %.3f  clocks %s' % (d['value'], d['ms_per_step'], r['achieved'], r['share_of_step'], d['clocks']['sm_mhz']))
"
  done
done > gpurun_out/r02/ab_pool_idx.txt 2>&1
for o in 0 1; do
python - <<PY >> gpurun_out/r02/ab_pool_idx.txt 2>&1
import subprocess, json
PY
done
cat gpurun_out/r02/ab_pool_idx.txt
