#!/bin/bash
# feature-leg sweep: images per GPU x batch per forward pass
mkdir -p gpurun_out/r02
for cfg in "512 32" "512 64" "1280 32" "1280 64" "1280 128"; do
  set -- $cfg
  python bench.py --config feat4 --feature-images $1 --feature-batch $2 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('feat4 images $1 batch $2 ->', round(d['value'], 1), 'img/s')
"
done > gpurun_out/r02/feat_sweep.txt 2>&1
python bench.py --config feat5 --feature-images 1280 --feature-batch 64 2>/dev/null | grep value | head -c 300 >> gpurun_out/r02/feat_sweep.txt
echo sweep-done
