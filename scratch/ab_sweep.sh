#!/bin/bash
# A/B on one box: conv1_2 forward on the tap-stacked sweep kernel (sweep64=1, default) vs conv_c64 (sweep64=0)
mkdir -p gpurun_out/r02
F="--steps 20 --no-cpu-baseline --no-gpu-library --no-e2e --no-features"
for rep in 1 2; do
  for o in 0 1; do
    python bench.py $F --opt sweep64=$o 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('nst640 sweep64=$o rep $rep: %.1f image-steps/s  %.2f ms/step  conv %.1f TF/s frac %.3f clocks %s' % (d['value'], d['ms_per_step'], r['achieved'], r['frac'], d['clocks']['sm_mhz']))
"
  done
done > gpurun_out/r02/ab_sweep.txt 2>&1
for o in 0 1; do
  python bench.py --config feat4 --opt sweep64=$o 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('feat4 sweep64=$o: %.1f images/s' % d['value'])
" >> gpurun_out/r02/ab_sweep.txt
done
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nst.py -m gpu -x -q 2>&1 | tail -2 >> gpurun_out/r02/ab_sweep.txt
cat gpurun_out/r02/ab_sweep.txt
