#!/bin/bash
F="--steps 20 --no-cpu-baseline --no-gpu-library --no-e2e --no-features"
for rep in 1 2; do
  for o in 1 2; do
    python bench.py $F --opt sweep64=$o 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('nst640 sweep64=$o rep $rep: %.1f image-steps/s  %.2f ms/step  conv %.1f TF/s frac %.3f clocks %s' % (d['value'], d['ms_per_step'], r['achieved'], r['frac'], d['clocks']['sm_mhz']))
"
  done
done
for o in 1 2; do python bench.py --config nst224 --steps 40 --no-cpu-baseline --no-gpu-library --no-e2e --opt sweep64=$o 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('nst224 sweep64=$o: %.1f image-steps/s' % d['value'])
"; done
