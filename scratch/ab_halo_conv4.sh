#!/bin/bash
# A/B on one box: conv4_x (80x50) on the generic kernel (default: pair tiles would be 22-26 % padding) vs forced onto conv_halo
# (--opt halo2=2 sends EVERY applicable 3x3 layer to the halo kernel, conv4_x and the tiny layers included)
F="--steps 30 --no-e2e --no-features --no-cpu-baseline --no-gpu-library"
for rep in 1 2; do
  for opt in "" "--opt halo2=2"; do
    python bench.py $F $opt 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][0])
print('$opt' or 'default', 'value %.1f' % d['value'], 'ms %.2f' % d['ms_per_step'], 'conv frac %.4f' % d['roofline']['frac'], 'conv ms/launch %.4f' % d['roofline']['avg_launch_ms'], d['clocks']['sm_mhz'])
"
  done
done
