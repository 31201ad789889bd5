#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== bench B=64 steps 40"; timeout 900 python bench.py --steps 40 --warmup 3 2>&1 | tail -3
echo "=== bench reference"; timeout 300 python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tail -2
