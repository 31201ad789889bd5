#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "=== bench 60"; timeout 900 python bench.py --steps 60 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1
echo "=== bench full"; timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1
