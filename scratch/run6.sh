#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "=== bench full"; timeout 900 python bench.py --no-cpu-baseline --no-e2e 2>&1 | tail -1
