#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "=== kernel table"; timeout 600 python scratch/kernel_table.py 32 2>&1 | grep -E "head|tail|gram fwd"
echo "=== bench full"; timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-features 2>&1 | tail -1 | cut -c1-400
