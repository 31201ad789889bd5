#!/bin/bash
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q 2>&1 | tail -3
echo "=== kernel table"; timeout 600 python scratch/kernel_table.py 32 2>&1 | grep -E "conv|tail"
echo "=== bench"; timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-features 2>&1 | tail -1 | cut -c1-250
