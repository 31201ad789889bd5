#!/bin/bash
echo "=== kernel table generic (c64 off)"; ISX_C64=0 timeout 600 python scratch/kernel_table.py 32 2>&1 | grep -E "conv|tail"
echo "=== kernel table c64"; ISX_C64=1 timeout 600 python scratch/kernel_table.py 32 2>&1 | grep -E "conv1_|tail"
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -3
echo "=== bench generic"; ISX_C64=0 timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-features 2>&1 | tail -1 | cut -c1-250
echo "=== bench c64"; ISX_C64=1 timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-features 2>&1 | tail -1 | cut -c1-250
