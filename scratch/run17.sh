#!/bin/bash
echo "=== kernel tests c64 forced"; ISX_C64=2 timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -8
echo "=== nst tests c64 forced"; ISX_C64=2 timeout 900 python -m pytest tests/test_gpu_nst.py -m gpu -q 2>&1 | tail -5
echo "=== kernel table c64"; ISX_C64=1 timeout 600 python scratch/kernel_table.py 32 2>&1 | grep -E "conv1_|tail"
echo "=== bench c64"; ISX_C64=1 timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-features 2>&1 | tail -1 | cut -c1-250
