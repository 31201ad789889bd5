#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests persist"; ISX_PERSIST=1 timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -6
echo "=== nst tests persist"; ISX_PERSIST=1 timeout 900 python -m pytest tests/test_gpu_nst.py -m gpu -q 2>&1 | tail -6
echo "=== kernel table persist"; ISX_PERSIST=1 timeout 600 python scratch/kernel_table.py 32 2>&1 | grep -E "conv|head|tail"
echo "=== bench persist"; ISX_PERSIST=1 timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-features 2>&1 | tail -1 | cut -c1-300
echo "=== pytest default"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
