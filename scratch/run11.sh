#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest default"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== persist kernel tests + table"; ISX_PERSIST=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q 2>&1 | tail -2
ISX_PERSIST=1 timeout 600 python scratch/kernel_table.py 32 2>&1 | grep -E "conv2_2|conv3_2|conv4_2|tail|head" 
echo "=== b1 timing"; timeout 600 python scratch/b1_timing.py 2>&1 | tail -10
echo "=== bench"; timeout 900 python bench.py --no-cpu-baseline --no-e2e --no-features 2>&1 | tail -1 | cut -c1-300
