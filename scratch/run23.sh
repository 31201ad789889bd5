#!/bin/bash
echo "=== tail tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv1_1 or tail" 2>&1 | tail -5
echo "=== kernel table"; timeout 600 python scratch/kernel_table.py 32 2>&1 | grep -E "conv1_1"
