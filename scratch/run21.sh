#!/bin/bash
echo "=== halo kernel tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "halo" 2>&1 | tail -3
echo "=== kernel table halo2 forced"; ISX_HALO2=2 timeout 600 python scratch/kernel_table.py 32 2>&1 | grep -E "conv[23]_"
