#!/bin/bash
echo "=== halo kernel tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "halo" 2>&1 | tail -15
