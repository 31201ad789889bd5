#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "sweep" 2>&1 | tail -2
timeout 200 python scratch/sweep_dbg4.py 2>&1 | tail -8
