#!/bin/bash
mkdir -p gpurun_out
for k in test_lbfgs test_eval test_gram_and test_nst; do
  echo "=== $k"
  timeout 600 python -m pytest tests/test_gpu_nst.py -q -k $k -m gpu -s 2>&1 | grep -v "^$" | tail -40
done
echo "=== layer bench B=8"; timeout 300 python scratch/bench_layers.py 8 2>&1 | tail -20
