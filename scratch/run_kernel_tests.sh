#!/bin/bash
# run each kernel-test group in its own process so one CUDA fault does not poison the rest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt; lscpu | grep "Model name" >> gpurun_out/gpu.txt
for k in test_maxpool test_content test_bn_stats test_conv1_1 test_conv3x3_fwd test_conv3x3_dgrad test_gram_fwd test_gram_bwd; do
  echo "=== $k" 
  timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k $k -m gpu -x 2>&1 | tail -25
done
