#!/bin/bash
# ncu --set full of the final conv kernels (one launch each of: conv_c64 fwd, conv_c64 dgrad+mask+Gram, conv_halo<128> conv2_2 fwd /
# dgrad+mask+Gram, conv_halo<128> conv3_2 fwd / dgrad+mask+Gram, conv1_1_tail) -> profiles/r01_ncu_new_kernels_full_raw.csv
mkdir -p gpurun_out
python scratch/ncu_target.py 1 > gpurun_out/plain_target.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv_halo|conv_c64|conv1_1_tail" -s 0 -c 7 -o gpurun_out/prof_r01_new_kernels -f python scratch/ncu_target.py 1 > gpurun_out/ncu_target.log 2>&1
tail -2 gpurun_out/ncu_target.log
ncu -i gpurun_out/prof_r01_new_kernels.ncu-rep --page raw --csv > gpurun_out/r01_ncu_new_kernels_full_raw.csv 2>/dev/null
ls -la gpurun_out/prof_r01_new_kernels.ncu-rep gpurun_out/r01_ncu_new_kernels_full_raw.csv
