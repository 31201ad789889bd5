#!/bin/bash
# duration of the L-BFGS control kernel at a full 100-pair history (nst224: one coupled problem)
mkdir -p gpurun_out/r02
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lbfgs_control -s 100 -c 4 --csv --log-file gpurun_out/r02/ncu_lbfgs_nst224.csv python bench.py --config nst224 --steps 5 --no-cpu-baseline --no-gpu-library --no-e2e --no-features > /dev/null 2>&1
grep lbfgs_control gpurun_out/r02/ncu_lbfgs_nst224.csv | cut -d, -f5,15 | cut -c1-60
