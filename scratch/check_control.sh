#!/bin/bash
mkdir -p gpurun_out/r02
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lbfgs_control -s 100 -c 1 -o gpurun_out/r02/control_full -f python bench.py --config nst224 --steps 5 --no-cpu-baseline --no-gpu-library --no-e2e --no-features > /dev/null 2>&1
ncu -i gpurun_out/r02/control_full.ncu-rep --page source --csv > gpurun_out/r02/control_source.csv 2>/dev/null
ncu -i gpurun_out/r02/control_full.ncu-rep --page raw --csv > gpurun_out/r02/control_raw.csv 2>/dev/null
ls -la gpurun_out/r02/control_*
