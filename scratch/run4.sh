#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --batch 16 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 30 -c 4 -o gpurun_out/prof_conv $CMD > gpurun_out/ncu2.log 2>&1
echo "=== full bench"; timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -2
