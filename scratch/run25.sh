#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline --no-e2e --no-features"
$CMD > gpurun_out/plain_b64.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_b64_v2.csv $CMD > gpurun_out/ncu_b64.log 2>&1
tail -1 gpurun_out/ncu_b64.log | cut -c1-200
python scratch/kernel_table.py 32 > gpurun_out/kernel_table_v2.txt 2>&1; tail -3 gpurun_out/kernel_table_v2.txt
