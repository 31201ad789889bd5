#!/bin/bash
mkdir -p gpurun_out/r02
CMD="python bench.py --config nst224 --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-library --no-e2e --no-prefill"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 120 -c 100 --csv --log-file gpurun_out/r02/ncu_launches_nst224.csv $CMD > /dev/null 2>&1
echo rc=$?
