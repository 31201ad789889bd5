#!/bin/bash
# refreshed round-2 evidence with the FINAL kernels (conv1_2 on conv_sweep64, pooled in registers): launch list of one NST
# tick at batch 64 (durations + DRAM bytes) and one --set full capture of conv_sweep64 (forward + dgrad), each after the
# same command exited 0 without ncu
mkdir -p gpurun_out/r02b
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-library --no-e2e --no-features --no-prefill"
$CMD > gpurun_out/r02b/plain_final.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02b/plain_final.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 150 -c 120 --csv \
  --log-file gpurun_out/r02b/ncu_launches_bench_b64_final.csv $CMD > gpurun_out/r02b/ncu_final.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_sweep64" -s 6 -c 2 -o gpurun_out/r02b/r02_conv_sweep64 -f $CMD > gpurun_out/r02b/ncu_full.log 2>&1
echo "set full rc=$?"
ncu -i gpurun_out/r02b/r02_conv_sweep64.ncu-rep --page raw --csv > gpurun_out/r02b/ncu_conv_sweep64_full_raw.csv 2>/dev/null
ls -la gpurun_out/r02b/ | tail -8
