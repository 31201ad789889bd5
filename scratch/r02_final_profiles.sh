#!/bin/bash
# final round-2 evidence: launch list of one NST tick at batch 64 (durations + DRAM bytes), after the same command exited 0 without ncu
mkdir -p gpurun_out/r02
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-library --no-e2e --no-features --no-prefill"
$CMD > gpurun_out/r02/plain_final.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02/plain_final.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 150 -c 120 --csv \
  --log-file gpurun_out/r02/ncu_launches_bench_b64_final.csv $CMD > gpurun_out/r02/ncu_final.log 2>&1
echo "launch list rc=$?"
# one --set full capture of the routing-byte pool backward and of conv_c64 forward without its full-resolution store
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"maxpool_bwd_idx|conv_c64" -s 12 -c 5 -o gpurun_out/r02/r02_pool_idx_c64 -f $CMD > gpurun_out/r02/ncu_full.log 2>&1
echo "set full rc=$?"
ncu -i gpurun_out/r02/r02_pool_idx_c64.ncu-rep --page raw --csv > gpurun_out/r02/ncu_pool_idx_c64_full_raw.csv 2>/dev/null
ls -la gpurun_out/r02/ | tail -8
