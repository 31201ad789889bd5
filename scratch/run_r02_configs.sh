#!/bin/bash
# round-2 bench lines of every --config (one GPU), saved under gpurun_out/r02/ and copied to profiles/ by hand
mkdir -p gpurun_out/r02
F="--no-cpu-baseline --no-gpu-library"
python bench.py > gpurun_out/r02/bench_default.json 2> gpurun_out/r02/bench_default.err
python bench.py --steps 300 --no-e2e --no-features $F > gpurun_out/r02/bench_300steps.json 2>/dev/null
python bench.py --config nst640_5tap --steps 30 $F > gpurun_out/r02/bench_nst640_5tap.json 2> gpurun_out/r02/5tap.err
python bench.py --config masked_gram --steps 30 $F > gpurun_out/r02/bench_masked_gram.json 2> gpurun_out/r02/masked.err
python bench.py --config nst224 --steps 60 $F > gpurun_out/r02/bench_nst224.json 2> gpurun_out/r02/nst224.err
python bench.py --config nst1024 --steps 30 --e2e-evals 60 $F > gpurun_out/r02/bench_nst1024.json 2> gpurun_out/r02/nst1024.err
python bench.py --config feat4 > gpurun_out/r02/bench_feat4.json 2> gpurun_out/r02/feat4.err
python bench.py --config feat5 > gpurun_out/r02/bench_feat5.json 2> gpurun_out/r02/feat5.err
python bench.py --config frames2020 > gpurun_out/r02/bench_frames2020.json 2> gpurun_out/r02/frames.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02/bench_reference.json 2>/dev/null
echo configs-done
