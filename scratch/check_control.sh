#!/bin/bash
mkdir -p gpurun_out/r02
python -m pytest tests/test_gpu_nst.py -m gpu -x -q -k "lbfgs or lean" 2>&1 | tail -2
for cfg in nst224 nst640; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lbfgs -s 400 -c 12 --csv --log-file gpurun_out/r02/ncu_lbfgs_$cfg.csv python bench.py --config $cfg --steps 5 --no-cpu-baseline --no-gpu-library --no-e2e --no-features > /dev/null 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/r02/ncu_lbfgs_$cfg.csv")))
hdr = None
print("$cfg")
for r in rows:
    if hdr is None:
        if "Kernel Name" in r: hdr = r
        continue
    d = dict(zip(hdr, r))
    print("  ", d["Kernel Name"][:34], d["Grid Size"], d["Metric Value"], d["Metric Unit"])
PY
done
for i in 1 2; do python bench.py --config nst224 --steps 40 --no-cpu-baseline --no-gpu-library --no-e2e 2>/dev/null | cut -c1-110; done
python bench.py --steps 20 --no-cpu-baseline --no-gpu-library --no-e2e --no-features 2>/dev/null | cut -c1-110
