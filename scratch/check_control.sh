#!/bin/bash
mkdir -p gpurun_out/r02
python -m pytest tests/test_gpu_nst.py -m gpu -x -q -k "lbfgs or lean" 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lbfgs -s 440 -c 8 --csv --log-file gpurun_out/r02/ncu_lbfgs_nst224.csv python bench.py --config nst224 --steps 5 --no-cpu-baseline --no-gpu-library --no-e2e > /dev/null 2>&1
python - <<PY
import csv
for f in ["gpurun_out/r02/ncu_lbfgs_nst224.csv"]:
    rows = list(csv.reader(open(f)))
    hdr = None
    print(f)
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r: hdr = r
            continue
        d = dict(zip(hdr, r))
        print("  ", d["Kernel Name"][:40], d["Grid Size"], d["Metric Value"], d["Metric Unit"])
PY
python bench.py --config nst224 --steps 40 --no-cpu-baseline --no-gpu-library --no-e2e 2>/dev/null | cut -c1-140
