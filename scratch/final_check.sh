#!/bin/bash
mkdir -p gpurun_out/r02
python -m pytest tests/ -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py > gpurun_out/r02/bench_default_final.json 2> gpurun_out/r02/bench_default_final.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r02/bench_reference_final.json 2>/dev/null
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02/bench_default_final.json") if l.startswith("{")][0])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "feat", d["secondary"]["value"], d["secondary_5tap"]["value"], "frac", d["roofline"]["frac"], d["clocks"], "cpu", d["cpu_baseline"]["value"], "lib", d["gpu_library_baseline"])
r = json.loads([l for l in open("gpurun_out/r02/bench_reference_final.json") if l.startswith("{")][0])
print("reference", r["value"], r["cpu_baseline"])
PY
