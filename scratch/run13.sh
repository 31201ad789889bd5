#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench default"; timeout 1200 python bench.py 2>&1 | tail -1 > gpurun_out/bench_default.json; cut -c1-3000 gpurun_out/bench_default.json
echo "=== bench reference"; timeout 600 python bench.py --impl reference 2>&1 | tail -1 | cut -c1-600
