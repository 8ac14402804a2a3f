#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_exp8.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"], d["phases_ms"], d.get("fit"))'
echo "== parity all"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
