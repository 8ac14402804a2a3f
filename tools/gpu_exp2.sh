#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench (balanced tiled)"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_exp2.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"], d["phases_ms"])'
echo "== parity"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== matvec"; timeout 600 python tools/dev_matvec.py 1000000 0 2>&1 | tail -4 | tee gpurun_out/matvec.log
