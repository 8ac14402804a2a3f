#!/bin/bash
# Experiment: Hadamard variants A/B on the bench workload + matvec breakdown + parity.
set -u
mkdir -p gpurun_out
for v in 0 16 12 8; do
  echo "== variant $v"
  PLT_HAD_VARIANT=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_v$v.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["phases_ms"])'
done
echo "== parity (default variant)"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== matvec"; timeout 600 python tools/dev_matvec.py 1000000 0 2>&1 | tail -8 | tee gpurun_out/matvec.log
