#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== matvec"; timeout 600 python tools/dev_matvec.py 1000000 0 2>&1 | tail -4 | tee gpurun_out/matvec.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["phases_ms"])'
echo "== parity (p2p-related)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
echo "== c4 inf"; timeout 600 python tools/dev_c4.py 500000 2>&1 | tail -10 | tee gpurun_out/c4.log
