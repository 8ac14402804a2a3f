#!/bin/bash
# Quick regression check on one B200: parity suite, bench phases, order-12 matvec.
set -u
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fit 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["phases_ms"], d["phase_rooflines"])'
echo "== matvec"; timeout 600 python tools/dev_matvec.py 1000000 0 2>&1 | tail -4 | head -1
