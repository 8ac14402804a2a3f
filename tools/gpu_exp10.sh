#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 0 1; do
  echo "== tmem=$v"
  if [ $v = 1 ]; then export PLT_HAD_TMEM=1; else unset PLT_HAD_TMEM; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fit 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["phases_ms"])'
  timeout 300 python tools/dev_matvec.py 1000000 0 2>&1 | tail -4 | head -1
done
echo "== parity tmem"; PLT_HAD_TMEM=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
