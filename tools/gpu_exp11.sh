#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in "4 0" "4 1" "6 0" "6 1"; do
  set -- $cfg
  echo "== G=$1 PF=$2"
  PLT_HAD_G=$1 PLT_HAD_PF=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fit 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["phases_ms"]["m2l_hadamard"])'
  PLT_HAD_G=$1 PLT_HAD_PF=$2 timeout 300 python tools/dev_matvec.py 1000000 0 2>&1 | tail -4 | head -1 | cut -c1-60
done
