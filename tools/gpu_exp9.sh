#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fit 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases_ms"])'
echo "== hadamard traffic"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_m2l_hadamard -s 18 -c 6 --csv --log-file gpurun_out/had_traffic.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-fit > /dev/null 2>&1
grep hadamard gpurun_out/had_traffic.csv | awk -F'","' '{print $13, $15}' | tr -d '"'
echo "== matvec"; timeout 600 python tools/dev_matvec.py 1000000 0 2>&1 | tail -4 | head -1
