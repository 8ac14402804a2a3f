#!/bin/bash
# Round-1 final evidence (tag r01_g): smoke, bench (both arms), ncu launch list, Hadamard full capture + traffic.
set -u
mkdir -p gpurun_out
export NCU_TAG=r01_g
SKIP_TESTS=1 bash tools/gpu_round.sh
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== hadamard dram traffic (6 launches)"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_m2l_hadamard -s 18 -c 6 --csv --log-file gpurun_out/had_traffic.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-fit > /dev/null 2>&1
tail -1 gpurun_out/had_traffic.csv | cut -c1-200
