#!/bin/bash
# Round-1 evidence refresh (tag r01_f): tests, smoke, bench (both arms), ncu launch list, full captures.
set -u
export NCU_TAG=r01_f
SKIP_TESTS=0 bash tools/gpu_round.sh
echo "== ncu full: p2p (matvec workload)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_p2p -s 2 -c 1 -f -o gpurun_out/prof_r01_f_p2p \
  python tools/dev_matvec.py 1000000 0 > gpurun_out/ncu_p2p.log 2>&1
tail -1 gpurun_out/ncu_p2p.log
echo "== hadamard leaf-level dram traffic (6 launches)"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_m2l_hadamard -s 18 -c 6 --csv --log-file gpurun_out/had_traffic.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-fit > /dev/null 2>&1
tail -3 gpurun_out/had_traffic.csv | cut -c1-300
