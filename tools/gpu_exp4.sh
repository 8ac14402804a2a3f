#!/bin/bash
# ncu captures of the secondary kernels: P2P (matvec workload) and the fused leaf pass (bench workload)
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_p2p -s 2 -c 1 -f -o gpurun_out/prof_p2p \
  python tools/dev_matvec.py 1000000 0 > gpurun_out/ncu_p2p.log 2>&1
tail -2 gpurun_out/ncu_p2p.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_l2l_l2p_leaf -s 2 -c 1 -f -o gpurun_out/prof_leaf \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-fit > gpurun_out/ncu_leaf.log 2>&1
tail -2 gpurun_out/ncu_leaf.log
ls -la gpurun_out
