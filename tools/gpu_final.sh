#!/bin/bash
# Final-state evidence of a tag: full GPU test suite, smoke, both bench arms.  usage: TAG=r02_j bash tools/gpu_final.sh
set -u
TAG=${TAG:-r02_j}
mkdir -p gpurun_out
echo "== tests"
( time timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/${TAG}_tests.log
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
echo "== bench"
( time timeout 900 python bench.py 2> gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench.json ) 2>&1 | tail -3
cut -c1-600 gpurun_out/${TAG}_bench.json
echo "== bench reference arm"
( time timeout 600 python bench.py --impl reference 2> gpurun_out/${TAG}_bench_reference.err | tail -1 > gpurun_out/${TAG}_bench_reference.json ) 2>&1 | tail -3
cut -c1-400 gpurun_out/${TAG}_bench_reference.json
[ -n "${SKIP_NCU:-}" ] && exit 0
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-fit --no-sampler > gpurun_out/${TAG}_ncu_launches.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_launches.log | cut -c1-120
echo "== ncu full: k_m2l_hadamard_tiled (leaf-level launch)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_m2l_hadamard_tiled -s 5 -c 1 -f -o gpurun_out/${TAG}_prof_hadamard \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-fit --no-sampler > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_full.log | cut -c1-120
