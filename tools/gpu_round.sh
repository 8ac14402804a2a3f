#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list, one full ncu capture.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 | tee gpurun_out/smoke.log
fi
if [ "${SKIP_BENCH:-0}" != "1" ]; then
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.log
fi
if [ "${SKIP_REF:-0}" != "1" ]; then
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -3 | tee gpurun_out/bench_ref.log
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
if [ "${SKIP_LAUNCHES:-0}" != "1" ]; then
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-fit > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c1-300
fi
echo "== ncu full: ${NCU_KERNEL:=k_m2l_hadamard}"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL} -s ${NCU_SKIP:-6} -c ${NCU_COUNT:-2} -f -o gpurun_out/prof_${NCU_TAG:-top} \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-fit > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
fi
ls -la gpurun_out
