#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== krylov/operator tests"; timeout 900 python -m pytest tests/test_gpu_krylov.py -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/pytest_krylov.log
echo "== fgmres 1M, 1 rank"; timeout 600 python tools/dev_fgmres.py 1000000 20 0 2>&1 | tail -5 | tee gpurun_out/fgmres_1.log
