#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== c4 inf"; timeout 600 python tools/dev_c4.py 500000 2>&1 | tail -6 | cut -c1-330 | tee gpurun_out/c4.log
echo "== matvec"; timeout 600 python tools/dev_matvec.py 1000000 0 2>&1 | tail -4 | head -1 | tee gpurun_out/matvec.log
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
