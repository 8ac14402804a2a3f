#!/bin/bash
# TMEM-assisted Hadamard kernel: parity tests + A/B against the shared-memory kernel on config #3.
set -u
mkdir -p gpurun_out
echo "== parity (TMEM kernel on)"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4 | tee gpurun_out/tmem_tests.log
for v in 0 1; do
  echo "== PLT_HAD_TMEM=$v"
  PLT_HAD_TMEM=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-fit --no-cpu-baseline --no-sampler --no-e2e 2> gpurun_out/tmem_$v.err \
    | python -c 'import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(json.dumps({"ms":d["ms_per_step"],"phases":d["phases_ms"],"roofline_frac":d["roofline"]["frac"]}))' | tee gpurun_out/tmem_$v.json
done
