#!/bin/bash
# Slab-streamed host evaluation (plt_eval_evaluate_points): parity tests + e2e A/B over the slab count.
set -u
mkdir -p gpurun_out
echo "== tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "evaluate_points or device_resident or smoke" 2>&1 | tail -5 | tee gpurun_out/slabs_tests.log
for k in 1 3 4 6 8; do
  echo "== PLT_SLABS=$k"
  PLT_SLABS=$k timeout 600 python bench.py --steps 10 --warmup 3 --no-fit --no-cpu-baseline --no-sampler 2> gpurun_out/slabs_$k.err \
    | python -c 'import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(json.dumps({"ms":d["ms_per_step"],"e2e":d["e2e"],"clocks":d["clocks"]}))' | tee gpurun_out/slabs_$k.json
done
