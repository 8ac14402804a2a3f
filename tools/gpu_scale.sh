#!/bin/bash
# Strong scaling of config #3 on one node: bench.py at N = 2, 4, 8 the way the driver launches it.
set -u
mkdir -p gpurun_out
TAG=${TAG:-r02}
for N in ${NS:-2 4 8}; do
  echo "== N=$N"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
  tail -c 400 gpurun_out/${TAG}_bench_${N}gpu.err | grep -v "^$" | tail -3
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${TAG}_bench_${N}gpu.json") if l.startswith("{")][-1])
    m = d["multi_gpu"]
    print("N=$N value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 3),
          "rel diff vs 1 GPU", m["n_gpu_vs_1_gpu_max_rel_diff"], "per-rank ms", m["step_ms_per_rank"], "upward", round(m["upward_ms_max"], 3),
          "allgather", round(m["allgather_ms_max"], 3), "replicas", round(m["weak_replicas"]["value"], 1))
    print({k: round(v, 3) for k, v in d["phases_ms"].items()})
except Exception as e:
    print("N=$N failed:", e)
PY
done
