#!/bin/bash
# A/B of library builds on the GPU box: bench.py device-resident step for each tools/ab/lib_*.so
# (plus the default build with the environment switches listed in AB_ENVS, e.g. "PLT_DEBUG_NO_BLK=1")
run() {
  python bench.py --steps 5 --warmup 3 --no-fit --no-cpu-baseline --no-e2e --no-sampler 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), {k: round(v,3) for k, v in d['phases_ms'].items() if v >= 0.2})"
}
for lib in default tools/ab/lib_*.so; do
  [ -e "$lib" ] || [ "$lib" = default ] || continue
  if [ "$lib" = default ]; then unset PLT_B200_LIB; else export PLT_B200_LIB=$PWD/$lib; fi
  run $lib
done
unset PLT_B200_LIB
for e in ${AB_ENVS:-}; do env $e bash -c "$(declare -f run); run $e"; done
