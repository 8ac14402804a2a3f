#!/bin/bash
# A/B of library builds on the GPU box: bench.py device-resident step for each tools/ab/lib_*.so
for lib in default tools/ab/lib_*.so; do
  if [ "$lib" = default ]; then unset PLT_B200_LIB; else export PLT_B200_LIB=$PWD/$lib; fi
  python bench.py --steps 5 --warmup 3 --no-fit --no-cpu-baseline --no-e2e --no-sampler 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phases_ms'].items() if k in ('m2l_hadamard','m2l_idft','l2l_l2p_leaf','p2p')})"
done
