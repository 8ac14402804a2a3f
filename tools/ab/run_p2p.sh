#!/bin/bash
for lib in default tools/ab/lib_p2p_*.so; do
  if [ "$lib" = default ]; then unset PLT_B200_LIB; else export PLT_B200_LIB=$PWD/$lib; fi
  echo "== $lib"; python tools/dev_matvec.py 1000000 0 2>&1 | grep -E "iter 4|vs direct" | sed -E "s/.*(p2p.: [0-9.]+).*/\1/"
done
