#!/bin/bash
# Dev tool: time the C2 workload with every tools/_abl/libaad_<mask>.so
cd "$(dirname "$0")/.."
for f in tools/_abl/libaad_*.so; do
  m=${f##*_}; m=${m%.so}
  echo -n "mask $m: "
  AAD_LIB_PATH=$PWD/$f python tools/gpu_time_c2.py --quick 2>&1 | head -1
done
