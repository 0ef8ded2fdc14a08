#!/bin/bash
# Dev tool: time every config shape (tools/gpu_time_c2.py) with every tools/_abl/libaad_<name>.so
cd "$(dirname "$0")/.."
for f in tools/_abl/libaad_*.so; do
  m=${f##*_}; m=${m%.so}
  echo "== $m"
  AAD_LIB_PATH=$PWD/$f python tools/gpu_time_c2.py 2>&1 | head -9
done
