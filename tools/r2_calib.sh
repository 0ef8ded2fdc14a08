#!/bin/bash
# Round-2 calibration of k_stft_ws against k_stft_fb (C2 workload, dev builds from tools/ablate_k1.sh)
cd "$(dirname "$0")/.."
run() { echo -n "$1: "; shift; timeout 120 env "$@" python tools/gpu_time_c2.py --quick 2>&1 | head -3 | tr '\n' '|'; echo; }
run "shipped ws      " AAD_LIB_PATH=$PWD/audioanalysisdetector_b200/libaad_b200.so
run "shipped fb      " AAD_LIB_PATH=$PWD/audioanalysisdetector_b200/libaad_b200.so AAD_K1=fb
for f in tools/_abl/libaad_*.so; do
  m=${f##*libaad_}; m=${m%.so}
  case $m in
    abl*) run "$m (fb)" AAD_LIB_PATH=$PWD/$f AAD_K1=fb ;;
    *) run "$m" AAD_LIB_PATH=$PWD/$f ;;
  esac
done
