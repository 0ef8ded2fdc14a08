#!/bin/bash
# A/B of the TMA-staged sample path (dev library tools/_abl/libaad_tma.so, built with -DAAD_TMA_STAGE=1)
cd "$(dirname "$0")/.."
echo "== shipped"; timeout 300 python tools/gpu_time_c2.py 2>&1 | head -3
echo "== TMA staged: parity"; timeout 600 env AAD_LIB_PATH=$PWD/tools/_abl/libaad_tma.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mfcc or logmel or config2 or torchaudio" 2>&1 | tail -2
echo "== TMA staged: timing"; timeout 300 env AAD_LIB_PATH=$PWD/tools/_abl/libaad_tma.so python tools/gpu_time_c2.py 2>&1 | head -3
