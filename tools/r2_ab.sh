#!/bin/bash
# A/B of the shipped library: default against the dev switches given as arguments (e.g. AAD_FB_SMEM=1)
cd "$(dirname "$0")/.."
echo "== default"; timeout 300 python tools/gpu_time_c2.py 2>&1 | head -9
for sw in "$@"; do echo "== $sw"; timeout 300 env $sw python tools/gpu_time_c2.py 2>&1 | head -9; done
