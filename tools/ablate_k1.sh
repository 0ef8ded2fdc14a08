#!/bin/bash
# Dev tool: build k_stft_fb with parts removed (-DAAD_ABLATE=mask; results are wrong) into
# tools/_abl/libaad_<mask>.so, in parallel.  Run on the GPU box with tools/ablate_run.sh.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_abl
# each argument: "<name>:<extra nvcc flags>", e.g. "abl8:-DAAD_ABLATE=8" or "fbu2:-DAAD_FBU=2"
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 --expt-relaxed-constexpr -O3 -lineinfo -shared \
    -Xcompiler -fPIC $flags -DAAD_DEV_BUILD -o tools/_abl/libaad_$name.so audioanalysisdetector_b200/csrc/aad_api.cu &
done
wait
ls -la tools/_abl
