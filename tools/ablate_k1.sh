#!/bin/bash
# Dev tool: build variants of the library (e.g. k_stft_fb with parts removed, -DAAD_ABLATE=mask; results are
# wrong) into tools/_abl/libaad_<name>.so.  Run on the GPU box with tools/ablate_run.sh / ablate_run_full.sh.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_abl
# each argument: "<name>:<extra nvcc flags>", e.g. "abl8:-DAAD_ABLATE=8 -DAAD_DEV_BUILD" or "fbu2:-DAAD_FBU=2"
# (-DAAD_DEV_BUILD keeps only the n_fft 2048 float32 variant of k_stft_fb: much faster to build)
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  python -m audioanalysisdetector_b200.build --out tools/_abl/libaad_$name.so --flags "$flags" &
done
wait
ls -la tools/_abl
