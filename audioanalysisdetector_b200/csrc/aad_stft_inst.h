// k_stft_fb is instantiated per transform size in its own translation unit (aad_stft_inst.cu compiled with
// -DAAD_INST_L=4 / 8 / 16 / 32) so that the 30 variants build in parallel; aad_api.cu reaches them through these
// accessors (host-side kernel handles: launchable and usable with cudaFuncSetAttribute from any unit).
#pragma once
#include <stddef.h>

#include "aad_kernels.cuh"

namespace aad {
typedef void (*stft_kernel_t)(const StftArgs);
// mode: InMode; pre: pre-emphasis variant; pair: second filter bank in the same launch (n_fft 2048 / 512, no pre-emphasis)
stft_kernel_t pick_stft_L4(int mode, bool pre, bool pair);
stft_kernel_t pick_stft_L8(int mode, bool pre, bool pair);
stft_kernel_t pick_stft_L8_dense(int mode, bool pre);  // dense filter bank (gammatone): n_fft 512 only
stft_kernel_t pick_stft_L16(int mode, bool pre, bool pair);
stft_kernel_t pick_stft_L32(int mode, bool pre, bool pair);
}  // namespace aad
