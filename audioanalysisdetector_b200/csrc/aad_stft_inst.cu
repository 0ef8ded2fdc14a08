// One translation unit per transform size: -DAAD_INST_L=4 | 8 | 16 | 32 (see aad_stft_inst.h, build.py).
#define AAD_STFT_ONLY
#include "aad_stft_inst.h"

#ifndef AAD_INST_L
#error "compile with -DAAD_INST_L=4|8|16|32"
#endif

namespace aad {

#define AAD_CAT2(a, b) a##b
#define AAD_CAT(a, b) AAD_CAT2(a, b)

stft_kernel_t AAD_CAT(pick_stft_L, AAD_INST_L)(int mode, bool pre, bool pair) {
  constexpr int L = AAD_INST_L, TILE = stft_tile(AAD_INST_L);
#if AAD_ABLATE || defined(AAD_DEV_BUILD)
  // dev builds: one variant
#if AAD_INST_L == 32
  return mode == 0 && !pre && !pair ? k_stft_fb<32, IN_F32, false, 32> : nullptr;
#else
  return nullptr;
#endif
#else
  if (pair) {  // two filter banks on one STFT: mel-type plans (no pre-emphasis), n_fft 2048 and 512
#if AAD_INST_L == 32 || AAD_INST_L == 8
    if (pre) return nullptr;
    switch (mode) {
      case 0: return k_stft_fb<L, IN_F32, false, TILE, true>;
      case 1: return k_stft_fb<L, IN_F32_Q16, false, TILE, true>;
      default: return k_stft_fb<L, IN_I16, false, TILE, true>;
    }
#else
    return nullptr;
#endif
  }
  switch (mode * 2 + (pre ? 1 : 0)) {
    case 0: return k_stft_fb<L, IN_F32, false, TILE>;
    case 1: return k_stft_fb<L, IN_F32, true, TILE>;
    case 2: return k_stft_fb<L, IN_F32_Q16, false, TILE>;
    case 3: return k_stft_fb<L, IN_F32_Q16, true, TILE>;
    case 4: return k_stft_fb<L, IN_I16, false, TILE>;
    default: return k_stft_fb<L, IN_I16, true, TILE>;
  }
#endif
}

#if AAD_INST_L == 8
stft_kernel_t pick_stft_L8_dense(int mode, bool pre) {
#if AAD_ABLATE || defined(AAD_DEV_BUILD)
  return nullptr;
#else
  constexpr int TILE = stft_tile(8, true);
  switch (mode * 2 + (pre ? 1 : 0)) {
    case 0: return k_stft_fb<8, IN_F32, false, TILE, false, 1>;
    case 1: return k_stft_fb<8, IN_F32, true, TILE, false, 1>;
    case 2: return k_stft_fb<8, IN_F32_Q16, false, TILE, false, 1>;
    case 3: return k_stft_fb<8, IN_F32_Q16, true, TILE, false, 1>;
    case 4: return k_stft_fb<8, IN_I16, false, TILE, false, 1>;
    default: return k_stft_fb<8, IN_I16, true, TILE, false, 1>;
  }
#endif
}
#endif

}  // namespace aad

#if defined(AAD_PHASE_TIMING) && AAD_INST_L == 32
// dev only: read and reset the phase counters of k_stft_fb (they live in this unit)
extern "C" int aad_dev_phase_cycles(unsigned long long* out4) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -3;
  if (cudaMemcpyFromSymbol(out4, aad::g_phase_cycles, 4 * sizeof(unsigned long long)) != cudaSuccess) return -3;
  unsigned long long z[4] = {0, 0, 0, 0};
  return cudaMemcpyToSymbol(aad::g_phase_cycles, z, sizeof(z)) == cudaSuccess ? 0 : -3;
}
#endif
