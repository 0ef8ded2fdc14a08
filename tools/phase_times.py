"""Dev tool: per-phase warp-cycles of k_stft_fb from a -DAAD_PHASE_TIMING build (AAD_LIB_PATH)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
from audioanalysisdetector_b200 import _lib as L
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
wav = (0.1 * torch.randn((4096, 64000), generator=g, device=dev)).clamp_(-1, 1)
fe = Frontend(FrontendParams.mfcc(16000, n_mfcc=40, n_delta=2), dev)
lib = L.load()
buf = (C.c_ulonglong * 4)()
for _ in range(3): fe(wav)
lib.aad_dev_phase_cycles(buf)
n = 5
for _ in range(n): fe(wav)
lib.aad_dev_phase_cycles(buf)
tot = sum(buf)
names = ["FFT phase", "barrier 1", "filterbank phase", "barrier 2"]
warps = 148 * 16
for nm, v in zip(names, buf):
    print(f"{nm:18s} {v / n / warps / 1e3:9.1f} kcycles per warp per launch   {100 * v / tot:5.1f} %")
print(f"tiles per CTA: {516096 / 32 / 148:.1f};  total {tot / n / warps / 1e3:.1f} kcycles per warp")
