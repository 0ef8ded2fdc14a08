import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
wav = (0.1 * torch.randn((4096, 64000), generator=g, device=dev)).clamp_(-1, 1)
for tile in ("16", "32"):
    for stag in ("0", "4000", "8000", "16000"):
        os.environ["AAD_TILE"] = tile; os.environ["AAD_STAGGER"] = stag
        fe = Frontend(FrontendParams.mfcc(16000, n_mfcc=40, n_delta=2), dev)
        for _ in range(3): fe(wav)
        fe.set_profiling(True)
        ts = []
        for _ in range(5):
            fe(wav); torch.cuda.synchronize(); ts.append(fe.kernel_times_ms()["stft_fb"])
        print(f"tile {tile} stagger {stag:6s}: stft_fb {min(ts):.3f} ms (median {sorted(ts)[2]:.3f})", flush=True)
