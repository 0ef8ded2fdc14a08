"""Minimal driver for ncu captures of the C2 workload (dev tool): 5 calls of MFCC-40+d+dd on 4096 x 4 s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wav = (0.1 * torch.randn((B, 64000), generator=g, device=dev)).clamp_(-1, 1)
fe = Frontend(FrontendParams.mfcc(16000, n_mfcc=40, n_delta=2), dev)
for _ in range(5):
    out, nf, st = fe(wav)
torch.cuda.synchronize()
print("ok", int(st.sum().item()), float(out[0, 0, :4].sum().item()))
