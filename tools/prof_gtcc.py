"""Minimal driver for ncu captures of the dense filter-bank kernel (dev tool): GTCC-13 on 4096 x 4 s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
wav = (0.1 * torch.randn((4096, 64000), generator=g, device=dev)).clamp_(-1, 1)
fe = Frontend(FrontendParams.gtcc(16000), dev)
for _ in range(4):
    out, nf, st = fe(wav)
torch.cuda.synchronize()
print("ok", int(st.sum().item()))
