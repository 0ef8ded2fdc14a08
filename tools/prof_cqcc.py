"""Minimal driver for ncu captures of the CQCC kernels (dev tool): 2 calls on 25 380 two-second chunks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audioanalysisdetector_b200 as aad
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 25380
wav = (0.1 * torch.randn((N, 32000), generator=g, device=dev)).clamp_(-1, 1)
fe = aad.CqccFrontend(16000, device=dev)
for _ in range(2):
    out, nf, st = fe(wav)
torch.cuda.synchronize()
print("ok", int(st.sum()))
