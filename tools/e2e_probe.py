"""e2e pipeline probe (dev tool): Frontend.extract_host on the C2 workload for several chunk sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
dev = torch.device("cuda:0")
B, Ls = 4096, 64000
fe = Frontend(FrontendParams.mfcc(16000, n_mfcc=40, n_delta=2), dev)
t_max, c_out, _ = fe.query(B, Ls)
hw = torch.empty((B, Ls), dtype=torch.float32, pin_memory=True); hw.normal_(); hw.mul_(0.1).clamp_(-1, 1)
ho = torch.empty((B, c_out, t_max), dtype=torch.float32, pin_memory=True)
hlen = np.full(B, Ls, dtype=np.int32)
for chunk in (0, 32, 64, 128, 256, 512, 1024):
    fe.extract_host(hw.numpy(), hlen, out=ho.numpy(), chunk_utts=chunk)
    t0 = time.perf_counter()
    for _ in range(3):
        fe.extract_host(hw.numpy(), hlen, out=ho.numpy(), chunk_utts=chunk)
    dt = (time.perf_counter() - t0) / 3
    print(f"chunk_utts {chunk:5d}: {dt*1e3:7.2f} ms/step  {B*4/3600/dt:7.1f} audio-h/s  ({B*Ls*4/dt/1e9:.1f} GB/s H2D-equivalent)", flush=True)
