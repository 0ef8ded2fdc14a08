"""Minimal driver for ncu captures of the CNN-BiLSTM inference kernels (dev tool): config-4 batch, fixture weights."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import audioanalysisdetector_b200 as aad
dev = torch.device("cuda:0")
g = np.load(os.path.join(ROOT, "tests", "golden", "consumer.npz"))
weights = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w::")}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 25380
gen = torch.Generator(device=dev); gen.manual_seed(1)
feats = 10.0 * torch.randn((B, 13, 63), generator=gen, device=dev) - 20.0
eng = aad.DetectorEngine(weights, feature_dim=13, device=dev)
for _ in range(5):
    s = eng(feats)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    s = eng(feats)
e1.record(); torch.cuda.synchronize()
print("ok", float(s.mean()), "ms/call", e0.elapsed_time(e1) / 20)
