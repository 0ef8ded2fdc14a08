import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import audioanalysisdetector_b200 as aad
from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
from audioanalysisdetector_b200 import _lib as L
from oracle import spafe_ref as sp
dev = torch.device("cuda:0")
rng = np.random.default_rng(3)
for sr in (16000, 22050, 48000):
    B = 5
    lens = [2 * sr, int(1.3 * sr), sr // 2, 401 * sr // 16000, 3 * sr]
    wav = np.zeros((B, max(lens)), np.float32)
    for i, n in enumerate(lens):
        t = np.arange(n) / sr
        wav[i, :n] = (0.3 * np.sin(2 * np.pi * (200 + 300 * i) * t) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    for spectrum in (L.SPEC_POWER, L.SPEC_MAGNITUDE):
        fe = Frontend(FrontendParams.gtcc(sr, spectrum=spectrum), dev)
        out, nf, st = fe(torch.from_numpy(wav).to(dev), torch.tensor(lens, dtype=torch.int32, device=dev))
        out = out.cpu().numpy(); nf = nf.cpu().numpy()
        worst = 0
        for i, n in enumerate(lens):
            ref = sp.gfcc(wav[i, :n], sr, 13, nfilts=40, spectrum="power" if spectrum == 0 else "magnitude")
            got = out[i, :nf[i]]
            assert got.shape == ref.shape, (got.shape, ref.shape)
            worst = max(worst, np.abs(got - ref).max() / max(1.0, np.abs(ref).max()))
        print(sr, spectrum, "max rel-to-peak err", worst, st.cpu().numpy())
y = wav[0, :lens[0]]
g = aad.extract_gtcc((y, sr))
print(g.shape, g.dtype, np.abs(g - sp.gfcc(y, sr, 13, nfilts=40)).max())
