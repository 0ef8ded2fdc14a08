"""Dev tool: k_stft_ws (warp-specialised, tensor-core filter bank) against k_stft_fb on the same inputs.
AAD_K1 is read at plan creation, so both kernels can be driven from one process."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audioanalysisdetector_b200.frontend import Frontend, FrontendParams

dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(5)
def both(params, wav, lens=None):
    outs = []
    for k in ("fb", "ws"):
        os.environ["AAD_K1"] = k
        fe = Frontend(params, dev)
        out, nf, st = fe(wav, lens)
        torch.cuda.synchronize()
        outs.append((out.clone(), nf.clone(), st.clone()))
    os.environ.pop("AAD_K1")
    (a, nfa, sa), (b, nfb, sb) = outs
    assert torch.equal(nfa, nfb) and torch.equal(sa, sb), "n_frames / status differ"
    return float((a - b).abs().max()), float(a.abs().max())

wav = (0.1 * torch.randn((300, 64000), generator=g, device=dev)).clamp_(-1, 1)
lens = torch.randint(3000, 64001, (300,), generator=torch.Generator().manual_seed(3)).to(torch.int32).to(dev)
for name, p in (("mfcc40+d+dd", FrontendParams.mfcc(16000, n_mfcc=40, n_delta=2)),
                ("logmel64", FrontendParams.logmel(16000)),
                ("logmel128", FrontendParams.logmel(16000, n_mels=128))):
    print(name, "fixed: max|ws-fb| %.3g (max|x| %.3g)" % both(p, wav), " ragged: %.3g (%.3g)" % both(p, wav, lens), flush=True)
w16 = (wav * 32767).to(torch.int16)
print("logmel64 int16: %.3g (%.3g)" % both(FrontendParams.logmel(16000), w16, lens))
