"""Dev tool: CQCC throughput on the configs[3] corpus (25 380 two-second 16 kHz chunks) and the CQCC -> CNN-BiLSTM
pipeline the reference trains (cnn_bilstm_hybrid.py: feature_dim 19, 63 frames)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import audioanalysisdetector_b200 as aad
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
N = 25380
wav = (0.1 * torch.randn((N, 32000), generator=g, device=dev)).clamp_(-1, 1)
fe = aad.CqccFrontend(16000, device=dev)
def timed(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
ms = timed(lambda: fe(wav))
feats, nf, st = fe(wav)
gw = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "consumer.npz"))
weights = {k[3:]: torch.from_numpy(gw[k]).to(dev) for k in gw.files if k.startswith("w::")}
# the fixture's model has feature_dim 13; the conv runs over the 63 frames as channels, so any F works
engine = aad.DetectorEngine(weights, feature_dim=19, device=dev)
ms_model = timed(lambda: engine(feats))
hours = N * 2 / 3600
print(json.dumps({"workload": "CQCC-19 of 25 380 x 2 s @16 kHz (the reference's slowest extractor: 17.1 min on its 8-worker CPU run, "
                  "ASV_deep_learning.ipynb:210)", "cqcc_ms": ms, "audio_hours_per_s": hours / (ms * 1e-3),
                  "model_ms": ms_model, "status_nonzero": int(st.ne(0).sum()), "shape": list(feats.shape)}))
