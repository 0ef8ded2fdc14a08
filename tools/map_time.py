"""Dev tool: the reference notebook's whole extractor map (ASV_deep_learning.ipynb:152-160: cqcc, gtcc, mel-spect, mfcc,
lfcc) over the configs[3] corpus -- 25 380 two-second 16 kHz chunks resident in HBM as int16 PCM -- on one GPU.
The reference's own run of this cell: 17.1 + 9.1 + 5.2 + 5.5 + 5.3 = 42.2 min for 28 408 chunks with 8 workers
(ASV_deep_learning.ipynb:210,238,265,291,317; I/O and process overhead included)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audioanalysisdetector_b200 as aad
from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
N = 25380
wav = (3276.8 * torch.randn((N, 32000), generator=g, device=dev)).clamp_(-32768, 32767).to(torch.int16)
cq = aad.CqccFrontend(16000, device=dev)
gt = Frontend(FrontendParams.gtcc(16000), dev)
mf = Frontend(FrontendParams.mfcc(16000, n_mfcc=13), dev)
ml = Frontend(FrontendParams.logmel(16000, n_mels=64), dev)
lf = Frontend(FrontendParams.lfcc(16000), dev)
steps = {"cqcc": lambda: cq(wav), "gtcc": lambda: gt(wav), "mfcc + mel-spect (one STFT)": lambda: mf.extract_pair(ml, wav),
         "lfcc": lambda: lf(wav)}
def timed(fn, iters=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
res = {k: timed(f) for k, f in steps.items()}
res["whole map"] = timed(lambda: [f() for f in steps.values()])
hours = N * 2 / 3600
print(json.dumps({"workload": "5-feature extractor map, 25 380 x 2 s @16 kHz int16 PCM in HBM", "ms": res,
                  "audio_hours_per_s_whole_map": hours / (res["whole map"] * 1e-3),
                  "reference_minutes_28408_chunks_8_workers": 42.2}))
