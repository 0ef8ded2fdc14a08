"""Dev tool: GTCC throughput on the configs[3] corpus (25 380 two-second 16 kHz chunks; the reference needs 9.1 min
for 28 408 chunks with 8 workers, ASV_deep_learning.ipynb:238)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
N = 25380
wav = (0.1 * torch.randn((N, 32000), generator=g, device=dev)).clamp_(-1, 1)
res = {}
for name, p in (("gtcc13 (40 gammatone filters, dense)", FrontendParams.gtcc(16000)), ("lfcc13 (24 linear filters, banded)", FrontendParams.lfcc(16000, quantize_i16=False))):
    fe = Frontend(p, dev)
    fe.set_profiling(True) if hasattr(fe, "set_profiling") else None
    for _ in range(3): fe(wav)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): out, nf, st = fe(wav)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res[name] = {"ms": ms, "audio_hours_per_s": N * 2 / 3600 / (ms * 1e-3), "shape": list(out.shape), "status_nonzero": int(st.ne(0).sum())}
print(json.dumps(res))
