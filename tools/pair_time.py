"""Dev tool: where the paired MFCC-13 + log-mel-64 call (configs[3]) spends its time, against the two single calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
wav = (0.1 * torch.randn((25380, 32000), generator=g, device=dev)).clamp_(-1, 1)
mf = Frontend(FrontendParams.mfcc(16000, n_mfcc=13), dev)
ml = Frontend(FrontendParams.logmel(16000, n_mels=64), dev)
def timed(fn, n=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("mfcc13 alone   %.3f ms" % timed(lambda: mf(wav)))
print("logmel64 alone %.3f ms" % timed(lambda: ml(wav)))
print("pair           %.3f ms" % timed(lambda: mf.extract_pair(ml, wav)))
for fe, name in ((mf, "mfcc13"), (ml, "logmel64")):
    fe.set_profiling(True)
    fe(wav); torch.cuda.synchronize()
    print(name, "kernel times (prep, stft_fb, epilogue):", fe.kernel_times_ms())
    fe.set_profiling(False)
