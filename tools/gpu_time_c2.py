"""Time configs on the GPU with per-kernel breakdown (dev tool). Usage: python tools/gpu_time_c2.py [--quick]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
from audioanalysisdetector_b200 import _lib as L
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)

def timeit(name, params, wav, lengths=None, iters=10):
    fe = Frontend(params, dev)
    for _ in range(3): fe(wav, lengths)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fe(wav, lengths)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fe.set_profiling(True)
    fe(wav, lengths); torch.cuda.synchronize()
    kt = fe.kernel_times_ms()
    fe.set_profiling(False)
    secs = float(wav.shape[0] * wav.shape[1]) if lengths is None else float(lengths.sum().item())
    sr = params.sample_rate
    print(f"{name:34s} {ms:8.3f} ms/step  {secs/sr/3600/(ms*1e-3):9.1f} audio-h/s   stft_fb {kt['stft_fb']:.3f}  epi {kt['epilogue']:.3f}  prep {kt['prepare']:.3f}", flush=True)
    return ms

wav = (0.1 * torch.randn((4096, 64000), generator=g, device=dev)).clamp_(-1, 1)
timeit("C2 mfcc40+d+dd 2048/512", FrontendParams.mfcc(16000, n_mfcc=40, n_delta=2), wav)
timeit("   mfcc13 2048/512", FrontendParams.mfcc(16000, n_mfcc=13), wav)
timeit("   logmel64 2048/512", FrontendParams.logmel(16000), wav)
if "--quick" in sys.argv:
    sys.exit(0)
timeit("C2b mfcc40+d+dd 512/160/80", FrontendParams.mfcc(16000, n_mfcc=40, n_mels=80, n_fft=512, hop_length=160, n_delta=2), wav)
timeit("C1 logmel80 512/160 (B=4096)", FrontendParams.logmel(16000, n_mels=80, n_fft=512, hop_length=160), wav)
timeit("C1 logmel80 512/160 (B=64)", FrontendParams.logmel(16000, n_mels=80, n_fft=512, hop_length=160), wav[:64].contiguous(), iters=50)
# the same as a CUDA graph replay on static buffers: what the three launches cost without per-call allocations
fe64 = Frontend(FrontendParams.logmel(16000, n_mels=80, n_fft=512, hop_length=160), dev)
g64 = fe64.graph(wav[:64].contiguous())
for _ in range(5): g64.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): g64.replay()
e1.record(); torch.cuda.synchronize()
ref64 = fe64(wav[:64].contiguous())[0]
print(f"C1 logmel80 512/160 (B=64) graph   {e0.elapsed_time(e1) / 200:8.4f} ms/step  (replay equals the direct call: {bool(torch.equal(ref64, g64.out))})", flush=True)
# C3: variable-length int16 LFCC 20x3
lens = torch.randint(16000, 128001, (4096,), generator=torch.Generator().manual_seed(3)).to(torch.int32)
w16 = (wav[:, :1].new_empty((4096, 128000)).normal_(generator=g) * 3000).clamp_(-32767, 32767).to(torch.int16)
timeit("C3 lfcc20x3 int16 var 1-8s", FrontendParams.lfcc(16000, n_ceps=20, nfilts=20, win_len=0.02, n_delta=2, layout=L.LAYOUT_CT), w16, lens.to(dev))
del w16
timeit("   lfcc13 default f32 2s (B=4096)", FrontendParams.lfcc(16000), wav[:, :32000].contiguous())
# C5: long form 48k
del wav
for B in (1, 8):
    wl = (0.1 * torch.randn((B, 28800000), generator=g, device=dev)).clamp_(-1, 1)
    timeit(f"C5 logmel128 48k 10min B={B}", FrontendParams.logmel(48000, n_mels=128, n_fft=2048, hop_length=480), wl, iters=5)
    del wl
