"""Bring-up check on a real B200: CUDA path vs oracle on a few configs, prints error stats."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import audioanalysisdetector_b200 as aad
from audioanalysisdetector_b200 import _lib as L
from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
import oracle
from oracle import librosa_ref as LR, spafe_ref as SR, delta_ref as DR

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)

def noise(n): return np.clip(0.1 * rng.standard_normal(n), -1, 1).astype(np.float32)
def speech(n, sr=16000):
    t = np.arange(n) / sr
    return (0.3*np.sin(2*np.pi*140*t)*(1+0.5*np.sin(2*np.pi*3*t)) + 0.05*np.sin(2*np.pi*2300*t)
            + 0.003*rng.standard_normal(n)).astype(np.float32)

def run(params, clips, dtype=torch.float32):
    fe = Frontend(params, dev)
    B = len(clips); Lmax = (max(len(c) for c in clips) + 3)//4*4
    w = np.zeros((B, Lmax), dtype=np.float32 if dtype == torch.float32 else np.int16)
    for i, c in enumerate(clips): w[i, :len(c)] = c
    out, nf, st = fe(torch.from_numpy(w).to(dev), torch.tensor([len(c) for c in clips], dtype=torch.int32, device=dev))
    torch.cuda.synchronize()
    return out.cpu().numpy(), nf.cpu().numpy(), st.cpu().numpy(), fe

ok = True
def report(name, got, want, tol):
    global ok
    err = np.abs(got - want).max()
    flag = "OK " if err <= tol else "BAD"
    if err > tol: ok = False
    print(f"{flag} {name:58s} max|err|={err:.3e} (tol {tol:g}) ref range [{want.min():.3g},{want.max():.3g}]", flush=True)

# ---- log-mel, several n_fft
for (n_fft, hop, n_mels, sr) in [(2048, 512, 64, 16000), (512, 160, 80, 16000), (1024, 256, 40, 16000), (256, 64, 32, 8000), (2048, 480, 128, 48000)]:
    clips = [noise(32000), speech(47999, sr), noise(5000), noise(300)]
    p = FrontendParams.logmel(sr, n_mels=n_mels, n_fft=n_fft, hop_length=hop)
    out, nf, st, fe = run(p, clips)
    for i, c in enumerate(clips):
        want = LR.logmel_db(c, sr, n_mels=n_mels, n_fft=n_fft, hop_length=hop)
        assert nf[i] == want.shape[1] and st[i] == 0, (nf[i], want.shape, st[i])
        report(f"logmel n_fft={n_fft} hop={hop} mels={n_mels} sr={sr} clip{i}", out[i, :, :nf[i]], want, 1e-3)
    fbt = fe.table(L.TABLE_FILTERBANK); fbo = LR.mel_filterbank(sr, n_fft, n_mels)
    print("   fb table max diff", np.abs(fbt - fbo).max(), "win diff", np.abs(fe.table(L.TABLE_WINDOW) - __import__('scipy.signal').signal.get_window('hann', n_fft, fftbins=True)).max())

# ---- linear mel energies through ln + exp
p = FrontendParams.logmel(16000, n_mels=128).replace(log_type=L.LOG_LN, top_db=-1.0)
clips = [noise(64000), speech(64000)]
out, nf, st, _ = run(p, clips)
for i, c in enumerate(clips):
    want = LR.melspectrogram(c, 16000, n_mels=128)
    got = np.exp(out[i, :, :nf[i]].astype(np.float64))
    print(f"    linear mel clip{i}: peak-normalised err {np.abs(got-want).max()/want.max():.3e}, elementwise rel max {np.abs(got/np.maximum(want,1e-30)-1).max():.3e}")

# ---- MFCC + deltas (config 2 shape)
for n_mfcc, nd in [(13, 0), (40, 2), (20, 1)]:
    clips = [noise(64000), speech(64000), noise(32000), speech(5000)]
    p = FrontendParams.mfcc(16000, n_mfcc=n_mfcc, n_delta=nd)
    out, nf, st, _ = run(p, clips)
    for i, c in enumerate(clips):
        want = oracle.mfcc_with_deltas_ref(c, 16000, n_mfcc=n_mfcc, n_delta=nd)
        assert nf[i] == want.shape[1] and st[i] == 0
        report(f"mfcc{n_mfcc} n_delta={nd} clip{i}", out[i, :, :nf[i]], want, 1e-3)

# ---- LFCC reference defaults (float -> int16 quantise), TC layout
clips = [noise(32000), speech(40001), noise(16000), noise(500)]
p = FrontendParams.lfcc(16000, n_ceps=13)
out, nf, st, fe = run(p, clips)
for i, c in enumerate(clips):
    want = SR.lfcc(SR.quantize_int16(c), fs=16000, num_ceps=13)
    assert nf[i] == want.shape[0] and st[i] == 0, (nf[i], want.shape, st[i])
    report(f"lfcc13 default clip{i}", out[i, :nf[i], :], want, 1e-3)
print("   lin fb diff", np.abs(fe.table(L.TABLE_FILTERBANK) - SR.linear_filter_banks(24, 512, 16000)[0] / 512).max())

# ---- LFCC config 3: int16 in, win 320, 20 filt, 20 ceps, deltas, CT layout
clips16 = [SR.quantize_int16(noise(n)) for n in (16000, 77777, 128000)]
p = FrontendParams.lfcc(16000, n_ceps=20, nfilts=20, win_len=0.02, n_delta=2, layout=L.LAYOUT_CT)
out, nf, st, _ = run(p, clips16, dtype=torch.int16)
for i, c in enumerate(clips16):
    want = oracle.lfcc_with_deltas_ref(c, 16000)
    assert nf[i] == want.shape[1] and st[i] == 0
    report(f"lfcc20x3 int16 clip{i}", out[i, :, :nf[i]], want, 1e-3)

# ---- status codes
p = FrontendParams.mfcc(16000, n_mfcc=13, n_delta=2)
out, nf, st, _ = run(p, [noise(32000), noise(0), noise(512*7)])
print("status", st, "n_frames", nf)
p = FrontendParams.lfcc(16000)
out, nf, st, _ = run(p, [noise(399), noise(400)])
print("status", st, "n_frames", nf)

# ---- timing
p = FrontendParams.mfcc(16000, n_mfcc=40, n_delta=2)
fe = Frontend(p, dev)
B, Ls = 4096, 64000
g = torch.Generator(device=dev); g.manual_seed(1)
wav = (0.1 * torch.randn((B, Ls), generator=g, device=dev)).clamp_(-1, 1)
for _ in range(3): fe(wav)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): fe(wav)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"C2 MFCC40+d+dd 4096x4s: {ms:.3f} ms/step -> {B*Ls/16000/3600/(ms*1e-3):.1f} audio-h/s")
print("fp32 peak TFLOP/s:", aad.fp32_peak_tflops(0))
print("ALL OK" if ok else "SOME BAD")
sys.exit(0 if ok else 1)
