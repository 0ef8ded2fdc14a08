"""Oracle: numpy/scipy restatement of spafe 0.3.x `spafe.features.lfcc.lfcc`.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED (spafe is
not installed in this image; the reference has no fixtures).

Reference call sites:
  ASV_dl_func.py:434-435   y_int16 = (y * 32767).astype(np.int16)
                           lfcc(sig=y_int16, fs=sr, num_ceps=n_ceps)
  ASV_func.py:68-69 ; train_fun.py:84-85   (same)

spafe 0.3.x chain restated (spafe/features/lfcc.py, spafe/utils/preprocessing.py,
spafe/fbanks/linear_fbanks.py), all float64:
  pre_emphasis (0.97, first sample kept) -> framing (win 25 ms / hop 10 ms, no
  padding, tail dropped) -> np.hamming(frame_len) -> |np.fft.fft(frames, nfft)|
  -> (1/nfft) * |X|^2 -> @ linear filterbank^T -> zero_handling (0 -> eps)
  -> np.log -> scipy.fftpack.dct(type=2, norm='ortho', axis=1)[:, :num_ceps].

Least-certain step: the filterbank construction.  `linear_filter_banks` restates
spafe >= 0.2 / 0.3.x (`requirements.txt:5` pins ~= 0.3.3): triangles on the
CONTINUOUS bin frequencies `np.linspace(low_freq, high_freq, nfft//2 + 1)` with
`(freqs >= lower) == (freqs <= center)` masks, returning the tuple
`(fbank, center_freqs)` that `lfcc` unpacks as `lin_fbanks_mat, _ = ...`.  The
python_speech_features-style construction on integer FFT bins
(`floor((nfft + 1) * hz / fs)`) belongs to spafe 0.1.x and is kept as
`linear_filter_banks_intbin`.  Evidence: two independent recollections of the
0.3.x source (round-1 builder's notes listed it as "a continuous-frequency variant
exists in other versions", the round-1 review recalled the tuple return and the
mask expression of 0.3.x); the wheel cannot be opened here (`pip download
spafe==0.3.3` has no index).  Everything downstream takes the matrix as data, so
parity of the CUDA path with this oracle holds for either (both are exercised in
tests), and a deployment that has spafe installed can pass spafe's own matrix as
a custom bank.
"""
from __future__ import annotations

import numpy as np
import scipy.fft

EPS = np.finfo(float).eps


def quantize_int16(y):
    """ASV_dl_func.py:434 -- float32 multiply, then C-style truncation toward zero.

    numpy's float32->int16 `astype` goes through a 32-bit integer and keeps the low
    16 bits (wrap-around) for out-of-range values on x86; in-range values are
    truncated toward zero.
    """
    y = np.asarray(y, dtype=np.float32)
    scaled = y * np.float32(32767)
    with np.errstate(invalid="ignore"):
        return scaled.astype(np.int32).astype(np.int16)


def pre_emphasis(sig, pre_emph_coeff=0.97):
    """spafe.utils.preprocessing.pre_emphasis: [s0, s[1:] - c*s[:-1]] (float64)."""
    sig = np.asarray(sig)
    return np.append(sig[0], sig[1:] - pre_emph_coeff * sig[:-1])


def frame_params(fs, win_len=0.025, win_hop=0.01):
    return int(win_len * fs), int(win_hop * fs)


def n_frames_uncentered(length, frame_length, frame_step):
    if length < frame_length:
        return 0
    return (int(length) - frame_length) // frame_step + 1


def framing(sig, fs=16000, win_len=0.025, win_hop=0.01):
    """spafe.utils.preprocessing.framing via stride_trick: no padding, tail dropped."""
    if win_len < win_hop:
        raise ValueError("win_len must be >= win_hop")
    frame_length, frame_step = frame_params(fs, win_len, win_hop)
    sig = np.asarray(sig)
    nrows = n_frames_uncentered(sig.size, frame_length, frame_step)
    if nrows <= 0:
        raise ValueError("signal shorter than one frame")
    idx = np.arange(frame_length)[None, :] + frame_step * np.arange(nrows)[:, None]
    return sig[idx], frame_length


def linear_filter_banks(nfilts=24, nfft=512, fs=16000, low_freq=0, high_freq=None,
                        scale="constant"):
    """spafe 0.3.x `spafe.fbanks.linear_fbanks.linear_filter_banks` -> (fbank, center_freqs).

    Edge frequencies `low + k * |high - low| / (nfilts + 1)`; bin frequencies
    `np.linspace(low_freq, high_freq, nfft//2 + 1)` (spafe spaces the bins over the
    band it was given, which equals the FFT bin centres for the default band
    0 .. fs/2); rising slope over lower <= f <= center, falling slope over
    center <= f <= upper, unit peak; `scale="constant"` leaves the heights at 1.
    """
    high_freq = high_freq or fs / 2
    low_freq = low_freq or 0
    if low_freq < 0 or high_freq > fs / 2:
        raise ValueError("bad frequency range")
    delta_hz = abs(high_freq - low_freq) / (nfilts + 1)
    scale_freqs = low_freq + delta_hz * np.arange(0, nfilts + 2)
    lower_edges, upper_edges, centers = scale_freqs[:-2], scale_freqs[2:], scale_freqs[1:-1]
    freqs = np.linspace(low_freq, high_freq, nfft // 2 + 1)
    fbank = np.zeros((nfilts, nfft // 2 + 1))
    c = 1.0 if scale in ("descendant", "constant") else 0.0
    for j, (center, lower, upper) in enumerate(zip(centers, lower_edges, upper_edges)):
        if scale == "descendant":
            c -= 1 / nfilts
            c = c * (c > 0) + 0 * (c < 0)
        elif scale == "ascendant":
            c += 1 / nfilts
            c = c * (c < 1) + 1 * (c > 1)
        left = (freqs >= lower) == (freqs <= center)
        fbank[j, left] = c * (freqs[left] - lower) / (center - lower)
        right = (freqs >= center) == (freqs <= upper)
        fbank[j, right] = c * (upper - freqs[right]) / (upper - center)
    return np.abs(fbank), centers


def linear_filter_banks_intbin(nfilts=24, nfft=512, fs=16000, low_freq=0, high_freq=None):
    """spafe 0.1.x (python_speech_features style): triangles on integer FFT bins, unit peak.

    bins = floor((nfft + 1) * hz / fs) for nfilts + 2 equally spaced edge
    frequencies; filter j rises over [b_j, b_{j+1}) and falls over [b_{j+1}, b_{j+2}).
    """
    high_freq = high_freq or fs / 2
    low_freq = low_freq or 0
    if low_freq < 0 or high_freq > fs / 2:
        raise ValueError("bad frequency range")
    linear_points = np.linspace(low_freq, high_freq, nfilts + 2)
    bins = np.floor((nfft + 1) * linear_points / fs)
    fbank = np.zeros([nfilts, nfft // 2 + 1])
    for j in range(nfilts):
        b0, b1, b2 = bins[j], bins[j + 1], bins[j + 2]
        fbank[j, int(b0):int(b1)] = (np.arange(int(b0), int(b1)) - int(b0)) / (b1 - b0)
        fbank[j, int(b1):int(b2)] = (int(b2) - np.arange(int(b1), int(b2))) / (b2 - b1)
    return np.abs(fbank)


def zero_handling(x):
    return np.where(x == 0, EPS, x)


def linear_spectrogram(sig, fs=16000, pre_emph=True, pre_emph_coeff=0.97,
                       win_len=0.025, win_hop=0.01, nfilts=24, nfft=512,
                       low_freq=0, high_freq=None, fbanks=None):
    """spafe.features.lfcc.linear_spectrogram -> (T, nfilts) float64 energies."""
    if fbanks is None:
        fbanks, _ = linear_filter_banks(nfilts, nfft, fs, low_freq, high_freq)
    sig = np.asarray(sig)
    if pre_emph:
        sig = pre_emphasis(sig, pre_emph_coeff)
    frames, frame_length = framing(sig, fs, win_len, win_hop)
    windows = np.hamming(frame_length) * frames
    mag = np.absolute(np.fft.fft(windows, nfft))[:, : nfft // 2 + 1]
    power = (1.0 / nfft) * np.square(mag)
    return np.dot(power, fbanks.T)


def lfcc(sig, fs=16000, num_ceps=13, pre_emph=True, pre_emph_coeff=0.97,
         win_len=0.025, win_hop=0.01, nfilts=24, nfft=512, low_freq=0,
         high_freq=None, fbanks=None):
    """spafe.features.lfcc.lfcc defaults (dct_type=2, no energy/lifter/normalise).

    Returns float64 (T, num_ceps), T = (L - frame_len)//frame_step + 1.
    """
    if nfilts < num_ceps:
        raise ValueError("nfilts must be >= num_ceps")
    feats = linear_spectrogram(sig, fs, pre_emph, pre_emph_coeff, win_len, win_hop,
                               nfilts, nfft, low_freq, high_freq, fbanks)
    log_feats = np.log(zero_handling(feats))
    return scipy.fft.dct(log_feats, type=2, axis=1, norm="ortho")[:, :num_ceps]


# ---------------------------------------------------------------------------------------------------------------
# spafe.features.gfcc.gfcc / spafe.fbanks.gammatone_fbanks (extract_gtcc, ASV_dl_func.py:484-499:
#   gtccs = gfcc(sig=y, fs=sr, num_ceps=n_ceps, nfilts=n_filters)   with the float waveform of librosa.load)
#
# PARITY UNPINNED, recalled from the 0.3.x source like the rest of this file.  The chain of 0.3.x `erb_spectrogram`
# has the same block as `linear_spectrogram` (|fft(windows, nfft)|[:, :nfft//2+1], then (1/nfft) * square) followed by
# `np.dot(abs_fft_values, fbanks.T)`, `np.power(features, 1/3)` and the ortho DCT-II; spafe 0.1.x multiplied the
# MAGNITUDE spectrum with the bank instead (`spectrum="magnitude"` here, AAD_SPEC_MAGNITUDE in the library).
# `gammatone_filter_banks` is Slaney's ERB filter bank (Auditory Toolbox MakeERBFilters) evaluated on the unit circle
# at the FFT bins, as in D. Ellis' fft2gammatonemx: least certain are the ERB order passed through as `order=4`
# (Slaney's own value is 1) and the per-filter normalisation to a maximum of 1.  A deployment that has spafe passes
# spafe's own matrix as AAD_FB_CUSTOM_DENSE, which makes both questions moot.
# ---------------------------------------------------------------------------------------------------------------
EAR_Q = 9.26449
MIN_BW = 24.7


def generate_center_frequencies(min_freq, max_freq, nfilts):
    """ERB-spaced centre frequencies (Slaney's ERBSpace), ascending."""
    m = np.arange(1, nfilts + 1)
    c = EAR_Q * MIN_BW
    cf = (max_freq + c) * np.exp((m / nfilts) * np.log((min_freq + c) / (max_freq + c))) - c
    return cf[::-1]


def gammatone_filter_banks(nfilts=24, nfft=512, fs=16000, low_freq=0, high_freq=None, order=4):
    """-> (fbank (nfilts, nfft//2 + 1) float64, centre frequencies); scale='constant'."""
    high_freq = high_freq or fs / 2
    low_freq = low_freq or 0
    T = 1.0 / fs
    u = np.exp(2j * np.pi * np.arange(nfft // 2 + 1) / nfft)[None, :]
    fcs = generate_center_frequencies(low_freq, high_freq, nfilts)
    erb = ((fcs / EAR_Q) ** order + MIN_BW ** order) ** (1.0 / order)
    B = 1.019 * 2 * np.pi * erb
    wT = 2 * fcs * np.pi * T
    K = np.exp(B * T)
    pole = (np.exp(1j * wT) / K)[:, None]
    smax, smin = np.sqrt(3 + 2 ** 1.5), np.sqrt(3 - 2 ** 1.5)
    A = [((np.cos(wT) + s * np.sin(wT)) / K)[:, None] for s in (smax, -smax, smin, -smin)]
    kj = np.exp(1j * wT)
    G = [2 * T * kj * (a[:, 0] - kj) for a in A]
    coe = -2 / K ** 2 - 2 * kj ** 2 + 2 * (1 + kj ** 2) / K
    gain = np.abs(G[0] * G[1] * G[2] * G[3] * coe ** -4)
    fb = (T ** 4 / gain[:, None]) * np.abs((u - A[0]) * (u - A[1]) * (u - A[2]) * (u - A[3])) \
        * np.abs((u - pole) * (u - pole.conj())) ** -4.0
    fb = fb / fb.max(axis=1, keepdims=True)          # "make sure all filters has max value = 1.0"
    return fb, fcs


def erb_spectrogram(sig, fs=16000, pre_emph=True, pre_emph_coeff=0.97, win_len=0.025, win_hop=0.01,
                    nfilts=24, nfft=512, low_freq=0, high_freq=None, fbanks=None, spectrum="power"):
    if fbanks is None:
        fbanks, _ = gammatone_filter_banks(nfilts, nfft, fs, low_freq, high_freq)
    sig = np.asarray(sig, dtype=np.float64)
    if pre_emph:
        sig = pre_emphasis(sig, pre_emph_coeff)
    frames, frame_length = framing(sig, fs, win_len, win_hop)
    windows = np.hamming(frame_length) * frames
    mag = np.absolute(np.fft.fft(windows, nfft))[:, : nfft // 2 + 1]
    spec = (1.0 / nfft) * np.square(mag) if spectrum == "power" else mag
    return np.dot(spec, fbanks.T)


def gfcc(sig, fs=16000, num_ceps=13, pre_emph=True, pre_emph_coeff=0.97, win_len=0.025, win_hop=0.01,
         nfilts=24, nfft=512, low_freq=0, high_freq=None, fbanks=None, spectrum="power"):
    """spafe.features.gfcc.gfcc defaults (dct_type=2, no energy / lifter / normalise) -> float64 (T, num_ceps)."""
    if nfilts < num_ceps:
        raise ValueError("nfilts must be >= num_ceps")
    feats = erb_spectrogram(sig, fs, pre_emph, pre_emph_coeff, win_len, win_hop, nfilts, nfft, low_freq, high_freq,
                            fbanks, spectrum)
    return scipy.fft.dct(np.power(feats, 1 / 3), type=2, axis=1, norm="ortho")[:, :num_ceps]
