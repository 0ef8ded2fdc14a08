"""Oracle: numpy restatement of the reference's CQCC extractor (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

PARITY UNPINNED, and not pinnable here: `extract_cqcc` (ASV_dl_func.py:442-481) is

    fmin   = librosa.note_to_hz('C1');  fmax = sr / 2 - 100
    n_bins = int(np.floor(np.log2(fmax / fmin)) * bins_per_octave)
    cqt    = librosa.cqt(y, sr=sr, n_bins=n_bins, bins_per_octave=bins_per_octave, fmin=fmin)      # hop 512
    cqt_db = librosa.amplitude_to_db(np.abs(cqt), ref=np.max)
    interp_cqt[:, t] = interp1d(cqt_frequencies, cqt_db[:, t], 'linear', fill_value='extrapolate')(linspace(f0, f1, n_bins))
    cqcc   = dct(np.log(np.square(interp_cqt) + 1e-12), type=2, axis=0, norm='ortho')[:n_ceps]

and `librosa.cqt` (0.10 / 0.11: `vqt` with gamma = 0) halves the sample rate octave by octave with
`res_type='soxr_hq'`.  The soxr library is not in this image and its polyphase filters cannot be restated, so
`resample2` below is a documented stand-in (Kaiser-windowed sinc half-band, 255 taps, > 120 dB stop band, flat to
0.8 of the new Nyquist; the CQT bins of an octave lie below 0.55 of it).  Everything else follows librosa's published
algorithm [recalled from the 0.10.x source; the reference pins ~= 0.11.0]:

  wavelet_lengths   Q = filter_scale / alpha, alpha = (2^(2/bpo) - 1) / (2^(2/bpo) + 1); lengths = Q sr / f
  wavelet           exp(+2 pi i f n / sr) for n in arange(-l//2, l//2), times a hann window of fractional length
                    (`__float_window`), L1-normalised, centred in a buffer of 2^ceil(log2(max length))
  __vqt_filter_fft  basis *= lengths / n_fft;  fft;  keep bins 0 .. n_fft/2;  sparsify_rows(quantile = 0.01)
  vqt               per octave (top first): fft_basis *= sqrt(sr / my_sr); response = fft_basis . stft(my_y, n_fft,
                    hop = my_hop, window = ones, centred, zero padding); then my_hop //= 2, my_sr /= 2,
                    my_y = resample(my_y, 2 -> 1, scale=True) (i.e. times sqrt 2); stack, trim to the shortest octave,
                    divide by sqrt(lengths) (scale=True)
  amplitude_to_db   20 log10(max(1e-5, S)) - 20 log10(max(1e-5, max S)), floored at -80 dB  (float32)

What the CUDA path is tested against is THIS restatement (same stand-in resampler on both sides); how close both
are to librosa + soxr is not established.  What IS established without librosa: the recursion agrees with the constant-Q
transform computed directly from its definition at the original rate (no octaves, no resampler, no sparsification)
within 1-4 % of the peak, equally in resampled and never-resampled octaves (tests/test_oracle_cqcc.py).  Note also that log(dB^2 + 1e-12) has derivative 2/|dB|: cells within a
few hundredths of a dB of the utterance maximum amplify any difference between implementations.
"""
from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.signal

FMIN_C1 = 32.70319566257483      # librosa.note_to_hz('C1')
HOP = 512
RESAMPLE_TAPS = 255
RESAMPLE_BETA = 14.0


def cqt_frequencies(n_bins, fmin, bins_per_octave=12):
    return fmin * 2.0 ** (np.arange(n_bins, dtype=np.float64) / bins_per_octave)


def n_bins_for(sr, bins_per_octave=12, fmin=FMIN_C1):
    fmax = sr / 2 - 100
    return int(np.floor(np.log2(fmax / fmin)) * bins_per_octave)


def relative_bandwidth(bins_per_octave):
    r = 2.0 ** (2.0 / bins_per_octave)
    return (r - 1) / (r + 1)


def wavelet_lengths(freqs, sr, filter_scale=1.0, bins_per_octave=12):
    """-> (lengths in samples (fractional), filter_cutoff); hann window bandwidth 1.50018310546875"""
    alpha = relative_bandwidth(bins_per_octave)
    q = float(filter_scale) / alpha
    lengths = q * sr / freqs
    cutoff = np.max(freqs * (1 + 0.5 * 1.50018310546875 / q))
    return lengths, cutoff


def float_window_hann(n):
    """librosa.filters.__float_window('hann'): periodic hann of floor(n) samples, zero-extended to ceil(n)."""
    n_min, n_max = int(np.floor(n)), int(np.ceil(n))
    w = scipy.signal.get_window("hann", n_min, fftbins=True)
    if len(w) < n_max:
        w = np.pad(w, [(0, n_max - len(w))], mode="constant")
    w[n_min:] = 0.0
    return w


def wavelet_basis(freqs, sr, bins_per_octave=12, filter_scale=1.0):
    """librosa.filters.wavelet(..., pad_fft=True, norm=1): complex (n_filters, n_fft) and the lengths."""
    lengths, _ = wavelet_lengths(freqs, sr, filter_scale, bins_per_octave)
    filters = []
    for ilen, f in zip(lengths, freqs):
        n = np.arange(-ilen // 2, ilen // 2, dtype=float)
        sig = np.exp(2j * np.pi * f * n / sr)
        sig = sig * float_window_hann(len(sig))
        sig = sig / np.sum(np.abs(sig))                       # util.normalize(norm=1)
        filters.append(sig)
    max_len = int(2.0 ** np.ceil(np.log2(max(lengths))))
    out = np.zeros((len(filters), max_len), dtype=np.complex128)
    for i, sig in enumerate(filters):
        lpad = (max_len - len(sig)) // 2                      # util.pad_center
        out[i, lpad:lpad + len(sig)] = sig
    return out, lengths


def sparsify_rows(x, quantile=0.01):
    mags = np.abs(x)
    norms = np.sum(mags, axis=1, keepdims=True)
    mag_sort = np.sort(mags, axis=1)
    cumulative = np.cumsum(mag_sort / norms, axis=1)
    thr_idx = np.argmin(cumulative < quantile, axis=1)
    out = np.zeros_like(x)
    for i in range(x.shape[0]):
        keep = mags[i] >= mag_sort[i, thr_idx[i]]
        out[i, keep] = x[i, keep]
    return out


def octave_fft_basis(freqs_oct, my_sr, sr, bins_per_octave=12, sparsity=0.01):
    """__vqt_filter_fft + the sqrt(sr / my_sr) rescale of vqt -> (fft_basis complex64 (12, n_fft/2+1), n_fft)."""
    basis, lengths = wavelet_basis(freqs_oct, my_sr, bins_per_octave)
    n_fft = basis.shape[1]
    basis = basis * (lengths[:, None] / float(n_fft))
    fft_basis = scipy.fft.fft(basis, n=n_fft, axis=1)[:, : n_fft // 2 + 1]
    fft_basis = sparsify_rows(fft_basis, quantile=sparsity)
    return (fft_basis * np.sqrt(sr / my_sr)).astype(np.complex64), n_fft


def resample_taps():
    """The stand-in for soxr_hq's 2 -> 1 stage: windowed-sinc low-pass at a quarter of the input rate."""
    n = np.arange(RESAMPLE_TAPS) - (RESAMPLE_TAPS - 1) / 2
    h = 0.5 * np.sinc(0.5 * n) * np.kaiser(RESAMPLE_TAPS, RESAMPLE_BETA)
    return h / h.sum()


def resample2(y):
    """y at rate 2 -> rate 1 with librosa.resample(scale=True) semantics: ceil(len / 2) samples, times sqrt(2)."""
    h = resample_taps().astype(np.float32)
    n_out = (len(y) + 1) // 2
    c = (RESAMPLE_TAPS - 1) // 2
    ypad = np.concatenate([np.zeros(c, np.float32), y.astype(np.float32), np.zeros(c + 2, np.float32)])
    idx = 2 * np.arange(n_out)[:, None] + np.arange(RESAMPLE_TAPS)[None, :]
    out = (ypad[idx].astype(np.float64) * h[None, :].astype(np.float64)).sum(axis=1)
    return (np.sqrt(2.0) * out).astype(np.float32)


def stft_ones(y, n_fft, hop):
    """librosa.stft(window='ones', center=True, pad_mode='constant') -> complex64 (n_fft/2+1, 1 + len(y)//hop)"""
    ypad = np.concatenate([np.zeros(n_fft // 2, np.float32), y.astype(np.float32), np.zeros(n_fft // 2, np.float32)])
    t = 1 + len(y) // hop
    idx = hop * np.arange(t)[:, None] + np.arange(n_fft)[None, :]
    if idx.max() >= len(ypad):
        raise ValueError("signal too short for the CQT")
    return np.fft.rfft(ypad[idx].astype(np.float64), axis=1).T.astype(np.complex64)


def cqt(y, sr, n_bins=None, bins_per_octave=12, fmin=FMIN_C1, hop_length=HOP):
    """|.|-ready complex CQT (n_bins, T) following librosa.cqt's defaults (see the module docstring)."""
    y = np.asarray(y, dtype=np.float32)
    if n_bins is None:
        n_bins = n_bins_for(sr, bins_per_octave, fmin)
    if n_bins <= 0 or len(y) == 0:
        raise ValueError("nothing to transform")
    n_octaves = int(np.ceil(float(n_bins) / bins_per_octave))
    n_filters = min(bins_per_octave, n_bins)
    freqs = cqt_frequencies(n_bins, fmin, bins_per_octave)
    lengths, cutoff = wavelet_lengths(freqs, sr, 1.0, bins_per_octave)
    if cutoff > sr / 2:
        raise ValueError("wavelet basis exceeds the Nyquist frequency")
    resp = []
    my_y, my_sr, my_hop = y, float(sr), int(hop_length)
    for i in range(n_octaves):
        sl = slice(-n_filters, None) if i == 0 else slice(-n_filters * (i + 1), -n_filters * i)
        fft_basis, n_fft = octave_fft_basis(freqs[sl], my_sr, sr, bins_per_octave)
        d = stft_ones(my_y, n_fft, my_hop)
        resp.append((fft_basis.astype(np.complex128) @ d.astype(np.complex128)).astype(np.complex64))
        if my_hop % 2 == 0:
            my_hop //= 2
            my_sr /= 2.0
            my_y = resample2(my_y)
    max_col = min(r.shape[1] for r in resp)
    out = np.zeros((n_bins, max_col), dtype=np.complex64)
    end = n_bins
    for r in resp:                                            # __trim_stack: octaves come top first
        n_oct = r.shape[0]
        if end < n_oct:
            out[:end] = r[-end:, :max_col]
        else:
            out[end - n_oct:end] = r[:, :max_col]
        end -= n_oct
    return out / np.sqrt(lengths)[:, None].astype(np.float32)


def amplitude_to_db(mag, amin=1e-5, top_db=80.0):
    mag = np.asarray(mag, dtype=np.float32)
    ref = np.float32(mag.max())
    p = np.maximum(np.float32(amin) ** 2, mag * mag)
    log_spec = (10.0 * np.log10(p) - 10.0 * np.log10(np.maximum(np.float32(amin) ** 2, ref * ref))).astype(np.float32)
    return np.maximum(log_spec, log_spec.max() - np.float32(top_db))


def interp_to_linear_freqs(cqt_db, freqs):
    """scipy.interpolate.interp1d(kind='linear') per frame onto np.linspace(f.min(), f.max(), n_bins), cast to float32"""
    lin = np.linspace(freqs.min(), freqs.max(), num=len(freqs))
    hi = np.clip(np.searchsorted(freqs, lin, side="left"), 1, len(freqs) - 1)
    lo = hi - 1
    w = ((lin - freqs[lo]) / (freqs[hi] - freqs[lo]))[:, None]
    x = cqt_db.astype(np.float64)
    return (x[lo] + w * (x[hi] - x[lo])).astype(np.float32)


def cqcc(y, sr, bins_per_octave=12, n_ceps=19):
    """ASV_dl_func.py:453-471 -> float32 (n_ceps, T), T = 1 + len(y) // 512."""
    n_bins = n_bins_for(sr, bins_per_octave)
    c = cqt(y, sr, n_bins=n_bins, bins_per_octave=bins_per_octave)
    db = amplitude_to_db(np.abs(c))
    interp = interp_to_linear_freqs(db, cqt_frequencies(n_bins, FMIN_C1, bins_per_octave))
    log_power = np.log(np.square(interp) + np.float32(1e-12)).astype(np.float32)
    return scipy.fft.dct(log_power, type=2, axis=0, norm="ortho")[:n_ceps].astype(np.float32)


def extract_cqcc_ref(y, sr, chunk_start=None, chunk_end=None, bins_per_octave=12, n_ceps=19, mean=False):
    """Reference call surface (in-memory clip instead of a path): ndarray, or None where the reference's blanket
    try/except would have returned None."""
    try:
        y = np.asarray(y, dtype=np.float32)
        if chunk_start is not None and chunk_end is not None:
            y = y[int(chunk_start * sr):min(int(chunk_end * sr), len(y))]
        out = cqcc(y, sr, bins_per_octave, n_ceps)
        return np.mean(out, axis=1) if mean else out
    except Exception:
        return None
