"""Oracle: numpy/scipy restatement of the librosa 0.11 calls on the hot path.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED (librosa
is not installed in this image and the reference has no fixtures); every
function below restates the published librosa 0.11.0 algorithm and is
cross-checked in tests/test_oracle_librosa.py against torchaudio / scipy and
analytic answers.

Reference call sites that reach this arithmetic:
  ASV_dl_func.py:533  librosa.feature.melspectrogram(y=y, sr=sr, n_mels=n_mels, fmax=fmax or sr/2)
  ASV_dl_func.py:534  librosa.power_to_db(S, ref=np.max)
  ASV_dl_func.py:416  librosa.feature.mfcc(y=y, sr=sr, n_mfcc=n_mfcc)
  ASV_func.py:52,151-152 ; train_fun.py:73   (same calls)

dtype chain (librosa 0.11): y float32 -> frames float32 * window float64 ->
rfft in float64 -> stored complex64 -> |.| float32 -> **2 float32 -> mel basis
float32 -> einsum float32 -> power_to_db float32 -> DCT float32.
`dtype="f64"` below switches every step to float64 (the "truth" variant).
"""
from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.signal

# ----------------------------------------------------------------------------
# mel scale (librosa.core.convert.hz_to_mel / mel_to_hz, htk=False -> Slaney)
# ----------------------------------------------------------------------------
_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(freq, htk=False):
    freq = np.asanyarray(freq, dtype=np.float64)
    if htk:
        return 2595.0 * np.log10(1.0 + freq / 700.0)
    mels = freq / _F_SP
    log_t = freq >= _MIN_LOG_HZ
    if freq.ndim:
        mels[log_t] = _MIN_LOG_MEL + np.log(freq[log_t] / _MIN_LOG_HZ) / _LOGSTEP
    elif log_t:
        mels = _MIN_LOG_MEL + np.log(freq / _MIN_LOG_HZ) / _LOGSTEP
    return mels


def mel_to_hz(mels, htk=False):
    mels = np.asanyarray(mels, dtype=np.float64)
    if htk:
        return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    freqs = _F_SP * mels
    log_t = mels >= _MIN_LOG_MEL
    if mels.ndim:
        freqs[log_t] = _MIN_LOG_HZ * np.exp(_LOGSTEP * (mels[log_t] - _MIN_LOG_MEL))
    elif log_t:
        freqs = _MIN_LOG_HZ * np.exp(_LOGSTEP * (mels - _MIN_LOG_MEL))
    return freqs


def mel_frequencies(n_mels, fmin, fmax, htk=False):
    """librosa.mel_frequencies: n_mels points uniformly spaced on the mel axis."""
    mels = np.linspace(hz_to_mel(fmin, htk), hz_to_mel(fmax, htk), n_mels)
    return mel_to_hz(mels, htk)


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False,
                   norm="slaney", dtype=np.float32):
    """librosa.filters.mel -> (n_mels, 1 + n_fft//2) of `dtype`.

    Triangles are built in float64 on rfftfreq bin centres, stored into a
    `dtype` array, then area-normalised in place (`weights *= enorm`), which for
    dtype=float32 rounds twice -- reproduced here.
    """
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=dtype)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_frequencies(n_mels + 2, fmin, fmax, htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if norm == "slaney":
        enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
        weights *= enorm[:, np.newaxis]
    elif norm is not None:
        raise ValueError("only norm='slaney' or None restated")
    return weights


# ----------------------------------------------------------------------------
# STFT (librosa.stft defaults: hann periodic, center=True, pad_mode='constant')
# ----------------------------------------------------------------------------
def n_frames_centered(length, hop_length):
    return 1 + int(length) // int(hop_length)


def stft(y, n_fft=2048, hop_length=None, win_length=None, window="hann",
         center=True, dtype="ref"):
    """librosa.stft -> complex (1 + n_fft//2, T).

    dtype="ref": float32 samples, float64 window, float64 rfft, result rounded to
    complex64 (what librosa 0.11 does).  dtype="f64": everything float64.
    """
    y = np.asarray(y)
    if y.ndim != 1:
        raise ValueError("mono only")
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = win_length // 4
    if y.size == 0:
        raise ValueError("empty signal")
    if not np.isfinite(y).all():
        raise ValueError("Audio buffer is not finite everywhere")
    fft_window = scipy.signal.get_window(window, win_length, fftbins=True)  # float64
    if win_length < n_fft:  # util.pad_center
        lpad = (n_fft - win_length) // 2
        fft_window = np.pad(fft_window, (lpad, n_fft - win_length - lpad))
    if dtype == "ref":
        y = y.astype(np.float32, copy=False)
    else:
        y = y.astype(np.float64)
    if center:
        y = np.pad(y, (n_fft // 2, n_fft // 2), mode="constant")
    elif y.size < n_fft:
        raise ValueError("too short")
    T = 1 + (y.size - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(T)[None, :]
    frames = y[idx]                                   # (n_fft, T)
    spec = scipy.fft.rfft(fft_window[:, None] * frames, axis=0)   # float64 math
    if dtype == "ref":
        spec = spec.astype(np.complex64)
    return spec


def power_spectrogram(y, n_fft=2048, hop_length=512, win_length=None,
                      window="hann", center=True, dtype="ref"):
    """librosa.core.spectrum._spectrogram with power=2: np.abs(D)**2.0."""
    D = stft(y, n_fft, hop_length, win_length, window, center, dtype)
    return np.abs(D) ** 2.0


def melspectrogram(y, sr, n_fft=2048, hop_length=512, win_length=None,
                   window="hann", center=True, n_mels=128, fmin=0.0, fmax=None,
                   htk=False, norm="slaney", dtype="ref"):
    """librosa.feature.melspectrogram(power=2.0) -> (n_mels, T)."""
    S = power_spectrogram(y, n_fft, hop_length, win_length, window, center, dtype)
    fb_dtype = np.float32 if dtype == "ref" else np.float64
    mel_basis = mel_filterbank(sr, n_fft, n_mels, fmin, fmax, htk, norm, fb_dtype)
    return np.einsum("ft,mf->mt", S, mel_basis, optimize=True)


def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
    """librosa.power_to_db.  `ref` may be a callable (np.max) or a scalar."""
    S = np.asarray(S)
    magnitude = S
    if callable(ref):
        ref_value = ref(magnitude)
    else:
        ref_value = np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        if top_db < 0:
            raise ValueError("top_db must be non-negative")
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def dct_ortho(x, axis, n_out=None):
    """scipy.fftpack.dct(x, type=2, norm='ortho', axis=axis)[:n_out] along axis."""
    out = scipy.fft.dct(x, type=2, norm="ortho", axis=axis)
    if n_out is not None:
        sl = [slice(None)] * out.ndim
        sl[axis] = slice(0, n_out)
        out = out[tuple(sl)]
    return out


def mfcc(y, sr, n_mfcc=20, n_fft=2048, hop_length=512, n_mels=128, fmin=0.0,
         fmax=None, win_length=None, dtype="ref"):
    """librosa.feature.mfcc(dct_type=2, norm='ortho', lifter=0) -> (n_mfcc, T).

    S = power_to_db(melspectrogram(...))  (ref=1.0, amin=1e-10, top_db=80), then
    scipy DCT-II ortho along the mel axis, first n_mfcc rows.
    """
    S = power_to_db(melspectrogram(y, sr, n_fft=n_fft, hop_length=hop_length,
                                   win_length=win_length, n_mels=n_mels, fmin=fmin,
                                   fmax=fmax, dtype=dtype))
    return dct_ortho(S, axis=-2, n_out=n_mfcc)


def logmel_db(y, sr, n_mels=64, n_fft=2048, hop_length=512, fmax=None, dtype="ref"):
    """What extract_mel_spectrogram computes after loading (ASV_dl_func.py:533-534)."""
    S = melspectrogram(y, sr, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels,
                       fmax=fmax or sr / 2, dtype=dtype)
    return power_to_db(S, ref=np.max)
