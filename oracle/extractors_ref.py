"""Oracle: the reference's per-clip extractors on in-memory waveforms.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Mirrors ASV_dl_func.py:404-439,522-538 after `librosa.load` (the decode itself is
outside the hot path): chunk slicing in samples, the library call, optional
time-mean, and the blanket try/except -> None error convention.
"""
from __future__ import annotations

import numpy as np

from . import librosa_ref, spafe_ref, delta_ref


def _slice(y, sr, chunk_start, chunk_end):
    # ASV_dl_func.py:408-411
    if chunk_start is not None and chunk_end is not None:
        start_sample = int(chunk_start * sr)
        end_sample = min(int(chunk_end * sr), len(y))
        y = y[start_sample:end_sample]
    return y


def extract_mel_spectrogram_ref(y, sr, chunk_start=None, chunk_end=None, n_mels=64,
                                fmax=None, mean=False, dtype="ref"):
    """ASV_dl_func.py:522-538 on a decoded float32 waveform -> (n_mels, T) float32."""
    try:
        y = _slice(np.asarray(y, dtype=np.float32), sr, chunk_start, chunk_end)
        S_db = librosa_ref.logmel_db(y, sr, n_mels=n_mels, fmax=fmax, dtype=dtype)
        return np.mean(S_db, axis=1) if mean else S_db
    except Exception:
        return None


def compute_melspec_ref(y, sr, n_mels=128, hop_length=512, n_fft=2048, dtype="ref"):
    """ASV_dataset.ipynb:1151 (cell [27]) `compute_melspec` on a decoded waveform: log-mel
    (melspectrogram(power=2.0) -> power_to_db(ref=np.max)) followed by a global z-normalisation
    `(S - S.mean()) / S.std()` over the whole (n_mels, T) matrix (numpy: population std)."""
    S_db = librosa_ref.logmel_db(np.asarray(y, dtype=np.float32), sr, n_mels=n_mels, n_fft=n_fft,
                                 hop_length=hop_length, dtype=dtype)
    return (S_db - S_db.mean()) / S_db.std()


def extract_mfcc_ref(y, sr, chunk_start=None, chunk_end=None, n_mfcc=13, mean=False,
                     dtype="ref"):
    """ASV_dl_func.py:404-420 on a decoded float32 waveform -> (n_mfcc, T) float32."""
    try:
        y = _slice(np.asarray(y, dtype=np.float32), sr, chunk_start, chunk_end)
        feat = librosa_ref.mfcc(y, sr, n_mfcc=n_mfcc, dtype=dtype)
        return np.mean(feat, axis=1) if mean else feat
    except Exception:
        return None


def extract_lfcc_ref(y, sr, chunk_start=None, chunk_end=None, n_ceps=13, mean=False,
                     mean_axis=1):
    """ASV_dl_func.py:423-439 -> (T, n_ceps) float64.  `mean_axis=1` is what
    ASV_dl_func.py:436 does; ASV_func.py:70 / train_fun.py:86 use axis 0."""
    try:
        y = _slice(np.asarray(y, dtype=np.float32), sr, chunk_start, chunk_end)
        y_int16 = spafe_ref.quantize_int16(y)
        lf = spafe_ref.lfcc(sig=y_int16, fs=sr, num_ceps=n_ceps)
        return np.mean(lf, axis=mean_axis) if mean else lf
    except Exception:
        return None


def extract_gtcc_ref(y, sr, chunk_start=None, chunk_end=None, n_filters=40, n_ceps=13, mean=False, **kw):
    """ASV_dl_func.py:484-499 -> (T, n_ceps) float64 (mean=True: axis 1, as the reference writes it)."""
    try:
        y = _slice(np.asarray(y, dtype=np.float32), sr, chunk_start, chunk_end)
        g = spafe_ref.gfcc(sig=y, fs=sr, num_ceps=n_ceps, nfilts=n_filters, **kw)
        return np.mean(g, axis=1) if mean else g
    except Exception:
        return None


def mfcc_with_deltas_ref(y, sr, n_mfcc=40, n_fft=2048, hop_length=512, n_mels=128,
                         n_delta=2, width=9, dtype="ref"):
    """BASELINE.json configs[1]: MFCC + delta + delta-delta -> (3*n_mfcc, T)."""
    feat = librosa_ref.mfcc(np.asarray(y, dtype=np.float32), sr, n_mfcc=n_mfcc,
                            n_fft=n_fft, hop_length=hop_length, n_mels=n_mels, dtype=dtype)
    return delta_ref.stack_deltas(feat, n_delta=n_delta, width=width)


def lfcc_with_deltas_ref(sig, sr, num_ceps=20, nfilts=20, nfft=512, win_len=0.02,
                         win_hop=0.01, n_delta=2, width=9, fbanks=None):
    """BASELINE.json configs[2]: ASVspoof-LA style LFCC + deltas -> (3*num_ceps, T).

    `sig` is int16 PCM (or anything spafe.lfcc would accept)."""
    lf = spafe_ref.lfcc(sig=sig, fs=sr, num_ceps=num_ceps, nfilts=nfilts, nfft=nfft,
                        win_len=win_len, win_hop=win_hop, fbanks=fbanks)     # (T, C)
    return delta_ref.stack_deltas(np.ascontiguousarray(lf.T), n_delta=n_delta, width=width)
