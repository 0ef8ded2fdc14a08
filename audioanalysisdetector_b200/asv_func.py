"""Drop-ins for the older extractor signatures of the reference's `ASV_func.py` (the classical notebook
pipeline): the same three extractors as `ASV_dl_func.py`, but `mean=True` by default, no `augment`
argument, and LFCC averaged over TIME (`np.mean(lfccs, axis=0)`, ASV_func.py:70) instead of over the
coefficient axis (ASV_dl_func.py:436).

  extract_mfcc(filepath, chunk_start=None, chunk_end=None, sr=None, n_mfcc=13, mean=True)             ASV_func.py:43-56
  extract_lfcc(filepath, chunk_start=None, chunk_end=None, n_ceps=13, mean=True)                      ASV_func.py:59-73
  extract_mel_spectrogram(filepath, chunk_start=None, chunk_end=None, sr=None, n_mels=64, fmax=None, mean=True)  :142-156
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import extractors as _x
from .frontend import FrontendParams


def extract_mfcc(filepath, chunk_start=None, chunk_end=None, sr=None, n_mfcc=13, mean=True):
    return _x.extract_mfcc(filepath, chunk_start=chunk_start, chunk_end=chunk_end, sr=sr, n_mfcc=n_mfcc, mean=mean)


def extract_mel_spectrogram(filepath, chunk_start=None, chunk_end=None, sr=None, n_mels=64, fmax=None, mean=True):
    return _x.extract_mel_spectrogram(filepath, chunk_start=chunk_start, chunk_end=chunk_end, sr=sr, n_mels=n_mels,
                                      fmax=fmax, mean=mean)


def extract_lfcc(filepath, chunk_start=None, chunk_end=None, n_ceps=13, mean=True):
    try:
        y, sr = _x._prepare_clip(filepath, chunk_start, chunk_end, None, None)
        out, status = _x._run_batch(FrontendParams.lfcc(sr, n_ceps=n_ceps, time_mean=bool(mean)), [y])
        if out[0] is None:
            raise ValueError(L.ITEM_STATUS_NAMES.get(int(status[0]), "item failed"))
        return out[0].astype(np.float64)       # (n_ceps,) mean over time, or (T, n_ceps)
    except Exception as e:
        print(f"[BŁĄD LFCC] {filepath if isinstance(filepath, str) else '<array>'}: {e}")
        return None
