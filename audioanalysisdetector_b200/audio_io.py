"""Minimal audio decode for the drop-in extractors (host side, outside the hot path).

The reference decodes with `librosa.load(filepath, sr=sr)` (ASV_dl_func.py:406,425,524):
float32 mono at the file's native rate when sr is None.  soundfile/librosa are not part
of this image, so PCM WAV is decoded with the standard library; other containers go
through `soundfile` when it is importable.  A custom loader can be installed with
`set_loader(fn)` where fn(path) -> (float32 mono ndarray, sample_rate).
"""
from __future__ import annotations

import wave
from typing import Callable, Optional, Tuple

import numpy as np

_loader: Optional[Callable[[str], Tuple[np.ndarray, int]]] = None


def set_loader(fn: Optional[Callable[[str], Tuple[np.ndarray, int]]]):
    global _loader
    _loader = fn


def _load_wav(path: str, keep_pcm16: bool = False) -> Tuple[np.ndarray, int]:
    with wave.open(path, "rb") as w:
        sr, nch, sw, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(n)
    if sw == 2 and nch == 1 and keep_pcm16:
        return np.frombuffer(raw, dtype="<i2").copy(), sr    # sample value = int16 / 32768 (as decoded below)
    if sw == 2:
        y = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif sw == 4:
        y = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif sw == 1:
        y = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"unsupported PCM sample width {sw}")
    if nch > 1:
        y = y.reshape(-1, nch).mean(axis=1).astype(np.float32)   # librosa.to_mono
    return y, sr


def load_pcm(source) -> Tuple[np.ndarray, int]:
    """Like `load` at the native rate, but mono 16-bit PCM WAV files come back as the raw int16 samples
    (value = int16 / 32768, exactly what `load` would return as float32): half the bytes to upload."""
    if _loader is None and not isinstance(source, tuple) and str(source).lower().endswith(".wav"):
        return _load_wav(str(source), keep_pcm16=True)
    return load(source)


def load(source, sr: Optional[int] = None) -> Tuple[np.ndarray, int]:
    """-> (float32 mono waveform, sample_rate).

    `source` is a file path, or an in-memory clip given as (waveform, sample_rate)."""
    if isinstance(source, tuple) and len(source) == 2:
        y, native = np.asarray(source[0], dtype=np.float32), int(source[1])
    elif _loader is not None:
        y, native = _loader(source)
        y = np.asarray(y, dtype=np.float32)
    else:
        path = str(source)
        if path.lower().endswith(".wav"):
            y, native = _load_wav(path)
        else:
            try:
                import soundfile as sf  # noqa: WPS433 (optional dependency)
            except ImportError as e:  # pragma: no cover
                raise RuntimeError(f"cannot decode {path}: soundfile is not installed") from e
            y, native = sf.read(path, dtype="float32", always_2d=False)
            if y.ndim > 1:
                y = y.mean(axis=1).astype(np.float32)
    if sr is not None and int(sr) != native:
        raise ValueError(f"resampling {native} -> {sr} Hz is outside the front-end (load at native rate)")
    return y, native
